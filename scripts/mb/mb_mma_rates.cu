// Issue rate of the legacy (mma.sync) tensor-core instructions on this GPU, per SM sub-partition: one warp per
// sub-partition (4 warps per block, one block per SM), ILP independent accumulator chains, cycles per instruction.
// experiments only.  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb_mma_rates mb_mma_rates.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int KIND, int ILP>
__global__ void rate_kernel(int iters, long long *cycles, int *sink)
{
    int acc_i[ILP][4]; float acc_f[ILP][4]; double acc_d[ILP][2];
    for (int i = 0; i < ILP; i++) { for (int j = 0; j < 4; j++) { acc_i[i][j] = 0; acc_f[i][j] = 0.f; } acc_d[i][0] = acc_d[i][1] = 0.0; }
    int a0 = threadIdx.x * 0x01010101, a1 = a0 ^ 0x7f, b0 = 0x01ff0001, b1 = 0x00010100;
    double da = 1.0 + threadIdx.x, db = 0.5;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                    : "+r"(acc_i[i][0]), "+r"(acc_i[i][1]), "+r"(acc_i[i][2]), "+r"(acc_i[i][3]) : "r"(a0), "r"(a1), "r"(a0), "r"(a1), "r"(b0), "r"(b1));
            else if (KIND == 1)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                    : "+f"(acc_f[i][0]), "+f"(acc_f[i][1]), "+f"(acc_f[i][2]), "+f"(acc_f[i][3]) : "r"(a0), "r"(a1), "r"(a0), "r"(a1), "r"(b0), "r"(b1));
            else if (KIND == 2)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                    : "+d"(acc_d[i][0]), "+d"(acc_d[i][1]) : "d"(da), "d"(db));
            else if (KIND == 3)
                asm volatile("mma.sync.aligned.m16n8k32.row.col.f32.e4m3.e4m3.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                    : "+f"(acc_f[i][0]), "+f"(acc_f[i][1]), "+f"(acc_f[i][2]), "+f"(acc_f[i][3]) : "r"(a0), "r"(a1), "r"(a0), "r"(a1), "r"(b0), "r"(b1));
            else
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                    : "+f"(acc_f[i][0]), "+f"(acc_f[i][1]), "+f"(acc_f[i][2]), "+f"(acc_f[i][3]) : "r"(a0), "r"(a1), "r"(a0), "r"(a1), "r"(b0), "r"(b1));
        }
    }
    const long long t1 = clock64();
    int s = 0;
    for (int i = 0; i < ILP; i++) s += acc_i[i][0] + (int)acc_f[i][0] + (int)acc_d[i][0];
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
    if (s == 123456789) *sink = s;
}

template <int KIND, int ILP>
void run(const char *name, double macs)
{
    long long *d_c; int *d_s; cudaMalloc(&d_c, 8); cudaMalloc(&d_s, 4);
    const int iters = 2000;
    for (int warps : {4, 8, 16}) {
        rate_kernel<KIND, ILP><<<148, 32 * warps>>>(iters, d_c, d_s);
        rate_kernel<KIND, ILP><<<148, 32 * warps>>>(iters, d_c, d_s);
        cudaDeviceSynchronize();
        long long c = 0; cudaMemcpy(&c, d_c, 8, cudaMemcpyDeviceToHost);
        const double per = (double)c / ((double)iters * ILP);          // cycles per instruction per warp
        const double per_smsp = per / (warps / 4.0);
        printf("%-28s ILP %d warps/SM %2d: %7.2f cycles per instruction per sub-partition  -> %8.1f TMAC/s per GPU at 1.965 GHz (%s)\n",
               name, ILP, warps, per_smsp, macs / per_smsp * 4 * 148 * 1.965e9 / 1e12, cudaGetErrorString(cudaGetLastError()));
    }
}

int main()
{
    run<0, 8>("IMMA m16n8k32 s8", 16.0 * 8 * 32);
    run<1, 8>("HMMA m16n8k16 bf16->f32", 16.0 * 8 * 16);
    run<3, 8>("QMMA m16n8k32 e4m3->f32", 16.0 * 8 * 32);
    run<4, 8>("HMMA m16n8k8 tf32->f32", 16.0 * 8 * 8);
    run<2, 8>("DMMA m8n8k4 f64", 8.0 * 8 * 4);
    return 0;
}
