// Microbenchmark: FP64 issue behaviour on sm_100a.  DMMA (mma.sync.m8n8k4.f64) and DFMA throughput per SM
// as a function of resident warps per SM and independent accumulator chains per warp (ILP).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb_dmma mb_dmma.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dmma_k(double *out, int iters)
{
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0000001;
    double c[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; i++) { c[i][0] = 0; c[i][1] = 0; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void dfma_k(double *out, int iters)
{
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) a[i] = threadIdx.x * 1e-9 + i;
    const double m = 1.0000001, c = 1e-7;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) a[i] = fma(a[i], m, c);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
double time_ms(F launch)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch();
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    const double clk = prop.clockRate * 1e3;
    double *d;
    cudaMalloc(&d, sizeof(double) * sms * 2048);
    const int iters = 1 << 14;
    printf("SMs %d clock %.0f MHz (nominal)\n", sms, clk / 1e6);
    printf("kind warps/SM ILP  per-SM-per-clk(FMA)  cycles-per-instr-per-SMSP  TFLOP/s\n");
    for (int wps : {4, 8, 16, 32, 64}) {
        const int threads = wps * 32 > 1024 ? 1024 : wps * 32, blocks = sms * (wps * 32 / threads);
#define RUN_DMMA(I) { double ms = time_ms([&] { dmma_k<I><<<blocks, threads>>>(d, iters); }); \
        double n = (double)iters * I * wps * sms; double fma = n * 256; \
        printf("DMMA %2d %d  %8.2f  %8.2f  %6.2f\n", wps, I, fma / (ms * 1e-3 * clk * sms), (ms * 1e-3 * clk) / ((double)iters * I * wps / 4.0), 2 * fma / (ms * 1e-3) / 1e12); }
        RUN_DMMA(1) RUN_DMMA(2) RUN_DMMA(4) RUN_DMMA(8)
#define RUN_DFMA(I) { double ms = time_ms([&] { dfma_k<I><<<blocks, threads>>>(d, iters); }); \
        double n = (double)iters * I * wps * sms; double fma = n * 32; \
        printf("DFMA %2d %d  %8.2f  %8.2f  %6.2f\n", wps, I, fma / (ms * 1e-3 * clk * sms), (ms * 1e-3 * clk) / ((double)iters * I * wps / 4.0), 2 * fma / (ms * 1e-3) / 1e12); }
        RUN_DFMA(1) RUN_DFMA(2) RUN_DFMA(4) RUN_DFMA(8)
    }
    return 0;
}
