// Micro-benchmark of contraction variants: 296 blocks (2 per SM) x 256 threads, every block repeatedly
// contracts its own fold matrix (int8, N x K) with R right-hand sides, like FullStat does.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I pareben_b200/csrc -o /tmp/mb scripts/mb/mb_contract.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "common.cuh"
#include "gauss_fit.cuh"
#include "contract_v7.cuh"
using namespace pareben;

__global__ void __launch_bounds__(256, 2) k_v4(const int8_t *X8, int ld8, int N, int K, int R, double *Vall, int ldv, double *Gall, int reps, const double *phi)
{
    __shared__ __align__(32) double sV[SV_DOUBLES];
    double *V = Vall + (size_t)blockIdx.x * (N + 64) * ldv;
    double *G = Gall + (size_t)blockIdx.x * R * K;
    for (int rep = 0; rep < reps; rep++)
        contract<false, int8_t>(X8, N, K, K, R, [&](int r, int h) { return phi[(size_t)r * N + h]; },
                                [&](int r, int c, double a) { G[(size_t)r * K + c] = a; }, V, ldv, sV, false);
}

__global__ void __launch_bounds__(256, 2) k_v7(const int8_t *X8, int ld8, int N, int K, int R, double *Vall, int ldv, double *Gall, int reps, const double *phi, double *credall)
{
    extern __shared__ __align__(128) unsigned char dsm[];
    double *V = Vall + (size_t)blockIdx.x * (N + 64) * ldv;
    double *G = Gall + (size_t)blockIdx.x * R * K;
    double *cred = credall + (size_t)blockIdx.x * 8192;
    const int Rp = (R + 7) / 8 * 8, Np = (N + 63) / 64 * 64;
    for (int rep = 0; rep < reps; rep++) {
        __syncthreads();
        for (int r = 0; r < Rp; r += 4)
            for (int h = threadIdx.x; h < Np; h += blockDim.x) {
                double v[4];
                for (int q = 0; q < 4; q++) v[q] = (r + q < R && h < N) ? phi[(size_t)(r + q) * N + h] : 0.0;
                *reinterpret_cast<double4 *>(V + (size_t)h * ldv + r) = make_double4(v[0], v[1], v[2], v[3]);
            }
        __syncthreads();
        v7::contract(X8, ld8, N, K, R, V, ldv, dsm, cred, [&](int r, int c, double a) { G[(size_t)r * K + c] = a; }, false);
    }
}

int main(int argc, char **argv)
{
    const int N = 400, K = 481, R = argc > 1 ? atoi(argv[1]) : 30, reps = 20, nblk = 296;
    const int ld8_v4 = K, ld8 = ((K + 15) & ~15) + 512, ldv = 64;
    std::vector<int8_t> hx4((size_t)N * ld8_v4), hx((size_t)N * ld8, 0);
    srand(1);
    for (int h = 0; h < N; h++) for (int c = 0; c < K; c++) { int8_t v = (int8_t)(rand() % 3 - 1); hx4[(size_t)h * K + c] = v; hx[(size_t)h * ld8 + c] = v; }
    std::vector<double> hphi((size_t)R * N);
    for (auto &v : hphi) v = (rand() % 2001 - 1000) / 1000.0;
    int8_t *dx4, *dx; double *dphi, *dV, *dG4, *dG7, *dcred;
    cudaMalloc(&dx4, hx4.size()); cudaMalloc(&dx, hx.size()); cudaMalloc(&dphi, hphi.size() * 8);
    cudaMalloc(&dV, (size_t)nblk * (N + 64) * ldv * 8); cudaMalloc(&dG4, (size_t)nblk * R * K * 8); cudaMalloc(&dG7, (size_t)nblk * R * K * 8);
    cudaMalloc(&dcred, (size_t)nblk * 8192 * 8);
    cudaMemcpy(dx4, hx4.data(), hx4.size(), cudaMemcpyHostToDevice); cudaMemcpy(dx, hx.data(), hx.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(dphi, hphi.data(), hphi.size() * 8, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k_v7, cudaFuncAttributeMaxDynamicSharedMemorySize, v7::SMEM_BYTES);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double fma_total = (double)nblk * reps * N * K * R;
    for (int which = 0; which < 2; which++) {
        for (int it = 0; it < 2; it++) {
            cudaEventRecord(e0);
            if (which == 0) k_v4<<<nblk, 256>>>(dx4, ld8_v4, N, K, R, dV, ldv, dG4, reps, dphi);
            else k_v7<<<nblk, 256, v7::SMEM_BYTES>>>(dx, ld8, N, K, R, dV, ldv, dG7, reps, dphi, dcred);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            cudaError_t err = cudaGetLastError();
            if (it == 1) printf("%s R=%d: %.3f ms  -> %.2f TFLOP/s (alg), %.0f cycles/call/block  [%s]\n", which == 0 ? "v4" : "v7", R, ms,
                                2 * fma_total / ms / 1e9, ms * 1e-3 * 1.965e9 / reps, cudaGetErrorString(err));
        }
    }
    std::vector<double> g4((size_t)R * K), g7((size_t)R * K);
    cudaMemcpy(g4.data(), dG4 + (size_t)5 * R * K, g4.size() * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(g7.data(), dG7 + (size_t)5 * R * K, g7.size() * 8, cudaMemcpyDeviceToHost);
    double md = 0, mx = 0; for (size_t i = 0; i < g4.size(); i++) { md = fmax(md, fabs(g4[i] - g7[i])); mx = fmax(mx, fabs(g4[i])); }
    printf("max |v4 - v7| = %.3e (max |v4| = %.3e)\n", md, mx);
    return 0;
}
