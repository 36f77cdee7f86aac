// Isolated timing of the int8 score contraction (imma_contract.cuh) at config-2 size: cycles of the slicing step and of
// the tensor-core pass, per call, with the fit kernel's launch shape (2 blocks of 256 threads per SM, 41 KB shared memory).
// experiments only.  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -std=c++17 -I../../pareben_b200/csrc -o mb_imma mb_imma.cu
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_pipeline.h>
#include "common.cuh"
namespace pareben { __device__ inline double div_by(double a, double b, double r) { const double q = a * r; return fma(fma(-q, b, a), r, q); } }
#include "imma_contract.cuh"
using namespace pareben;

__global__ void __launch_bounds__(256, 2)
mb_kernel(const int8_t *A8f, const int8_t *A8sqf, int ldt, int Kc, int N, int M, const double *phi, const double *w, const double *e,
          int8_t *bs_all, double *bscale_all, double *G_all, double *SQ_all, int reps, long long *cyc)
{
    extern __shared__ __align__(32) double s_buf[];
    const int R = M + 1, Rp = (R + 7) & ~7, LD = (N + 3) & ~3;
    int8_t *bs = bs_all + (size_t)blockIdx.x * IM_SLICES * Rp * ldt;
    double *bscale = bscale_all + (size_t)blockIdx.x * Rp;
    double *G = G_all + (size_t)blockIdx.x * M * Kc, *S = SQ_all + (size_t)blockIdx.x * 2 * Kc, *Q = S + Kc;
    long long t_slice = 0, t_pass = 0;
    for (int rep = 0; rep < reps; rep++) {
        __syncthreads();
        const long long t0 = clock64();
        imma_slice(R, Rp, N, ldt, [&](int r, int h) { return r < M ? w[h] * phi[(size_t)r * LD + h] : e[h]; }, bs, bscale);
        __syncthreads();
        const long long t1 = clock64();
        imma_pass(A8f, A8sqf, ldt, Kc, R, Rp, bs, bscale, s_buf,
            [&](int r, int c, double val, int) { if (r < M) G[(size_t)r * Kc + c] = val; else Q[c] = val; },
            [&](int c, double val) { S[c] = val; });
        __syncthreads();
        const long long t2 = clock64();
        t_slice += t1 - t0; t_pass += t2 - t1;
    }
    if (threadIdx.x == 0) { atomicAdd((unsigned long long *)cyc, (unsigned long long)t_slice); atomicAdd((unsigned long long *)cyc + 1, (unsigned long long)t_pass); }
}

int main()
{
    const int Kc = 481, N = 400, M = 34, ldt = 448, KP = ldt / 64, LD = (N + 3) & ~3, R = M + 1, Rp = (R + 7) & ~7;
    const int nblk = 296, reps = 20;
    // random ternary design in row positions (any order: the check below uses the same positions), fragment-major
    std::vector<int8_t> X((size_t)Kc * ldt, 0);
    srand(1);
    for (int c = 0; c < Kc; c++) for (int p = 0; p < ldt; p++) { const int h = (p & ~15) | ((p & 3) << 2) | ((p >> 2) & 3); X[(size_t)c * ldt + p] = h < N ? (int8_t)(rand() % 3 - 1) : 0; }
    const int n_mt = (Kc + 15) / 16;
    std::vector<int8_t> Af((size_t)n_mt * KP * 1024, 0), Asq(Af.size(), 0);
    for (int mt = 0; mt < n_mt; mt++) for (int kp = 0; kp < KP; kp++) for (int lane = 0; lane < 32; lane++) for (int half = 0; half < 2; half++) {
        const int c = mt * 16 + half * 8 + lane / 4;
        for (int b = 0; b < 16; b++) {
            const int8_t x = c < Kc ? X[(size_t)c * ldt + kp * 64 + (lane % 4) * 16 + b] : 0;
            const size_t o = ((((size_t)mt * KP + kp) * 32 + lane) * 2 + half) * 16 + b;
            Af[o] = x; Asq[o] = (int8_t)(x * x);
        }
    }
    std::vector<double> phi((size_t)M * LD, 0.0), w(N), e(N);
    for (int h = 0; h < N; h++) { w[h] = 0.05 + 0.2 * (rand() / (double)RAND_MAX); e[h] = rand() / (double)RAND_MAX - 0.5; phi[h] = 1.0; }
    for (int r = 1; r < M; r++) for (int h = 0; h < N; h++) phi[(size_t)r * LD + h] = (rand() % 3 - 1) / 17.3;
    int8_t *dA, *dAsq, *dbs; double *dphi, *dw, *de, *dscale, *dG, *dSQ; long long *dcyc;
    cudaMalloc(&dA, Af.size()); cudaMalloc(&dAsq, Af.size()); cudaMalloc(&dbs, (size_t)nblk * IM_SLICES * Rp * ldt);
    cudaMalloc(&dphi, phi.size() * 8); cudaMalloc(&dw, N * 8); cudaMalloc(&de, N * 8); cudaMalloc(&dscale, (size_t)nblk * Rp * 8);
    cudaMalloc(&dG, (size_t)nblk * M * Kc * 8); cudaMalloc(&dSQ, (size_t)nblk * 2 * Kc * 8); cudaMalloc(&dcyc, 16);
    cudaMemcpy(dA, Af.data(), Af.size(), cudaMemcpyHostToDevice); cudaMemcpy(dAsq, Asq.data(), Af.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(dphi, phi.data(), phi.size() * 8, cudaMemcpyHostToDevice); cudaMemcpy(dw, w.data(), N * 8, cudaMemcpyHostToDevice); cudaMemcpy(de, e.data(), N * 8, cudaMemcpyHostToDevice);
    cudaMemset(dcyc, 0, 16);
    const int smem = 41984;
    cudaFuncSetAttribute(mb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    mb_kernel<<<nblk, 256, smem>>>(dA, dAsq, ldt, Kc, N, M, dphi, dw, de, dbs, dscale, dG, dSQ, 2, dcyc);
    cudaMemset(dcyc, 0, 16);
    cudaEventRecord(e0);
    mb_kernel<<<nblk, 256, smem>>>(dA, dAsq, ldt, Kc, N, M, dphi, dw, de, dbs, dscale, dG, dSQ, reps, dcyc);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    long long cyc[2]; cudaMemcpy(cyc, dcyc, 16, cudaMemcpyDeviceToHost);
    printf("%s: %d blocks x %d reps in %.2f ms; per call: slice %.0f cycles, pass %.0f cycles (DMMA path: ~330000 per call in the fit kernel)\n",
           cudaGetErrorString(cudaGetLastError()), nblk, reps, ms, (double)cyc[0] / (nblk * reps), (double)cyc[1] / (nblk * reps));
    // check block 0 against a host FP64 evaluation
    std::vector<double> G((size_t)M * Kc), SQ(2 * Kc);
    cudaMemcpy(G.data(), dG, G.size() * 8, cudaMemcpyDeviceToHost); cudaMemcpy(SQ.data(), dSQ, SQ.size() * 8, cudaMemcpyDeviceToHost);
    double worst = 0;
    for (int c = 0; c < Kc; c++) {
        long double bb = 0, ze = 0;
        for (int p = 0; p < ldt; p++) { const int h = (p & ~15) | ((p & 3) << 2) | ((p >> 2) & 3); if (h >= N) continue; const int x = X[(size_t)c * ldt + p]; bb += (long double)x * x * w[h]; ze += (long double)x * e[h]; }
        worst = fmax(worst, fabs((double)bb - SQ[c]) / fabs((double)bb)); worst = fmax(worst, fabs((double)ze - SQ[Kc + c]) / (fabs((double)ze) + 1e-3));
        for (int r = 0; r < M; r += 7) {
            long double z = 0;
            for (int p = 0; p < ldt; p++) { const int h = (p & ~15) | ((p & 3) << 2) | ((p >> 2) & 3); if (h >= N) continue; z += (long double)X[(size_t)c * ldt + p] * (w[h] * phi[(size_t)r * LD + h]); }
            worst = fmax(worst, fabs((double)z - G[(size_t)r * Kc + c]) / (fabs((double)z) + 1e-3));
        }
    }
    printf("worst relative difference against the host long-double evaluation: %.3e\n", worst);
    return 0;
}
