// fit_kernel.cuh -- the persistent batched-fit kernel and its per-variant host launchers.
// Each of the four (Epis, prior) variants is compiled in its own translation unit (fit_gm.cu,
// fit_ge.cu, fit_bm.cu, fit_be.cu) so that the build runs four ptxas jobs in parallel; the host
// side (pareben.cu) reaches them through the plain functions declared in fit_launch.h.
#pragma once
#include "common.cuh"
#include "gauss_fit.cuh"
#include "binom_fit.cuh"
#include "fit_launch.h"

namespace pareben {

// the persistent batched-fit kernel: one block = one fit at a time, fits pulled from a queue
#ifndef PAREBEN_MIN_BLOCKS
#define PAREBEN_MIN_BLOCKS 2
#endif
constexpr int FIT_THREADS = FIT_T;     // compile-time maximum (register budget: 2 x 256 or 4 x 128 threads per SM)

constexpr int cmax(int a, int b) { return a > b ? a : b; }
constexpr int SWEEP_DOUBLES = SWEEP_SMEM_M * SWEEP_SMEM_M + 2 * SWEEP_SMEM_M;
constexpr int QUAD_DOUBLES = cmax(2 * QT * LDS_V + 784, cmax(4096 + 1040, GRAM_PIPE_DOUBLES));
constexpr int S_BUF_DOUBLES = cmax(cmax(cmax(SWEEP_DOUBLES, SWEEP_PANEL_DOUBLES), cmax(SV_DOUBLES, COL_MAX_ROWS)), cmax(GRAM_DOUBLES, QUAD_DOUBLES));
constexpr size_t S_BUF_BYTES = (size_t)S_BUF_DOUBLES * sizeof(double);

template <bool EPIS, bool BINOMIAL>
__global__ void __launch_bounds__(FIT_THREADS, PAREBEN_MIN_BLOCKS)
eben_fit_kernel(Problem P, Variant v, const FitTask *__restrict__ tasks, int n_tasks, int *queue, char *slabs,
                size_t slab_stride, FitOutputs out)
{
    // One shared buffer (dynamic: it exceeds the 48 KB static limit), used by phases that never overlap: the
    // cp.async ring of the contraction, the resident SIGMA tile of the quadratic forms, the Gram partial tiles
    // and the pivot columns of the register sweep.
    extern __shared__ __align__(32) double s_buf[];
    double *sV = s_buf;
    __shared__ double red[66];
    __shared__ int redi[66];
    __shared__ int s_task;
    Scratch sc{red, redi, s_buf};
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_task = atomicAdd(queue, 1);
        __syncthreads();
        const int ti = s_task;
        if (ti >= n_tasks) break;
        const FitTask task = tasks[ti];
        const FoldData F = P.folds[task.fold];
        Slab s = carve_slab(slabs + (size_t)blockIdx.x * slab_stride, P.cap, P.nmax, P.Kc);
#ifdef PAREBEN_PHASE_TIMING
        if (threadIdx.x == 0 && task.out_index < FIT_TRACE_MAX) { g_fit_t0[task.out_index] = global_ns(); g_fit_block[task.out_index] = blockIdx.x; }
#endif
        if (BINOMIAL) binom_fit<EPIS>(P, F, v, s, task.lambda, task.alpha, task, out, sV, sc);
        else gauss_fit<EPIS>(P, F, v, s, task.lambda, task.alpha, task, out, sV, sc);
#ifdef PAREBEN_PHASE_TIMING
        if (threadIdx.x == 0 && task.out_index < FIT_TRACE_MAX) g_fit_t1[task.out_index] = global_ns();
#endif
    }
}


template <bool EPIS, bool BINOMIAL>
inline cudaError_t launch_fit_variant(int grid, int threads, cudaStream_t stream, const Problem &P, const Variant &v,
                                      const FitTask *tasks, int n_tasks, int *queue, char *slabs, size_t slab_stride,
                                      const FitOutputs &out)
{
    cudaError_t e = cudaFuncSetAttribute(eben_fit_kernel<EPIS, BINOMIAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S_BUF_BYTES);
    if (e != cudaSuccess) return e;
    eben_fit_kernel<EPIS, BINOMIAL><<<grid, threads, S_BUF_BYTES, stream>>>(P, v, tasks, n_tasks, queue, slabs, slab_stride, out);
    return cudaGetLastError();
}

template <bool EPIS, bool BINOMIAL>
inline cudaError_t occupancy_variant(int *blocks_per_sm, int threads)
{
    cudaError_t e = cudaFuncSetAttribute(eben_fit_kernel<EPIS, BINOMIAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S_BUF_BYTES);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, eben_fit_kernel<EPIS, BINOMIAL>, threads, S_BUF_BYTES);
}

// experiments only (timing build): copy out / reset this translation unit's phase counters and fit trace
inline void timing_variant(unsigned long long *cycles_calls, int reset, unsigned long long *t0, unsigned long long *t1, int *block, int n)
{
#ifdef PAREBEN_PHASE_TIMING
    if (cycles_calls) { cudaMemcpyFromSymbol(cycles_calls, g_phase_cycles, sizeof(unsigned long long) * PH_COUNT);
                        cudaMemcpyFromSymbol(cycles_calls + PH_COUNT, g_phase_calls, sizeof(unsigned long long) * PH_COUNT); }
    if (reset) { unsigned long long z[PH_COUNT] = {0}; cudaMemcpyToSymbol(g_phase_cycles, z, sizeof z); cudaMemcpyToSymbol(g_phase_calls, z, sizeof z); }
    if (t0 && n > 0) {
        if (n > FIT_TRACE_MAX) n = FIT_TRACE_MAX;
        cudaMemcpyFromSymbol(t0, g_fit_t0, sizeof(unsigned long long) * n);
        cudaMemcpyFromSymbol(t1, g_fit_t1, sizeof(unsigned long long) * n);
        cudaMemcpyFromSymbol(block, g_fit_block, sizeof(int) * n);
    }
#else
    (void)cycles_calls; (void)reset; (void)t0; (void)t1; (void)block; (void)n;
#endif
}

}  // namespace pareben

#define PAREBEN_DEFINE_VARIANT(NAME, EPIS, BINOMIAL)                                                                      \
    namespace pareben {                                                                                                   \
    cudaError_t launch_fit_##NAME(int grid, int threads, cudaStream_t stream, const Problem &P, const Variant &v,        \
                                  const FitTask *tasks, int n_tasks, int *queue, char *slabs, size_t slab_stride,         \
                                  const FitOutputs &out)                                                                  \
    { return launch_fit_variant<EPIS, BINOMIAL>(grid, threads, stream, P, v, tasks, n_tasks, queue, slabs, slab_stride, out); } \
    cudaError_t occupancy_##NAME(int *blocks_per_sm, int threads) { return occupancy_variant<EPIS, BINOMIAL>(blocks_per_sm, threads); } \
    void timing_##NAME(unsigned long long *cc, int reset, unsigned long long *t0, unsigned long long *t1, int *block, int n) \
    { timing_variant(cc, reset, t0, t1, block, n); }                                                                      \
    }
