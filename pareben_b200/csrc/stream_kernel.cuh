// stream_kernel.cuh -- launchers of the two streaming-mode kernels, one translation unit per variant.
#pragma once
#include "stream_fit.cuh"
#include "stream_scan.cuh"
#include "stream_launch.h"

namespace pareben {

constexpr size_t ADV_SMEM_BYTES = (size_t)SWEEP_PANEL_DOUBLES * sizeof(double);
constexpr size_t SCAN_SMEM_BYTES = (size_t)SCAN_STAGES * STAGE_D * sizeof(double);

template <bool EPIS>
inline cudaError_t launch_stream_advance(int n_fits, cudaStream_t stream, const Problem &P, const Variant &v, StreamFit *fits,
                                         const StreamShared &sh, const FitOutputs &out)
{
    stream_advance_kernel<EPIS><<<n_fits, ADV_THREADS, ADV_SMEM_BYTES, stream>>>(P, v, fits, n_fits, sh, out);
    return cudaGetLastError();
}

template <bool EPIS>
inline cudaError_t launch_stream_scan(int grid, cudaStream_t stream, const Problem &P, StreamFit *fits, const StreamShared &sh)
{
    cudaError_t e = cudaFuncSetAttribute(stream_scan_kernel<EPIS, !EPIS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCAN_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    stream_scan_kernel<EPIS, !EPIS><<<grid, SCAN_THREADS, SCAN_SMEM_BYTES, stream>>>(P, fits, sh);
    return cudaGetLastError();
}

}  // namespace pareben

#define PAREBEN_DEFINE_STREAM(NAME, EPIS)                                                                                 \
    namespace pareben {                                                                                                   \
    cudaError_t launch_stream_advance_##NAME(int n_fits, cudaStream_t stream, const Problem &P, const Variant &v,         \
                                             StreamFit *fits, const StreamShared &sh, const FitOutputs &out)              \
    { return launch_stream_advance<EPIS>(n_fits, stream, P, v, fits, sh, out); }                                          \
    cudaError_t launch_stream_scan_##NAME(int grid, cudaStream_t stream, const Problem &P, StreamFit *fits, const StreamShared &sh) \
    { return launch_stream_scan<EPIS>(grid, stream, P, fits, sh); }                                                       \
    }
