// stream_launch.h -- host-callable entry points of the streaming-mode translation units (stream_gm.cu, stream_ge.cu).
#pragma once
#include "stream.cuh"

namespace pareben {

#define PAREBEN_DECLARE_STREAM(NAME)                                                                                      \
    cudaError_t launch_stream_advance_##NAME(int n_fits, cudaStream_t stream, const Problem &P, const Variant &v,         \
                                             StreamFit *fits, const StreamShared &sh, const FitOutputs &out);             \
    cudaError_t launch_stream_scan_##NAME(int grid, cudaStream_t stream, const Problem &P, StreamFit *fits, const StreamShared &sh);

PAREBEN_DECLARE_STREAM(gm)   // Gaussian, main effects
PAREBEN_DECLARE_STREAM(ge)   // Gaussian, Epis

}  // namespace pareben
