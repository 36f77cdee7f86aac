// stream_gm.cu -- streaming-mode kernels, Gaussian main effects (see stream.cuh).
#include "stream_kernel.cuh"
PAREBEN_DEFINE_STREAM(gm, false)
