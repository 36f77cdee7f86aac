// stream_ge.cu -- streaming-mode kernels, Gaussian Epis (see stream.cuh).
#include "stream_kernel.cuh"
PAREBEN_DEFINE_STREAM(ge, true)
