// Single translation unit of libqasr.so: the kernels live in headers (several are non-template __global__
// functions), so the two API files are compiled together.
#include "qasr_api.cu"
#include "decoder_api.cu"
