// Kernels and launch helpers of the likelihood pipeline for supersampling factor S = 0 (run-time S: every factor without a compiled-in loop)
// (explicit instantiation; see api_internal.cuh).
#define NMB_INSTANTIATE_S 0
#include "api_internal.cuh"
