// Kernels and launch helpers of the likelihood pipeline for supersampling factor S = 1
// (explicit instantiation; see api_internal.cuh).
#define NMB_INSTANTIATE_S 1
#include "api_internal.cuh"
