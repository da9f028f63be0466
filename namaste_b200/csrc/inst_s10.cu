// Kernels and launch helpers of the likelihood pipeline for supersampling factor S = 10
// (explicit instantiation; see api_internal.cuh).
#define NMB_INSTANTIATE_S 10
#include "api_internal.cuh"
