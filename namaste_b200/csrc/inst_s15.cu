// Kernels and launch helpers of the likelihood pipeline for supersampling factor S = 15
// (explicit instantiation; see api_internal.cuh).
#define NMB_INSTANTIATE_S 15
#include "api_internal.cuh"
