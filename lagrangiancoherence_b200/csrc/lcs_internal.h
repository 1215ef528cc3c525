// Host-side helpers shared by the translation units of liblcs_b200.so.
#pragma once
#include <cuda_runtime.h>
#include "../../include/lcs_b200.h"

int lcs_fail(int code, const char* msg);                 // records msg, returns code
int lcs_fail_cuda(cudaError_t e, const char* where);     // records the CUDA error string, returns LCS_E_CUDA
int lcs_env_int(const char* name, int dflt);             // tuning knobs (LCS_ADVECT_BAND, ...)
int lcs_sm_count();                                      // multiprocessors of the current device (cached)
void lcs_count_launches(int n);                          // bookkeeping behind lcs_kernel_launches()
