#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include "../../include/spkemb.h"
#include "common.cuh"
namespace spk {
int optim_step(const spk_optim_tensors& t, int kind, int64_t step, float lr, float beta1, float beta2, float eps,
               float wd, float max_norm, float grad_scale, float* norm_scratch, int phase, int chunk, int nchunks,
               cudaStream_t st);
}
