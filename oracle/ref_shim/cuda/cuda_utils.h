// Shadows PatchGen/include/cuda/cuda_utils.h (which includes <cuda.h> / <cuda_runtime.h> and blocks on getchar() on errors).
#ifndef CUDA_UTILS_H
#define CUDA_UTILS_H
#include <iostream>
#include <hf6d_shim_cuda.hpp>
#endif
