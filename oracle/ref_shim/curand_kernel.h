#include <hf6d_shim_cuda.hpp>
