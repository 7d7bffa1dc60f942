#include <hf6d_shim_cv.hpp>
