// <math.h> as the reference's tested toolchain saw it (README.md:16: Ubuntu 14.04, i.e. gcc 4.8 / glibc 2.19): the C
// header only.  libstdc++ ships its own <math.h> wrapper since gcc 6 that pulls the float / long double overloads of
// <cmath> into the global namespace; with it the reference's unqualified calls change meaning -- cos(float) at
// HFTest.cpp:45-50 would become cosf instead of cos(double) narrowed to float, and pow(float, int) at :527-534 would (in
// gnu++98) become a float multiplication instead of pow(double, double).  TEST INFRASTRUCTURE, see Eigen/Dense.
#ifndef HF6D_SHIM_MATH_H
#define HF6D_SHIM_MATH_H
#ifdef _GLIBCXX_INCLUDE_NEXT_C_HEADERS
#include_next <math.h>
#else
#define _GLIBCXX_INCLUDE_NEXT_C_HEADERS
#include_next <math.h>
#undef _GLIBCXX_INCLUDE_NEXT_C_HEADERS
#endif
#endif
