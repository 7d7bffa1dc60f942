// Stand-in for glog + gflags declarations (not installed).  TEST INFRASTRUCTURE, see Eigen/Dense.
#ifndef HF6D_SHIM_GLOG_H
#define HF6D_SHIM_GLOG_H
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>

struct Hf6dShimCheck {
    bool ok;
    std::ostringstream s;
    Hf6dShimCheck(bool ok_, const char* expr, const char* file, int line) : ok(ok_) {
        if (!ok) s << "Check failed: " << expr << " (" << file << ":" << line << ") ";
    }
    ~Hf6dShimCheck() {
        if (!ok) {
            std::fprintf(stderr, "%s\n", s.str().c_str());
            std::abort();
        }
    }
    template <typename T>
    Hf6dShimCheck& operator<<(const T& v) {
        if (!ok) s << v;
        return *this;
    }
};
#define CHECK(c) Hf6dShimCheck((c) ? true : false, #c, __FILE__, __LINE__)
#define CHECK_GT(a, b) Hf6dShimCheck((a) > (b), #a " > " #b, __FILE__, __LINE__)
#define CHECK_EQ(a, b) Hf6dShimCheck((a) == (b), #a " == " #b, __FILE__, __LINE__)
#define DECLARE_bool(name) extern bool FLAGS_##name
#define DECLARE_string(name) extern std::string FLAGS_##name
#define DECLARE_int32(name) extern int FLAGS_##name
using std::cin;
using std::cout;
#endif
