// Host stand-in for the CUDA runtime pieces PatchGen/src/cuda/*.cu use, so that the reference's kernels and their host
// drivers compile with g++ and run on the CPU: the legacy texture-reference API they are written against was removed in
// CUDA 12 and there is no GPU in the build container.  TEST INFRASTRUCTURE, see Eigen/Dense.
//
// oracle/build_ref.py rewrites ONE construct of the .cu text at build time: `k<<<grid, block>>>(args)` becomes
// `HF6D_SHIM_LAUNCH(k, grid, block, args)`, which runs the kernel body for every (block, thread) in order on the host.
// (Thread 0 runs first, which is all `__shared__` + `__syncthreads()` need in these kernels.)
//
// Two statements are made here because the reference leaves them to hardware (both are choices C1 / C2 of the oracle):
//  * tex3D / tex2D: unnormalised coordinates, border address mode, linear filter with the fraction rounded to 8 bits,
//    S = ((w00*T00 + w10*T10) + w01*T01) + w11*T11 in fp32.  tests/test_gpu_parity.py compares this filter with the B200's
//    texture unit.
//  * clock64()-seeded cuRAND: a counter-based generator keyed on (seed, block index).
#ifndef HF6D_SHIM_CUDA_HPP
#define HF6D_SHIM_CUDA_HPP
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __shared__ static
inline void __syncthreads() {}

struct Hf6dShimDim { unsigned x, y, z; };
extern Hf6dShimDim threadIdx, blockIdx, blockDim, gridDim;
extern const char* hf6d_shim_kernel_name;
extern unsigned long long hf6d_shim_fill_seed;

typedef int cudaError_t;
typedef int cudaError;
enum { cudaSuccess = 0 };
inline const char* cudaGetErrorString(int) { return "shim"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaThreadSynchronize() { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
typedef int cudaEvent_t;
struct cudaDeviceProp { int maxThreadsPerBlock; };

enum cudaTextureReadMode { cudaReadModeElementType = 0 };
enum cudaTextureAddressMode { cudaAddressModeWrap = 0, cudaAddressModeClamp = 1, cudaAddressModeMirror = 2, cudaAddressModeBorder = 3 };
enum cudaTextureFilterMode { cudaFilterModePoint = 0, cudaFilterModeLinear = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
struct cudaChannelFormatDesc { int x, y, z, w, f; };
struct cudaExtent { size_t width, height, depth; };
inline cudaExtent make_cudaExtent(size_t w, size_t h, size_t d) { cudaExtent e; e.width = w; e.height = h; e.depth = d; return e; }
struct cudaPitchedPtr { void* ptr; size_t pitch, xsize, ysize; };
inline cudaPitchedPtr make_cudaPitchedPtr(void* p, size_t pitch, size_t xs, size_t ys) { cudaPitchedPtr r; r.ptr = p; r.pitch = pitch; r.xsize = xs; r.ysize = ys; return r; }
struct cudaArray { std::vector<float> d; size_t w, h, dep; };
struct cudaPos { size_t x, y, z; };
struct cudaMemcpy3DParms {
    cudaArray* srcArray; cudaPos srcPos; cudaPitchedPtr srcPtr;
    cudaArray* dstArray; cudaPos dstPos; cudaPitchedPtr dstPtr;
    cudaExtent extent; cudaMemcpyKind kind;
};

template <typename T, int DIM, cudaTextureReadMode MODE>
struct texture {
    int normalized;
    cudaTextureFilterMode filterMode;
    cudaTextureAddressMode addressMode[3];
    cudaChannelFormatDesc channelDesc;
    const cudaArray* bound;
};

inline cudaError_t cudaMalloc3DArray(cudaArray** a, const cudaChannelFormatDesc*, cudaExtent e) {
    *a = new cudaArray; (*a)->w = e.width; (*a)->h = e.height; (*a)->dep = e.depth; (*a)->d.assign(e.width * e.height * e.depth, 0.f); return cudaSuccess;
}
inline cudaError_t cudaMallocArray(cudaArray** a, const cudaChannelFormatDesc*, size_t w, size_t h) {
    *a = new cudaArray; (*a)->w = w; (*a)->h = h; (*a)->dep = 1; (*a)->d.assign(w * h, 0.f); return cudaSuccess;
}
inline cudaError_t cudaMemcpy3D(const cudaMemcpy3DParms* p) {  // host pitched pointer -> array, rows contiguous
    cudaArray* a = p->dstArray;
    const char* src = (const char*)p->srcPtr.ptr;
    for (size_t z = 0; z < p->extent.depth; ++z)
        for (size_t y = 0; y < p->extent.height; ++y)
            std::memcpy(&a->d[(z * a->h + y) * a->w], src + (z * p->srcPtr.ysize + y) * p->srcPtr.pitch, p->extent.width * sizeof(float));
    return cudaSuccess;
}
inline cudaError_t cudaMemcpyToArray(cudaArray* a, size_t, size_t, const void* src, size_t bytes, cudaMemcpyKind) {
    std::memcpy(&a->d[0], src, bytes); return cudaSuccess;
}
template <typename T, int DIM, cudaTextureReadMode MODE>
inline cudaError_t cudaBindTextureToArray(texture<T, DIM, MODE>& t, const cudaArray* a, const cudaChannelFormatDesc&) { t.bound = a; return cudaSuccess; }
template <typename T, int DIM, cudaTextureReadMode MODE>
inline cudaError_t cudaBindTextureToArray(texture<T, DIM, MODE>& t, const cudaArray* a) { t.bound = a; return cudaSuccess; }
template <typename T, int DIM, cudaTextureReadMode MODE>
inline cudaError_t cudaUnbindTexture(texture<T, DIM, MODE>& t) { t.bound = 0; return cudaSuccess; }
inline cudaError_t cudaFreeArray(cudaArray* a) { delete a; return cudaSuccess; }
template <typename T>
inline cudaError_t cudaMalloc(T** p, size_t bytes) { *p = (T*)std::calloc(bytes ? bytes : 1, 1); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, cudaMemcpyKind) { std::memcpy(dst, src, bytes); return cudaSuccess; }
inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }

// ---- the texture unit's linear filter (choice C1)
inline float hf6d_shim_frac8(float a) { return std::floor(a * 256.0f + 0.5f) / 256.0f; }
inline float hf6d_shim_texel(const cudaArray* a, long x, long y, long z) {
    if (x < 0 || y < 0 || z < 0 || x >= (long)a->w || y >= (long)a->h || z >= (long)a->dep) return 0.0f;  // border
    return a->d[((size_t)z * a->h + (size_t)y) * a->w + (size_t)x];
}
// filter over the two coordinates (p, q); `fixed` addresses a texel centre exactly (the channel index of the 3-D texture)
inline float hf6d_shim_bilinear(const cudaArray* a, bool three_d, float fixed, float p, float q) {
    long c = 0;
    if (three_d) {
        const float cb = fixed - 0.5f;
        c = (long)std::floor(cb);
        if (cb - std::floor(cb) != 0.0f) { std::fprintf(stderr, "shim: channel coordinate off a texel centre\n"); std::abort(); }
    }
    const float u = p - 0.5f, v = q - 0.5f;
    const float fu = std::floor(u), fv = std::floor(v);
    const long i = (long)fu, j = (long)fv;
    const float al = hf6d_shim_frac8(u - fu), be = hf6d_shim_frac8(v - fv);
    const float w00 = (1.0f - al) * (1.0f - be), w10 = al * (1.0f - be), w01 = (1.0f - al) * be, w11 = al * be;
    float t00, t10, t01, t11;
    if (three_d) {
        t00 = hf6d_shim_texel(a, c, i, j); t10 = hf6d_shim_texel(a, c, i + 1, j);
        t01 = hf6d_shim_texel(a, c, i, j + 1); t11 = hf6d_shim_texel(a, c, i + 1, j + 1);
    } else {
        t00 = hf6d_shim_texel(a, i, j, 0); t10 = hf6d_shim_texel(a, i + 1, j, 0);
        t01 = hf6d_shim_texel(a, i, j + 1, 0); t11 = hf6d_shim_texel(a, i + 1, j + 1, 0);
    }
    float s = w00 * t00;
    s = s + w10 * t10;
    s = s + w01 * t01;
    s = s + w11 * t11;
    return s;
}
template <cudaTextureReadMode MODE>
inline float tex3D(const texture<float, 3, MODE>& t, float x, float y, float z) { return hf6d_shim_bilinear(t.bound, true, x, y, z); }
template <cudaTextureReadMode MODE>
inline float tex2D(const texture<float, 2, MODE>& t, float x, float y) { return hf6d_shim_bilinear(t.bound, false, 0.f, x, y); }

// ---- clock64() + cuRAND (choice C2): draws are the fields of mix64(fill_seed + golden * (block + 1)) in call order
inline unsigned long long hf6d_shim_mix64(unsigned long long z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
struct curandState { unsigned long long z, z2, z3; int n; };
extern int hf6d_shim_clock_calls;  // clock64() calls by thread 0 of the current block (retries of the normals kernel)
inline long long clock64() { return (long long)(++hf6d_shim_clock_calls); }
inline void curand_init(unsigned long long seed_from_clock, unsigned long long, unsigned long long, curandState* s) {
    const unsigned long long G = 0x9E3779B97F4A7C15ULL;
    if (seed_from_clock <= 1) s->z = hf6d_shim_mix64(hf6d_shim_fill_seed + G * (unsigned long long)(blockIdx.x + 1));
    else s->z = s->z3;  // a retry continues the sequence (never reached in practice: needs a zero-length draw)
    s->z2 = hf6d_shim_mix64(s->z + G);
    s->z3 = hf6d_shim_mix64(s->z2 + G);
    s->n = 0;
}
inline unsigned int curand(curandState* s) {
    const int i = s->n++;
    if (i < 3) return (unsigned int)((s->z >> (16 * i)) & 0xFFFF);  // r, g, b
    if (std::strcmp(hf6d_shim_kernel_name, "extract_rgbd") == 0) return (unsigned int)((s->z >> 48) & 0xFFFF);  // d
    if (i == 3) return (unsigned int)(s->z2 & 0xFFFFFFFFu);         // x
    if (i == 4) return (unsigned int)(s->z2 >> 32);                 // y
    return (unsigned int)(s->z3 & 0xFFFFFFFFu);                     // z
}

class cuda_timer {
  public:
    void start_timer() {}
    void print_timer(const std::string&) {}
};
class cuda_initializer {
  public:
    cudaDeviceProp deviceProp;
    void deviceInit(int = 0) { deviceProp.maxThreadsPerBlock = 1024; }
};

#define CUDA_CHECK_ERROR(msg) do { } while (0)
#define CUDA_SAFE_CALL_NO_SYNC(call) do { (void)(call); } while (0)
#define CUDA_SAFE_CALL(call) do { (void)(call); } while (0)

#define HF6D_SHIM_LAUNCH(kernel, grid, block, ...)                                   \
    do {                                                                             \
        hf6d_shim_kernel_name = #kernel;                                             \
        gridDim.x = (unsigned)(grid); gridDim.y = gridDim.z = 1;                     \
        blockDim.x = (unsigned)(block); blockDim.y = blockDim.z = 1;                 \
        for (unsigned hf6d_b = 0; hf6d_b < gridDim.x; ++hf6d_b) {                    \
            blockIdx.x = hf6d_b;                                                     \
            hf6d_shim_clock_calls = 0;                                               \
            for (unsigned hf6d_t = 0; hf6d_t < blockDim.x; ++hf6d_t) {               \
                threadIdx.x = hf6d_t;                                                \
                kernel(__VA_ARGS__);                                                 \
            }                                                                        \
        }                                                                            \
    } while (0)
#endif
