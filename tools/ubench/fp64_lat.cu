// Micro-benchmark: latency / throughput of the FP64 and conversion instructions the gather's variance chain uses (B200).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o fp64_lat fp64_lat.cu && ./fp64_lat
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, long long* cyc, int n, float seed, double t) {
    float v = seed + threadIdx.x * 1e-3f;
    double d = v;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
        if (MODE == 0) v = __double2float_rn(__dadd_rn((double)v, t));          // the chain: F2F -> DADD -> F2F
        if (MODE == 1) d = __dadd_rn(d, t);                                       // DADD alone
        if (MODE == 2) v = __double2float_rn((double)v * 1.0000001);             // F2F -> DMUL -> F2F
        if (MODE == 3) v = __fadd_rn(v, seed);                                    // FADD
        if (MODE == 4) d = __fma_rn(d, 1.0000001, t);                             // DFMA
    }
    long long t1 = clock64();
    if (MODE == 1 || MODE == 4) v = (float)d;
    out[blockIdx.x * blockDim.x + threadIdx.x] = v;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE>
void run(const char* name, int warps_per_sm) {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 2048 * 4); cudaMalloc(&cyc, 8);
    const int n = 4096;
    int threads = 32 * (warps_per_sm > 32 ? 32 : warps_per_sm), ctas = 148 * (warps_per_sm > 32 ? warps_per_sm / 32 : 1);
    k<MODE><<<ctas, threads>>>(out, cyc, n, 1.0f, 1e-9);
    cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<MODE><<<ctas, threads>>>(out, cyc, n, 1.0f, 1e-9);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s warps/SM %2d: %.1f cycles/iter (warp 0), kernel %.3f ms\n", name, warps_per_sm, (double)c / n, ms);
}
int main() {
    for (int w : {1, 4, 8, 16, 32, 64}) {
        run<0>("F2F->DADD->F2F chain", w);
        run<1>("DADD chain", w);
        run<2>("F2F->DMUL->F2F chain", w);
        run<3>("FADD chain", w);
        run<4>("DFMA chain", w);
    }
    return 0;
}
