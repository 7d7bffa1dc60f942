// Stand-alone bring-up check for the tcgen05 encoder layer kernel: compares each of the three layer shapes
// against a naive CUDA-core kernel on the same bf16 operands and times it with CUDA events.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o build/encoder_test tools/encoder_test.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../object_detector_6d_b200/csrc/encoder.cuh"

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(2);                                                                   \
        }                                                                              \
    } while (0)

__global__ void naive_layer(const __nv_bfloat16* A, const __nv_bfloat16* W, const float* bias, float* out, int M, int N,
                            int K) {
    int n = blockIdx.y * blockDim.x + threadIdx.x;
    int m = blockIdx.x;
    if (n >= N || m >= M) return;
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc += __bfloat162float(A[(size_t)m * K + k]) * __bfloat162float(W[(size_t)n * K + k]);
    acc += bias[n];
    out[(size_t)m * N + n] = 1.0f / (1.0f + expf(-acc));
}

static uint32_t rng_state = 12345;
static float frand() {
    rng_state = rng_state * 1664525u + 1013904223u;
    return (rng_state >> 8) * (1.0f / 16777216.0f);
}

static int run_case(int M, int K, int n_pad, int n_valid, int block_n, bool last, int iters) {
    int m_cap = (M + 127) / 128 * 128;
    std::vector<__nv_bfloat16> hA((size_t)m_cap * K), hW((size_t)n_pad * K);
    std::vector<float> hb(n_pad);
    for (auto& a : hA) a = __float2bfloat16(frand());
    float sc = 2.0f / sqrtf((float)K);
    for (auto& w : hW) w = __float2bfloat16((frand() - 0.5f) * sc);
    for (auto& b : hb) b = (frand() - 0.5f) * 0.2f;
    __nv_bfloat16 *dA, *dW;
    float *db, *dref;
    void* dout;
    int* dM;
    size_t out_elem = last ? 4 : 2;
    int out_ld = last ? n_valid : n_pad;
    CK(cudaMalloc(&dA, hA.size() * 2));
    CK(cudaMalloc(&dW, hW.size() * 2));
    CK(cudaMalloc(&db, n_pad * 4));
    CK(cudaMalloc(&dref, (size_t)M * n_pad * 4));
    CK(cudaMalloc(&dout, (size_t)m_cap * out_ld * out_elem));
    CK(cudaMalloc(&dM, 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dW, hW.data(), hW.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, hb.data(), n_pad * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dM, &M, 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dout, 0, (size_t)m_cap * out_ld * out_elem));

    hf6d::EncoderLayerLaunch L;
    if (!hf6d::make_bf16_kmajor_map(&L.tmA, dA, m_cap, K, 128) ||
        !hf6d::make_bf16_kmajor_map(&L.tmB, dW, n_pad, K, block_n)) {
        printf("tensor map creation failed\n");
        return 1;
    }
    L.bias = db;
    L.out = dout;
    L.out_ld = out_ld;
    L.n_valid = n_valid;
    L.K = K;
    L.n_pad = n_pad;
    L.block_n = block_n;
    L.last = last;
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));

    CK(hf6d::launch_encoder_layer(L, dM, sms, 0));
    CK(cudaDeviceSynchronize());
    naive_layer<<<dim3(M, (n_pad + 127) / 128), 128>>>(dA, dW, db, dref, M, n_pad, K);
    CK(cudaDeviceSynchronize());

    std::vector<float> href((size_t)M * n_pad);
    CK(cudaMemcpy(href.data(), dref, href.size() * 4, cudaMemcpyDeviceToHost));
    double max_err = 0;
    size_t bad = 0;
    if (last) {
        std::vector<float> ho((size_t)m_cap * out_ld);
        CK(cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost));
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < n_valid; ++n) {
                double e = fabs((double)ho[(size_t)m * out_ld + n] - href[(size_t)m * n_pad + n]);
                if (e > max_err) max_err = e;
                if (!(e < 1e-4)) ++bad;
            }
    } else {
        std::vector<__nv_bfloat16> ho((size_t)m_cap * out_ld);
        CK(cudaMemcpy(ho.data(), dout, ho.size() * 2, cudaMemcpyDeviceToHost));
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < n_pad; ++n) {
                double e = fabs((double)__bfloat162float(ho[(size_t)m * out_ld + n]) - href[(size_t)m * n_pad + n]);
                if (e > max_err) max_err = e;
                if (!(e < 6e-3)) ++bad;
            }
    }
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) CK(hf6d::launch_encoder_layer(L, dM, sms, 0));
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) CK(hf6d::launch_encoder_layer(L, dM, sms, 0));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= iters;
    double tflops = 2.0 * M * (double)n_pad * K / (ms * 1e-3) / 1e12;
    printf("M=%d K=%d Npad=%d BN=%d last=%d : max_err=%.3g bad=%zu  %.3f ms  %.1f TFLOP/s\n", M, K, n_pad, block_n,
           (int)last, max_err, bad, ms, tflops);
    cudaFree(dA); cudaFree(dW); cudaFree(db); cudaFree(dref); cudaFree(dout); cudaFree(dM);
    return bad ? 1 : 0;
}

int main(int argc, char** argv) {
    int M = argc > 1 ? atoi(argv[1]) : 69600;
    int iters = argc > 2 ? atoi(argv[2]) : 20;
    int rc = 0;
    rc |= run_case(300, 256, 1536, 1536, 256, false, 2);   // tiny: partial last M block
    rc |= run_case(M, 256, 1536, 1536, 256, false, iters);
    rc |= run_case(M, 1536, 1024, 1024, 256, false, iters);
    rc |= run_case(M, 1024, 800, 800, 160, true, iters);
    printf(rc ? "ENCODER_TEST FAIL\n" : "ENCODER_TEST OK\n");
    return rc;
}
