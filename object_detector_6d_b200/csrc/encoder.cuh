// Patch auto-encoder forward, one dense layer per launch:  Y = sigmoid(X * W^T + b)
//
// Replaces the Caffe InnerProduct+Sigmoid triple the reference runs per batch of 100 patches
// (reference: HoughForest/src/HFTest.cpp:585-596, net definition generate_scripts.sh:424-524).
//
// B200 design: persistent, warp-specialised tcgen05 kernel.
//   warp 0      : TMA producer  (A tile 128x64 bf16, W tile BLOCK_Nx64 bf16, 128-byte swizzle, 4-stage ring)
//   warp 1      : MMA issuer    (tcgen05.mma cta_group::1 kind::f16, M=128, N=BLOCK_N, K=16; fp32 accum in TMEM)
//   warp 2      : TMEM allocator (512 columns = two accumulator buffers, so epilogue(i) overlaps mma(i+1))
//   warps 4..11 : epilogue      (tcgen05.ld -> +bias -> sigmoid -> bf16 / fp32 store)
// M (= number of processed patches P') is read from device memory so the launch is graph-capturable and needs no
// host round trip after the centre scan.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "ptx_sm100.cuh"

namespace hf6d {

constexpr int ENC_BLOCK_M = 128;
constexpr int ENC_BLOCK_K = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int ENC_STAGES = 4;
constexpr int ENC_THREADS = 384;
constexpr int ENC_EPI_WARPS = 8;
constexpr int ENC_TMEM_COLS = 512;
constexpr int ENC_MAX_N = 1536;

template <int BLOCK_N>
struct EncSmem {
    static constexpr int A_BYTES = ENC_BLOCK_M * ENC_BLOCK_K * 2;
    static constexpr int B_BYTES = BLOCK_N * ENC_BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int BIAS_OFF = ENC_STAGES * STAGE_BYTES;
    static constexpr int BAR_OFF = BIAS_OFF + ENC_MAX_N * 4;
    static constexpr int TOTAL = BAR_OFF + 16 * 8 + 16;
    static constexpr int DYN_BYTES = TOTAL + 1024;  // slack for manual 1024-B alignment
};

__device__ __forceinline__ float sigmoid_fast(float x) {
    // sigma(x) = 0.5*tanh(x/2)+0.5 : one MUFU op per element (hidden layers; result is rounded to bf16 anyway)
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
    return fmaf(0.5f, t, 0.5f);
}
__device__ __forceinline__ float sigmoid_accurate(float x) {
    // 1/(1+e^-x) with ex2.approx + IEEE reciprocal: ~2 ulp, used for the feature layer the forest thresholds
    return __frcp_rn(1.0f + __expf(-x));
}

// LAST=false : out is bf16 [m_cap][out_ld], all BLOCK_N columns stored (padded columns hold sigma(0)=0.5 and meet
//              zero weight columns in the next layer)
// LAST=true  : out is fp32 [m_cap][out_ld], columns < n_valid stored
template <int BLOCK_N, bool LAST>
__global__ void __launch_bounds__(ENC_THREADS, 1)
encoder_layer_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const float* __restrict__ bias, void* __restrict__ out, int out_ld, int n_valid,
                     const int* __restrict__ m_ptr, int K, int n_pad) {
    using S = EncSmem<BLOCK_N>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

    float* s_bias = reinterpret_cast<float*>(smem + S::BIAS_OFF);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
    uint64_t* full = bars;                     // [STAGES]
    uint64_t* empty = bars + ENC_STAGES;       // [STAGES]
    uint64_t* tfull = bars + 2 * ENC_STAGES;   // [2]
    uint64_t* tempty = tfull + 2;              // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int M = *m_ptr;
    const int m_blocks = (M + ENC_BLOCK_M - 1) / ENC_BLOCK_M;
    const int n_blocks = n_pad / BLOCK_N;
    const int k_blocks = K / ENC_BLOCK_K;
    const int tiles = m_blocks * n_blocks;

    for (int i = threadIdx.x; i < n_pad; i += ENC_THREADS) s_bias[i] = bias[i];

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmA);
        ptx::prefetch_tensormap(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < ENC_STAGES; ++s) {
            ptx::mbar_init(&full[s], 1);
            ptx::mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            ptx::mbar_init(&tfull[a], 1);
            ptx::mbar_init(&tempty[a], ENC_EPI_WARPS);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(tmem_slot, ENC_TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
                const int mb = t / n_blocks, nb = t % n_blocks;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    ptx::mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * S::STAGE_BYTES;
                    uint8_t* sb = sa + S::A_BYTES;
                    ptx::mbar_arrive_expect_tx(&full[stage], S::STAGE_BYTES);
                    ptx::tma_load_2d(sa, &tmA, &full[stage], kb * ENC_BLOCK_K, mb * ENC_BLOCK_M);
                    ptx::tma_load_2d(sb, &tmB, &full[stage], kb * ENC_BLOCK_K, nb * BLOCK_N);
                    if (++stage == ENC_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (single thread)
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(ENC_BLOCK_M, BLOCK_N);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
                ptx::mbar_wait(&tempty[acc], acc_phase ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * 256;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    ptx::mbar_wait(&full[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t sa = ptx::smem_u32(smem + stage * S::STAGE_BYTES);
                    const uint64_t adesc = ptx::make_kmajor_sw128_desc(sa);
                    const uint64_t bdesc = ptx::make_kmajor_sw128_desc(sa + S::A_BYTES);
#pragma unroll
                    for (int k = 0; k < ENC_BLOCK_K / 16; ++k) {
                        // +32 B per K=16 step inside the 128-B swizzle row: +2 in the (addr>>4) field
                        ptx::umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                    }
                    ptx::umma_commit(&empty[stage]);
                    if (++stage == ENC_STAGES) { stage = 0; phase ^= 1; }
                }
                ptx::umma_commit(&tfull[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ------------------------------------------------------------ epilogue
        const int q = warp & 3;               // TMEM lane quadrant this warp may touch
        const int half = (warp - 4) >> 2;     // which half of the tile's columns
        constexpr int HALF_N = BLOCK_N / 2;
        constexpr int CHUNKS = HALF_N / 16;
        static_assert(HALF_N % 16 == 0, "column half must be a multiple of 16");
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
            const int mb = t / n_blocks, nb = t % n_blocks;
            const int row = mb * ENC_BLOCK_M + q * 32 + lane;
            ptx::mbar_wait(&tfull[acc], acc_phase);
            ptx::tc_fence_after();
            const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256 + half * HALF_N;
#pragma unroll 1
            for (int c = 0; c < CHUNKS; ++c) {
                uint32_t v[16];
                ptx::tmem_ld_32x32b_x16(taddr0 + c * 16, v);
                ptx::tmem_ld_wait();
                const int col = nb * BLOCK_N + half * HALF_N + c * 16;
                if constexpr (!LAST) {
                    uint32_t packed[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float a = sigmoid_fast(__uint_as_float(v[2 * j]) + s_bias[col + 2 * j]);
                        float b = sigmoid_fast(__uint_as_float(v[2 * j + 1]) + s_bias[col + 2 * j + 1]);
                        __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
                        packed[j] = *reinterpret_cast<uint32_t*>(&p);
                    }
                    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out) + (size_t)row * out_ld + col;
                    uint4* o4 = reinterpret_cast<uint4*>(o);
                    o4[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                    o4[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
                } else {
                    float f[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) f[j] = sigmoid_accurate(__uint_as_float(v[j]) + s_bias[col + j]);
                    float* o = reinterpret_cast<float*>(out) + (size_t)row * out_ld + col;
                    if (col + 16 <= n_valid) {
                        float4* o4 = reinterpret_cast<float4*>(o);
#pragma unroll
                        for (int j = 0; j < 4; ++j) o4[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (col + j < n_valid) o[j] = f[j];
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tempty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, ENC_TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------------- host side
typedef CUresult (*PFN_tensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                             CUtensorMapFloatOOBfill);

inline PFN_tensorMapEncodeTiled get_tensor_map_encoder() {
    static PFN_tensorMapEncodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<PFN_tensorMapEncodeTiled>(p);
    }
    return fn;
}

// bf16 row-major [rows][cols] matrix, box = 64 columns x box_rows rows, 128-byte swizzle.
inline bool make_bf16_kmajor_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    PFN_tensorMapEncodeTiled enc = get_tensor_map_encoder();
    if (!enc) return false;
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)ENC_BLOCK_K, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

struct EncoderLayerLaunch {
    CUtensorMap tmA, tmB;
    const float* bias;
    void* out;
    int out_ld, n_valid, K, n_pad, block_n;
    bool last;
};

template <int BLOCK_N, bool LAST>
inline cudaError_t launch_encoder_layer_t(const EncoderLayerLaunch& L, const int* m_ptr, int grid, cudaStream_t st) {
    auto kern = encoder_layer_kernel<BLOCK_N, LAST>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             EncSmem<BLOCK_N>::DYN_BYTES);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    kern<<<grid, ENC_THREADS, EncSmem<BLOCK_N>::DYN_BYTES, st>>>(L.tmA, L.tmB, L.bias, L.out, L.out_ld, L.n_valid,
                                                                 m_ptr, L.K, L.n_pad);
    return cudaGetLastError();
}

inline cudaError_t launch_encoder_layer(const EncoderLayerLaunch& L, const int* m_ptr, int grid, cudaStream_t st) {
    if (L.block_n == 256 && !L.last) return launch_encoder_layer_t<256, false>(L, m_ptr, grid, st);
    if (L.block_n == 160 && L.last) return launch_encoder_layer_t<160, true>(L, m_ptr, grid, st);
    if (L.block_n == 256 && L.last) return launch_encoder_layer_t<256, true>(L, m_ptr, grid, st);
    return cudaErrorInvalidValue;
}

}  // namespace hf6d
