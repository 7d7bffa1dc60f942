// Patch auto-encoder forward, one dense layer per launch:  Y = sigmoid(X * W^T + b)
//
// Replaces the Caffe InnerProduct+Sigmoid triple the reference runs per batch of 100 patches
// (reference: HoughForest/src/HFTest.cpp:585-596, net definition generate_scripts.sh:424-524).
//
// B200 design: persistent, warp-specialised tcgen05 kernel.
//   warp 0      : TMA producer  (A tile 128x64 bf16, W tile BLOCK_Nx64 bf16, 128-byte swizzle, 4-stage ring)
//   warp 1      : MMA issuer    (tcgen05.mma cta_group::1 kind::f16, M=128, N=BLOCK_N, K=16; fp32 accum in TMEM)
//   warp 2      : TMEM allocator (512 columns = two accumulator buffers, so epilogue(i) overlaps mma(i+1))
//   warps 4..11 : epilogue      (tcgen05.ld -> +bias -> sigmoid -> swizzled shared-memory staging -> TMA store)
// M (= number of processed patches P') is read from device memory so the launch is graph-capturable and needs no
// host round trip after the centre scan.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "ptx_sm100.cuh"

namespace hf6d {

constexpr int ENC_BLOCK_M = 128;
constexpr int ENC_BLOCK_K = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int ENC_THREADS = 384;
constexpr int ENC_EPI_WARPS = 8;
constexpr int ENC_TMEM_COLS = 512;
constexpr int ENC_MAX_N = 1536;
constexpr int ENC_EPI_BUF_BYTES = 32 * 64;  // one staged output chunk of a warp: 32 rows x 64 bytes

// STAGES: depth of the TMA->MMA operand ring.  EPI_BUFS: staged output chunks per epilogue warp (ring).
template <int BLOCK_N, int STAGES, int EPI_BUFS>
struct EncSmem {
    static constexpr int A_BYTES = ENC_BLOCK_M * ENC_BLOCK_K * 2;
    static constexpr int B_BYTES = BLOCK_N * ENC_BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int EPI_OFF = STAGES * STAGE_BYTES;
    static constexpr int BIAS_OFF = EPI_OFF + ENC_EPI_WARPS * EPI_BUFS * ENC_EPI_BUF_BYTES;
    static constexpr int BAR_OFF = BIAS_OFF + ENC_MAX_N * 4;
    static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 4) * 8 + 16;
    static constexpr int DYN_BYTES = TOTAL + 1024;  // slack for manual 1024-B alignment
    static_assert(DYN_BYTES <= 232448, "shared memory budget of one SM exceeded");
};

// sigma(x) = 0.5*tanh(x/2)+0.5 : one MUFU op per element (hidden layers; the result is rounded to bf16 anyway).
// h0, h1 are x/2.  (tanh.approx.f16x2 would not help: it issues one MUFU.TANH.F16 per half.)
__device__ __forceinline__ uint32_t sigmoid_pair_bf16(float h0, float h1) {
    float t0, t1;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(h0));
    asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(h1));
    const __nv_bfloat162 b = __floats2bfloat162_rn(fmaf(0.5f, t0, 0.5f), fmaf(0.5f, t1, 0.5f));
    return *reinterpret_cast<const uint32_t*>(&b);
}
__device__ __forceinline__ float sigmoid_accurate(float x) {
    // 1/(1+e^-x) with ex2.approx and rcp.approx: two MUFU ops, relative error ~2^-21 -- three orders of magnitude below
    // the bf16 operand rounding of the GEMM that feeds it.  Used for the feature layer the forest thresholds.
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return r;
}

// LAST=false : out is bf16 [m_cap][n_pad], all BLOCK_N columns stored (padded columns hold sigma(0)=0.5 and meet
//              zero weight columns in the next layer); `bias` holds 0.5*b (the tanh form wants x/2)
// LAST=true  : out is fp32 [m_cap][n_valid]; columns >= n_valid are clipped by the TMA store
// tmC is the output tensor map: boxes of 32 rows x 64 bytes, 64-byte swizzle.
template <int BLOCK_N, bool LAST, int STAGES, int EPI_BUFS>
__global__ void __launch_bounds__(ENC_THREADS, 1)
encoder_layer_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmC, const float* __restrict__ bias,
                     const int* __restrict__ m_ptr, int K, int n_pad) {
    using S = EncSmem<BLOCK_N, STAGES, EPI_BUFS>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

    float* s_bias = reinterpret_cast<float*>(smem + S::BIAS_OFF);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
    uint64_t* full = bars;                 // [STAGES]
    uint64_t* empty = bars + STAGES;       // [STAGES]
    uint64_t* tfull = bars + 2 * STAGES;   // [2]
    uint64_t* tempty = tfull + 2;          // [2]
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
        ptx::prefetch_tensormap(&tmC);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
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
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
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
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                ptx::umma_commit(&tfull[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ------------------------------------------------------------ epilogue
        // TMEM -> registers -> bias + sigmoid -> 64-byte row chunks staged in shared memory (64-byte swizzle, so the
        // 16-byte vector stores of a warp spread over all banks) -> TMA store: full-sector, fully coalesced writes
        // issued by the copy engine, off the LSU.
        const int q = warp & 3;               // TMEM lane quadrant this warp may touch
        const int half = (warp - 4) >> 2;     // which half of the tile's columns
        constexpr int HALF_N = BLOCK_N / 2;
        constexpr int CHUNK_COLS = LAST ? 16 : 32;  // 64 bytes of output per row
        constexpr int CHUNKS = HALF_N / CHUNK_COLS;
        static_assert(HALF_N % CHUNK_COLS == 0, "column half must be a whole number of 64-byte chunks");
        const uint32_t stage_base = ptx::smem_u32(smem + S::EPI_OFF + (warp - 4) * EPI_BUFS * ENC_EPI_BUF_BYTES);
        const uint32_t row_off = (uint32_t)lane * 64u;
        const uint32_t sw = ((uint32_t)lane >> 1) & 3u;  // 64-byte swizzle: 16-byte unit index ^= bits [7,9) of the address
        int it = 0;                                      // chunks staged by this warp so far
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
            const int mb = t / n_blocks, nb = t % n_blocks;
            const int row0 = mb * ENC_BLOCK_M + q * 32;
            ptx::mbar_wait(&tfull[acc], acc_phase);
            ptx::tc_fence_after();
            const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256 + half * HALF_N;
#pragma unroll 1
            for (int c = 0; c < CHUNKS; ++c, ++it) {
                const int col = nb * BLOCK_N + half * HALF_N + c * CHUNK_COLS;
                uint32_t o[16];
                if constexpr (!LAST) {
                    uint32_t v[32];
                    ptx::tmem_ld_32x32b_x32(taddr0 + c * CHUNK_COLS, v);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        o[j] = sigmoid_pair_bf16(fmaf(__uint_as_float(v[2 * j]), 0.5f, s_bias[col + 2 * j]),
                                                 fmaf(__uint_as_float(v[2 * j + 1]), 0.5f, s_bias[col + 2 * j + 1]));
                } else {
                    uint32_t v[16];
                    ptx::tmem_ld_32x32b_x16(taddr0 + c * CHUNK_COLS, v);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) o[j] = __float_as_uint(sigmoid_accurate(__uint_as_float(v[j]) + s_bias[col + j]));
                }
                if (c == CHUNKS - 1) {  // the accumulator is drained: hand it back before the stores
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&tempty[acc]);
                }
                const uint32_t buf = stage_base + (uint32_t)(it % EPI_BUFS) * ENC_EPI_BUF_BYTES;
                if (it >= EPI_BUFS) {  // the TMA store that last used this buffer must have read it
                    if (lane == 0) ptx::tma_store_wait_read<EPI_BUFS - 1>();
                    __syncwarp();
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    ptx::st_shared_v4(buf + row_off + (((uint32_t)u ^ sw) << 4), o[4 * u], o[4 * u + 1], o[4 * u + 2], o[4 * u + 3]);
                ptx::fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    ptx::tma_store_2d(&tmC, reinterpret_cast<const void*>(smem + (buf - ptx::smem_u32(smem))), col, row0);
                    ptx::tma_store_commit();
                }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (lane == 0) ptx::tma_store_wait_all<0>();  // shared memory must outlive the copies; writes complete before exit
        __syncwarp();
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

// Output map: row-major [rows][cols] of elem_bytes-wide elements, boxes of 32 rows x 64 bytes, 64-byte swizzle.
inline bool make_out_map(CUtensorMap* map, void* base, uint64_t rows, uint64_t cols, int elem_bytes) {
    PFN_tensorMapEncodeTiled enc = get_tensor_map_encoder();
    if (!enc) return false;
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {cols * (uint64_t)elem_bytes};
    cuuint32_t box[2] = {(cuuint32_t)(64 / elem_bytes), 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, gdim,
                     gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

struct EncoderLayerLaunch {
    CUtensorMap tmA, tmB, tmC;
    const float* bias;  // hidden layers: 0.5 * b
    int K, n_pad, block_n;
    bool last, short_k;  // short_k: few K blocks per tile -> 3 operand stages, deeper output staging
};

template <int BLOCK_N, bool LAST, int STAGES, int EPI_BUFS>
inline cudaError_t launch_encoder_layer_t(const EncoderLayerLaunch& L, const int* m_ptr, int grid, cudaStream_t st) {
    auto kern = encoder_layer_kernel<BLOCK_N, LAST, STAGES, EPI_BUFS>;
    using S = EncSmem<BLOCK_N, STAGES, EPI_BUFS>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::DYN_BYTES);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    kern<<<grid, ENC_THREADS, S::DYN_BYTES, st>>>(L.tmA, L.tmB, L.tmC, L.bias, m_ptr, L.K, L.n_pad);
    return cudaGetLastError();
}

inline cudaError_t launch_encoder_layer(const EncoderLayerLaunch& L, const int* m_ptr, int grid, cudaStream_t st) {
    if (L.block_n == 256 && !L.last && L.short_k) return launch_encoder_layer_t<256, false, 3, 4>(L, m_ptr, grid, st);
    if (L.block_n == 256 && !L.last) return launch_encoder_layer_t<256, false, 4, 1>(L, m_ptr, grid, st);
    if (L.block_n == 160 && L.last) return launch_encoder_layer_t<160, true, 4, 4>(L, m_ptr, grid, st);
    if (L.block_n == 256 && L.last) return launch_encoder_layer_t<256, true, 4, 1>(L, m_ptr, grid, st);
    return cudaErrorInvalidValue;
}

}  // namespace hf6d
