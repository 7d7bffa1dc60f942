// Patch auto-encoder forward, one dense layer per launch:  Y = sigmoid(X * W^T + b)
//
// Replaces the Caffe InnerProduct+Sigmoid triple the reference runs per batch of 100 patches
// (reference: HoughForest/src/HFTest.cpp:585-596, net definition generate_scripts.sh:424-524).
//
// B200 design: persistent, warp-specialised tcgen05 kernel, launched as CTA pairs (thread-block clusters of 2 = one TPC).
//   warp 0       : TMA producer  (A tile 128x64 bf16 + this CTA's HALF of the W tile, 128-byte swizzle, mbarrier ring)
//   warp 1       : MMA issuer    (even CTA only: tcgen05.mma cta_group::2 kind::f16, M=256 over the pair, N=BLOCK_N, K=16;
//                                 each CTA's 128 rows accumulate in fp32 in its own TMEM)
//   warp 2       : TMEM allocator (512 columns = two accumulator buffers, so epilogue(i) overlaps mma(i+1))
//   warps 4..    : epilogue      (tcgen05.ld -> +bias -> sigmoid -> swizzled shared-memory staging -> TMA store)
// Why pairs: with cta_group::1 every MMA streams A (4 KB) + the whole W slice (N x 32 B) out of shared memory while TMA
// writes A + W into it; at N = 160 that is more than the shared-memory pipe delivers (tensor pipe 47 % busy).  In a pair
// each CTA holds and reads only half of W, and pulls A + W/2 per k-block from L2.
// M (= number of processed patches P') is read from device memory so the launch is graph-capturable and needs no
// host round trip after the centre scan.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>

#include "ptx_sm100.cuh"

namespace hf6d {

constexpr int ENC_BLOCK_M = 128;
constexpr int ENC_BLOCK_K = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int ENC_TMEM_COLS = 512;
constexpr int ENC_MAX_N = 1536;
// STAGES: depth of the TMA->MMA operand ring.  EPI_BUFS: staged output chunks per epilogue warp (ring), each 32 rows x
// CHUNK_BYTES (64 or 128: the inner extent of the TMA store box; 128-byte rows halve the number of stores).
// PAIR: 1 = stand-alone CTAs (cta_group::1), 2 = CTA pairs (cta_group::2).  EPI_WARPS: 8 or 16 epilogue warps.
template <int BLOCK_N, int STAGES, int EPI_BUFS, int PAIR, int CHUNK_BYTES, int EPI_WARPS>
struct EncSmem {
    static constexpr int THREADS = (4 + EPI_WARPS) * 32;
    static constexpr int EPI_BUF_BYTES = 32 * CHUNK_BYTES;
    static constexpr int A_BYTES = ENC_BLOCK_M * ENC_BLOCK_K * 2;
    static constexpr int B_BYTES = (BLOCK_N / PAIR) * ENC_BLOCK_K * 2;  // this CTA's share of the weight tile
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int EPI_OFF = STAGES * STAGE_BYTES;
    static constexpr int BIAS_OFF = EPI_OFF + EPI_WARPS * EPI_BUFS * EPI_BUF_BYTES;
    static constexpr int BAR_OFF = BIAS_OFF + ENC_MAX_N * 4;
    static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 4) * 8 + 16;
    static constexpr int DYN_BYTES = TOTAL + 1024;  // slack for manual 1024-B alignment
    static_assert(STAGE_BYTES % 1024 == 0, "operand tiles must stay 1024-byte aligned (128-byte swizzle atoms)");
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
// the same with fp16 results (encoder mode 2: fp16 operands, 11-bit significands instead of bf16's 8)
__device__ __forceinline__ uint32_t sigmoid_pair_f16(float h0, float h1) {
    float t0, t1;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(h0));
    asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(h1));
    const __half2 b = __floats2half2_rn(fmaf(0.5f, t0, 0.5f), fmaf(0.5f, t1, 0.5f));
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

// 1 / (1 + e^-x) in full fp32 (expf and an IEEE division): the split-bf16 mode's activation, where the tanh / ex2 forms'
// 2^-11 .. 2^-21 relative error would be the largest error left.
__device__ __forceinline__ float sigmoid_fp32(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }
// y -> (hi, lo) bf16 with hi + lo = y to ~2^-17 relative
__device__ __forceinline__ void split_pair_bf16(float y0, float y1, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(y0, y1);
    const float2 hf = __bfloat1622float2(h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(__fsub_rn(y0, hf.x), __fsub_rn(y1, hf.y));
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// SPLIT (hf6d_set_encoder_mode(ctx, 1)): near-fp32 products out of bf16 tensor-core operands.  Every operand x is held as
// two bf16 numbers x_hi = bf16(x), x_lo = bf16(x - x_hi), and a product a*w is accumulated (fp32, TMEM) as
//     a_hi*w_hi + a_lo*w_hi + a_hi*w_lo            (a_lo*w_lo ~ 2^-18 |a w| is dropped)
// -- three passes over K instead of one, expressed as ONE k-loop of n_seg * K / 64 blocks whose segments address different
// column ranges of the operand matrices: A = [a_hi | a_lo] ([M][2K]), W = [w_hi | w_lo] ([N][2K]);
//     segment 0: A cols [0,K) x W cols [0,K);  1: A [K,2K) x W [0,K);  2 (the last): A [0,K) x W [K,2K).
// The first layer's input (quantised patch values 0..255) is exact in bf16, so it runs two segments (A x w_hi, A x w_lo).
// Hidden layers store sigmoid(x) as hi and lo halves ([M][2 n_pad]: lo at column lo_off + n), the feature layer fp32.
// Same ring, same barriers, same MMA shape as the bf16 mode; measured against the fp32 oracle in tests/test_gpu_e2e.py.
//
// LAST=false : out is bf16 [m_cap][n_pad], all BLOCK_N columns stored (padded columns hold sigma(0)=0.5 and meet
//              zero weight columns in the next layer); `bias` holds 0.5*b (the tanh form wants x/2)
// LAST=true  : out is fp32 [m_cap][n_valid]; columns >= n_valid are clipped by the TMA store
// OUT16      : (LAST only) out is fp16 [m_cap][n_valid] -- the feature storage the traversal can read at half the HBM bytes
// tmC is the output tensor map: boxes of 32 rows x CHUNK_BYTES, swizzle = CHUNK_BYTES.
// PAIR = 2   : the two CTAs of a cluster take the m-blocks 2*g and 2*g + 1 of the SAME n-block (a missing last m-block is
//              computed on whatever the padded rows hold and clipped by the TMA store).  tmB has a box of BLOCK_N / 2 rows:
//              CTA r loads weight rows [r * BLOCK_N / 2, (r + 1) * BLOCK_N / 2) of the tile.  Barrier protocol:
//                full[s]   lives in the even CTA: one arrive.expect_tx by its producer for the bytes of BOTH CTAs; both
//                          producers' TMA loads complete on it (cp.async.bulk.tensor.cta_group::2)
//                empty[s]  one per CTA: tcgen05.commit.cta_group::2 multicast frees the stage in both
//                tfull[a]  one per CTA, same multicast commit; tempty[a] in the even CTA counts the epilogue warps of both
template <int BLOCK_N, bool LAST, int STAGES, int EPI_BUFS, int PAIR, int CHUNK_BYTES, int EPI_WARPS, bool SPLIT = false,
          bool OUT16 = false>
__global__ void __launch_bounds__((4 + EPI_WARPS) * 32, 1)
encoder_layer_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmC, const float* __restrict__ bias,
                     const int* __restrict__ m_ptr, int K, int n_pad, int reverse_m, int n_seg, int lo_off, int shard_rank,
                     int shard_world, int fp16, float in_scale) {
    using S = EncSmem<BLOCK_N, STAGES, EPI_BUFS, PAIR, CHUNK_BYTES, EPI_WARPS>;
    static_assert(PAIR == 1 || PAIR == 2, "stand-alone CTAs or CTA pairs");
    static_assert(EPI_WARPS % 4 == 0, "every TMEM lane quadrant needs the same number of epilogue warps");
    static_assert(!OUT16 || (LAST && !SPLIT), "fp16 storage is for the feature layer of the bf16 / fp16 operand modes");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

    float* s_bias = reinterpret_cast<float*>(smem + S::BIAS_OFF);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
    uint64_t* full = bars;                 // [STAGES]
    uint64_t* empty = bars + STAGES;       // [STAGES]
    uint64_t* tfull = bars + 2 * STAGES;   // [2]
    uint64_t* tempty = tfull + 2;          // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    // warp-uniform by construction, and known to be so by the compiler (see ptx::elect_one)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;

    const int M = __shfl_sync(0xffffffffu, *m_ptr, 0);
    // patch sharding: this rank's row blocks [mb_lo, mb_lo + m_blocks) of the frame's ceil(M / 128) (common.cuh, PatchShard)
    const int m_blocks_all = (M + ENC_BLOCK_M - 1) / ENC_BLOCK_M;
    const int mb_lo = (int)((long long)m_blocks_all * shard_rank / shard_world);
    const int m_blocks = (int)((long long)m_blocks_all * (shard_rank + 1) / shard_world) - mb_lo;
    const int n_blocks = n_pad / BLOCK_N;
    const int k_seg = K / ENC_BLOCK_K;                 // k-blocks per segment
    const int k_blocks = (SPLIT ? n_seg : 1) * k_seg;  // k-blocks per tile
    // tile schedule: work item -> (m-block group, n-block); both CTAs of a pair walk the same items
    const int cta_rank = PAIR > 1 ? (int)ptx::cluster_ctarank() : 0;
    const int m_groups = (m_blocks + PAIR - 1) / PAIR;
    const int tiles = m_groups * n_blocks;
    const int first = blockIdx.x / PAIR, step = gridDim.x / PAIR;
    // reverse_m: walk the m-blocks from the end.  The layers alternate direction so that each one starts on the rows its
    // producer wrote LAST -- the part of the activation matrix (145-227 MB) that is still in the 126 MB L2.
    auto m_group_of = [&](int g) { return reverse_m ? m_groups - 1 - g : g; };
    // (the m-block index below is relative to mb_lo: `mb_lo +` is added where it becomes a row coordinate)

    for (int i = threadIdx.x; i < n_pad; i += S::THREADS) s_bias[i] = bias[i];

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
            ptx::mbar_init(&tempty[a], PAIR * EPI_WARPS);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        if constexpr (PAIR > 1) {
            ptx::tmem_alloc_pair(tmem_slot, ENC_TMEM_COLS);
            ptx::tmem_relinquish_pair();
        } else {
            ptx::tmem_alloc(tmem_slot, ENC_TMEM_COLS);
            ptx::tmem_relinquish();
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (PAIR > 1) ptx::cluster_sync();  // the peer's barriers are initialised before anything signals them
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (whole warp walks, one lane issues)
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t full0 = PAIR > 1 ? ptx::mapa_shared(ptx::smem_u32(&full[0]), 0) : 0;  // the even CTA's barriers
        for (int t = first; t < tiles; t += step) {
            const int mb = mb_lo + m_group_of(t / n_blocks) * PAIR + cta_rank, nb = t % n_blocks;
            for (int kb = 0; kb < k_blocks; ++kb) {
                // operand columns of this k-block: the bf16 mode walks both matrices in step; the split mode's segments
                // pair (a_hi, w_hi), (a_lo, w_hi), (a_hi, w_lo) -- the last segment reads the lo half of W, the middle one
                // of three the lo half of A
                int ka = kb, kw = kb;
                if constexpr (SPLIT) {
                    const int seg = kb / k_seg, j = kb - seg * k_seg;
                    ka = (n_seg == 3 && seg == 1) ? k_seg + j : j;
                    kw = (seg == n_seg - 1) ? k_seg + j : j;
                }
                ptx::mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* sa = smem + stage * S::STAGE_BYTES;
                uint8_t* sb = sa + S::A_BYTES;
                if (ptx::elect_one()) {
                    if constexpr (PAIR > 1) {
                        if (cta_rank == 0) ptx::mbar_arrive_expect_tx(&full[stage], PAIR * S::STAGE_BYTES);
                        ptx::tma_load_2d_pair(sa, &tmA, full0 + stage * 8, ka * ENC_BLOCK_K, mb * ENC_BLOCK_M);
                        ptx::tma_load_2d_pair(sb, &tmB, full0 + stage * 8, kw * ENC_BLOCK_K,
                                              nb * BLOCK_N + cta_rank * (BLOCK_N / 2));
                    } else {
                        ptx::mbar_arrive_expect_tx(&full[stage], S::STAGE_BYTES);
                        ptx::tma_load_2d(sa, &tmA, &full[stage], ka * ENC_BLOCK_K, mb * ENC_BLOCK_M);
                        ptx::tma_load_2d(sb, &tmB, &full[stage], kw * ENC_BLOCK_K, nb * BLOCK_N);
                    }
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (the even CTA of a pair; one lane issues)
        if (cta_rank == 0) {
            // operand format: bf16, or fp16 in encoder mode 2 (same tiles, same rate; only the descriptor's format fields differ)
            const uint32_t idesc = fp16 ? ptx::make_idesc_bf16_f32(ENC_BLOCK_M * PAIR, BLOCK_N, true)
                                        : ptx::make_idesc_bf16_f32(ENC_BLOCK_M * PAIR, BLOCK_N, false);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int t = first; t < tiles; t += step) {
                ptx::mbar_wait<(PAIR > 1)>(&tempty[acc], acc_phase ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * 256;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    ptx::mbar_wait(&full[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t sa = ptx::smem_u32(smem + stage * S::STAGE_BYTES);
                    const uint64_t adesc = ptx::make_kmajor_sw128_desc(sa);
                    const uint64_t bdesc = ptx::make_kmajor_sw128_desc(sa + S::A_BYTES);
                    if (ptx::elect_one()) {
#pragma unroll
                        for (int k = 0; k < ENC_BLOCK_K / 16; ++k) {
                            // +32 B per K=16 step inside the 128-B swizzle row: +2 in the (addr>>4) field
                            if constexpr (PAIR > 1) ptx::umma_bf16_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                            else ptx::umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                        }
                        if constexpr (PAIR > 1) ptx::umma_commit_pair(&empty[stage], (uint16_t)0x3);
                        else ptx::umma_commit(&empty[stage]);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (ptx::elect_one()) {
                    if constexpr (PAIR > 1) ptx::umma_commit_pair(&tfull[acc], (uint16_t)0x3);
                    else ptx::umma_commit(&tfull[acc]);
                }
                __syncwarp();
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------ epilogue
        // TMEM -> registers -> bias + sigmoid -> row chunks staged in shared memory (swizzled, so the 16-byte vector
        // stores of a warp spread over all banks) -> TMA store: full-sector, fully coalesced writes issued by the copy
        // engine, off the LSU.
        const int q = warp & 3;               // TMEM lane quadrant this warp may touch
        const int sub = (warp - 4) >> 2;      // the warps of a quadrant take the column chunks round-robin
        constexpr int SUBS = EPI_WARPS / 4;
        constexpr int CHUNK_COLS = CHUNK_BYTES / ((LAST && !OUT16) ? 4 : 2);  // output columns per staged chunk
        constexpr int TILE_CHUNKS = BLOCK_N / CHUNK_COLS;
        constexpr int UNITS = CHUNK_BYTES / 16;                   // 16-byte units per staged row
        static_assert(BLOCK_N % CHUNK_COLS == 0, "the tile must be a whole number of chunks");
        static_assert(CHUNK_COLS == 16 || CHUNK_COLS == 32 || CHUNK_COLS == 64, "TMEM loads come in x16 / x32");
        const uint32_t stage_base = ptx::smem_u32(smem + S::EPI_OFF + (warp - 4) * EPI_BUFS * S::EPI_BUF_BYTES);
        const uint32_t row_off = (uint32_t)lane * CHUNK_BYTES;
        // TMA swizzle: the 16-byte unit index is XORed with address bits [7, 7 + log2(UNITS))
        const uint32_t sw = UNITS == 8 ? ((uint32_t)lane & 7u) : (((uint32_t)lane >> 1) & 3u);
        const uint32_t tempty0 = PAIR > 1 ? ptx::mapa_shared(ptx::smem_u32(&tempty[0]), 0) : 0;  // the even CTA's barriers
        int it = 0;                                      // chunks staged by this warp so far
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int t = first; t < tiles; t += step) {
            const int mb = mb_lo + m_group_of(t / n_blocks) * PAIR + cta_rank, nb = t % n_blocks;
            const int row0 = mb * ENC_BLOCK_M + q * 32;
            ptx::mbar_wait(&tfull[acc], acc_phase);
            ptx::tc_fence_after();
            const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256;
#pragma unroll 1
            for (int c = sub; c < TILE_CHUNKS; c += SUBS, ++it) {
                const int tcol = c * CHUNK_COLS;             // column inside the tile
                const int col = nb * BLOCK_N + tcol;
                uint32_t v[CHUNK_COLS];
                if constexpr (CHUNK_COLS == 16) {
                    ptx::tmem_ld_32x32b_x16(taddr0 + tcol, v);
                } else {
#pragma unroll
                    for (int h = 0; h < CHUNK_COLS / 32; ++h)
                        ptx::tmem_ld_32x32b_x32(taddr0 + tcol + h * 32, *reinterpret_cast<uint32_t(*)[32]>(&v[h * 32]));
                }
                ptx::tmem_ld_wait();
                if (c + SUBS >= TILE_CHUNKS) {  // this warp has drained its part of the accumulator: hand it back now
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if constexpr (PAIR > 1) ptx::mbar_arrive_cluster(tempty0 + acc * 8);
                        else ptx::mbar_arrive(&tempty[acc]);
                    }
                }
                uint32_t o[UNITS * 4];
                uint32_t o_lo[(SPLIT && !LAST) ? UNITS * 4 : 1];  // split mode, hidden layers: the lo halves of the same chunk
                const uint32_t bias_addr = ptx::smem_u32(s_bias + col);  // broadcast reads, 16 bytes each
                if constexpr (!LAST && SPLIT) {
#pragma unroll
                    for (int j = 0; j < CHUNK_COLS / 4; ++j) {
                        const float4 b = ptx::ld_shared_f4(bias_addr + 16 * j);
                        split_pair_bf16(sigmoid_fp32(__fadd_rn(__uint_as_float(v[4 * j]), b.x)),
                                        sigmoid_fp32(__fadd_rn(__uint_as_float(v[4 * j + 1]), b.y)), o[2 * j], o_lo[2 * j]);
                        split_pair_bf16(sigmoid_fp32(__fadd_rn(__uint_as_float(v[4 * j + 2]), b.z)),
                                        sigmoid_fp32(__fadd_rn(__uint_as_float(v[4 * j + 3]), b.w)), o[2 * j + 1], o_lo[2 * j + 1]);
                    }
                } else if constexpr (!LAST) {
                    // x/2 = acc * (in_scale / 2) + b/2: in_scale is 1 except for the first layer of the fp16 mode, whose
                    // weights are NOT pre-divided by 255 (they would drop into fp16's subnormals)
                    const float hs = 0.5f * in_scale;
                    if (fp16) {
#pragma unroll
                        for (int j = 0; j < CHUNK_COLS / 4; ++j) {
                            const float4 b = ptx::ld_shared_f4(bias_addr + 16 * j);
                            o[2 * j] = sigmoid_pair_f16(fmaf(__uint_as_float(v[4 * j]), hs, b.x), fmaf(__uint_as_float(v[4 * j + 1]), hs, b.y));
                            o[2 * j + 1] = sigmoid_pair_f16(fmaf(__uint_as_float(v[4 * j + 2]), hs, b.z), fmaf(__uint_as_float(v[4 * j + 3]), hs, b.w));
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < CHUNK_COLS / 4; ++j) {
                            const float4 b = ptx::ld_shared_f4(bias_addr + 16 * j);
                            o[2 * j] = sigmoid_pair_bf16(fmaf(__uint_as_float(v[4 * j]), 0.5f, b.x),
                                                         fmaf(__uint_as_float(v[4 * j + 1]), 0.5f, b.y));
                            o[2 * j + 1] = sigmoid_pair_bf16(fmaf(__uint_as_float(v[4 * j + 2]), 0.5f, b.z),
                                                             fmaf(__uint_as_float(v[4 * j + 3]), 0.5f, b.w));
                        }
                    }
                } else if constexpr (OUT16) {  // the feature layer stored as fp16 (feature storage 1): same sigmoid, half the bytes
#pragma unroll
                    for (int j = 0; j < CHUNK_COLS / 4; ++j) {
                        const float4 b = ptx::ld_shared_f4(bias_addr + 16 * j);
                        const __half2 p0 = __floats2half2_rn(sigmoid_accurate(__uint_as_float(v[4 * j]) + b.x),
                                                             sigmoid_accurate(__uint_as_float(v[4 * j + 1]) + b.y));
                        const __half2 p1 = __floats2half2_rn(sigmoid_accurate(__uint_as_float(v[4 * j + 2]) + b.z),
                                                             sigmoid_accurate(__uint_as_float(v[4 * j + 3]) + b.w));
                        o[2 * j] = *reinterpret_cast<const uint32_t*>(&p0);
                        o[2 * j + 1] = *reinterpret_cast<const uint32_t*>(&p1);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < CHUNK_COLS / 4; ++j) {
                        const float4 b = ptx::ld_shared_f4(bias_addr + 16 * j);
                        if constexpr (SPLIT) {
                            o[4 * j] = __float_as_uint(sigmoid_fp32(__fadd_rn(__uint_as_float(v[4 * j]), b.x)));
                            o[4 * j + 1] = __float_as_uint(sigmoid_fp32(__fadd_rn(__uint_as_float(v[4 * j + 1]), b.y)));
                            o[4 * j + 2] = __float_as_uint(sigmoid_fp32(__fadd_rn(__uint_as_float(v[4 * j + 2]), b.z)));
                            o[4 * j + 3] = __float_as_uint(sigmoid_fp32(__fadd_rn(__uint_as_float(v[4 * j + 3]), b.w)));
                        } else {
                            o[4 * j] = __float_as_uint(sigmoid_accurate(__uint_as_float(v[4 * j]) + b.x));
                            o[4 * j + 1] = __float_as_uint(sigmoid_accurate(__uint_as_float(v[4 * j + 1]) + b.y));
                            o[4 * j + 2] = __float_as_uint(sigmoid_accurate(__uint_as_float(v[4 * j + 2]) + b.z));
                            o[4 * j + 3] = __float_as_uint(sigmoid_accurate(__uint_as_float(v[4 * j + 3]) + b.w));
                        }
                    }
                }
                // one staged chunk + TMA store per half (the bf16 mode and the feature layer have a single "half")
#pragma unroll
                for (int half = 0; half < ((SPLIT && !LAST) ? 2 : 1); ++half) {
                    const uint32_t* ov = half ? o_lo : o;
                    const uint32_t buf = stage_base + (uint32_t)(it % EPI_BUFS) * S::EPI_BUF_BYTES;
                    if (it >= EPI_BUFS) {  // the TMA store that last used this buffer must have read it
                        if (ptx::elect_one()) ptx::tma_store_wait_read<EPI_BUFS - 1>();  // always the same lane: bulk groups are per thread
                        __syncwarp();
                    }
#pragma unroll
                    for (int u = 0; u < UNITS; ++u)
                        ptx::st_shared_v4(buf + row_off + (((uint32_t)u ^ sw) << 4), ov[4 * u], ov[4 * u + 1], ov[4 * u + 2], ov[4 * u + 3]);
                    ptx::fence_proxy_async();
                    __syncwarp();
                    if (ptx::elect_one()) {
                        ptx::tma_store_2d(&tmC, reinterpret_cast<const void*>(smem + (buf - ptx::smem_u32(smem))), col + half * lo_off, row0);
                        ptx::tma_store_commit();
                    }
                    if (half + 1 < ((SPLIT && !LAST) ? 2 : 1)) ++it;
                }
            }
            if (sub >= TILE_CHUNKS) {  // a warp without any chunk still has to release the accumulator
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if constexpr (PAIR > 1) ptx::mbar_arrive_cluster(tempty0 + acc * 8);
                    else ptx::mbar_arrive(&tempty[acc]);
                }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (ptx::elect_one()) ptx::tma_store_wait_all<0>();  // shared memory must outlive the copies; writes complete before exit
        __syncwarp();
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (PAIR > 1) ptx::cluster_sync();  // no CTA leaves while its peer's MMAs read its operands or signal its barriers
    if (warp == 2) {
        ptx::tc_fence_after();
        if constexpr (PAIR > 1) ptx::tmem_dealloc_pair(tmem_base, ENC_TMEM_COLS);
        else ptx::tmem_dealloc(tmem_base, ENC_TMEM_COLS);
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

// Output map: row-major [rows][cols] of elem_bytes-wide elements, boxes of 32 rows x chunk_bytes (64 or 128), swizzle =
// chunk_bytes.
inline bool make_out_map(CUtensorMap* map, void* base, uint64_t rows, uint64_t cols, int elem_bytes, int chunk_bytes) {
    PFN_tensorMapEncodeTiled enc = get_tensor_map_encoder();
    if (!enc) return false;
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {cols * (uint64_t)elem_bytes};
    cuuint32_t box[2] = {(cuuint32_t)(chunk_bytes / elem_bytes), 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, gdim,
                     gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, chunk_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

struct EncoderLayerLaunch {
    CUtensorMap tmA, tmB, tmC;  // tmB: box of block_n / pair rows; tmC: boxes of 32 rows x chunk bytes
    const float* bias;          // hidden layers: 0.5 * b (bf16 mode), b (split mode)
    int K, n_pad, block_n;      // K: columns of ONE half of the operands in split mode
    bool last, short_k;  // short_k: few K blocks per tile
    int variant;         // row of HF6D_ENC_CONFIGS for this layer's shape class
    int reverse_m;       // walk the m-blocks from the end (see the kernel)
    bool split;          // split-bf16 arithmetic (see the kernel): variant then selects pairs (0) or stand-alone CTAs (1)
    int n_seg, lo_off;   // split: k segments (2: exact A, 3: hi/lo A); column offset of the lo halves in the hidden output
    int shard_rank, shard_world;  // patch sharding: the row blocks this rank encodes (0 / 0 or 1 = all)
    int fp16;                     // operands are fp16 instead of bf16 (encoder mode 2)
    bool out16;                   // feature layer only: fp16 output (HF6D_ENC_OUT16_CONFIGS)
    float in_scale;               // hidden layers: accumulator scale before the bias (0 = 1.0; 1/255 for fp16 layer 1)
};

template <int BLOCK_N, bool LAST, int STAGES, int EPI_BUFS, int PAIR, int CHUNK_BYTES, int EPI_WARPS, bool SPLIT = false,
          bool OUT16 = false>
inline cudaError_t launch_encoder_layer_t(const EncoderLayerLaunch& L, const int* m_ptr, int sms, cudaStream_t st, bool probe) {
    auto kern = encoder_layer_kernel<BLOCK_N, LAST, STAGES, EPI_BUFS, PAIR, CHUNK_BYTES, EPI_WARPS, SPLIT, OUT16>;
    using S = EncSmem<BLOCK_N, STAGES, EPI_BUFS, PAIR, CHUNK_BYTES, EPI_WARPS>;
    static int grid = 0;
    if (!grid) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::DYN_BYTES);
        if (e != cudaSuccess) return e;
        grid = sms;
        if (PAIR > 1) {  // as many co-resident pairs as the GPCs can hold
            cudaLaunchConfig_t q{};
            q.gridDim = dim3(sms / PAIR * PAIR);
            q.blockDim = dim3(S::THREADS);
            q.dynamicSmemBytes = S::DYN_BYTES;
            cudaLaunchAttribute a[1];
            a[0].id = cudaLaunchAttributeClusterDimension;
            a[0].val.clusterDim.x = PAIR; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
            q.attrs = a; q.numAttrs = 1;
            int n_clusters = 0;
            e = cudaOccupancyMaxActiveClusters(&n_clusters, kern, &q);
            if (e != cudaSuccess) return e;
            if (n_clusters < 1) return cudaErrorInvalidConfiguration;
            grid = std::min(n_clusters, sms / PAIR) * PAIR;
        }
    }
    if (probe) return cudaSuccess;  // the kernel fits and (for pairs) the GPCs can co-schedule the clusters
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(S::THREADS);
    cfg.dynamicSmemBytes = S::DYN_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attrs[1];
    attrs[0].id = cudaLaunchAttributeClusterDimension;
    attrs[0].val.clusterDim.x = PAIR; attrs[0].val.clusterDim.y = 1; attrs[0].val.clusterDim.z = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, L.tmA, L.tmB, L.tmC, L.bias, m_ptr, L.K, L.n_pad, L.reverse_m, SPLIT ? L.n_seg : 1,
                              L.lo_off, L.shard_world > 1 ? L.shard_rank : 0, L.shard_world > 1 ? L.shard_world : 1, L.fp16,
                              L.in_scale == 0.f ? 1.0f : L.in_scale);
}

// The kernel configurations, one table for the launcher and for the slot's tensor maps (W box rows, output chunk width).
//   shape class: 0 = hidden layer, short K (<= 6 k-blocks);  1 = hidden layer;  2 = feature layer N % 160 == 0;
//                3 = feature layer, 256-wide tiles;  4 = feature layer, 208-wide tiles (pairs only)
//   variant 0 is the default of its class; the others are kept for HF6D_ENC_VARIANT="a,b,c" (per layer) experiments and
//   as the fallback when CTA pairs cannot be scheduled (variant 1: stand-alone CTAs, cta_group::1; class 4 has none).
//   Variant 0 is the PIPELINED default (contexts with several frame slots), variant 2 of classes 0 / 1 / 4 the one that is
//   fastest alone (one more ring stage; one-slot contexts use it): with several frames in flight the smaller footprints
//   (160-180 KB instead of 192-218 KB of shared memory) leave room for other frames' gather / vote / blur CTAs beside the
//   persistent encoder CTA, which is worth +3-5 % frames/s for 2-5 % of encoder time.
#define HF6D_ENC_CONFIGS(X)                                        \
    /*  cls var  N   LAST  ST EB PAIR CHUNK EPI      B200, 70.9 k patches: us per layer (alone) */ \
    X(0, 0, 256, false, 4, 1, 2, 128, 8)  /* 59.5 */               \
    X(0, 1, 256, false, 3, 2, 1, 128, 8)  /* 65.5 */               \
    X(0, 2, 256, false, 5, 1, 2, 128, 8)  /* 57.7 */               \
    X(0, 3, 256, false, 4, 2, 2, 64, 16)  /* 59.6 */               \
    X(0, 4, 256, false, 6, 1, 2, 64, 8)   /* 63.7 */               \
    X(0, 5, 256, false, 4, 2, 2, 128, 8)  /* 60.1 */               \
    X(1, 0, 256, false, 5, 1, 2, 64, 8)   /* 131.5 */              \
    X(1, 1, 256, false, 4, 1, 1, 64, 8)   /* 151.6 */              \
    X(1, 2, 256, false, 6, 1, 2, 64, 8)   /* 129.0 */              \
    X(1, 3, 256, false, 5, 2, 2, 64, 8)   /* 135.1 */              \
    X(1, 4, 256, false, 4, 2, 2, 128, 8)  /* 146.2 */              \
    X(2, 0, 160, true, 7, 1, 2, 128, 8)   /* 98.3 */               \
    X(2, 1, 160, true, 4, 2, 1, 128, 8)   /* 118.8 */              \
    X(2, 2, 160, true, 5, 2, 2, 128, 8)   /* 104.5 */              \
    X(2, 3, 160, true, 7, 2, 2, 64, 8)    /* 98.5 */               \
    X(2, 4, 160, true, 4, 2, 2, 128, 12)  /* 122.9 */              \
    X(4, 0, 208, true, 5, 2, 2, 64, 8)    /* 95.5 */               \
    X(4, 1, 208, true, 5, 4, 2, 64, 8)    /* 96.2 */               \
    X(4, 2, 208, true, 6, 2, 2, 64, 8)    /* 94.2 */               \
    X(4, 3, 208, true, 4, 2, 2, 64, 8)    /* 113 */                \
    X(3, 0, 256, true, 4, 2, 2, 64, 8)                             \
    X(3, 1, 256, true, 4, 1, 1, 64, 8)

// Split-bf16 mode (hf6d_set_encoder_mode 1): one configuration per tile shape -- it is the accuracy mode, three passes over
// K, so there is no tuning table.  EPI_BUFS = 2: a hidden-layer chunk is staged and stored twice (hi and lo halves).
//   variant 0 = CTA pairs, variant 1 = stand-alone CTAs (devices that cannot co-schedule clusters; no 208-wide tiles).
#define HF6D_ENC_SPLIT_CONFIGS(X)                  \
    /*  var  N   LAST  ST EB PAIR CHUNK EPI */     \
    X(0, 256, false, 4, 2, 2, 128, 8)              \
    X(0, 208, true, 5, 2, 2, 64, 8)                \
    X(0, 160, true, 5, 2, 2, 128, 8)               \
    X(0, 256, true, 4, 2, 2, 64, 8)                \
    X(1, 256, false, 3, 2, 1, 128, 8)              \
    X(1, 160, true, 4, 2, 1, 128, 8)               \
    X(1, 256, true, 3, 2, 1, 64, 8)

// Feature layer with fp16 output (feature storage 1): 32-column chunks (64-byte store boxes), so the tile is 160 or 256 wide.
//   variant 0 = CTA pairs, pipelined footprint; 1 = stand-alone CTAs; 2 = CTA pairs, deepest ring (one-slot contexts).
#define HF6D_ENC_OUT16_CONFIGS(X)                  \
    /*  var  N   ST EB PAIR CHUNK EPI */           \
    X(0, 160, 5, 2, 2, 64, 8)                      \
    X(1, 160, 4, 2, 1, 64, 8)                      \
    X(2, 160, 7, 2, 2, 64, 8)                      \
    X(0, 256, 4, 2, 2, 64, 8)                      \
    X(1, 256, 3, 2, 1, 64, 8)                      \
    X(2, 256, 4, 2, 2, 64, 8)

inline int encoder_shape_class(int block_n, bool last, bool short_k) {
    if (!last) return short_k ? 0 : 1;
    return block_n == 160 ? 2 : block_n == 208 ? 4 : 3;
}

struct EncoderConfig {
    int pair, chunk_bytes;
};
// pair == 0: no such variant
inline EncoderConfig encoder_config(int block_n, bool last, bool short_k, int variant, bool split = false, bool out16 = false) {
    if (out16) {
#define X(VAR, N, ST, EB, PAIR, CHUNK, EPI) \
    if (variant == VAR && block_n == N && last) return EncoderConfig{PAIR, CHUNK};
        HF6D_ENC_OUT16_CONFIGS(X)
#undef X
        return EncoderConfig{0, 0};
    }
    if (split) {
#define X(VAR, N, LAST, ST, EB, PAIR, CHUNK, EPI) \
    if (variant == VAR && block_n == N && last == LAST) return EncoderConfig{PAIR, CHUNK};
        HF6D_ENC_SPLIT_CONFIGS(X)
#undef X
        return EncoderConfig{0, 0};
    }
    const int cls = encoder_shape_class(block_n, last, short_k);
#define X(CLS, VAR, N, LAST, ST, EB, PAIR, CHUNK, EPI) \
    if (cls == CLS && variant == VAR && block_n == N) return EncoderConfig{PAIR, CHUNK};
    HF6D_ENC_CONFIGS(X)
#undef X
    return EncoderConfig{0, 0};
}

// probe = true: only check that this variant can run on the current device (shared memory opt-in, cluster occupancy).
inline cudaError_t launch_encoder_layer(const EncoderLayerLaunch& L, const int* m_ptr, int sms, cudaStream_t st,
                                        bool probe = false) {
    if (L.out16) {
#define X(VAR, N, ST, EB, PAIR, CHUNK, EPI)                         \
    if (L.variant == VAR && L.block_n == N && L.last && !L.split)   \
        return launch_encoder_layer_t<N, true, ST, EB, PAIR, CHUNK, EPI, false, true>(L, m_ptr, sms, st, probe);
        HF6D_ENC_OUT16_CONFIGS(X)
#undef X
        return cudaErrorInvalidValue;
    }
    if (L.split) {
#define X(VAR, N, LAST, ST, EB, PAIR, CHUNK, EPI)                   \
    if (L.variant == VAR && L.block_n == N && L.last == LAST)       \
        return launch_encoder_layer_t<N, LAST, ST, EB, PAIR, CHUNK, EPI, true>(L, m_ptr, sms, st, probe);
        HF6D_ENC_SPLIT_CONFIGS(X)
#undef X
        return cudaErrorInvalidValue;
    }
    const int cls = encoder_shape_class(L.block_n, L.last, L.short_k);
#define X(CLS, VAR, N, LAST, ST, EB, PAIR, CHUNK, EPI)                 \
    if (cls == CLS && L.variant == VAR && L.block_n == N)              \
        return launch_encoder_layer_t<N, LAST, ST, EB, PAIR, CHUNK, EPI>(L, m_ptr, sms, st, probe);
    HF6D_ENC_CONFIGS(X)
#undef X
    return cudaErrorInvalidValue;
}

}  // namespace hf6d
