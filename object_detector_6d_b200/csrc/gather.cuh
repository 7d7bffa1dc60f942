// Stage SCAN + GATHER: valid patch centres (ordered compaction) and the fused
// bilinear RGB-D gather + local normalisation + uint8 quantisation.
//
// Replaces  patch_extractor_gpu::extract_patches_rgbd  (PatchGen/src/cuda/patch_extractor.cu:230-309, 318-433: host
// loop over the stride grid, 3-D texture upload, one 64-thread block per patch, 71 MB D2H of fp32 patches) and the
// single-threaded normalise/quantise loop of HFTest::test_image (HoughForest/src/HFTest.cpp:500-570).
//
// B200 design: the frame stays in its 5 B/pixel form (uint8 BGR + uint16 depth, 1.5 MB: L2 resident), the
// stride-grid scan is an ordered two-kernel compaction (so patch indices equal the reference's row-major push order),
// and one kernel produces the encoder's A operand directly: bf16 [P'][256] holding the quantised value q as an exact
// integer (1/255 is folded into the first layer's weights).  The fp32 patches never exist in memory.
//
// Arithmetic is the oracle's, operation for operation (IEEE _rn intrinsics, no FMA contraction), including the strictly
// sequential mean / variance accumulation order c -> row -> col of the reference and the double-precision variance terms
// its `pow(x - mean, 2) / N` compiles to (oracle choice C3), so the uint8 patches are bit-exact.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace hf6d {

__device__ __forceinline__ bool centre_valid(const uint16_t* __restrict__ depth, const FrameGeom& g, int w, int h) {
    const float d = (float)depth[(size_t)h * g.W + w];
    if (d == 0.0f) return false;
    const float dm = __fdiv_rn(d, 1000.0f);
    if (!(dm < g.dist_thr)) return false;
    const int a = adaptive_size(g, dm);
    const int x0 = w - a / 2, x1 = x0 + a - 1;
    const int y0 = h - a / 2, y1 = y0 + a - 1;
    return x0 >= 0 && y0 >= 0 && x1 < g.W && y1 < g.H;
}

// One warp per stride-grid row: number of valid centres in the row.
__global__ void scan_count_kernel(const uint16_t* __restrict__ depth, FrameGeom g, int* __restrict__ row_count) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= g.gh) return;
    const int h = row * g.stride;
    int n = 0;
    for (int c0 = 0; c0 < g.gw; c0 += 32) {
        const int c = c0 + lane;
        const bool ok = c < g.gw && centre_valid(depth, g, c * g.stride, h);
        n += __popc(__ballot_sync(0xffffffffu, ok));
    }
    if (lane == 0) row_count[row] = n;
}

// One warp per stride-grid row: exclusive offset = sum of the counts of earlier rows, then ballot compaction.
// counts[0] = P, counts[1] = P' = floor(P / batch) * batch (the reference drops the partial batch, HFTest.cpp:433).
__global__ void scan_compact_kernel(const uint16_t* __restrict__ depth, FrameGeom g, const int* __restrict__ row_count,
                                    int* __restrict__ locs, int* __restrict__ counts, int* __restrict__ row_off) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= g.gh) return;
    int before = 0, total = 0;
    for (int r = lane; r < g.gh; r += 32) {
        const int n = row_count[r];
        total += n;
        if (r < row) before += n;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        before += __shfl_xor_sync(0xffffffffu, before, o);
        total += __shfl_xor_sync(0xffffffffu, total, o);
    }
    if (row == 0 && lane == 0) {
        counts[0] = total;
        counts[1] = (total / g.batch) * g.batch;
    }
    if (lane == 0 && row_off) row_off[row] = before;  // index of the row's first patch (gather_tile_kernel)
    const int h = row * g.stride;
    int pos = before;
    for (int c0 = 0; c0 < g.gw; c0 += 32) {
        const int c = c0 + lane;
        const bool ok = c < g.gw && centre_valid(depth, g, c * g.stride, h);
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (ok) {
            const int i = pos + __popc(m & ((1u << lane) - 1u));
            if (i < g.cap) {
                locs[2 * i] = c * g.stride;
                locs[2 * i + 1] = h;
            }
        }
        pos += __popc(m);
    }
}

// Two quantised values 0..255 as the encoder's A operand: exact integers in bf16 (8-bit significand: up to 256) or fp16.
__device__ __forceinline__ uint32_t pack_q_pair(uint32_t q0, uint32_t q1, int fp16) {
    if (fp16) {
        const __half2 h = __floats2half2_rn((float)q0, (float)q1);
        return *reinterpret_cast<const uint32_t*>(&h);
    }
    const __nv_bfloat162 h = __floats2bfloat162_rn((float)q0, (float)q1);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// Software restatement of the texture fetch the reference uses (3-D float texture, unnormalised coordinates,
// linear filter, border address mode): fraction rounded to 8 bits, out-of-range texel = 0,
// S = ((w00*T00 + w10*T10) + w01*T01) + w11*T11 in fp32.
struct Bilerp {
    int i, j;
    float w00, w10, w01, w11;
};
__device__ __forceinline__ float frac8(float a) { return floorf(__fadd_rn(__fmul_rn(a, 256.0f), 0.5f)) * (1.0f / 256.0f); }
__device__ __forceinline__ Bilerp make_bilerp(float u, float v) {
    Bilerp b;
    const float fu = floorf(u), fv = floorf(v);
    b.i = (int)fu;
    b.j = (int)fv;
    const float a = frac8(__fsub_rn(u, fu)), c = frac8(__fsub_rn(v, fv));
    const float na = __fsub_rn(1.0f, a), nc = __fsub_rn(1.0f, c);
    b.w00 = __fmul_rn(na, nc);
    b.w10 = __fmul_rn(a, nc);
    b.w01 = __fmul_rn(na, c);
    b.w11 = __fmul_rn(a, c);
    return b;
}
__device__ __forceinline__ float blend(const Bilerp& b, float t00, float t10, float t01, float t11) {
    float s = __fmul_rn(b.w00, t00);
    s = __fadd_rn(s, __fmul_rn(b.w10, t10));
    s = __fadd_rn(s, __fmul_rn(b.w01, t01));
    s = __fadd_rn(s, __fmul_rn(b.w11, t11));
    return s;
}

// The reference's texture holds (B/255, G/255, R/255, depth mm) per texel (HFTest.cpp:370-379).  Here a texel is 8 bytes,
// {B | G<<8 | R<<16, depth}: one 8-byte load instead of three byte loads and a 16-bit load (the gather is bound by load
// instructions, not by bytes), converted to the reference's float values in registers.  2.4 MB at 640x480: L2 resident.
__global__ void pack_frame_kernel(const uint8_t* __restrict__ bgr, const uint16_t* __restrict__ depth, int n_px,
                                  uint2* __restrict__ tex) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_px) return;
    const uint8_t* p = bgr + (size_t)i * 3;
    tex[i] = make_uint2((unsigned)p[0] | ((unsigned)p[1] << 8) | ((unsigned)p[2] << 16), (unsigned)depth[i]);
}

constexpr int GATHER_PATCHES_PER_CTA = 16;
constexpr int GATHER_THREADS = GATHER_PATCHES_PER_CTA * 8;
constexpr int GATHER_ROW = 258;    // floats per patch row of the term buffer: 2 (mod 32), so the 32 chains below hit 32 banks
constexpr int GATHER_D_OFF = 193;  // depth terms start here: 1 (mod 32)

// The reference's strictly sequential float accumulation (HFTest.cpp:508-534: `mean += x / N` element by element, c ->
// row -> col) for all 16 patches of the CTA at once: lane 2p walks the 192 colour terms of patch p, lane 2p+1 its 64
// depth terms -- 32 independent chains, one warp, conflict-free shared-memory reads.
__device__ __forceinline__ void sequential_sums(const float (*term)[GATHER_ROW], float (*stat)[4], int first_stat) {
    if (threadIdx.x < 32) {
        const int pl = threadIdx.x >> 1, which = threadIdx.x & 1;
        const float* src = term[pl] + (which ? GATHER_D_OFF : 0);
        float m = 0.f;
#pragma unroll 16
        for (int j = 0; j < 64; ++j) m = __fadd_rn(m, src[j]);
        if (!which) {
#pragma unroll 16
            for (int j = 64; j < 192; ++j) m = __fadd_rn(m, src[j]);
        }
        stat[pl][first_stat + which] = m;
    }
}

// a / 3 correctly rounded without the IEEE division sequence (Markstein: y = RN(1/3), q = RN(a y), r = a - 3 q exactly,
// q' = RN(q + r y)); tools/check_const_division.c compares it with a / 3.0 on 10^9 squares of floats.
__device__ __forceinline__ double div3_rn(double a) {
    const double y = 0x1.5555555555555p-2;
    const double q = __dmul_rn(a, y);
    return __fma_rn(__fma_rn(-3.0, q, a), y, q);
}

// The reference's variance accumulation (HFTest.cpp:527-534: `std += pow(x - mean, 2) / N`, which its toolchain evaluates
// as pow(double, double): the float deviation squared in double, divided by (double)N, added to (double)std and narrowed
// back to the float accumulator, element by element).  dev holds the float deviations x - mean; lanes as in
// sequential_sums.  N = 192 = 3 * 64 and N = 64: the power of two is an exact scaling.
__device__ __forceinline__ void sequential_variances(const float (*dev)[GATHER_ROW], float (*stat)[4]) {
    if (threadIdx.x < 32) {
        const int pl = threadIdx.x >> 1, which = threadIdx.x & 1;
        const float* src = dev[pl] + (which ? GATHER_D_OFF : 0);
        float m = 0.f;
#pragma unroll 8
        for (int j = 0; j < 64; ++j) {
            const double d = (double)src[j], dd = __dmul_rn(d, d);
            const double t = __dmul_rn(which ? dd : div3_rn(dd), 1.0 / 64.0);
            m = __double2float_rn(__dadd_rn((double)m, t));
        }
        if (!which) {
#pragma unroll 8
            for (int j = 64; j < 192; ++j) {
                const double d = (double)src[j];
                m = __double2float_rn(__dadd_rn((double)m, __dmul_rn(div3_rn(__dmul_rn(d, d)), 1.0 / 64.0)));
            }
        }
        stat[pl][2 + which] = m;
    }
}

// 8 threads per patch (one per patch row), 16 patches per CTA.  ps == 8 only (the 256-input encoder).
// a_out : bf16 [cap][256], CHW order, value = q (exact integer 0..255)
// q_out : optional uint8 [cap][256] (debug capture / parity)
__global__ void __launch_bounds__(GATHER_THREADS)
gather_normalise_kernel(const uint2* __restrict__ tex, FrameGeom g, const int* __restrict__ locs,
                        const int* __restrict__ counts, PatchShard shard,
                        __nv_bfloat16* __restrict__ a_out, uint8_t* __restrict__ q_out, int a_fp16) {
    __shared__ float s_term[GATHER_PATCHES_PER_CTA][GATHER_ROW];
    __shared__ float s_stat[GATHER_PATCHES_PER_CTA][4];  // mean_rgb, mean_d, var_rgb, var_d

    int p_lo, Pp;  // this rank's patches [p_lo, Pp): cuts on multiples of 128, so a CTA's 16 patches are all in or all out
    patch_shard_range(counts[1], shard, p_lo, Pp);
    const int pl = threadIdx.x >> 3;
    const int ty = threadIdx.x & 7;
    const int p = blockIdx.x * GATHER_PATCHES_PER_CTA + pl;
    if (blockIdx.x * GATHER_PATCHES_PER_CTA >= Pp || (blockIdx.x + 1) * GATHER_PATCHES_PER_CTA <= p_lo) return;  // whole CTA out of range
    const bool live = p < Pp;

    float val[4][8];
    if (live) {
        const int2 ctr = *reinterpret_cast<const int2*>(locs + 2 * p);
        const int cx = ctr.x, cy = ctr.y;
        const float dc = div_const<1000, 1>((float)tex[(size_t)cy * g.W + cx].y);  // exact texel fetch, :253
        const int a = adaptive_size(g, dc);
        const int x0 = cx - a / 2, y0 = cy - a / 2;
        const float step = __fdiv_rn((float)a, (float)g.ps);
        float fill[4] = {0.f, 0.f, 0.f, 0.f};
        if (g.fill_random) {
            // counter-based stand-in for the clock64()-seeded cuRAND draw of patch_extractor.cu:236-244
            const unsigned long long z = mix64(g.fill_seed + 0x9E3779B97F4A7C15ULL * (unsigned long long)(p + 1));
            fill[2] = div_const<255, 1>((float)((z & 0xFFFF) % 255));
            fill[1] = div_const<255, 1>((float)(((z >> 16) & 0xFFFF) % 255));
            fill[0] = div_const<255, 1>((float)(((z >> 32) & 0xFFFF) % 255));
            fill[3] = div_const<255, 1>((float)(((z >> 48) & 0xFFFF) % 255));
        }
        const float v = __fadd_rn((float)y0, __fmul_rn((float)ty, step));
        // the row pair (j, j+1) and the vertical weights are shared by the 8 samples of this thread
        const float fv = floorf(v);
        const int j = (int)fv;
        const float c = frac8(__fsub_rn(v, fv)), nc = __fsub_rn(1.0f, c);
        const bool in_y0 = j >= 0 && j < g.H, in_y1 = j + 1 >= 0 && j + 1 < g.H;
#pragma unroll
        for (int tx = 0; tx < 8; ++tx) {
            const float u = __fadd_rn((float)x0, __fmul_rn((float)tx, step));
            const float fu = floorf(u);
            const int i = (int)fu;
            const float al = frac8(__fsub_rn(u, fu)), na = __fsub_rn(1.0f, al);
            Bilerp b;
            b.w00 = __fmul_rn(na, nc);
            b.w10 = __fmul_rn(al, nc);
            b.w01 = __fmul_rn(na, c);
            b.w11 = __fmul_rn(al, c);
            const bool in_x0 = i >= 0 && i < g.W, in_x1 = i + 1 >= 0 && i + 1 < g.W;
            const uint2* t0 = tex + ((size_t)j * g.W + i);
            const bool k00 = in_x0 && in_y0, k10 = in_x1 && in_y0, k01 = in_x0 && in_y1, k11 = in_x1 && in_y1;
            const uint2 zero = make_uint2(0u, 0u);  // border texel
            const uint2 q00 = k00 ? __ldg(t0) : zero, q10 = k10 ? __ldg(t0 + 1) : zero;
            const uint2 q01 = k01 ? __ldg(t0 + g.W) : zero, q11 = k11 ? __ldg(t0 + g.W + 1) : zero;
            const float d = div_const<1000, 1>(blend(b, (float)q00.y, (float)q10.y, (float)q01.y, (float)q11.y));
            if (d > 0.f) {
#pragma unroll
                for (int ch = 0; ch < 3; ++ch)
                    val[ch][tx] = blend(b, div_const<255, 1>((float)((q00.x >> (8 * ch)) & 0xFFu)),
                                        div_const<255, 1>((float)((q10.x >> (8 * ch)) & 0xFFu)),
                                        div_const<255, 1>((float)((q01.x >> (8 * ch)) & 0xFFu)),
                                        div_const<255, 1>((float)((q11.x >> (8 * ch)) & 0xFFu)));
                float td = __fadd_rn(__fdiv_rn(__fsub_rn(d, dc), g.range), 0.5f);
                if (td > 1.0f) td = 1.0f;
                if (td < 0.0f) td = 0.0f;
                val[3][tx] = td;
            } else {
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) val[ch][tx] = fill[ch];
            }
        }
    } else {
#pragma unroll
        for (int ch = 0; ch < 4; ++ch)
#pragma unroll
            for (int tx = 0; tx < 8; ++tx) val[ch][tx] = 0.f;
    }

    // ---- means: terms x/N in parallel, then the reference's sequential accumulation (HFTest.cpp:508-524)
#pragma unroll
    for (int ch = 0; ch < 4; ++ch)
#pragma unroll
        for (int tx = 0; tx < 8; ++tx)
            s_term[pl][(ch < 3 ? ch * 64 : GATHER_D_OFF) + ty * 8 + tx] =
                ch < 3 ? div_const<192, 1>(val[ch][tx]) : __fmul_rn(val[ch][tx], 1.0f / 64.0f);
    __syncthreads();
    sequential_sums(s_term, s_stat, 0);
    __syncthreads();
    const float mean_rgb = s_stat[pl][0], mean_d = s_stat[pl][1];
    // ---- "std" (variance, never sqrt'ed; terms in double) (HFTest.cpp:527-534)
#pragma unroll
    for (int ch = 0; ch < 4; ++ch)
#pragma unroll
        for (int tx = 0; tx < 8; ++tx)
            s_term[pl][(ch < 3 ? ch * 64 : GATHER_D_OFF) + ty * 8 + tx] = __fsub_rn(val[ch][tx], ch < 3 ? mean_rgb : mean_d);
    __syncthreads();
    sequential_variances(s_term, s_stat);
    __syncthreads();
    if (!live) return;
    const float lim_rgb = __fmul_rn(3.0f, s_stat[pl][2]), lim_d = __fmul_rn(3.0f, s_stat[pl][3]);
    // ---- clip to +-3 "std", scale to [0.1, 0.9], quantise (HFTest.cpp:538-565); NaN (lim == 0) -> 0
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
        const float m = ch < 3 ? mean_rgb : mean_d;
        const float lim = ch < 3 ? lim_rgb : lim_d;
        uint32_t qb[8];
#pragma unroll
        for (int tx = 0; tx < 8; ++tx) {
            float x = __fsub_rn(val[ch][tx], m);
            if (x > lim) x = lim;
            if (x < -lim) x = -lim;
            x = __fdiv_rn(x, lim);
            x = __fadd_rn(__fmul_rn(__fadd_rn(x, 1.0f), 0.4f), 0.1f);
            qb[tx] = (uint32_t)(f2i_x86(__fmul_rn(x, 255.0f)) & 0xFF);
        }
        const size_t o = (size_t)p * 256 + ch * 64 + ty * 8;
        uint32_t pk[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) pk[k] = pack_q_pair(qb[2 * k], qb[2 * k + 1], a_fp16);
        *reinterpret_cast<uint4*>(a_out + o) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        if (q_out) {
            const uint32_t lo = qb[0] | (qb[1] << 8) | (qb[2] << 16) | (qb[3] << 24);
            const uint32_t hi = qb[4] | (qb[5] << 8) | (qb[6] << 16) | (qb[7] << 24);
            *reinterpret_cast<uint2*>(q_out + o) = make_uint2(lo, hi);
        }
    }
}

// ------------------------------------------------------------------------------------------------ tiled gather
// gather_normalise_kernel above issues 256 scattered 8-byte loads per patch (ncu, round 1: 1.1e7 load sectors = 366 MB of L1
// traffic for a 2.4 MB texel array) and runs the sequential statistics of 16 patches on one warp while the CTA's other three
// wait.  gather_tile_kernel is the same arithmetic, reorganised around what neighbouring patches share:
//   * a CTA takes up to 32 CONSECUTIVE patches of one stride-grid row (grid = row segments x rows; scan_compact_kernel
//     publishes the index of every row's first patch): their footprints overlap almost completely, so the CTA loads the
//     bounding box of all of them into shared memory once -- coalesced rows of 8-byte texels, zero outside the image (the
//     texture unit's border mode) -- and every bilinear tap is a shared-memory read;
//   * the strictly sequential sums (HFTest.cpp:508-534) are one lane per chain: warp 0 runs the 32 colour chains, warp 1
//     the 32 depth chains (uniform trip counts, conflict-free rows), everything else of the patch stays in registers;
//   * a segment whose bounding box does not fit the tile (a row interrupted by holes, patches much larger than planned for)
//     takes its taps from global memory like the old kernel.
constexpr int GT_PATCHES = 32;
constexpr int GT_THREADS = GT_PATCHES * 8;
constexpr int GT_PITCH = 257;   // floats per patch row of the value buffer: 1 (mod 32), one bank per lane in the chains
constexpr int GT_VAL_BYTES = GT_PATCHES * GT_PITCH * 4;

template <bool STAGED>
__device__ __forceinline__ uint2 gt_texel(const uint2* __restrict__ tex, const uint2* s_tile, const FrameGeom& g, int x, int y,
                                          int tx0, int ty0, int pitch) {
    if (STAGED) return s_tile[(y - ty0) * pitch + (x - tx0)];
    return (x >= 0 && x < g.W && y >= 0 && y < g.H) ? __ldg(tex + (size_t)y * g.W + x) : make_uint2(0u, 0u);
}

// The 8 samples of patch row ty (same operations, same order as gather_normalise_kernel).
template <bool STAGED>
__device__ __forceinline__ void gt_sample_row(const uint2* __restrict__ tex, const uint2* s_tile, const FrameGeom& g, int x0, int y0,
                                              float step, float dc, int ty, const float fill[4], int tx0, int ty0, int pitch,
                                              float (&val)[4][8]) {
    const float v = __fadd_rn((float)y0, __fmul_rn((float)ty, step));
    const float fv = floorf(v);
    const int j = (int)fv;
    const float c = frac8(__fsub_rn(v, fv)), nc = __fsub_rn(1.0f, c);
#pragma unroll
    for (int tx = 0; tx < 8; ++tx) {
        const float u = __fadd_rn((float)x0, __fmul_rn((float)tx, step));
        const float fu = floorf(u);
        const int i = (int)fu;
        const float al = frac8(__fsub_rn(u, fu)), na = __fsub_rn(1.0f, al);
        Bilerp b;
        b.w00 = __fmul_rn(na, nc);
        b.w10 = __fmul_rn(al, nc);
        b.w01 = __fmul_rn(na, c);
        b.w11 = __fmul_rn(al, c);
        const uint2 q00 = gt_texel<STAGED>(tex, s_tile, g, i, j, tx0, ty0, pitch), q10 = gt_texel<STAGED>(tex, s_tile, g, i + 1, j, tx0, ty0, pitch);
        const uint2 q01 = gt_texel<STAGED>(tex, s_tile, g, i, j + 1, tx0, ty0, pitch), q11 = gt_texel<STAGED>(tex, s_tile, g, i + 1, j + 1, tx0, ty0, pitch);
        const float d = div_const<1000, 1>(blend(b, (float)q00.y, (float)q10.y, (float)q01.y, (float)q11.y));
        if (d > 0.f) {
#pragma unroll
            for (int ch = 0; ch < 3; ++ch)
                val[ch][tx] = blend(b, div_const<255, 1>((float)((q00.x >> (8 * ch)) & 0xFFu)),
                                    div_const<255, 1>((float)((q10.x >> (8 * ch)) & 0xFFu)),
                                    div_const<255, 1>((float)((q01.x >> (8 * ch)) & 0xFFu)),
                                    div_const<255, 1>((float)((q11.x >> (8 * ch)) & 0xFFu)));
            float td = __fadd_rn(__fdiv_rn(__fsub_rn(d, dc), g.range), 0.5f);
            if (td > 1.0f) td = 1.0f;
            if (td < 0.0f) td = 0.0f;
            val[3][tx] = td;
        } else {
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) val[ch][tx] = fill[ch];
        }
    }
}

// grid = (ceil(gw / GT_PATCHES), gh); dynamic shared memory = max(tile_cap_texels * 8, GT_VAL_BYTES).
__global__ void __launch_bounds__(GT_THREADS)
gather_tile_kernel(const uint2* __restrict__ tex, FrameGeom g, const int* __restrict__ locs, const int* __restrict__ counts,
                   const int* __restrict__ row_count, const int* __restrict__ row_off, int tile_cap_texels, PatchShard shard,
                   __nv_bfloat16* __restrict__ a_out, uint8_t* __restrict__ q_out, int a_fp16) {
    extern __shared__ __align__(16) uint8_t gt_smem[];
    uint2* s_tile = reinterpret_cast<uint2*>(gt_smem);
    float* s_val = reinterpret_cast<float*>(gt_smem);  // overlays the tile once every tap has been read
    __shared__ int s_box[4];                           // x min, y min, x max, y max of the segment's footprints
    __shared__ float s_stat[GT_PATCHES][4];            // mean_rgb, mean_d, var_rgb, var_d

    const int row = blockIdx.y, seg0 = blockIdx.x * GT_PATCHES;
    const int n_row = row_count[row];
    if (seg0 >= n_row) return;
    int p_lo, Pp;  // this rank's patches [p_lo, Pp)
    patch_shard_range(counts[1], shard, p_lo, Pp);
    const int p0 = row_off[row] + seg0;
    if (p0 >= Pp) return;
    const int n_here = min(min(GT_PATCHES, n_row - seg0), Pp - p0);
    if (p0 + n_here <= p_lo) return;
    const int pl = threadIdx.x >> 3, ty = threadIdx.x & 7;
    const int p = p0 + pl;
    const bool live = pl < n_here && p >= p_lo;
    if (threadIdx.x == 0) { s_box[0] = INT_MAX; s_box[1] = INT_MAX; s_box[2] = INT_MIN; s_box[3] = INT_MIN; }
    __syncthreads();

    int x0 = 0, y0 = 0;
    float step = 0.f, dc = 0.f;
    float fill[4] = {0.f, 0.f, 0.f, 0.f};
    if (live) {
        const int2 ctr = *reinterpret_cast<const int2*>(locs + 2 * p);
        dc = div_const<1000, 1>((float)tex[(size_t)ctr.y * g.W + ctr.x].y);  // exact texel fetch, patch_extractor.cu:253
        const int a = adaptive_size(g, dc);
        x0 = ctr.x - a / 2;
        y0 = ctr.y - a / 2;
        step = __fdiv_rn((float)a, (float)g.ps);
        if (g.fill_random) {
            const unsigned long long z = mix64(g.fill_seed + 0x9E3779B97F4A7C15ULL * (unsigned long long)(p + 1));
            fill[2] = div_const<255, 1>((float)((z & 0xFFFF) % 255));
            fill[1] = div_const<255, 1>((float)(((z >> 16) & 0xFFFF) % 255));
            fill[0] = div_const<255, 1>((float)(((z >> 32) & 0xFFFF) % 255));
            fill[3] = div_const<255, 1>((float)(((z >> 48) & 0xFFFF) % 255));
        }
        // texels this thread's row of samples touches: columns floor(u_0) .. floor(u_7) + 1, rows floor(v) .. floor(v) + 1
        const int xa = x0, xb = (int)floorf(__fadd_rn((float)x0, __fmul_rn(7.0f, step))) + 1;
        const int ya = (int)floorf(__fadd_rn((float)y0, __fmul_rn((float)ty, step))), yb = ya + 1;
        atomicMin(&s_box[0], xa);
        atomicMin(&s_box[1], ya);
        atomicMax(&s_box[2], xb);
        atomicMax(&s_box[3], yb);
    }
    __syncthreads();
    const int tx0 = s_box[0], ty0 = s_box[1];
    const int tw = s_box[2] - tx0 + 1, th = s_box[3] - ty0 + 1;
    const int pitch = tw | 1;  // odd: the rows of one patch spread over the banks
    const bool staged = (long long)pitch * th <= (long long)tile_cap_texels;
    if (staged) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int r = warp; r < th; r += GT_THREADS / 32) {
            const int y = ty0 + r;
            const bool row_in = y >= 0 && y < g.H;
            const uint2* src = tex + (size_t)(row_in ? y : 0) * g.W;
            for (int x = lane; x < tw; x += 32) {
                const int gx = tx0 + x;
                s_tile[r * pitch + x] = (row_in && gx >= 0 && gx < g.W) ? __ldg(src + gx) : make_uint2(0u, 0u);  // border texel = 0
            }
        }
    }
    __syncthreads();

    float val[4][8];
    if (live) {
        if (staged) gt_sample_row<true>(tex, s_tile, g, x0, y0, step, dc, ty, fill, tx0, ty0, pitch, val);
        else gt_sample_row<false>(tex, s_tile, g, x0, y0, step, dc, ty, fill, tx0, ty0, pitch, val);
    } else {
#pragma unroll
        for (int ch = 0; ch < 4; ++ch)
#pragma unroll
            for (int tx = 0; tx < 8; ++tx) val[ch][tx] = 0.f;
    }
    __syncthreads();  // every tap has been read: the tile's memory becomes the value buffer
#pragma unroll
    for (int ch = 0; ch < 4; ++ch)
#pragma unroll
        for (int tx = 0; tx < 8; ++tx) s_val[pl * GT_PITCH + ch * 64 + ty * 8 + tx] = val[ch][tx];
    __syncthreads();
    // ---- the reference's sequential sums, one lane per chain: warp 0 = colour (192 terms), warp 1 = depth (64 terms)
    if (threadIdx.x < 64) {
        const int cp = threadIdx.x & 31, which = threadIdx.x >> 5;
        const float* src = s_val + cp * GT_PITCH + which * 192;
        const int n_terms = which ? 64 : 192;
        float m = 0.f;  // means: `mean += x / N` in float (HFTest.cpp:508-524)
        if (which) {
#pragma unroll 16
            for (int j = 0; j < 64; ++j) m = __fadd_rn(m, __fmul_rn(src[j], 1.0f / 64.0f));
        } else {
#pragma unroll 16
            for (int j = 0; j < 192; ++j) m = __fadd_rn(m, div_const<192, 1>(src[j]));
        }
        float var = 0.f;  // "std": `std += pow(x - mean, 2) / N` in double, narrowed per element (oracle choice C3, :527-534)
#pragma unroll 8
        for (int j = 0; j < n_terms; ++j) {
            const double d = (double)__fsub_rn(src[j], m), dd = __dmul_rn(d, d);
            const double t = __dmul_rn(which ? dd : div3_rn(dd), 1.0 / 64.0);
            var = __double2float_rn(__dadd_rn((double)var, t));
        }
        s_stat[cp][which] = m;
        s_stat[cp][2 + which] = var;
    }
    __syncthreads();
    if (!live) return;
    const float mean_rgb = s_stat[pl][0], mean_d = s_stat[pl][1];
    const float lim_rgb = __fmul_rn(3.0f, s_stat[pl][2]), lim_d = __fmul_rn(3.0f, s_stat[pl][3]);
    // ---- clip to +-3 "std", scale to [0.1, 0.9], quantise (HFTest.cpp:538-565); NaN (lim == 0) -> 0
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
        const float m = ch < 3 ? mean_rgb : mean_d;
        const float lim = ch < 3 ? lim_rgb : lim_d;
        uint32_t qb[8];
#pragma unroll
        for (int tx = 0; tx < 8; ++tx) {
            float x = __fsub_rn(val[ch][tx], m);
            if (x > lim) x = lim;
            if (x < -lim) x = -lim;
            x = __fdiv_rn(x, lim);
            x = __fadd_rn(__fmul_rn(__fadd_rn(x, 1.0f), 0.4f), 0.1f);
            qb[tx] = (uint32_t)(f2i_x86(__fmul_rn(x, 255.0f)) & 0xFF);
        }
        const size_t o = (size_t)p * 256 + ch * 64 + ty * 8;
        uint32_t pk[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) pk[k] = pack_q_pair(qb[2 * k], qb[2 * k + 1], a_fp16);
        *reinterpret_cast<uint4*>(a_out + o) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        if (q_out) {
            const uint32_t lo = qb[0] | (qb[1] << 8) | (qb[2] << 16) | (qb[3] << 24);
            const uint32_t hi = qb[4] | (qb[5] << 8) | (qb[6] << 16) | (qb[7] << 24);
            *reinterpret_cast<uint2*>(q_out + o) = make_uint2(lo, hi);
        }
    }
}

// ================================================================================================ A2c: normals variant
// The patch mode the reference keeps next to the RGB-D one (HFTest.cpp:322-363 and :443-470, commented out there):
// surface normals from depth by central differences (surface_normals.cu:11-73), a 7-channel texture B,G,R,D,nx,ny,nz
// (patch_extractor.cu:12-111), 6-channel patches B,G,R + re-normalised interpolated normal, and plain uint8
// quantisation (no local normalisation) into a 384-input encoder.
//
// normals: float4 [H][W] = (nx, ny, nz, 0), one 16-byte texel.  Products and sums without FMA contraction (oracle C12).
__global__ void normals_kernel(const uint16_t* __restrict__ depth, int W, int H, float focal, float4* __restrict__ normals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W * H) return;
    const int x = i % W, y = i / W;
    float4 n = make_float4(0.f, 0.f, 0.f, 0.f);
    if (x > 0 && x < W - 1 && y > 0 && y < H - 1) {
        const float z = __fdiv_rn((float)depth[i], 1000.0f);
        const float z_left = __fdiv_rn((float)depth[i - 1], 1000.0f), z_right = __fdiv_rn((float)depth[i + 1], 1000.0f);
        const float z_up = __fdiv_rn((float)depth[i - W], 1000.0f), z_down = __fdiv_rn((float)depth[i + W], 1000.0f);
        if (z != 0.f && z_left != 0.f && z_right != 0.f && z_up != 0.f && z_down != 0.f) {
            const float hw = __fdiv_rn((float)W, 2.0f), hh = __fdiv_rn((float)H, 2.0f);
            const float xm = __fsub_rn(__fsub_rn((float)x, 1.0f), hw), xp = __fsub_rn(__fadd_rn((float)x, 1.0f), hw);
            const float x0 = __fsub_rn((float)x, hw);
            const float ym = __fsub_rn(__fsub_rn((float)y, 1.0f), hh), yp = __fsub_rn(__fadd_rn((float)y, 1.0f), hh);
            const float y0 = __fsub_rn((float)y, hh);
            const float x_left = __fdiv_rn(__fmul_rn(xm, z_left), focal), x_right = __fdiv_rn(__fmul_rn(xp, z_right), focal);
            const float x_up = __fdiv_rn(__fmul_rn(x0, z_up), focal), x_down = __fdiv_rn(__fmul_rn(x0, z_down), focal);
            const float y_left = __fdiv_rn(__fmul_rn(y0, z_left), focal), y_right = __fdiv_rn(__fmul_rn(y0, z_right), focal);
            const float y_up = __fdiv_rn(__fmul_rn(ym, z_up), focal), y_down = __fdiv_rn(__fmul_rn(yp, z_down), focal);
            const float ax = __fmul_rn(__fsub_rn(x_left, x_right), 0.5f), ay = __fmul_rn(__fsub_rn(y_left, y_right), 0.5f);
            const float az = __fmul_rn(__fsub_rn(z_left, z_right), 0.5f);
            const float bx = __fmul_rn(__fsub_rn(x_down, x_up), 0.5f), by = __fmul_rn(__fsub_rn(y_down, y_up), 0.5f);
            const float bz = __fmul_rn(__fsub_rn(z_down, z_up), 0.5f);
            const float nx = -__fsub_rn(__fmul_rn(ay, bz), __fmul_rn(az, by));
            const float ny = -__fsub_rn(__fmul_rn(az, bx), __fmul_rn(ax, bz));
            const float nz = -__fsub_rn(__fmul_rn(ax, by), __fmul_rn(ay, bx));
            const float mag = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(nx, nx), __fmul_rn(ny, ny)), __fmul_rn(nz, nz)));
            n = make_float4(__fdiv_rn(nx, mag), __fdiv_rn(ny, mag), __fdiv_rn(nz, mag), 0.f);
        }
    }
    normals[i] = n;
}

// x86 cvttsd2si semantics for the double expression of HFTest.cpp:466
__device__ __forceinline__ int d2i_x86_dev(double y) { return fabs(y) < 2147483648.0 ? (int)y : INT_MIN; }

// 8 threads per patch (one per patch row), 16 patches per CTA, like gather_normalise_kernel; no statistics pass.
// a_out : bf16 [cap][384], CHW order (B,G,R,nx,ny,nz planes), value = q (exact integer 0..255)
__global__ void __launch_bounds__(GATHER_THREADS)
gather_normals_kernel(const uint2* __restrict__ tex, const float4* __restrict__ normals, FrameGeom g,
                      const int* __restrict__ locs, const int* __restrict__ counts, __nv_bfloat16* __restrict__ a_out,
                      uint8_t* __restrict__ q_out) {
    const int Pp = counts[1];
    const int ty = threadIdx.x & 7;
    const int p = blockIdx.x * GATHER_PATCHES_PER_CTA + (threadIdx.x >> 3);
    if (p >= Pp) return;
    const int2 ctr = *reinterpret_cast<const int2*>(locs + 2 * p);
    const float dc = div_const<1000, 1>((float)tex[(size_t)ctr.y * g.W + ctr.x].y);
    const int a = adaptive_size(g, dc);
    const int x0 = ctr.x - a / 2, y0 = ctr.y - a / 2;
    const float step = __fdiv_rn((float)a, (float)g.ps);
    float fill[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (g.fill_random) {
        unsigned long long z = mix64(g.fill_seed + 0x9E3779B97F4A7C15ULL * (unsigned long long)(p + 1));
        fill[2] = div_const<255, 1>((float)((z & 0xFFFF) % 255));
        fill[1] = div_const<255, 1>((float)(((z >> 16) & 0xFFFF) % 255));
        fill[0] = div_const<255, 1>((float)(((z >> 32) & 0xFFFF) % 255));
        fill[5] = 1.0f;
        for (int attempt = 0; attempt < 4; ++attempt) {
            const unsigned long long z2 = mix64(z + 0x9E3779B97F4A7C15ULL), z3 = mix64(z2 + 0x9E3779B97F4A7C15ULL);
            z = z3;
            const float xr = __fsub_rn((float)((unsigned)z2 % 100000u), 50000.0f);
            const float yr = __fsub_rn((float)((unsigned)(z2 >> 32) % 100000u), 50000.0f);
            const float zr = (float)((unsigned)z3 % 50000u);
            const float norm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(xr, xr), __fmul_rn(yr, yr)), __fmul_rn(zr, zr)));
            if (norm != 0.f) { fill[3] = __fdiv_rn(xr, norm); fill[4] = __fdiv_rn(yr, norm); fill[5] = __fdiv_rn(zr, norm); break; }
        }
    }
    const float v = __fadd_rn((float)y0, __fmul_rn((float)ty, step));
    const float fv = floorf(v);
    const int j = (int)fv;
    const float c = frac8(__fsub_rn(v, fv)), nc = __fsub_rn(1.0f, c);
    const bool in_y0 = j >= 0 && j < g.H, in_y1 = j + 1 >= 0 && j + 1 < g.H;
    uint32_t qb[6][8];
#pragma unroll
    for (int tx = 0; tx < 8; ++tx) {
        const float u = __fadd_rn((float)x0, __fmul_rn((float)tx, step));
        const float fu = floorf(u);
        const int i = (int)fu;
        const float al = frac8(__fsub_rn(u, fu)), na = __fsub_rn(1.0f, al);
        Bilerp b;
        b.w00 = __fmul_rn(na, nc);
        b.w10 = __fmul_rn(al, nc);
        b.w01 = __fmul_rn(na, c);
        b.w11 = __fmul_rn(al, c);
        const bool in_x0 = i >= 0 && i < g.W, in_x1 = i + 1 >= 0 && i + 1 < g.W;
        const size_t o00 = (size_t)j * g.W + i;
        const bool k00 = in_x0 && in_y0, k10 = in_x1 && in_y0, k01 = in_x0 && in_y1, k11 = in_x1 && in_y1;
        const uint2 zero = make_uint2(0u, 0u);
        const uint2 q00 = k00 ? __ldg(tex + o00) : zero, q10 = k10 ? __ldg(tex + o00 + 1) : zero;
        const uint2 q01 = k01 ? __ldg(tex + o00 + g.W) : zero, q11 = k11 ? __ldg(tex + o00 + g.W + 1) : zero;
        float val[6];
        bool in_object = false;
        const float d = div_const<1000, 1>(blend(b, (float)q00.y, (float)q10.y, (float)q01.y, (float)q11.y));
        if (d > 0.f) {
            const float4 zf = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 n00 = k00 ? __ldg(normals + o00) : zf, n10 = k10 ? __ldg(normals + o00 + 1) : zf;
            const float4 n01 = k01 ? __ldg(normals + o00 + g.W) : zf, n11 = k11 ? __ldg(normals + o00 + g.W + 1) : zf;
            const float x = blend(b, n00.x, n10.x, n01.x, n11.x), y = blend(b, n00.y, n10.y, n01.y, n11.y);
            const float z = blend(b, n00.z, n10.z, n01.z, n11.z);
            const float norm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
            if (norm > 0.f) {
#pragma unroll
                for (int ch = 0; ch < 3; ++ch)
                    val[ch] = blend(b, div_const<255, 1>((float)((q00.x >> (8 * ch)) & 0xFFu)),
                                    div_const<255, 1>((float)((q10.x >> (8 * ch)) & 0xFFu)),
                                    div_const<255, 1>((float)((q01.x >> (8 * ch)) & 0xFFu)),
                                    div_const<255, 1>((float)((q11.x >> (8 * ch)) & 0xFFu)));
                val[3] = __fdiv_rn(x, norm);
                val[4] = __fdiv_rn(y, norm);
                val[5] = __fdiv_rn(z, norm);
                in_object = true;
            }
        }
        if (!in_object) {
#pragma unroll
            for (int ch = 0; ch < 6; ++ch) val[ch] = fill[ch];
        }
        // HFTest.cpp:463-466: colour (uchar)(v * 255.0f); normals (uchar)((v / 2.0 + 0.5f) * 255.0f), a double expression
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) qb[ch][tx] = (uint32_t)(f2i_x86(__fmul_rn(val[ch], 255.0f)) & 0xFF);
#pragma unroll
        for (int ch = 3; ch < 6; ++ch)
            qb[ch][tx] = (uint32_t)(d2i_x86_dev(__dmul_rn(__dadd_rn(__dmul_rn((double)val[ch], 0.5), 0.5), 255.0)) & 0xFF);
    }
#pragma unroll
    for (int ch = 0; ch < 6; ++ch) {
        const size_t o = (size_t)p * 384 + ch * 64 + ty * 8;
        uint32_t pk[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            __nv_bfloat162 h2 = __floats2bfloat162_rn((float)qb[ch][2 * k], (float)qb[ch][2 * k + 1]);
            pk[k] = *reinterpret_cast<uint32_t*>(&h2);
        }
        *reinterpret_cast<uint4*>(a_out + o) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        if (q_out) {
            const uint32_t lo = qb[ch][0] | (qb[ch][1] << 8) | (qb[ch][2] << 16) | (qb[ch][3] << 24);
            const uint32_t hi = qb[ch][4] | (qb[ch][5] << 8) | (qb[ch][6] << 16) | (qb[ch][7] << 24);
            *reinterpret_cast<uint2*>(q_out + o) = make_uint2(lo, hi);
        }
    }
}

}  // namespace hf6d
