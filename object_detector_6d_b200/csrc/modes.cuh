// Stages CENTRES and POSE: mode seeking on the vote accumulators.
//
// The reference's "mode seeking" is a normalised box blur followed by a sliding-window non-max suppression, applied
// hierarchically: centre map -> z histogram -> yaw/pitch map -> roll histogram
// (HoughForest/src/HFTest.cpp:702-707, 742-925; NMS HFTest.cpp:219-268).  These kernels are that neighbour reduction,
// tiled:
//  * box blur   = exact 64-bit integer window sums (row pass on a per-row prefix sum in shared memory, column pass as a
//                 running sum), scaled once in double:  (float)((double)S / 65536 * 1/(kx*ky))  -- what cv::blur does
//                 on CV_32F (double accumulation, single scale) without its summation-order dependence.
//  * NMS        = tiled neighbour reduction: key-maxima of 8x8 blocks, candidates pruned against the blocks inside their
//                 window, exact warp-cooperative verification of the survivors.  A window emits only if no element
//                 beats its centre under the order (value desc, row asc, col asc), which is the reference's
//                 monotonic-deque tie-breaking.  The reference's loop-bound quirk (the vertical pass stops at rows-wy,
//                 so the bottom wy-1 window rows are never produced) is kept.
//  * top-N      = block-wide selection on a sort key (score desc, then emission order x-major).
#pragma once
#include "common.cuh"

namespace hf6d {

__host__ __device__ __forceinline__ int reflect101(int i, int n) {  // cv::BORDER_REFLECT_101
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}

// A rectangle [r0, r0+nr) x [c0, c0+nc) of a virtual (rows x cols) map, stored densely (nr x nc) per map.  Cells of the
// virtual map outside the rectangle are exactly zero (no vote can land there), so reads outside return 0.
struct MapRect {
    int r0, c0, nr, nc;
};
struct MapDims {
    int rows, cols;
};

// ------------------------------------------------------------------------------------------------ accumulator reset
// One launch clears the pose accumulators of every class that is sought: segment g of class k starts at
// base[g] + k * HF6D_MAX_CENTRES * slot_bytes[g] and spans n_slots[k] * slot_bytes[g] bytes (slot_bytes multiples of 4,
// bases 256-byte aligned, so class starts are 16-byte aligned).  grid = (x, K, segments).
struct ClearPlan {
    void* base[4];
    unsigned long long slot_bytes[4];
    int n_slots[HF6D_MAX_CLASSES];
};
__global__ void __launch_bounds__(256)
clear_accumulators_kernel(const __grid_constant__ ClearPlan plan) {
    const int k = blockIdx.y, gseg = blockIdx.z;
    const unsigned long long bytes = (unsigned long long)plan.n_slots[k] * plan.slot_bytes[gseg];
    if (bytes == 0 || plan.base[gseg] == nullptr) return;
    uint8_t* p = static_cast<uint8_t*>(plan.base[gseg]) + (unsigned long long)k * HF6D_MAX_CENTRES * plan.slot_bytes[gseg];
    const unsigned long long n16 = bytes / 16;
    uint4* p16 = reinterpret_cast<uint4*>(p);
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16;
         i += (unsigned long long)gridDim.x * blockDim.x)
        p16[i] = make_uint4(0u, 0u, 0u, 0u);
    if (blockIdx.x == 0) {
        unsigned* tail = reinterpret_cast<unsigned*>(p + n16 * 16);
        if (threadIdx.x < (bytes - n16 * 16) / 4) tail[threadIdx.x] = 0u;
    }
}

// ------------------------------------------------------------------------------------------------ box blur
constexpr int BLUR_WARPS = 4;
constexpr int ROW_RUN = 8;     // consecutive columns a lane sums serially in the row pass
// The row pass keeps its prefix sums in shared memory with one slot skipped after every 8: lanes that write runs of 8
// consecutive 8-byte entries would otherwise all hit the same four banks (stride 64 B); with the skew the stride is 72 B.
__host__ __device__ __forceinline__ int pre_slot(int i) { return i + (i >> 3); }
__host__ __device__ __forceinline__ size_t box_rows_smem_bytes(int nc) { return (size_t)BLUR_WARPS * (pre_slot(nc + 1) + 1) * 8; }
constexpr int COL_BATCH = 8;   // rows loaded together while the column pass builds its first window sum

// Accumulators of the same layout that live in OTHER GPUs' memory (tree-sharded mode, csrc/hf6d_api.cu "peer exchange"):
// the row pass adds them to its own while it loads them, so summing the ranks' vote maps over NVLink costs no kernel, no
// staging buffer and no second pass -- only the maps of the classes this rank seeks ever cross the link.
struct PeerMaps {
    const unsigned long long* base[HF6D_MAX_PEERS - 1];
    int n;
};

// tmp[m][r][c] = sum_k acc[m][r][reflect(c - kx/2 + k)]   for r in in.rows, c in out.cols  (tmp is in.nr x out.nc)
__global__ void __launch_bounds__(BLUR_WARPS * 32)
box_rows_kernel(const unsigned long long* __restrict__ acc, unsigned long long* __restrict__ tmp, MapDims md, MapRect in,
                MapRect out, int kx, const uint8_t* __restrict__ map_active, const __grid_constant__ PeerMaps peers) {
    extern __shared__ unsigned long long s_pre[];  // [BLUR_WARPS][pre_slot(in.nc + 1) + 1]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m = blockIdx.y;
    if (map_active && !map_active[m]) return;
    const int r = blockIdx.x * BLUR_WARPS + warp;
    if (r >= in.nr) return;
    unsigned long long* pre = s_pre + (size_t)warp * (pre_slot(in.nc + 1) + 1);
    const unsigned long long* src = acc + ((size_t)m * in.nr + r) * in.nc;
    // Inclusive prefix sums of the row into pre[1..nc].  Loads are coalesced (lane l reads columns c0 + 32 j + l; the peers'
    // maps cross NVLink, where a lane-private run of 8-byte loads would fetch every 32-byte sector eight times), the raw
    // values pass through the prefix array so that each lane can then take ROW_RUN CONSECUTIVE columns, sum them serially in
    // registers, and only the 32 lane totals go through the shuffle scan -- one scan per 256 columns instead of one per 32.
    unsigned long long carry = 0;
    if (lane == 0) pre[pre_slot(0)] = 0;
    for (int c0 = 0; c0 < in.nc; c0 += 32 * ROW_RUN) {
        const int cb = c0 + lane * ROW_RUN;
        unsigned long long vv[ROW_RUN];
#pragma unroll
        for (int j = 0; j < ROW_RUN; ++j) {
            const int c = c0 + j * 32 + lane;
            vv[j] = c < in.nc ? src[c] : 0ull;
        }
        for (int q = 0; q < peers.n; ++q) {  // the other ranks' partial sums of the same cells, read in place
            const unsigned long long* psrc = peers.base[q] + ((size_t)m * in.nr + r) * in.nc;
#pragma unroll
            for (int j = 0; j < ROW_RUN; ++j) {
                const int c = c0 + j * 32 + lane;
                if (c < in.nc) vv[j] += psrc[c];
            }
        }
#pragma unroll
        for (int j = 0; j < ROW_RUN; ++j) {
            const int c = c0 + j * 32 + lane;
            if (c < in.nc) pre[pre_slot(c + 1)] = vv[j];
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < ROW_RUN; ++j) vv[j] = cb + j < in.nc ? pre[pre_slot(cb + j + 1)] : 0ull;
        __syncwarp();  // every lane has its run before any lane overwrites slots with prefix sums
#pragma unroll
        for (int j = 1; j < ROW_RUN; ++j) vv[j] += vv[j - 1];
        unsigned long long tot = vv[ROW_RUN - 1];  // inclusive scan of the lane totals
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long nb = __shfl_up_sync(0xffffffffu, tot, o);
            if (lane >= o) tot += nb;
        }
        const unsigned long long base = carry + tot - vv[ROW_RUN - 1];  // everything before this lane's run
#pragma unroll
        for (int j = 0; j < ROW_RUN; ++j)
            if (cb + j < in.nc) pre[pre_slot(cb + j + 1)] = base + vv[j];
        carry += __shfl_sync(0xffffffffu, tot, 31);
    }
    __syncwarp();
    auto seg = [&](int a, int b) -> unsigned long long {  // sum over global columns [a, b], clipped to the input rectangle
        a = max(a, in.c0);
        b = min(b, in.c0 + in.nc - 1);
        return b >= a ? pre[pre_slot(b - in.c0 + 1)] - pre[pre_slot(a - in.c0)] : 0ull;
    };
    unsigned long long* dst = tmp + ((size_t)m * in.nr + r) * out.nc;
    const int n = md.cols;
    for (int c = lane; c < out.nc; c += 32) {
        const int a = out.c0 + c - kx / 2, b = a + kx - 1;
        unsigned long long s;
        if (n > 1 && a > -n && b < 2 * n - 1) {
            s = seg(max(a, 0), min(b, n - 1));
            if (a < 0) s += seg(1, -a);                        // -1..a  reflect to  1..-a
            if (b > n - 1) s += seg(2 * n - 2 - b, n - 2);     // n..b   reflect to  n-2..2n-2-b
        } else {  // kernel wider than the map: plain loop
            s = 0;
            for (int k = a; k <= b; ++k) s += seg(reflect101(k, n), reflect101(k, n));
        }
        dst[c] = s;
    }
}

// out[m][r][c] = (float)( (double)(sum_k tmp[m][reflect(r - ky/2 + k)][c]) / 65536 * scale )  for (r, c) in out
// Also emits bmax[m][r/8][c/8], the key-maximum (value, ~row, ~col) of every aligned 8x8 block of the output -- the
// first step of the NMS below -- while the values are still in registers.
constexpr int BLUR_COL_CHUNK = 32;
constexpr int NMS_BLOCK = 8;
static_assert(BLUR_COL_CHUNK % NMS_BLOCK == 0, "a column chunk must hold whole NMS blocks");

__device__ __forceinline__ unsigned long long nms_key(float v, int gy, int gx) {
    return ((unsigned long long)__float_as_uint(v) << 32) | ((unsigned long long)(0xFFFFu - (unsigned)gy) << 16) |
           (unsigned long long)(0xFFFFu - (unsigned)gx);
}

__global__ void __launch_bounds__(128)
box_cols_kernel(const unsigned long long* __restrict__ tmp, float* __restrict__ dst_all, MapDims md, MapRect in,
                MapRect out, int ky, double scale, const uint8_t* __restrict__ map_active,
                unsigned long long* __restrict__ bmax) {
    const int m = blockIdx.z;
    if (map_active && !map_active[m]) return;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int rbeg = blockIdx.y * BLUR_COL_CHUNK;
    if (rbeg >= out.nr) return;
    const bool live = c < out.nc;
    const int cc = live ? c : 0;
    const unsigned long long* src = tmp + (size_t)m * in.nr * out.nc + cc;
    float* dst = dst_all + (size_t)m * out.nr * out.nc + cc;
    const int n = md.rows;
    auto at = [&](int gr) -> unsigned long long {  // global row, reflected; rows outside the input rectangle are zero
        const int rr = reflect101(gr, n) - in.r0;
        return (rr >= 0 && rr < in.nr) ? src[(size_t)rr * out.nc] : 0ull;
    };
    const int rend = min(rbeg + BLUR_COL_CHUNK, out.nr);
    const int bx = (out.nc + NMS_BLOCK - 1) / NMS_BLOCK, by = (out.nr + NMS_BLOCK - 1) / NMS_BLOCK;
    unsigned long long s = 0;
    {
        const int a = out.r0 + rbeg - ky / 2;
        const bool interior = a >= 0 && a + ky - 1 < n && a - in.r0 >= 0 && a - in.r0 + ky - 1 < in.nr;
        const unsigned long long* p0 = src + (size_t)(interior ? a - in.r0 : 0) * out.nc;
        for (int k0 = 0; k0 < ky; k0 += COL_BATCH) {  // batches of independent loads (a plain loop waits for every one)
            unsigned long long v[COL_BATCH];
#pragma unroll
            for (int i = 0; i < COL_BATCH; ++i)
                v[i] = k0 + i < ky ? (interior ? p0[(size_t)(k0 + i) * out.nc] : at(a + k0 + i)) : 0ull;
#pragma unroll
            for (int i = 0; i < COL_BATCH; ++i) s += v[i];
        }
    }
    const double scale16 = scale * (1.0 / 65536.0);  // a power of two: (s / 65536) * scale == s * (scale / 65536) bit for bit
    for (int r0 = rbeg; r0 < rend; r0 += NMS_BLOCK) {  // one block row at a time: its 16 loads are issued together
        unsigned long long add[NMS_BLOCK], sub[NMS_BLOCK];
        const int a0 = out.r0 + r0 - ky / 2;                          // global row leaving the window at step 0
        const int lo = a0 - in.r0, hi = lo + NMS_BLOCK - 1 + ky;      // rectangle-local rows this block row touches
        if (a0 >= 0 && a0 + NMS_BLOCK - 1 + ky < n && lo >= 0 && hi < in.nr) {
            // interior (almost every block row): no reflection, no clipping -- two pointers and a constant stride
            const unsigned long long* ps = src + (size_t)lo * out.nc;
            const unsigned long long* pa = ps + (size_t)ky * out.nc;
#pragma unroll
            for (int i = 0; i < NMS_BLOCK; ++i) {
                add[i] = pa[(size_t)i * out.nc];
                sub[i] = ps[(size_t)i * out.nc];
            }
        } else {
#pragma unroll
            for (int i = 0; i < NMS_BLOCK; ++i) {
                add[i] = at(a0 + i + ky);
                sub[i] = at(a0 + i);
            }
        }
        float best_v = 0.f;  // maximum of this column inside the block row (values are >= 0), topmost on ties
        int best_r = r0;
#pragma unroll
        for (int i = 0; i < NMS_BLOCK; ++i) {
            const int r = r0 + i;
            if (r < rend) {
                const float v = (float)((double)s * scale16);
                if (live) dst[(size_t)r * out.nc] = v;
                if (v > best_v) { best_v = v; best_r = r; }
                s += add[i];
                s -= sub[i];
            }
        }
        unsigned long long b = live ? nms_key(best_v, out.r0 + best_r, out.c0 + c) : 0ull;
#pragma unroll
        for (int o = 1; o < NMS_BLOCK; o <<= 1) b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
        if (bmax && live && (threadIdx.x & (NMS_BLOCK - 1)) == 0) bmax[((size_t)m * by + r0 / NMS_BLOCK) * bx + c / NMS_BLOCK] = b;
    }
}

// ------------------------------------------------------------------------------------------------ NMS
// The reference's sliding-window NMS (HFTest.cpp:219-268) emits a window iff the window maximum is non-zero and sits
// at the window centre, where its two monotonic deques resolve ties to the topmost row, then the leftmost column.
// Equivalently: candidate (ccx, ccy) = (left + wx/2, top + wy/2) is emitted iff its value v != 0 and NO element of the
// window  [left, left+wx) x [top, top+wy)  "beats" it under the total order  key = (value, ~row, ~col).
// Tiled neighbour reduction in two kernels:
//   1. the key-maximum of every aligned 8x8 block of the map (written by box_cols_kernel as it produces the map);
//   2. nms_select_kernel  : one lane per block.  A window wider than 15 contains its centre's own block, so only the
//      block maximum can be a window maximum (1 candidate in 64 survives); it is then compared with the maxima of the
//      blocks that overlap its window: a larger maximum that itself lies in the window beats it, a block whose
//      maximum is not larger cannot hold a beating element, and the (rare) remaining "suspect" blocks -- larger
//      maximum outside the window -- have their in-window elements checked by the warp, 64 elements at a time.
// The pruning steps only ever reject a candidate because a concrete window element beats it, so the result is exactly
// the reference's.  Windows narrower than 16 skip the pruning (every element of the block is verified directly).
// The reference's loop bounds (lefts 0..cols-wx, tops 0..rows-2*wy+1; the bottom wy-1 window rows are never produced)
// are applied by the caller through the origin ranges.
constexpr int NMS_LIST_CAP = 4096;

struct BlockGrid {
    int by, bx;  // blocks per map: ceil(nr / 8), ceil(nc / 8); block (i, j) covers rectangle rows [8i, 8i+8), cols [8j, 8j+8)
};
__host__ __device__ __forceinline__ BlockGrid make_block_grid(MapRect R) {
    return BlockGrid{(R.nr + NMS_BLOCK - 1) / NMS_BLOCK, (R.nc + NMS_BLOCK - 1) / NMS_BLOCK};
}

// Window origins (left, top) in global coordinates: left in [left0, left0+n_left), top in [top0, top0+n_top).
// Emits sort keys (score | ~x | ~y) into list[m].
constexpr int NMS_SELECT_THREADS = 64;
constexpr int NMS_ROW_BATCH = 6;      // block maxima of one block row fetched together (a 40-wide window spans <= 6 blocks)
constexpr int NMS_VERIFY_LOADS = 16;  // window elements per lane in flight during the exact verification
__global__ void __launch_bounds__(NMS_SELECT_THREADS)
nms_select_kernel(const float* __restrict__ in, const unsigned long long* __restrict__ bmax, MapRect R, int wx, int wy,
                  int left0, int n_left, int top0, int n_top, unsigned long long* __restrict__ list,
                  int* __restrict__ list_n, const uint8_t* __restrict__ map_active) {
    const int m = blockIdx.z;
    if (map_active && !map_active[m]) return;
    const BlockGrid bg = make_block_grid(R);
    const int lane = threadIdx.x & 31;
    const float* src = in + (size_t)m * R.nr * R.nc;
    const unsigned long long* bm = bmax + (size_t)m * bg.by * bg.bx;
    const int lo_x = -(wx / 2), hi_x = wx - 1 - wx / 2, lo_y = -(wy / 2), hi_y = wy - 1 - wy / 2;
    const bool prune = wx >= 2 * NMS_BLOCK && wy >= 2 * NMS_BLOCK;  // the window then contains the centre's own block
    auto emit = [&](int ex, int ey, float ev) {
        const int idx = atomicAdd(list_n + m, 1);
        if (idx < NMS_LIST_CAP)
            list[(size_t)m * NMS_LIST_CAP + idx] = ((unsigned long long)__float_as_uint(ev) << 32) |
                                                   ((unsigned long long)(0xFFFFu - (unsigned)ex) << 16) |
                                                   (unsigned long long)(0xFFFFu - (unsigned)ey);
    };
    const int b = blockIdx.x * blockDim.x + threadIdx.x;           // block id within the map (warp-uniform bound below)
    const int nb = bg.by * bg.bx;
    const int bi = b / bg.bx, bj = b % bg.bx;
    // candidates of this lane: the block maximum (pruned mode) or every element of the block
    for (int e = 0; e < (prune ? 1 : NMS_BLOCK * NMS_BLOCK); ++e) {
        int cx = 0, cy = 0, i0 = 0, j0 = 0, nj = 1;
        float v = 0.f;
        bool alive = false, scan_all = false;
        unsigned long long suspects = 0ull;
        if (b < nb) {
            if (prune) {
                const unsigned long long k = bm[b];
                v = __uint_as_float((unsigned)(k >> 32));
                cy = 0xFFFF - (int)((k >> 16) & 0xFFFFu);
                cx = 0xFFFF - (int)(k & 0xFFFFu);
            } else {
                const int row = bi * NMS_BLOCK + e / NMS_BLOCK, col = bj * NMS_BLOCK + e % NMS_BLOCK;
                if (row < R.nr && col < R.nc) {
                    v = __ldg(src + (size_t)row * R.nc + col);
                    cy = R.r0 + row;
                    cx = R.c0 + col;
                }
            }
            const int left = cx + lo_x, top = cy + lo_y;
            alive = v != 0.f && left >= left0 && left < left0 + n_left && top >= top0 && top < top0 + n_top;
            if (alive && prune) {
                // Every block that overlaps the window [cx+lo_x, cx+hi_x] x [cy+lo_y, cy+hi_y] (rectangle-local block
                // coordinates).  A block whose maximum does not beat the candidate holds no element that does.  A block
                // with a larger maximum beats the candidate at once if that maximum lies in the window; otherwise it
                // is a "suspect": only its in-window elements still have to be looked at.
                const unsigned long long mine = nms_key(v, cy, cx);
                const int wx_lo = cx + lo_x, wx_hi = cx + hi_x, wy_lo = cy + lo_y, wy_hi = cy + hi_y;  // global
                j0 = max(0, (wx_lo - R.c0) >> 3);
                i0 = max(0, (wy_lo - R.r0) >> 3);
                const int j1 = min(bg.bx - 1, (wx_hi - R.c0) >> 3), i1 = min(bg.by - 1, (wy_hi - R.r0) >> 3);
                nj = j1 - j0 + 1;
                const bool fits = nj * (i1 - i0 + 1) <= 64;
                for (int i = i0; i <= i1 && alive; ++i)
                    for (int jb = j0; jb <= j1 && alive; jb += NMS_ROW_BATCH) {  // a batch of independent loads per step
                        unsigned long long kk[NMS_ROW_BATCH];
#pragma unroll
                        for (int u = 0; u < NMS_ROW_BATCH; ++u) kk[u] = jb + u <= j1 ? bm[(size_t)i * bg.bx + jb + u] : 0ull;
#pragma unroll
                        for (int u = 0; u < NMS_ROW_BATCH; ++u) {
                            const unsigned long long k = kk[u];
                            const int j = jb + u;
                            if (!alive || j > j1 || k <= mine) continue;
                            const int ky = 0xFFFF - (int)((k >> 16) & 0xFFFFu), kx = 0xFFFF - (int)(k & 0xFFFFu);
                            if (ky >= wy_lo && ky <= wy_hi && kx >= wx_lo && kx <= wx_hi) { alive = false; continue; }
                            if (fits) suspects |= 1ull << ((i - i0) * nj + (j - j0));
                            else scan_all = true;
                        }
                    }
            } else if (alive) {
                scan_all = true;
            }
        }
        const float cv = v;
        if (alive && !scan_all && suspects == 0ull) {  // nothing left that could beat it
            emit(cx, cy, cv);
            alive = false;
        }
        unsigned todo = __ballot_sync(0xffffffffu, alive);
        while (todo) {  // exact verification of what is left, warp-cooperative
            const int src_lane = __ffs(todo) - 1;
            todo &= todo - 1;
            const int ccx = __shfl_sync(0xffffffffu, cx, src_lane), ccy = __shfl_sync(0xffffffffu, cy, src_lane);
            const float ccv = __shfl_sync(0xffffffffu, cv, src_lane);
            const bool all = __shfl_sync(0xffffffffu, (int)scan_all, src_lane) != 0;
            bool beaten = false;
            if (!all) {
                unsigned long long sus = __shfl_sync(0xffffffffu, suspects, src_lane);
                const int bi0 = __shfl_sync(0xffffffffu, i0, src_lane), bj0 = __shfl_sync(0xffffffffu, j0, src_lane);
                const int bnj = __shfl_sync(0xffffffffu, nj, src_lane);
                while (sus && !beaten) {  // the 64 elements of one suspect block, two per lane
                    const int bit = __ffsll((long long)sus) - 1;
                    sus &= sus - 1;
                    const int bi = bi0 + bit / bnj, bj = bj0 + bit % bnj;
                    bool bt = false;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int rr = bi * NMS_BLOCK + h * 4 + (lane >> 3), rc = bj * NMS_BLOCK + (lane & 7);
                        const int dy = R.r0 + rr - ccy, dx = R.c0 + rc - ccx;
                        if (rr < R.nr && rc < R.nc && dy >= lo_y && dy <= hi_y && dx >= lo_x && dx <= hi_x) {
                            const float el = __ldg(src + (size_t)rr * R.nc + rc);
                            bt |= el > ccv || (el == ccv && (dy < 0 || (dy == 0 && dx < 0)));
                        }
                    }
                    beaten = __any_sync(0xffffffffu, bt);
                }
            } else {
                const int n_el = wx * wy;
                for (int e0 = 0; e0 < n_el && !beaten; e0 += 32 * NMS_VERIFY_LOADS) {
                    float el[NMS_VERIFY_LOADS];
                    int early[NMS_VERIFY_LOADS];
#pragma unroll
                    for (int r = 0; r < NMS_VERIFY_LOADS; ++r) {
                        const int i = e0 + r * 32 + lane;
                        const int wy_i = i / wx, dx = lo_x + (i - wy_i * wx), dy = lo_y + wy_i;
                        const int rr = ccy + dy - R.r0, rc = ccx + dx - R.c0;
                        early[r] = (dy < 0 || (dy == 0 && dx < 0)) ? 1 : 0;
                        el[r] = (i < n_el && rr >= 0 && rr < R.nr && rc >= 0 && rc < R.nc) ? __ldg(src + (size_t)rr * R.nc + rc) : 0.f;
                    }
                    bool bt = false;
#pragma unroll
                    for (int r = 0; r < NMS_VERIFY_LOADS; ++r) bt |= el[r] > ccv || (el[r] == ccv && early[r]);
                    beaten = __any_sync(0xffffffffu, bt);
                }
            }
            if (!beaten && lane == 0) emit(ccx, ccy, ccv);
        }
    }
}

// Block-wide arg-max over shared keys; returns the winning key (0 if none) and clears it.
__device__ __forceinline__ unsigned long long block_pop_max(unsigned long long* keys, int n, unsigned long long* s_red) {
    unsigned long long best = 0;
    int bi = -1;
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        if (keys[i] > best) { best = keys[i]; bi = i; }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const unsigned long long ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best) { best = ob; bi = oi; }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    if (lane == 0) { s_red[warp * 2] = best; s_red[warp * 2 + 1] = (unsigned long long)(long long)bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long b = 0;
        long long i = -1;
        for (int w = 0; w < nw; ++w)
            if (s_red[w * 2] > b) { b = s_red[w * 2]; i = (long long)s_red[w * 2 + 1]; }
        if (i >= 0) keys[i] = 0;
        s_red[64] = b;
    }
    __syncthreads();
    const unsigned long long r = s_red[64];
    __syncthreads();
    return r;
}

// Centre lists: per class the top max_loc maxima, and the ratio gate of HFTest.cpp:726.
struct ObjectLimits {
    int max_loc[HF6D_MAX_CLASSES];
    uint8_t should_detect[HF6D_MAX_CLASSES];
};

__global__ void __launch_bounds__(256)
select_centres_kernel(const unsigned long long* __restrict__ list, const int* __restrict__ list_n, ObjectLimits lim,
                      float min_ratio, hf6d_centre_list* __restrict__ out, uint8_t* __restrict__ active) {
    __shared__ unsigned long long s_k[NMS_LIST_CAP];
    __shared__ unsigned long long s_red[65];
    const int c = blockIdx.x;
    const int n = min(list_n[c], NMS_LIST_CAP);
    for (int i = threadIdx.x; i < n; i += blockDim.x) s_k[i] = list[(size_t)c * NMS_LIST_CAP + i];
    __syncthreads();
    const int want = lim.should_detect[c] ? min(min(lim.max_loc[c], n), HF6D_MAX_CENTRES) : 0;
    float top = 0.f;
    for (int k = 0; k < want; ++k) {
        const unsigned long long key = block_pop_max(s_k, n, s_red);
        if (threadIdx.x == 0) {
            const float sc = __uint_as_float((unsigned)(key >> 32));
            if (k == 0) top = sc;
            out[c].c[k].score = sc;
            out[c].c[k].x = 0xFFFF - (int)((key >> 16) & 0xFFFFu);
            out[c].c[k].y = 0xFFFF - (int)(key & 0xFFFFu);
            active[c * HF6D_MAX_CENTRES + k] = !(__fdiv_rn(sc, top) < min_ratio);
        }
    }
    if (threadIdx.x == 0) {
        out[c].n = want;
        for (int k = want; k < HF6D_MAX_CENTRES; ++k) {
            out[c].c[k].score = 0.f; out[c].c[k].x = 0; out[c].c[k].y = 0;
            active[c * HF6D_MAX_CENTRES + k] = 0;
        }
    }
}

// z mode per slot (HFTest.cpp:803-812): NMS (1 wide, z_nms tall) on the 300-bin histogram; best = highest score,
// earliest on ties.  A slot with no z mode produces no hypotheses (its active flag is cleared).
__global__ void __launch_bounds__(128)
z_mode_kernel(const unsigned long long* __restrict__ zacc, int z_nms, uint8_t* __restrict__ active,
              float* __restrict__ mode_z) {
    __shared__ float zf[HF6D_Z_BINS];
    __shared__ unsigned long long s_best;
    const int s = blockIdx.x;
    if (!active[s]) return;
    if (threadIdx.x == 0) s_best = 0;
    for (int i = threadIdx.x; i < HF6D_Z_BINS; i += blockDim.x)
        zf[i] = (float)((double)zacc[(size_t)s * HF6D_Z_BINS + i] / 65536.0);
    __syncthreads();
    const int n_top = HF6D_Z_BINS - 2 * z_nms + 2;  // tops 0 .. rows-2*wy+1
    for (int top = threadIdx.x; top < n_top; top += blockDim.x) {
        int brow = top;
        float bv = zf[top];
        for (int r = top + 1; r < top + z_nms; ++r)
            if (zf[r] > bv) { bv = zf[r]; brow = r; }
        if (bv != 0.f && brow == top + z_nms / 2)
            atomicMax(&s_best, ((unsigned long long)__float_as_uint(bv) << 32) | (unsigned long long)(0xFFFFu - (unsigned)brow));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_best == 0) active[s] = 0;
        else mode_z[s] = __fmul_rn((float)(0xFFFF - (int)(s_best & 0xFFFFu)), 0.01f);
    }
}

// yaw/pitch peaks per slot (HFTest.cpp:836-848): top max_yp maxima, stop at the first whose score/top < ratio.
__global__ void __launch_bounds__(256)
select_peaks_kernel(const unsigned long long* __restrict__ list, const int* __restrict__ list_n,
                    const uint8_t* __restrict__ active, int max_yp, float min_ratio, int* __restrict__ n_peaks,
                    int* __restrict__ peak_yx, float* __restrict__ peak_score) {
    __shared__ unsigned long long s_k[NMS_LIST_CAP];
    __shared__ unsigned long long s_red[65];
    __shared__ int s_stop;
    const int s = blockIdx.x;
    if (!active[s]) { if (threadIdx.x == 0) n_peaks[s] = 0; return; }
    const int n = min(list_n[s], NMS_LIST_CAP);
    for (int i = threadIdx.x; i < n; i += blockDim.x) s_k[i] = list[(size_t)s * NMS_LIST_CAP + i];
    if (threadIdx.x == 0) s_stop = 0;
    __syncthreads();
    const int want = min(max_yp, n);
    float top = 0.f;
    int got = 0;
    for (int k = 0; k < want; ++k) {
        const unsigned long long key = block_pop_max(s_k, n, s_red);
        if (threadIdx.x == 0) {
            const float sc = __uint_as_float((unsigned)(key >> 32));
            if (k == 0) top = sc;
            const float ratio = __fdiv_rn(sc, top);
            if (ratio < min_ratio) s_stop = 1;
            else {
                peak_yx[(s * max_yp + k) * 2] = 0xFFFF - (int)(key & 0xFFFFu);              // row = yaw bin
                peak_yx[(s * max_yp + k) * 2 + 1] = 0xFFFF - (int)((key >> 16) & 0xFFFFu);  // col = pitch bin
                peak_score[s * max_yp + k] = ratio;
                got = k + 1;
            }
        }
        __syncthreads();
        if (s_stop) break;
    }
    if (threadIdx.x == 0) n_peaks[s] = got;
}

// Roll modes per (slot, peak) (HFTest.cpp:874-925): blur (1 x pose_blur), NMS (1 x pose_nms), keep [180, 540],
// greedily up to max_roll modes more than 7 degrees from the previously accepted one (sep_ok table = the reference's
// acos(dot) test, evaluated on the host with libm for every integer pair).
struct HypRecord {
    int32_t valid, roll_bin;
    float roll_score;
};

__global__ void __launch_bounds__(128)
roll_modes_kernel(const unsigned long long* __restrict__ racc, const int* __restrict__ n_peaks, int max_yp,
                  int blur, int nms, int max_roll, const uint8_t* __restrict__ sep_ok /*[361][361]*/,
                  HypRecord* __restrict__ out /*[S][max_yp][max_roll]*/) {
    __shared__ unsigned long long s_acc[HF6D_POSE_BINS];
    __shared__ float s_blur[HF6D_POSE_BINS];
    __shared__ unsigned long long s_keys[HF6D_POSE_BINS];
    __shared__ int s_n;
    const int s = blockIdx.x, h = blockIdx.y;
    HypRecord* o = out + ((size_t)s * max_yp + h) * max_roll;
    if (h >= n_peaks[s]) {
        for (int i = threadIdx.x; i < max_roll; i += blockDim.x) o[i].valid = 0;
        return;
    }
    const int NB = HF6D_POSE_BINS;
    for (int i = threadIdx.x; i < NB; i += blockDim.x) s_acc[i] = racc[((size_t)s * max_yp + h) * NB + i];
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    const double scale = 1.0 / (double)blur;
    for (int i = threadIdx.x; i < NB; i += blockDim.x) {
        unsigned long long sum = 0;
        const int a = i - blur / 2;
        if (a >= 0 && a + blur <= NB) {  // interior bins: no border reflection
#pragma unroll 5
            for (int k = 0; k < blur; ++k) sum += s_acc[a + k];
        } else {
            for (int k = 0; k < blur; ++k) sum += s_acc[reflect101(a + k, NB)];
        }
        s_blur[i] = (float)(((double)sum / 65536.0) * scale);
    }
    __syncthreads();
    const int n_top = NB - 2 * nms + 2;
    for (int top = threadIdx.x; top < n_top; top += blockDim.x) {
        int brow = top;
        float bv = s_blur[top];
        for (int r = top + 1; r < top + nms; ++r)
            if (s_blur[r] > bv) { bv = s_blur[r]; brow = r; }
        const int cy = top + nms / 2;
        if (bv != 0.f && brow == cy && cy >= 180 && cy <= 540) {
            const int idx = atomicAdd(&s_n, 1);
            s_keys[idx] = ((unsigned long long)__float_as_uint(bv) << 32) | (unsigned long long)(0xFFFFu - (unsigned)cy);
        }
    }
    __syncthreads();
    // sort by key, descending: keys are distinct (they carry the bin), so a key's rank is the number of larger keys
    {
        const int n = s_n;
        unsigned long long mine[(HF6D_POSE_BINS + 127) / 128];
        int rank[(HF6D_POSE_BINS + 127) / 128];
        int cnt = 0;
        for (int i = threadIdx.x; i < n; i += blockDim.x, ++cnt) {
            const unsigned long long k = s_keys[i];
            int r = 0;
            for (int j = 0; j < n; ++j) r += s_keys[j] > k;
            mine[cnt] = k;
            rank[cnt] = r;
        }
        __syncthreads();
        for (int q = 0; q < cnt; ++q) s_keys[rank[q]] = mine[q];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int n = s_n;
        int got = 0, prev = -1;
        const float top = n ? __uint_as_float((unsigned)(s_keys[0] >> 32)) : 0.f;
        for (int i = 0; i < n && got < max_roll; ++i) {
            const int ry = 0xFFFF - (int)(s_keys[i] & 0xFFFFu);
            if (got == 0 || sep_ok[(prev - 180) * 361 + (ry - 180)]) {
                o[got].valid = 1;
                o[got].roll_bin = ry;
                o[got].roll_score = __fdiv_rn(__uint_as_float((unsigned)(s_keys[i] >> 32)), top);
                prev = ry;
                ++got;
            }
        }
        for (int i = got; i < max_roll; ++i) o[i].valid = 0;
    }
}

}  // namespace hf6d
