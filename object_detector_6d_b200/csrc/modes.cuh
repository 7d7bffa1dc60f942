// Stages CENTRES and POSE: mode seeking on the vote accumulators.
//
// The reference's "mode seeking" is a normalised box blur followed by a sliding-window non-max suppression, applied
// hierarchically: centre map -> z histogram -> yaw/pitch map -> roll histogram
// (HoughForest/src/HFTest.cpp:702-707, 742-925; NMS HFTest.cpp:219-268).  These kernels are that neighbour reduction,
// tiled:
//  * box blur   = exact 64-bit integer window sums (row pass on a per-row prefix sum in shared memory, column pass as a
//                 running sum), scaled once in double:  (float)((double)S / 65536 * 1/(kx*ky))  -- what cv::blur does
//                 on CV_32F (double accumulation, single scale) without its summation-order dependence.
//  * NMS        = separable sliding-window maximum over 64-bit keys  (value bits | ~row | ~col)  computed by window
//                 doubling in shared memory (row pass, then column pass); the key order reproduces the reference's
//                 monotonic-deque tie-breaking, and a window emits only if its maximum sits at the window centre.
//                 The reference's loop-bound quirk (the vertical pass stops at rows-wy, so the bottom wy-1 window rows
//                 are never produced) is kept.
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

// ------------------------------------------------------------------------------------------------ box blur
constexpr int BLUR_WARPS = 4;

// tmp[m][r][c] = sum_k acc[m][r][reflect(c - kx/2 + k)]   for r in in.rows, c in out.cols  (tmp is in.nr x out.nc)
__global__ void __launch_bounds__(BLUR_WARPS * 32)
box_rows_kernel(const unsigned long long* __restrict__ acc, unsigned long long* __restrict__ tmp, MapDims md, MapRect in,
                MapRect out, int kx, const uint8_t* __restrict__ map_active) {
    extern __shared__ unsigned long long s_pre[];  // [BLUR_WARPS][in.nc + 1]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m = blockIdx.y;
    if (map_active && !map_active[m]) return;
    const int r = blockIdx.x * BLUR_WARPS + warp;
    if (r >= in.nr) return;
    unsigned long long* pre = s_pre + (size_t)warp * (in.nc + 1);
    const unsigned long long* src = acc + ((size_t)m * in.nr + r) * in.nc;
    unsigned long long carry = 0;
    if (lane == 0) pre[0] = 0;
    for (int c0 = 0; c0 < in.nc; c0 += 32) {
        const int c = c0 + lane;
        unsigned long long v = c < in.nc ? src[c] : 0ull;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long n = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += n;
        }
        if (c < in.nc) pre[c + 1] = carry + v;
        carry += __shfl_sync(0xffffffffu, v, 31);
    }
    __syncwarp();
    auto seg = [&](int a, int b) -> unsigned long long {  // sum over global columns [a, b], clipped to the input rectangle
        a = max(a, in.c0);
        b = min(b, in.c0 + in.nc - 1);
        return b >= a ? pre[b - in.c0 + 1] - pre[a - in.c0] : 0ull;
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
constexpr int BLUR_COL_CHUNK = 32;
__global__ void __launch_bounds__(128)
box_cols_kernel(const unsigned long long* __restrict__ tmp, float* __restrict__ dst_all, MapDims md, MapRect in,
                MapRect out, int ky, double scale, const uint8_t* __restrict__ map_active) {
    const int m = blockIdx.z;
    if (map_active && !map_active[m]) return;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int rbeg = blockIdx.y * BLUR_COL_CHUNK;
    if (c >= out.nc || rbeg >= out.nr) return;
    const unsigned long long* src = tmp + (size_t)m * in.nr * out.nc + c;
    float* dst = dst_all + (size_t)m * out.nr * out.nc + c;
    const int n = md.rows;
    auto at = [&](int gr) -> unsigned long long {  // global row, reflected; rows outside the input rectangle are zero
        const int rr = reflect101(gr, n) - in.r0;
        return (rr >= 0 && rr < in.nr) ? src[(size_t)rr * out.nc] : 0ull;
    };
    const int rend = min(rbeg + BLUR_COL_CHUNK, out.nr);
    unsigned long long s = 0;
    {
        const int a = out.r0 + rbeg - ky / 2;
        for (int k = 0; k < ky; ++k) s += at(a + k);
    }
    for (int r = rbeg; r < rend; ++r) {
        dst[(size_t)r * out.nc] = (float)(((double)s / 65536.0) * scale);
        const int a = out.r0 + r - ky / 2;
        s += at(a + ky);
        s -= at(a);
    }
}

// ------------------------------------------------------------------------------------------------ NMS
// Separable sliding-window maximum over 64-bit keys  (value bits | ~row | ~col).  Values are >= 0, so float bits order
// like unsigned integers; ~row / ~col make the maximum pick the topmost row, then the leftmost column among equal
// values -- the element the reference's two monotonic deques keep at their front (HFTest.cpp:232-262).
//   1. rowkey[r][left] = max key of in[r][left .. left+wx-1]         one warp per row, window doubling in shared memory
//   2. key(left, top)  = max of rowkey[top .. top+wy-1][left]        32 x 64 origins per CTA, doubling along y;
//      the window emits iff its value != 0 and the winning element is the window centre.
constexpr int NMS_LIST_CAP = 4096;
constexpr int NMS_ROW_WARPS = 4;

__device__ __forceinline__ unsigned long long nms_key(float v, int gy, int gx) {
    return ((unsigned long long)__float_as_uint(v) << 32) | ((unsigned long long)(0xFFFFu - (unsigned)gy) << 16) |
           (unsigned long long)(0xFFFFu - (unsigned)gx);
}

// in: float [M][R.nr][R.nc]; rowkey: u64 [M][R.nr][n_left] for window lefts left0 .. left0+n_left-1 (global coords)
__global__ void __launch_bounds__(NMS_ROW_WARPS * 32)
nms_rowmax_kernel(const float* __restrict__ in, unsigned long long* __restrict__ rowkey, MapRect R, int wx, int left0,
                  int n_left, const uint8_t* __restrict__ map_active) {
    extern __shared__ unsigned long long s_row[];  // [NMS_ROW_WARPS][2][span], span = n_left + wx - 1
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m = blockIdx.y;
    if (map_active && !map_active[m]) return;
    const int r = blockIdx.x * NMS_ROW_WARPS + warp;
    if (r >= R.nr) return;
    const int span = n_left + wx - 1;
    unsigned long long* A = s_row + (size_t)warp * 2 * span;
    unsigned long long* B = A + span;
    const float* src = in + ((size_t)m * R.nr + r) * R.nc;
    const int gy = R.r0 + r;
    for (int x = lane; x < span; x += 32) {
        const int gx = left0 + x, rc = gx - R.c0;
        A[x] = nms_key((rc >= 0 && rc < R.nc) ? src[rc] : 0.f, gy, gx);
    }
    __syncwarp();
    int p = 1;
    while (p * 2 <= wx) {
        for (int x = lane; x < span; x += 32) B[x] = x + p < span ? max(A[x], A[x + p]) : A[x];
        __syncwarp();
        unsigned long long* t = A; A = B; B = t;
        p *= 2;
    }
    unsigned long long* dst = rowkey + ((size_t)m * R.nr + r) * n_left;
    for (int x = lane; x < n_left; x += 32) dst[x] = max(A[x], A[x + (wx - p)]);
}

constexpr int NMS_COL_TW = 32;   // window origins per CTA in x
constexpr int NMS_COL_TH = 64;   // and in y
constexpr int NMS_COL_THREADS = 256;
inline size_t nms_col_smem_bytes(int wy) { return (size_t)(NMS_COL_TH + wy - 1) * NMS_COL_TW * 8 * 2; }

// Window origins (left, top) in global coordinates: left in [left0, left0+n_left), top in [top0, top0+n_top).
// Rows outside the rectangle R are zero.  Emits sort keys (score | ~x | ~y) into list[m].
__global__ void __launch_bounds__(NMS_COL_THREADS)
nms_emit_kernel(const unsigned long long* __restrict__ rowkey, MapRect R, int wx, int wy, int left0, int n_left,
                int top0, int n_top, unsigned long long* __restrict__ list, int* __restrict__ list_n,
                const uint8_t* __restrict__ map_active) {
    extern __shared__ unsigned long long s_col[];
    const int m = blockIdx.z;
    if (map_active && !map_active[m]) return;
    const int th = NMS_COL_TH + wy - 1;
    unsigned long long* A = s_col;
    unsigned long long* B = s_col + (size_t)th * NMS_COL_TW;
    const int lx0 = blockIdx.x * NMS_COL_TW, ty0 = blockIdx.y * NMS_COL_TH;
    const unsigned long long* rk = rowkey + (size_t)m * R.nr * n_left;
    for (int i = threadIdx.x; i < th * NMS_COL_TW; i += NMS_COL_THREADS) {
        const int y = i / NMS_COL_TW, x = i % NMS_COL_TW;
        const int rr = top0 + ty0 + y - R.r0, lx = lx0 + x;
        A[i] = (rr >= 0 && rr < R.nr && lx < n_left) ? rk[(size_t)rr * n_left + lx] : 0ull;
    }
    __syncthreads();
    int p = 1;
    while (p * 2 <= wy) {
        for (int i = threadIdx.x; i < th * NMS_COL_TW; i += NMS_COL_THREADS) {
            const int y = i / NMS_COL_TW;
            B[i] = y + p < th ? max(A[i], A[i + p * NMS_COL_TW]) : A[i];
        }
        __syncthreads();
        unsigned long long* t = A; A = B; B = t;
        p *= 2;
    }
    for (int i = threadIdx.x; i < NMS_COL_TH * NMS_COL_TW; i += NMS_COL_THREADS) {
        const int y = i / NMS_COL_TW, x = i % NMS_COL_TW;
        if (lx0 + x >= n_left || ty0 + y >= n_top) continue;
        const unsigned long long k = max(A[y * NMS_COL_TW + x], A[(y + (wy - p)) * NMS_COL_TW + x]);
        const unsigned vb = (unsigned)(k >> 32);
        if (vb == 0u) continue;
        const int gy = 0xFFFF - (int)((k >> 16) & 0xFFFFu), gx = 0xFFFF - (int)(k & 0xFFFFu);
        const int ccx = left0 + lx0 + x + wx / 2, ccy = top0 + ty0 + y + wy / 2;
        if (gx != ccx || gy != ccy) continue;
        const int idx = atomicAdd(list_n + m, 1);
        if (idx < NMS_LIST_CAP)
            list[(size_t)m * NMS_LIST_CAP + idx] = ((unsigned long long)vb << 32) |
                                                   ((unsigned long long)(0xFFFFu - (unsigned)ccx) << 16) |
                                                   (unsigned long long)(0xFFFFu - (unsigned)ccy);
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
        for (int k = 0; k < blur; ++k) sum += s_acc[reflect101(i - blur / 2 + k, NB)];
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
    if (threadIdx.x == 0) {
        const int n = s_n;
        // selection sort by key desc (n <= ~40)
        for (int a = 0; a < n; ++a) {
            int b = a;
            for (int i = a + 1; i < n; ++i) if (s_keys[i] > s_keys[b]) b = i;
            const unsigned long long t = s_keys[a]; s_keys[a] = s_keys[b]; s_keys[b] = t;
        }
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
