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
//                 doubling in shared memory; the key order reproduces the reference's monotonic-deque tie-breaking
//                 (leftmost in a row, then topmost), and a window emits only if its maximum sits at the window centre.
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

// A rectangular region [r0, r0+nr) x [c0, c0+nc) of a (rows x cols) map, stored densely (nr x nc) per map.
struct MapRegion {
    int rows, cols;  // full (virtual) map
    int r0, c0, nr, nc;
};

// ------------------------------------------------------------------------------------------------ box blur
constexpr int BLUR_WARPS = 4;

// tmp[m][r][c] = sum_k acc[m][r][reflect(c - kx/2 + k)]   (region coordinates; out-of-region terms count as 0)
__global__ void __launch_bounds__(BLUR_WARPS * 32)
box_rows_kernel(const unsigned long long* __restrict__ acc, unsigned long long* __restrict__ tmp, MapRegion mr, int kx,
                const uint8_t* __restrict__ map_active) {
    extern __shared__ unsigned long long s_pre[];  // [BLUR_WARPS][nc + 1]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m = blockIdx.y;
    if (map_active && !map_active[m]) return;
    const int r = blockIdx.x * BLUR_WARPS + warp;
    if (r >= mr.nr) return;
    unsigned long long* pre = s_pre + (size_t)warp * (mr.nc + 1);
    const unsigned long long* src = acc + ((size_t)m * mr.nr + r) * mr.nc;
    unsigned long long carry = 0;
    if (lane == 0) pre[0] = 0;
    for (int c0 = 0; c0 < mr.nc; c0 += 32) {
        const int c = c0 + lane;
        unsigned long long v = c < mr.nc ? src[c] : 0ull;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long n = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += n;
        }
        if (c < mr.nc) pre[c + 1] = carry + v;
        carry += __shfl_sync(0xffffffffu, v, 31);
    }
    __syncwarp();
    auto seg = [&](int a, int b) -> unsigned long long {  // sum over global columns [a, b], clipped to the region
        a = max(a, mr.c0);
        b = min(b, mr.c0 + mr.nc - 1);
        return b >= a ? pre[b - mr.c0 + 1] - pre[a - mr.c0] : 0ull;
    };
    unsigned long long* dst = tmp + ((size_t)m * mr.nr + r) * mr.nc;
    const int n = mr.cols;
    for (int c = lane; c < mr.nc; c += 32) {
        const int a = mr.c0 + c - kx / 2, b = a + kx - 1;
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

// out[m][r][c] = (float)( (double)(sum_k tmp[m][reflect(r - ky/2 + k)][c]) / 65536 * scale )
constexpr int BLUR_COL_CHUNK = 16;
__global__ void __launch_bounds__(128)
box_cols_kernel(const unsigned long long* __restrict__ tmp, float* __restrict__ out, MapRegion mr, int ky, double scale,
                const uint8_t* __restrict__ map_active) {
    const int m = blockIdx.z;
    if (map_active && !map_active[m]) return;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int rbeg = blockIdx.y * BLUR_COL_CHUNK;
    if (c >= mr.nc || rbeg >= mr.nr) return;
    const unsigned long long* src = tmp + (size_t)m * mr.nr * mr.nc + c;
    float* dst = out + (size_t)m * mr.nr * mr.nc + c;
    const int n = mr.rows;
    auto at = [&](int gr) -> unsigned long long {  // global row, reflected; out-of-region rows count as 0
        const int rr = reflect101(gr, n) - mr.r0;
        return (rr >= 0 && rr < mr.nr) ? src[(size_t)rr * mr.nc] : 0ull;
    };
    const int rend = min(rbeg + BLUR_COL_CHUNK, mr.nr);
    unsigned long long s = 0;
    {
        const int a = mr.r0 + rbeg - ky / 2;
        for (int k = 0; k < ky; ++k) s += at(a + k);
    }
    for (int r = rbeg; r < rend; ++r) {
        dst[(size_t)r * mr.nc] = (float)(((double)s / 65536.0) * scale);
        const int a = mr.r0 + r - ky / 2;
        s += at(a + ky);
        s -= at(a);
    }
}

// ------------------------------------------------------------------------------------------------ NMS
constexpr int NMS_TILE = 32;
constexpr int NMS_THREADS = 256;
constexpr int NMS_LIST_CAP = 4096;

__device__ __forceinline__ unsigned long long nms_key(float v, int gy, int gx) {
    return ((unsigned long long)__float_as_uint(v) << 32) | ((unsigned long long)(0xFFFFu - (unsigned)gy) << 16) |
           (unsigned long long)(0xFFFFu - (unsigned)gx);
}

inline size_t nms_smem_bytes(int wx, int wy) { return (size_t)(NMS_TILE + wx - 1) * (NMS_TILE + wy - 1) * 8 * 2; }

// in: float [M][nr][nc] (region of a rows x cols map).  Window origins (left, top) in global coordinates:
//   left in [0, cols-wx], top in [0, rows-2*wy+1]  (reference loop bounds), further clipped to centres inside
//   [keep_lo, keep_hi]^2 when keep_lo >= 0.  Emits sort keys (score | ~x | ~y) into list[m].
__global__ void __launch_bounds__(NMS_THREADS)
nms_tile_kernel(const float* __restrict__ in, MapRegion mr, int wx, int wy, int keep_lo, int keep_hi, int left0,
                int top0, int n_left, int n_top, unsigned long long* __restrict__ list, int* __restrict__ list_n,
                const uint8_t* __restrict__ map_active) {
    extern __shared__ unsigned long long s_keys[];
    const int m = blockIdx.z;
    if (map_active && !map_active[m]) return;
    const int tw = NMS_TILE + wx - 1, th = NMS_TILE + wy - 1;
    unsigned long long* A = s_keys;
    unsigned long long* B = s_keys + (size_t)tw * th;
    const int left_base = left0 + blockIdx.x * NMS_TILE, top_base = top0 + blockIdx.y * NMS_TILE;
    const float* src = in + (size_t)m * mr.nr * mr.nc;
    for (int i = threadIdx.x; i < tw * th; i += NMS_THREADS) {
        const int y = i / tw, x = i % tw;
        const int gy = top_base + y, gx = left_base + x;
        const int ry = gy - mr.r0, rx = gx - mr.c0;
        unsigned long long k = 0;
        if (gy < mr.rows && gx < mr.cols && ry >= 0 && ry < mr.nr && rx >= 0 && rx < mr.nc)
            k = nms_key(src[(size_t)ry * mr.nc + rx], gy, gx);
        A[i] = k;
    }
    __syncthreads();
    // horizontal windows of wx by doubling: after the loop A[y][x] = max over [x, x+p)
    int p = 1;
    while (p * 2 <= wx) {
        for (int i = threadIdx.x; i < tw * th; i += NMS_THREADS) {
            const int x = i % tw;
            unsigned long long k = A[i];
            if (x + p < tw) k = max(k, A[i + p]);
            B[i] = k;
        }
        __syncthreads();
        unsigned long long* t = A; A = B; B = t;
        p *= 2;
    }
    for (int i = threadIdx.x; i < tw * th; i += NMS_THREADS) {
        const int x = i % tw;
        unsigned long long k = A[i];
        if (x + (wx - p) < tw) k = max(k, A[i + (wx - p)]);
        B[i] = k;  // valid for x < NMS_TILE
    }
    __syncthreads();
    { unsigned long long* t = A; A = B; B = t; }
    // vertical windows of wy
    p = 1;
    while (p * 2 <= wy) {
        for (int i = threadIdx.x; i < tw * th; i += NMS_THREADS) {
            const int y = i / tw;
            unsigned long long k = A[i];
            if (y + p < th) k = max(k, A[i + p * tw]);
            B[i] = k;
        }
        __syncthreads();
        unsigned long long* t = A; A = B; B = t;
        p *= 2;
    }
    for (int i = threadIdx.x; i < NMS_TILE * NMS_TILE; i += NMS_THREADS) {
        const int y = i / NMS_TILE, x = i % NMS_TILE;
        const int left = left_base + x, top = top_base + y;
        if (left >= left0 + n_left || top >= top0 + n_top) continue;
        unsigned long long k = A[y * tw + x];
        if (y + (wy - p) < th) k = max(k, A[(y + (wy - p)) * tw + x]);
        const unsigned vb = (unsigned)(k >> 32);
        const int gy = 0xFFFF - (int)((k >> 16) & 0xFFFFu), gx = 0xFFFF - (int)(k & 0xFFFFu);
        const int ccx = left + wx / 2, ccy = top + wy / 2;
        if (vb != 0u && gx == ccx && gy == ccy) {
            if (keep_lo >= 0 && (ccx < keep_lo || ccx > keep_hi || ccy < keep_lo || ccy > keep_hi)) continue;
            const int idx = atomicAdd(list_n + m, 1);
            if (idx < NMS_LIST_CAP)
                list[(size_t)m * NMS_LIST_CAP + idx] = ((unsigned long long)vb << 32) |
                                                       ((unsigned long long)(0xFFFFu - (unsigned)ccx) << 16) |
                                                       (unsigned long long)(0xFFFFu - (unsigned)ccy);
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
