// Hough-forest TRAINING on the GPU: HFTrain (HoughForest/src/HFTrain.cpp:87-1265) level by level.
//
// The reference optimises one tree depth at a time over ALL samples of ALL nodes of that depth (HFTrain.h:5-10), with OpenMP
// threads filling per-thread hash maps of per-node, per-test statistics that are merged under a critical section.  Here a level
// is a handful of kernels over the same decomposition, made for the machine:
//  * the training samples are kept as an index array sorted by node (a stable partition per level), cut into CHUNKS of at most
//    TR_CHUNK samples of one node; a CTA owns a chunk and ONE THREAD OWNS ONE TEST: it keeps its (mode, f1, f2, threshold) and
//    its left-child statistics in registers while the CTA streams the chunk's feature rows through shared memory (coalesced
//    800-float rows, TR_TILE at a time).  No atomics on the hot loop, no hash maps, every feature row read once per pass.
//  * integer statistics (class counts, sample counts) are added to the node's totals with integer atomics (exact, order
//    free); floating-point sums go to a per-chunk partial and are reduced per node in chunk order, so a forest is
//    reproducible bit for bit -- the reference's is not even reproducible by itself (rand() from OpenMP threads, float
//    sums in schedule order).
//  * the random draws come from a counter-based generator keyed by (seed, tree, level, node, test, draw): any thread can
//    compute any draw, and oracle/train.py computes the same ones.
// The host keeps the tree (it is a few thousand nodes), decides which children become leaves (HFTrain.cpp:1159-1176) and
// builds the chunk table of the next level; it never touches a feature.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace hf6d {

constexpr int TR_CHUNK = 1024;   // samples of one node per CTA
constexpr int TR_TILE = 8;       // feature rows staged at a time
constexpr int TR_MAX_K = 32;
constexpr int TR_V = 6;          // floating-point sums per (test, child): location x, y, z, |v|^2 ; pose 3 x (cos, sin)

__host__ __device__ __forceinline__ unsigned long long tr_mix64(unsigned long long z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
// oracle/train.py::rng_u64 (choice T1)
__host__ __device__ __forceinline__ unsigned long long tr_rng(unsigned long long seed, int tree, int level, unsigned node, int test,
                                                               int draw) {
    unsigned long long h = tr_mix64(seed ^ (0x9E3779B97F4A7C15ULL * (unsigned long long)(tree + 1)));
    h = tr_mix64(h + (unsigned long long)(unsigned)level);
    h = tr_mix64(h + (unsigned long long)node);
    return tr_mix64(h + (unsigned long long)((test << 3) | draw));
}
constexpr int TR_DRAW_MODE = 0, TR_DRAW_F1 = 1, TR_DRAW_F2 = 2, TR_DRAW_THR = 3;
constexpr int TR_LEVEL_SHUFFLE = -1;
constexpr unsigned TR_NODE_OBJECTIVE = 0xFFFFFFFFu;

struct TrTest {
    int mode, f1, f2;
    float thr;
};

struct TrChunk {
    int node;    // index of the node in the level
    int first;   // position of its first sample in the order array
    int count;
    int left_before, right_before;  // samples of the node that earlier chunks sent left / right (scatter pass)
};

// order-preserving map float -> uint for atomicMin / atomicMax
__device__ __forceinline__ unsigned tr_f2o(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float tr_o2f(unsigned o) {
    const unsigned u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}

// HFTrain::get_random_features (HFTrain.cpp:231-262): tests_per_node draws of (mode, f1, f2) per node.
__global__ void tr_features_kernel(unsigned long long seed, int tree, int level, int n_nodes, int tpn, int F, TrTest* __restrict__ base) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes * tpn) return;
    const int node = i / tpn, t = i - node * tpn;
    TrTest q;
    q.mode = (int)(tr_rng(seed, tree, level, (unsigned)node, t, TR_DRAW_MODE) % 2ull);
    q.f1 = (int)(tr_rng(seed, tree, level, (unsigned)node, t, TR_DRAW_F1) % (unsigned long long)F);
    q.f2 = (int)(tr_rng(seed, tree, level, (unsigned)node, t, TR_DRAW_F2) % (unsigned long long)F);
    q.thr = 0.f;
    base[i] = q;
}

__device__ __forceinline__ float tr_value(const TrTest& q, const float* row) {
    return q.mode == 0 ? __fsub_rn(row[q.f1], row[q.f2]) : row[q.f1];
}

// HFTrain::get_min_max_count_samples (HFTrain.cpp:267-362): value range of every base test over its node's samples, and the
// node's class histogram.  Thread t < tpn owns base test t; all threads stage the rows.
__global__ void tr_minmax_kernel(const float* __restrict__ feat, int F, const int* __restrict__ cls, const int* __restrict__ order,
                                 const TrChunk* __restrict__ chunks, const TrTest* __restrict__ base, int tpn, int K,
                                 unsigned* __restrict__ omin, unsigned* __restrict__ omax, int* __restrict__ class_cnt) {
    extern __shared__ float tr_rows[];  // [TR_TILE][F]
    __shared__ int s_cls[TR_TILE];
    __shared__ int s_hist[TR_MAX_K];
    const TrChunk ch = chunks[blockIdx.x];
    const int t = threadIdx.x;
    TrTest q{0, 0, 0, 0.f};
    if (t < tpn) q = base[(size_t)ch.node * tpn + t];
    if (t < TR_MAX_K) s_hist[t] = 0;
    float lo = 3.402823466e38f, hi = -3.402823466e38f;
    for (int s0 = 0; s0 < ch.count; s0 += TR_TILE) {
        const int ns = min(TR_TILE, ch.count - s0);
        __syncthreads();
        for (int r = 0; r < ns; ++r) {
            const int smp = order[ch.first + s0 + r];
            const float* src = feat + (size_t)smp * F;
            for (int c = t; c < F; c += blockDim.x) tr_rows[r * F + c] = src[c];
            if (t == 0) s_cls[r] = cls[smp];
        }
        __syncthreads();
        if (t < tpn)
            for (int r = 0; r < ns; ++r) {
                const float v = tr_value(q, tr_rows + r * F);
                lo = fminf(lo, v);  // the reference compares with < and >: NaN never replaces a bound; fminf / fmaxf agree
                hi = fmaxf(hi, v);
            }
        if (t < ns) atomicAdd(&s_hist[s_cls[t]], 1);
    }
    __syncthreads();
    if (t < tpn) {
        atomicMin(omin + (size_t)ch.node * tpn + t, tr_f2o(lo));
        atomicMax(omax + (size_t)ch.node * tpn + t, tr_f2o(hi));
    }
    if (t < K && s_hist[t]) atomicAdd(class_cnt + (size_t)ch.node * K + t, s_hist[t]);
}

// HFTrain::get_random_thresholds (HFTrain.cpp:365-392): every base test repeated tpt times with its own threshold in the
// value range; entry nt of a node ("everything goes left": its statistics are the node's totals) closes the list.
__global__ void tr_thresholds_kernel(unsigned long long seed, int tree, int level, int n_nodes, int tpn, int tpt,
                                     const TrTest* __restrict__ base, const unsigned* __restrict__ omin,
                                     const unsigned* __restrict__ omax, TrTest* __restrict__ tests) {
    const int nt = tpn * tpt;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n_nodes * (nt + 1)) return;
    const int node = (int)(i / (nt + 1)), t = (int)(i - (long long)node * (nt + 1));
    TrTest q;
    if (t == nt) {
        q.mode = 1; q.f1 = 0; q.f2 = 0;
        q.thr = __int_as_float(0x7f800000);  // +inf: val < thr for every finite value
    } else {
        const int b = t / tpt;
        q = base[(size_t)node * tpn + b];
        const float lo = tr_o2f(omin[(size_t)node * tpn + b]), hi = tr_o2f(omax[(size_t)node * tpn + b]);
        const float u = __fdiv_rn((float)(tr_rng(seed, tree, level, (unsigned)node, t, TR_DRAW_THR) >> 40), 16777216.0f);
        q.thr = __fadd_rn(__fmul_rn(u, __fsub_rn(hi, lo)), lo);
    }
    tests[i] = q;
}

// The statistics pass of find_classification_split / find_regression_location_split / find_regression_pose_split
// (HFTrain.cpp:399-470, 531-600, 762-840): METHOD 0: left class histogram per test; 1: left count, sum of (x, y, z) and of
// |(x, y, z)|^2; 2: left count and sums of (cos, sin) of yaw, pitch, roll.  vec = per-sample [9] doubles: x, y, z, then the six
// pose terms.  Integer results go to the node with atomics, double sums to the chunk's partial [chunk][nt + 1][TR_V].
template <int METHOD>
__global__ void tr_stats_kernel(const float* __restrict__ feat, int F, const int* __restrict__ cls, const double* __restrict__ vec,
                                const int* __restrict__ order, const TrChunk* __restrict__ chunks, const TrTest* __restrict__ tests,
                                int nt, int K, int* __restrict__ left_cnt /*[L][nt+1][K or 1]*/, double* __restrict__ partial) {
    extern __shared__ float tr_rows[];  // [TR_TILE][F]
    __shared__ int s_cls[TR_TILE];
    __shared__ double s_vec[TR_TILE][9];
    const TrChunk ch = chunks[blockIdx.x];
    const int t = threadIdx.x;
    const bool mine = t <= nt;
    TrTest q{1, 0, 0, 0.f};
    if (mine) q = tests[(size_t)ch.node * (nt + 1) + t];
    int cnt[METHOD == 0 ? TR_MAX_K : 1];
#pragma unroll
    for (int k = 0; k < (METHOD == 0 ? TR_MAX_K : 1); ++k) cnt[k] = 0;
    double sum[TR_V] = {0, 0, 0, 0, 0, 0};
    for (int s0 = 0; s0 < ch.count; s0 += TR_TILE) {
        const int ns = min(TR_TILE, ch.count - s0);
        __syncthreads();
        for (int r = 0; r < ns; ++r) {
            const int smp = order[ch.first + s0 + r];
            const float* src = feat + (size_t)smp * F;
            for (int c = t; c < F; c += blockDim.x) tr_rows[r * F + c] = src[c];
            if (t == 0) s_cls[r] = cls[smp];
            if (METHOD != 0 && t < 9) s_vec[r][t] = vec[(size_t)smp * 9 + t];
        }
        __syncthreads();
        if (mine)
            for (int r = 0; r < ns; ++r) {
                if (!(tr_value(q, tr_rows + r * F) < q.thr)) continue;
                if (METHOD == 0) {
                    const int c = s_cls[r];
#pragma unroll
                    for (int k = 0; k < TR_MAX_K; ++k) cnt[k] += (k == c);  // static indexing: the histogram stays in registers
                } else if (METHOD == 1) {
                    ++cnt[0];
                    const double x = s_vec[r][0], y = s_vec[r][1], z = s_vec[r][2];
                    sum[0] += x; sum[1] += y; sum[2] += z; sum[3] += x * x + y * y + z * z;
                } else {
                    ++cnt[0];
#pragma unroll
                    for (int k = 0; k < 6; ++k) sum[k] += s_vec[r][3 + k];
                }
            }
    }
    if (!mine) return;
    if (METHOD == 0) {
#pragma unroll
        for (int k = 0; k < TR_MAX_K; ++k)
            if (k < K && cnt[k]) atomicAdd(left_cnt + ((size_t)ch.node * (nt + 1) + t) * K + k, cnt[k]);
    } else {
        if (cnt[0]) atomicAdd(left_cnt + (size_t)ch.node * (nt + 1) + t, cnt[0]);
        double* dst = partial + ((size_t)blockIdx.x * (nt + 1) + t) * TR_V;
#pragma unroll
        for (int k = 0; k < TR_V; ++k) dst[k] = sum[k];
    }
}

// Per node: the chunks' partial sums added in chunk order (chunk_first[node] .. chunk_first[node + 1]).
__global__ void tr_reduce_kernel(const double* __restrict__ partial, const int* __restrict__ chunk_first, int nt,
                                 double* __restrict__ node_sum /*[L][nt+1][TR_V]*/) {
    const int node = blockIdx.x;
    for (int i = threadIdx.x; i < (nt + 1) * TR_V; i += blockDim.x) {
        double s = 0.0;
        for (int c = chunk_first[node]; c < chunk_first[node + 1]; ++c) s += partial[(size_t)c * (nt + 1) * TR_V + i];
        node_sum[(size_t)node * (nt + 1) * TR_V + i] = s;
    }
}

// The choice of the best test per node (HFTrain.cpp:472-523, 700-757, 940-994): first minimum of the objective over the tests
// that send at least one sample each way; best[node] = test index or -1 (the node becomes a leaf).
template <int METHOD>
__global__ void tr_best_kernel(const int* __restrict__ left_cnt, const double* __restrict__ node_sum, const int* __restrict__ class_cnt,
                               int nt, int K, const TrTest* __restrict__ tests, int* __restrict__ best, TrTest* __restrict__ best_test) {
    __shared__ float s_obj[1024];
    __shared__ int s_idx[1024];
    const int node = blockIdx.x, t = threadIdx.x;
    float obj = 3.402823466e38f;
    int idx = 0x7fffffff;
    if (t < nt) {
        if (METHOD == 0) {
            const int* lc = left_cnt + ((size_t)node * (nt + 1) + t) * K;
            const int* cc = class_cnt + (size_t)node * K;
            int nl = 0, n = 0;
            for (int k = 0; k < K; ++k) { nl += lc[k]; n += cc[k]; }
            const int nr = n - nl;
            if (nl != 0 && nr != 0) {
                float el = 0.f, er = 0.f;
                for (int k = 0; k < K; ++k) {
                    float p = __fdiv_rn((float)lc[k], (float)nl);
                    if (p != 0.f) el = (float)((double)el - (double)p * log((double)p));  // `entropy_left -= p * log(p)`
                    p = __fdiv_rn((float)(cc[k] - lc[k]), (float)nr);
                    if (p != 0.f) er = (float)((double)er - (double)p * log((double)p));
                }
                obj = __fadd_rn(__fmul_rn(el, (float)nl), __fmul_rn(er, (float)nr));
                idx = t;
            }
        } else {
            const int nl = left_cnt[(size_t)node * (nt + 1) + t], n = left_cnt[(size_t)node * (nt + 1) + nt];
            const int nr = n - nl;
            if (nl > 0 && nr > 0) {
                const double* sl = node_sum + ((size_t)node * (nt + 1) + t) * TR_V;
                const double* st = node_sum + ((size_t)node * (nt + 1) + nt) * TR_V;
                double o;
                if (METHOD == 1) {
                    const double lx = sl[0], ly = sl[1], lz = sl[2], rx = st[0] - lx, ry = st[1] - ly, rz = st[2] - lz;
                    o = (sl[3] - (lx * lx + ly * ly + lz * lz) / nl) + ((st[3] - sl[3]) - (rx * rx + ry * ry + rz * rz) / nr);
                } else {
                    double dl = 0.0, dr = 0.0;
                    for (int k = 0; k < 6; ++k) { const double a = sl[k], b = st[k] - sl[k]; dl += a * a; dr += b * b; }
                    o = (3.0 * nl - dl / nl) + (3.0 * nr - dr / nr);  // |(cos, sin) x 3|^2 = 3 per sample
                }
                obj = (float)o;
                idx = t;
            }
        }
    }
    s_obj[t] = obj;
    s_idx[t] = idx;
    __syncthreads();
    for (int o = blockDim.x >> 1; o; o >>= 1) {
        if (t < o) {
            const float a = s_obj[t], b = s_obj[t + o];
            const int ia = s_idx[t], ib = s_idx[t + o];
            if (ib != 0x7fffffff && (ia == 0x7fffffff || b < a || (b == a && ib < ia))) { s_obj[t] = b; s_idx[t] = ib; }
        }
        __syncthreads();
    }
    if (t == 0) {
        const int b = s_idx[0] == 0x7fffffff ? -1 : s_idx[0];
        best[node] = b;
        if (b >= 0) best_test[node] = tests[(size_t)node * (nt + 1) + b];
    }
}

// HFTrain::apply_tests_to_train_samples (HFTrain.cpp:999-1047), first half: which way every sample of a split node goes, and
// how many of a chunk go left.
__global__ void tr_apply_kernel(const float* __restrict__ feat, int F, const int* __restrict__ order, const TrChunk* __restrict__ chunks,
                                const int* __restrict__ best, const TrTest* __restrict__ best_test, uint8_t* __restrict__ go_left,
                                int* __restrict__ chunk_left) {
    __shared__ int s_n;
    const TrChunk ch = chunks[blockIdx.x];
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    int n = 0;
    if (best[ch.node] >= 0) {
        const TrTest q = best_test[ch.node];
        for (int s = threadIdx.x; s < ch.count; s += blockDim.x) {
            const float* row = feat + (size_t)order[ch.first + s] * F;
            const bool l = tr_value(q, row) < q.thr;
            go_left[ch.first + s] = l;
            n += l;
        }
    }
    atomicAdd(&s_n, n);
    __syncthreads();
    if (threadIdx.x == 0) chunk_left[blockIdx.x] = s_n;
}

// Second half: the stable partition.  child_next[node][side] = first position of the child's segment in the next level's order
// array, or -1 when the child is a leaf / the node did not split; child_id[node][side] = tree node the sample ends in then
// (the node itself when it did not split).  One CTA per chunk, positions by a block-wide scan of the flags.
__global__ void __launch_bounds__(TR_CHUNK)
tr_scatter_kernel(const int* __restrict__ order, const TrChunk* __restrict__ chunks, const int* __restrict__ best,
                  const uint8_t* __restrict__ go_left, const int* __restrict__ child_next, const int* __restrict__ child_id,
                  int* __restrict__ order_out, int* __restrict__ final_node) {
    __shared__ int s_warp[TR_CHUNK / 32];
    const TrChunk ch = chunks[blockIdx.x];
    const int s = threadIdx.x, lane = s & 31, warp = s >> 5;
    const bool have = s < ch.count;
    const int smp = have ? order[ch.first + s] : 0;
    const bool split = best[ch.node] >= 0;
    const int l = have && split && go_left[ch.first + s] ? 1 : 0;
    int incl = l;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int v = s_warp[lane], iv = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, iv, o);
            if (lane >= o) iv += u;
        }
        s_warp[lane] = iv - v;
    }
    __syncthreads();
    if (!have) return;
    const int lefts_before = s_warp[warp] + incl - l;  // lefts among samples 0 .. s-1 of the chunk
    if (!split) { final_node[smp] = child_id[2 * ch.node]; return; }
    const int side = l ? 0 : 1;
    const int nxt = child_next[2 * ch.node + side];
    if (nxt < 0) { final_node[smp] = child_id[2 * ch.node + side]; return; }
    const int rank = l ? ch.left_before + lefts_before : ch.right_before + (s - lefts_before);
    order_out[nxt + rank] = smp;
}

// per-sample regression vectors: x, y, z, cos / sin of yaw, pitch, roll in double (HFTrain.cpp:591-593, 824-829)
__global__ void tr_vectors_kernel(const float* __restrict__ dof, int n, double* __restrict__ vec) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* d = dof + (size_t)i * 6;
    double* v = vec + (size_t)i * 9;
    v[0] = d[3]; v[1] = d[4]; v[2] = d[5];
    v[3] = cos((double)d[0]); v[4] = sin((double)d[0]);
    v[5] = cos((double)d[1]); v[6] = sin((double)d[1]);
    v[7] = cos((double)d[2]); v[8] = sin((double)d[2]);
}

}  // namespace hf6d
