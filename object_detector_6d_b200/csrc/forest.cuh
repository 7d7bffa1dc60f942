// Stage TRAVERSE: every processed patch through every owned Hough tree to its leaf.
//
// Replaces HFTest::get_leaf (HoughForest/src/HFTest.cpp:144-163): a recursive pointer chase over heap TreeNodes,
// called P'*T times per frame from inside the OpenMP loop of test_image (HFTest.cpp:612-628).
//
// B200 design
//  * forest: internal nodes only, breadth-first per tree, TWO levels per 48-byte record {3 tests, 4 grandchild
//    entries} (model.hpp::PackedRecord); an entry is a record index or ~leaf.  The descent is a chain of dependent
//    L2 loads, so halving its length (D/2 record fetches of three 16-byte loads issued together) is what shortens it;
//    leaf payload is never touched.  measure_mode 1 is encoded as f2 = F, a shared-memory slot that holds 0.0f
//    (x - 0 == x), so the test is branch-free:  val = f[f1] - f[f2];  val < thr -> left  (strict <, NaN -> right, as
//    the reference).
//  * features: fp32 rows are streamed through shared memory exactly once (the stage's algorithmic HBM bytes) by the
//    TMA engine's 1-D bulk copy into a ring of 8-row buffers with full/empty mbarriers: a producer warp keeps every
//    free buffer loading while each consumer warp descends the rows of its own buffer.
//  * the first records of every tree (4 record levels = 8 tree levels when they fit 16 KB) are copied to shared memory
//    once per CTA, which takes the top of every descent off the L2 round trip.
//  * warp-cooperative descent: 16 lanes of a warp walk the SAME tree for 16 different patches, so the top levels are
//    broadcast loads and each lane's feature reads stay in its own shared-memory row.  A thread interleaves up to 4
//    trees (independent dependent-load chains) to hide L2 latency.
#pragma once
#include <algorithm>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace hf6d {

constexpr int TRV_ROWS = 8;                     // rows per ring buffer: a consumer warp covers 8 rows x 4 trees
constexpr int TRV_SLOTS = 32 / TRV_ROWS;        // trees descended concurrently by one warp per interleave step
constexpr int TRV_MAX_BUFS = 15;                // ring buffers = consumer warps (one more warp produces)
constexpr int TRV_CACHE_BYTES = 16 * 1024;      // shared-memory copy of the top records of every tree

struct TraversePlan {
    int n_bufs;      // ring depth
    int n_cache;     // records cached per tree
    size_t smem;     // dynamic shared memory
};
// Shared-memory row of one patch: F features + a constant-zero slot (fp32: 4 floats, fp16: 8 halves -- rows stay 16-byte
// aligned for the bulk copies).
inline int traverse_pitch(int F, int elem_bytes) { return F + 16 / elem_bytes; }
inline TraversePlan traverse_plan(int F, int T, int elem_bytes = 4) {
    TraversePlan p;
    const size_t buf = (size_t)TRV_ROWS * traverse_pitch(F, elem_bytes) * elem_bytes;
    p.n_cache = (int)std::min<size_t>(85, TRV_CACHE_BYTES / (sizeof(PackedRecord) * (size_t)std::max(T, 1)));  // 85 = 4 record levels
    const size_t cache = (size_t)p.n_cache * T * sizeof(PackedRecord);
    const size_t avail = 220 * 1024 - cache - 512;
    p.n_bufs = (int)std::max<size_t>(1, std::min<size_t>(TRV_MAX_BUFS, avail / buf));
    p.smem = cache + (size_t)p.n_bufs * buf + 2 * TRV_MAX_BUFS * 8 + 16;
    return p;
}

// One CTA per SM: warp n_bufs is the producer (it keeps every free ring buffer loading: 8 feature rows per buffer, one
// 1-D TMA bulk copy per row, completion on the buffer's `full` mbarrier); warps 0..n_bufs-1 each own one ring buffer:
// wait for it, descend its 8 rows through the owned trees, hand it back through the `empty` mbarrier.  With ~56 rows
// resident per SM the loads never pause for the descents and the descents never wait for a whole tile.
// FT = float: the reference's fp32 features; FT = __half: feature storage 1 (the feature layer wrote fp16; the value compared
// is the fp32 widening of the stored half, i.e. exactly what hf6d_fetch returns for the row).
__device__ __forceinline__ float trv_value(float v) { return v; }
__device__ __forceinline__ float trv_value(__half v) { return __half2float(v); }
template <int NCH, class FT = float>
__global__ void __launch_bounds__((TRV_MAX_BUFS + 1) * 32, 1)
traverse_kernel(const FT* __restrict__ features, DevForest f, const int* __restrict__ counts,
                int* __restrict__ leaf_ord, int shard_rank, int shard_world, int n_bufs, int n_cache, int n_recs,
                int reverse, PatchShard pshard) {
    extern __shared__ __align__(16) uint8_t trv_smem[];
    constexpr int ZPAD = 16 / (int)sizeof(FT);
    const int F = f.F, pitch = F + ZPAD;
    PackedRecord* cache = reinterpret_cast<PackedRecord*>(trv_smem);  // [T][n_cache]
    FT* ring = reinterpret_cast<FT*>(trv_smem + (size_t)n_cache * f.T * sizeof(PackedRecord));
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)n_bufs * TRV_ROWS * pitch);
    uint64_t* empty = full + TRV_MAX_BUFS;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int p_lo, Pp;  // patch sharding: rows [p_lo, Pp) (p_lo is a multiple of 128, so of TRV_ROWS); tiles are counted from p_lo
    patch_shard_range(counts[1], pshard, p_lo, Pp);
    const int tiles = (Pp - p_lo + TRV_ROWS - 1) / TRV_ROWS;
    const int n_owned = (f.T - shard_rank + shard_world - 1) / shard_world;

    for (int i = threadIdx.x; i < n_bufs * TRV_ROWS; i += blockDim.x) {
#pragma unroll
        for (int k = 0; k < ZPAD; ++k) ring[(size_t)i * pitch + F + k] = FT(0.0f);  // the constant-zero feature slot
    }
    {   // top records of every tree (16-byte words; a tree shorter than n_cache just caches a neighbour's records, unused)
        const uint4* src = reinterpret_cast<const uint4*>(f.recs);
        uint4* dst = reinterpret_cast<uint4*>(cache);
        for (int i = threadIdx.x; i < f.T * n_cache * 3; i += blockDim.x) {
            const int t = i / (n_cache * 3), w = i % (n_cache * 3);
            const int root = __ldg(f.root + t);
            const long long rec = (long long)max(root, 0) + w / 3;
            dst[i] = rec < n_recs ? __ldg(src + rec * 3 + w % 3) : make_uint4(0u, 0u, 0u, 0u);
        }
    }
    if (threadIdx.x == 0) {
        for (int b = 0; b < n_bufs; ++b) {
            ptx::mbar_init(&full[b], 1);
            ptx::mbar_init(&empty[b], 1);
        }
        ptx::fence_barrier_init();
    }
    __syncthreads();

    if (warp == n_bufs) {
        // ------------------------------------------------------------ producer
        int b = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            ptx::mbar_wait(&empty[b], phase ^ 1);
            const int p0 = p_lo + (reverse ? tiles - 1 - tile : tile) * TRV_ROWS;  // last rows first: see the consumers
            const int nrows = min(TRV_ROWS, Pp - p0);
            if (lane == 0) ptx::mbar_arrive_expect_tx(&full[b], (uint32_t)nrows * F * (uint32_t)sizeof(FT));
            __syncwarp();
            if (lane < nrows)
                ptx::bulk_load_1d(ring + ((size_t)b * TRV_ROWS + lane) * pitch, features + (size_t)(p0 + lane) * F, F * (int)sizeof(FT), &full[b]);
            if (++b == n_bufs) { b = 0; phase ^= 1; }
        }
    } else if (warp < n_bufs) {
        // ------------------------------------------------------------ consumers: warp w owns ring buffer w
        const int row = lane % TRV_ROWS, slot = lane / TRV_ROWS;
        const FT* my = ring + ((size_t)warp * TRV_ROWS + row) * pitch;
        uint32_t phase = 0;
        for (int tile = blockIdx.x + warp * gridDim.x; tile < tiles; tile += n_bufs * gridDim.x) {
            // The rows are walked from the END of the matrix: the feature layer wrote them in ascending order just before
            // this kernel, so the tail of the 227 MB is what the 126 MB L2 still holds -- reading it first turns those
            // rows into L2 hits instead of letting the head's misses evict them.
            const int p0 = p_lo + (reverse ? tiles - 1 - tile : tile) * TRV_ROWS;
            const int nrows = min(TRV_ROWS, Pp - p0);
            // trees this rank does not own: mark, so that a max-reduce across ranks assembles the full table
            if (shard_world > 1) {
                for (int i = lane; i < nrows * f.T; i += 32)
                    if ((i % f.T) % shard_world != shard_rank) leaf_ord[(size_t)p0 * f.T + i] = -1;
            }
            ptx::mbar_wait(&full[warp], phase);
            phase ^= 1;
            if (row < nrows) {
                // lane group `slot` owns the owned-tree indices slot, slot + 4, ...; NCH of them descend interleaved
                for (int k0 = slot; k0 < n_owned; k0 += TRV_SLOTS * NCH) {
                    int e[NCH], tr[NCH], base[NCH];
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        const int k = k0 + c * TRV_SLOTS;
                        tr[c] = k < n_owned ? shard_rank + k * shard_world : -1;
                        e[c] = tr[c] >= 0 ? __ldg(f.root + tr[c]) : -1;
                        base[c] = e[c];  // first record of the tree (breadth-first numbering starts at the root)
                    }
                    bool any = true;
                    while (any) {
                        uint4 ra[NCH], rb[NCH], rc[NCH];
#pragma unroll
                        for (int c = 0; c < NCH; ++c)
                            if (e[c] >= 0) {
                                const int local = e[c] - base[c];
                                if (local < n_cache) {
                                    const uint4* rp = reinterpret_cast<const uint4*>(cache + (size_t)tr[c] * n_cache + local);
                                    ra[c] = rp[0]; rb[c] = rp[1]; rc[c] = rp[2];
                                } else {
                                    const uint4* rp = reinterpret_cast<const uint4*>(f.recs + e[c]);
                                    ra[c] = __ldg(rp);      // f1f2[0..2], thr[0]
                                    rb[c] = __ldg(rp + 1);  // thr[1..2], next[0..1]
                                    rc[c] = __ldg(rp + 2);  // next[2..3]
                                }
                            }
                        any = false;
#pragma unroll
                        for (int c = 0; c < NCH; ++c)
                            if (e[c] >= 0) {
                                const float v0 = __fsub_rn(trv_value(my[ra[c].x & 0xFFFFu]), trv_value(my[ra[c].x >> 16]));
                                const bool right0 = !(v0 < __uint_as_float(ra[c].w));
                                const unsigned t1 = right0 ? ra[c].z : ra[c].y;
                                const float thr1 = __uint_as_float(right0 ? rb[c].y : rb[c].x);
                                const float v1 = __fsub_rn(trv_value(my[t1 & 0xFFFFu]), trv_value(my[t1 >> 16]));
                                const bool right1 = !(v1 < thr1);
                                e[c] = right0 ? (int)(right1 ? rc[c].y : rc[c].x) : (int)(right1 ? rb[c].w : rb[c].z);
                                any |= e[c] >= 0;
                            }
                    }
#pragma unroll
                    for (int c = 0; c < NCH; ++c)
                        if (tr[c] >= 0) leaf_ord[(size_t)(p0 + row) * f.T + tr[c]] = (~e[c]) - __ldg(f.leaf_base + tr[c]);
                }
            }
            __syncwarp();  // every lane is done reading the buffer
            if (lane == 0) ptx::mbar_arrive(&empty[warp]);
        }
    }
}

}  // namespace hf6d
