// Stage TRAVERSE: every processed patch through every owned Hough tree to its leaf.
//
// Replaces HFTest::get_leaf (HoughForest/src/HFTest.cpp:144-163): a recursive pointer chase over heap TreeNodes,
// called P'*T times per frame from inside the OpenMP loop of test_image (HFTest.cpp:612-628).
//
// B200 design
//  * forest: internal nodes only, breadth-first per tree, 16 bytes each {f1|f2, thr, left, right}; a child entry is a
//    node index or ~leaf, so a descent of depth D costs D 16-byte loads (L1/L2 resident: a few MB) and never touches
//    leaf payload.  measure_mode 1 is encoded as f2 = F, a shared-memory slot that holds 0.0f (x - 0 == x), so the test
//    is branch-free:  val = f[f1] - f[f2];  val < thr -> left  (strict <, NaN -> right, as the reference).
//  * features: fp32 rows are streamed through shared memory exactly once (the stage's algorithmic HBM bytes) by the
//    TMA engine's 1-D bulk copy, 32 rows (102 KB) per tile, completion on an mbarrier; two CTAs per SM alternate
//    between "tile in flight" and "descending", which is the load/compute overlap.
//  * warp-cooperative descent: the 32 lanes of a warp walk the SAME tree for 32 different patches, so the top five
//    levels are broadcast loads (1+2+4+8+16 distinct nodes instead of 160) and each lane's feature reads stay in its own
//    shared-memory row.  A thread interleaves up to 4 trees (independent dependent-load chains) to hide L2 latency.
#pragma once
#include "common.cuh"
#include "ptx_sm100.cuh"

namespace hf6d {

constexpr int TRV_TILE = 32;
constexpr int TRV_WARPS = 4;
constexpr int TRV_THREADS = TRV_WARPS * 32;

inline size_t traverse_smem_bytes(int F) { return (size_t)TRV_TILE * (F + 4) * 4 + 16; }

template <int NCH>
__global__ void __launch_bounds__(TRV_THREADS)
traverse_kernel(const float* __restrict__ features, DevForest f, const int* __restrict__ counts,
                int* __restrict__ leaf_ord, int shard_rank, int shard_world) {
    extern __shared__ __align__(16) uint8_t trv_smem[];
    const int F = f.F, pitch = F + 4;
    float* rows = reinterpret_cast<float*>(trv_smem);
    uint64_t* bar = reinterpret_cast<uint64_t*>(trv_smem + (size_t)TRV_TILE * pitch * 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Pp = counts[1];
    const int tiles = (Pp + TRV_TILE - 1) / TRV_TILE;
    const int n_owned = (f.T - shard_rank + shard_world - 1) / shard_world;

    if (threadIdx.x < TRV_TILE) {
#pragma unroll
        for (int k = 0; k < 4; ++k) rows[threadIdx.x * pitch + F + k] = 0.0f;  // the constant-zero feature slot
    }
    if (threadIdx.x == 0) {
        ptx::mbar_init(bar, 1);
        ptx::fence_barrier_init();
    }
    __syncthreads();

    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int p0 = tile * TRV_TILE;
        const int nrows = min(TRV_TILE, Pp - p0);
        if (warp == 0) {
            if (lane == 0) ptx::mbar_arrive_expect_tx(bar, (uint32_t)nrows * F * 4);
            __syncwarp();
            if (lane < nrows) ptx::bulk_load_1d(rows + lane * pitch, features + (size_t)(p0 + lane) * F, F * 4, bar);
        }
        // trees this rank does not own: mark, so that a max-reduce across ranks assembles the full table
        if (shard_world > 1) {
            for (int i = threadIdx.x; i < nrows * f.T; i += TRV_THREADS) {
                const int t = i % f.T;
                if (t % shard_world != shard_rank) leaf_ord[(size_t)p0 * f.T + i] = -1;
            }
        }
        ptx::mbar_wait(bar, phase);
        phase ^= 1;

        const float* my = rows + lane * pitch;
        if (lane < nrows) {
            // warp w owns the owned-tree indices w, w + TRV_WARPS, ...; NCH of them descend interleaved
            for (int k0 = warp; k0 < n_owned; k0 += TRV_WARPS * NCH) {
                int e[NCH], tr[NCH];
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    const int k = k0 + c * TRV_WARPS;
                    tr[c] = k < n_owned ? shard_rank + k * shard_world : -1;
                    e[c] = tr[c] >= 0 ? __ldg(f.root + tr[c]) : -1;
                }
                bool any = true;
                while (any) {
                    uint4 nd[NCH];
#pragma unroll
                    for (int c = 0; c < NCH; ++c)
                        if (e[c] >= 0) nd[c] = __ldg(reinterpret_cast<const uint4*>(f.nodes) + e[c]);
                    any = false;
#pragma unroll
                    for (int c = 0; c < NCH; ++c)
                        if (e[c] >= 0) {
                            const float val = __fsub_rn(my[nd[c].x & 0xFFFFu], my[nd[c].x >> 16]);
                            e[c] = (val < __uint_as_float(nd[c].y)) ? (int)nd[c].z : (int)nd[c].w;
                            any |= e[c] >= 0;
                        }
                }
#pragma unroll
                for (int c = 0; c < NCH; ++c)
                    if (tr[c] >= 0) leaf_ord[(size_t)(p0 + lane) * f.T + tr[c]] = (~e[c]) - __ldg(f.leaf_base + tr[c]);
            }
        }
        __syncthreads();  // everyone is done with this tile before the next bulk copy lands
    }
}

}  // namespace hf6d
