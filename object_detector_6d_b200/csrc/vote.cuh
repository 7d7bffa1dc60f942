// Stage VOTE (and the vote re-enumeration passes of stage POSE).
//
// Replaces HFTest::detect's voting loop (HoughForest/src/HFTest.cpp:184-214: per vote 6 libm calls, three 4x4 matrix
// products, a float map increment and a boost::unordered_map push_back), the per-batch per-thread map allocation and
// merge (HFTest.cpp:605-654) and, for the pose stage, the hash-map walks of HFTest.cpp:757-802 and :857-872.
//
// B200 design
//  * the patch-independent part of the vote geometry, R(yaw,pitch,roll)*(-x,-y,-z), is folded into the forest at load
//    time (model.hpp), so a vote is 12 bytes (3 floats) and casting it is 3 adds, 2 divides and a truncation -- the
//    same fp32 operations, in the same order, as the reference performs after its matrix products.
//  * weights are Q16 integers accumulated with 64-bit integer atomics: maps are bit-reproducible whatever the atomics
//    order and whatever the number of GPUs the trees are sharded over (the reference's float maps depend on the OpenMP
//    schedule, HFTest.cpp:645-654).
//  * the centre->leaf back-map (center_leaf_map) is never materialised: the pose stage re-enumerates the votes (same
//    arithmetic, so the same pixels) and keeps only those that fall in a centre window.  The reference's n^2
//    multiplicity (a leaf is pushed once per vote and every entry re-walks all the leaf's votes) is reproduced by
//    multiplying with the per-window hit counts.
//  * warp-cooperative: a warp owns 32 (patch, tree) items; lane l resolves item l's leaf and patch geometry, then the
//    warp walks the items' vote lists together, lanes striding over consecutive votes (coalesced 12-byte SoA reads).
#pragma once
#include "common.cuh"

namespace hf6d {

struct ObjectSwitches {
    uint8_t should_detect[HF6D_MAX_CLASSES];
};

// One accumulator slot per (class, centre rank).
struct PoseRegion {
    int lo, size;  // yaw/pitch accumulators cover global bins [lo, lo+size) in both dimensions
};

struct ItemCtx {
    int gb, ge;       // vote-group range of the item's leaf
    float tx, ty, tz; // back-projected patch centre (HFTest.cpp:83-88)
};

__device__ __forceinline__ ItemCtx load_item(const DevForest& f, const FrameGeom& g, const int* __restrict__ locs,
                                             const uint16_t* __restrict__ depth, const int* __restrict__ leaf_ord,
                                             long long item, long long n_items) {
    ItemCtx it;
    it.gb = it.ge = 0;
    it.tx = it.ty = it.tz = 0.f;
    if (item < n_items) {
        const int p = (int)(item / f.T), t = (int)(item % f.T);
        const int ord = leaf_ord[item];
        if (ord >= 0) {
            const int gl = __ldg(f.leaf_base + t) + ord;
            it.gb = __ldg(f.group_off + gl);
            it.ge = __ldg(f.group_off + gl + 1);
            if (it.ge > it.gb) {
                const int px = locs[2 * p], py = locs[2 * p + 1];
                const float z = __fdiv_rn((float)depth[(size_t)py * g.W + px], 1000.0f);  // HFTest.cpp:628
                it.tz = z;
                it.tx = __fdiv_rn(__fmul_rn(__fsub_rn((float)px, g.cx), z), g.fx);
                it.ty = __fdiv_rn(__fmul_rn(__fsub_rn((float)py, g.cy), z), g.fy);
            }
        }
    }
    return it;
}

// Point3DToImage, HFTest.cpp:21-37
__device__ __forceinline__ void project(const FrameGeom& g, float x, float y, float z, int& u, int& v) {
    if (z == 0.f) { u = 0; v = 0; return; }
    u = f2i_x86(__fadd_rn(__fadd_rn(__fmul_rn(__fdiv_rn(x, z), g.fx), g.cx), 0.5f));
    v = f2i_x86(__fadd_rn(__fadd_rn(__fmul_rn(__fdiv_rn(y, z), g.fy), g.cy), 0.5f));
}

constexpr int VOTE_THREADS = 256;

__global__ void __launch_bounds__(VOTE_THREADS)
vote_kernel(DevForest f, FrameGeom g, const __grid_constant__ ObjectSwitches sw, const int* __restrict__ locs,
            const uint16_t* __restrict__ depth, const int* __restrict__ leaf_ord, const int* __restrict__ counts,
            unsigned long long* __restrict__ maps) {
    const int lane = threadIdx.x & 31;
    const long long n_items = (long long)counts[1] * f.T;
    const long long warp0 = ((long long)blockIdx.x * (VOTE_THREADS >> 5) + (threadIdx.x >> 5)) * 32;
    const long long stride = (long long)gridDim.x * VOTE_THREADS;
    for (long long base = warp0; base < n_items; base += stride) {
        const ItemCtx it = load_item(f, g, locs, depth, leaf_ord, base + lane, n_items);
        unsigned todo = __ballot_sync(0xffffffffu, it.ge > it.gb);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int gb = __shfl_sync(0xffffffffu, it.gb, src), ge = __shfl_sync(0xffffffffu, it.ge, src);
            const float tx = __shfl_sync(0xffffffffu, it.tx, src), ty = __shfl_sync(0xffffffffu, it.ty, src);
            const float tz = __shfl_sync(0xffffffffu, it.tz, src);
            for (int gi = gb; gi < ge; ++gi) {
                const int4 grp = __ldg(reinterpret_cast<const int4*>(f.groups) + gi);  // cls, w, vbeg, vcnt
                if (!sw.should_detect[grp.x]) continue;
                unsigned long long* map = maps + (size_t)grp.x * g.H * g.W;
                for (int v = lane; v < grp.w; v += 32) {
                    const int vi = grp.z + v;
                    int uu, vv;
                    project(g, __fadd_rn(__ldg(f.ox + vi), tx), __fadd_rn(__ldg(f.oy + vi), ty),
                            __fadd_rn(__ldg(f.oz + vi), tz), uu, vv);
                    if (uu >= 0 && uu < g.W && vv >= 0 && vv < g.H)
                        atomicAdd(map + (size_t)vv * g.W + uu, (unsigned long long)(unsigned)grp.y);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ pose pass A
// For every cast vote that lands in the 40x40 window of a kept centre (HFTest.cpp:757-762) accumulate, for ALL votes
// of that leaf/class: the z histogram (using the WINDOW pixel's depth as if it were the patch centre, :766-775) and the
// yaw/pitch map with its +-360 wrap copies (:778-791).
struct CentreTable {  // device copy of the per-class centre lists
    const hf6d_centre_list* lists;  // [K]
    const uint8_t* active;          // [K][HF6D_MAX_CENTRES]
};

__global__ void __launch_bounds__(VOTE_THREADS)
pose_accum_kernel(DevForest f, FrameGeom g, const __grid_constant__ ObjectSwitches sw, const int* __restrict__ locs,
                  const uint16_t* __restrict__ depth, const int* __restrict__ leaf_ord,
                  const int* __restrict__ counts, CentreTable ct, int half_win, PoseRegion reg,
                  unsigned long long* __restrict__ zacc /*[S][Z_BINS]*/,
                  unsigned long long* __restrict__ ypacc /*[S][size][size]*/) {
    __shared__ hf6d_centre_list s_ctr[HF6D_MAX_CLASSES];
    __shared__ uint8_t s_act[HF6D_MAX_CLASSES][HF6D_MAX_CENTRES];
    for (int i = threadIdx.x; i < f.K * (int)(sizeof(hf6d_centre_list) / 4); i += blockDim.x)
        reinterpret_cast<int*>(s_ctr)[i] = reinterpret_cast<const int*>(ct.lists)[i];
    for (int i = threadIdx.x; i < f.K * HF6D_MAX_CENTRES; i += blockDim.x) (&s_act[0][0])[i] = ct.active[i];
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const long long n_items = (long long)counts[1] * f.T;
    const long long warp0 = ((long long)blockIdx.x * (VOTE_THREADS >> 5) + (threadIdx.x >> 5)) * 32;
    const long long stride = (long long)gridDim.x * VOTE_THREADS;
    const size_t yp_slot = (size_t)reg.size * reg.size;
    for (long long base = warp0; base < n_items; base += stride) {
        const ItemCtx it = load_item(f, g, locs, depth, leaf_ord, base + lane, n_items);
        unsigned todo = __ballot_sync(0xffffffffu, it.ge > it.gb);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int gb = __shfl_sync(0xffffffffu, it.gb, src), ge = __shfl_sync(0xffffffffu, it.ge, src);
            const float tx = __shfl_sync(0xffffffffu, it.tx, src), ty = __shfl_sync(0xffffffffu, it.ty, src);
            const float tz = __shfl_sync(0xffffffffu, it.tz, src);
            for (int gi = gb; gi < ge; ++gi) {
                const int4 grp = __ldg(reinterpret_cast<const int4*>(f.groups) + gi);
                const int c = grp.x;
                if (!sw.should_detect[c]) continue;
                const int nctr = s_ctr[c].n;
                if (nctr == 0) continue;
                const unsigned long long w = (unsigned)grp.y;
                int hits[HF6D_MAX_CENTRES];
#pragma unroll
                for (int k = 0; k < HF6D_MAX_CENTRES; ++k) hits[k] = 0;
                for (int v0 = 0; v0 < grp.w; v0 += 32) {  // entries (cast votes) of this leaf, 32 at a time
                    const int v = v0 + lane;
                    int uu = INT_MIN, vv = INT_MIN;
                    if (v < grp.w) {
                        const int vi = grp.z + v;
                        project(g, __fadd_rn(__ldg(f.ox + vi), tx), __fadd_rn(__ldg(f.oy + vi), ty),
                                __fadd_rn(__ldg(f.oz + vi), tz), uu, vv);
                    }
#pragma unroll
                    for (int k = 0; k < HF6D_MAX_CENTRES; ++k) {
                        if (k >= nctr) break;
                        if (!s_act[c][k]) continue;
                        const int ccx = s_ctr[c].c[k].x, ccy = s_ctr[c].c[k].y;
                        const bool hit = v < grp.w && vv >= ccy - half_win && vv < ccy + half_win &&
                                         uu >= ccx - half_win && uu < ccx + half_win;
                        unsigned hm = __ballot_sync(0xffffffffu, hit);
                        hits[k] += __popc(hm);
                        // z histogram: per entry (window pixel), all votes of the leaf
                        unsigned long long* zs = zacc + (size_t)(c * HF6D_MAX_CENTRES + k) * HF6D_Z_BINS;
                        while (hm) {
                            const int j = __ffs(hm) - 1;
                            hm &= hm - 1;
                            const int col = __shfl_sync(0xffffffffu, uu, j), row = __shfl_sync(0xffffffffu, vv, j);
                            if (row < 0 || row >= g.H || col < 0 || col >= g.W) continue;  // reference reads out of bounds
                            const unsigned d = depth[(size_t)row * g.W + col];
                            if (d == 0) continue;
                            const float zz = __fdiv_rn((float)d, 1000.0f);
                            for (int q = lane; q < grp.w; q += 32) {
                                const int zb = f2i_x86(__fdiv_rn(__fadd_rn(__ldg(f.oz + grp.z + q), zz), 0.01f));
                                if (zb >= 0 && zb < HF6D_Z_BINS) atomicAdd(zs + zb, w);
                            }
                        }
                    }
                }
                // yaw/pitch: every entry re-walks all votes -> weight = hits * w
#pragma unroll
                for (int k = 0; k < HF6D_MAX_CENTRES; ++k) {
                    if (k >= nctr) break;
                    if (hits[k] == 0) continue;
                    const unsigned long long wk = w * (unsigned long long)hits[k];
                    unsigned long long* yps = ypacc + (size_t)(c * HF6D_MAX_CENTRES + k) * yp_slot;
                    for (int q = lane; q < grp.w; q += 32) {
                        const int yaw = __ldg(f.yaw + grp.z + q), pit = __ldg(f.pitch + grp.z + q);
                        const int sy = yaw < 0 ? -1 : 1, sp = pit < 0 ? -1 : 1;  // copysign(1, (float)int): sign(0) = +1
#pragma unroll
                        for (int k1 = 0; k1 < 2; ++k1)
#pragma unroll
                            for (int k2 = 0; k2 < 2; ++k2) {
                                const int Y = yaw - sy * k1 * 360 + 360, Pp = pit - sp * k2 * 360 + 360;
                                if (Y < 0 || Y >= HF6D_POSE_BINS || Pp < 0 || Pp >= HF6D_POSE_BINS) continue;  // ref. would write OOB
                                const int ry = Y - reg.lo, rp = Pp - reg.lo;
                                if (ry < 0 || ry >= reg.size || rp < 0 || rp >= reg.size) continue;  // never read later
                                atomicAdd(yps + (size_t)ry * reg.size + rp, wk);
                            }
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ pose pass B
// Roll histograms (HFTest.cpp:857-872): for a yaw/pitch peak (Yp, Pp) every roll_nodemap entry inside the +-17 box
// re-walks all votes of its leaf.  roll_nodemap holds one entry per (window entry, vote) pair (k1 = k2 = 0 copy), so the
// weight of a leaf is  hits(window) * #votes-in-box * w.
struct PeakTable {
    const int* n_peaks;  // [S]
    const int* peak_yx;  // [S][MAXP][2] = (row = yaw bin, col = pitch bin)
    int max_peaks;
};

__global__ void __launch_bounds__(VOTE_THREADS)
roll_accum_kernel(DevForest f, FrameGeom g, const __grid_constant__ ObjectSwitches sw, const int* __restrict__ locs,
                  const uint16_t* __restrict__ depth, const int* __restrict__ leaf_ord,
                  const int* __restrict__ counts, CentreTable ct, int half_win, PeakTable pk, int half_box,
                  unsigned long long* __restrict__ racc /*[S][max_peaks][POSE_BINS]*/) {
    __shared__ hf6d_centre_list s_ctr[HF6D_MAX_CLASSES];
    __shared__ uint8_t s_act[HF6D_MAX_CLASSES][HF6D_MAX_CENTRES];
    for (int i = threadIdx.x; i < f.K * (int)(sizeof(hf6d_centre_list) / 4); i += blockDim.x)
        reinterpret_cast<int*>(s_ctr)[i] = reinterpret_cast<const int*>(ct.lists)[i];
    for (int i = threadIdx.x; i < f.K * HF6D_MAX_CENTRES; i += blockDim.x) (&s_act[0][0])[i] = ct.active[i];
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const long long n_items = (long long)counts[1] * f.T;
    const long long warp0 = ((long long)blockIdx.x * (VOTE_THREADS >> 5) + (threadIdx.x >> 5)) * 32;
    const long long stride = (long long)gridDim.x * VOTE_THREADS;
    for (long long base = warp0; base < n_items; base += stride) {
        const ItemCtx it = load_item(f, g, locs, depth, leaf_ord, base + lane, n_items);
        unsigned todo = __ballot_sync(0xffffffffu, it.ge > it.gb);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int gb = __shfl_sync(0xffffffffu, it.gb, src), ge = __shfl_sync(0xffffffffu, it.ge, src);
            const float tx = __shfl_sync(0xffffffffu, it.tx, src), ty = __shfl_sync(0xffffffffu, it.ty, src);
            const float tz = __shfl_sync(0xffffffffu, it.tz, src);
            for (int gi = gb; gi < ge; ++gi) {
                const int4 grp = __ldg(reinterpret_cast<const int4*>(f.groups) + gi);
                const int c = grp.x;
                if (!sw.should_detect[c]) continue;
                const int nctr = s_ctr[c].n;
                if (nctr == 0) continue;
                int hits[HF6D_MAX_CENTRES];
#pragma unroll
                for (int k = 0; k < HF6D_MAX_CENTRES; ++k) hits[k] = 0;
                for (int v0 = 0; v0 < grp.w; v0 += 32) {
                    const int v = v0 + lane;
                    int uu = INT_MIN, vv = INT_MIN;
                    if (v < grp.w) {
                        const int vi = grp.z + v;
                        project(g, __fadd_rn(__ldg(f.ox + vi), tx), __fadd_rn(__ldg(f.oy + vi), ty),
                                __fadd_rn(__ldg(f.oz + vi), tz), uu, vv);
                    }
#pragma unroll
                    for (int k = 0; k < HF6D_MAX_CENTRES; ++k) {
                        if (k >= nctr) break;
                        if (!s_act[c][k]) continue;
                        const int ccx = s_ctr[c].c[k].x, ccy = s_ctr[c].c[k].y;
                        const bool hit = v < grp.w && vv >= ccy - half_win && vv < ccy + half_win &&
                                         uu >= ccx - half_win && uu < ccx + half_win;
                        hits[k] += __popc(__ballot_sync(0xffffffffu, hit));
                    }
                }
#pragma unroll
                for (int k = 0; k < HF6D_MAX_CENTRES; ++k) {
                    if (k >= nctr) break;
                    if (hits[k] == 0) continue;
                    const int s = c * HF6D_MAX_CENTRES + k;
                    const int np = pk.n_peaks[s];
                    for (int h = 0; h < np; ++h) {
                        const int Yp = pk.peak_yx[(s * pk.max_peaks + h) * 2], Pp = pk.peak_yx[(s * pk.max_peaks + h) * 2 + 1];
                        int inbox = 0;
                        for (int q0 = 0; q0 < grp.w; q0 += 32) {
                            const int q = q0 + lane;
                            bool in = false;
                            if (q < grp.w) {
                                const int Y = (int)__ldg(f.yaw + grp.z + q) + 360, P0 = (int)__ldg(f.pitch + grp.z + q) + 360;
                                in = Y >= Yp - half_box && Y < Yp + half_box && P0 >= Pp - half_box && P0 < Pp + half_box;
                            }
                            inbox += __popc(__ballot_sync(0xffffffffu, in));
                        }
                        if (inbox == 0) continue;
                        const unsigned long long wk =
                            (unsigned long long)(unsigned)grp.y * (unsigned long long)hits[k] * (unsigned long long)inbox;
                        unsigned long long* rs = racc + ((size_t)s * pk.max_peaks + h) * HF6D_POSE_BINS;
                        for (int q = lane; q < grp.w; q += 32) {
                            const int r = __ldg(f.roll + grp.z + q);
                            const int b0 = r + 360, b1 = r < 0 ? r + 720 : r;
                            if (b0 >= 0 && b0 < HF6D_POSE_BINS) atomicAdd(rs + b0, wk);
                            if (b1 >= 0 && b1 < HF6D_POSE_BINS) atomicAdd(rs + b1, wk);
                        }
                    }
                }
            }
        }
    }
}

}  // namespace hf6d
