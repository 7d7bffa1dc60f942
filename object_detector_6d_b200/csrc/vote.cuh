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

// Yaw/pitch accumulators of one slot (class, centre rank) cover only the bins that can hold a vote AND can influence a
// kept peak: rows = yaw bins [y0, y0+ny), columns = pitch bins [p0, p0+np) of the 720x720 map (model bounding box of
// all vote copies, intersected with [180,540] +- nms/2 +- blur/2).  Everything outside is exactly zero.
struct PoseRegion {
    int y0, ny, p0, np;
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

struct SharedCentres {
    hf6d_centre_list ctr[HF6D_MAX_CLASSES];
    uint8_t act[HF6D_MAX_CLASSES][HF6D_MAX_CENTRES];
};

__device__ __forceinline__ void load_centres(SharedCentres& sc, const CentreTable& ct, int K) {
    for (int i = threadIdx.x; i < K * (int)(sizeof(hf6d_centre_list) / 4); i += blockDim.x)
        reinterpret_cast<int*>(sc.ctr)[i] = reinterpret_cast<const int*>(ct.lists)[i];
    for (int i = threadIdx.x; i < K * HF6D_MAX_CENTRES; i += blockDim.x) (&sc.act[0][0])[i] = ct.active[i];
    __syncthreads();
}

// bit k set <=> (u, v) lies in the window of active centre k of class c
__device__ __forceinline__ unsigned window_mask(const SharedCentres& sc, int c, int nctr, int half_win, int uu, int vv) {
    unsigned m = 0;
    for (int k = 0; k < nctr; ++k) {
        const int ccx = sc.ctr[c].c[k].x, ccy = sc.ctr[c].c[k].y;
        const bool hit = sc.act[c][k] && vv >= ccy - half_win && vv < ccy + half_win && uu >= ccx - half_win &&
                         uu < ccx + half_win;
        m |= (unsigned)hit << k;
    }
    return m;
}

// Coarse lookup in shared memory: the image (plus a margin of one window) is cut into cells of half_win pixels; a cell
// holds the OR of the centres whose window touches it.  Most votes hit an empty cell and skip the exact window tests.
struct CellGrid {
    int cell, gx, gy, x0, y0;  // cell size, grid size, pixel of cell (0,0)
};
__host__ __device__ __forceinline__ CellGrid make_cell_grid(int W, int H, int half_win) {
    CellGrid cg;
    cg.cell = half_win > 0 ? half_win : 1;
    cg.x0 = -2 * cg.cell;
    cg.y0 = -2 * cg.cell;
    cg.gx = (W + 4 * cg.cell + cg.cell - 1) / cg.cell;
    cg.gy = (H + 4 * cg.cell + cg.cell - 1) / cg.cell;
    return cg;
}
inline size_t cell_grid_bytes(int W, int H, int half_win, int K) {
    const CellGrid cg = make_cell_grid(W, H, half_win);
    return ((size_t)cg.gx * cg.gy * K * sizeof(uint16_t) + 7) / 8 * 8;
}

// Pass A.1: enumerate the cast votes again (same arithmetic as vote_kernel, so the same pixels).  For every vote that
// falls in the window of an active centre ("entry" of the reference's center_leaf_map):
//   * cnt[slot][group] += 1                  -- everything the entry contributes to the yaw/pitch and roll maps depends
//                                               only on its leaf, so those maps are built later from these counts;
//   * z histogram of the slot: the WINDOW pixel's depth stands in for the patch centre and every vote of the leaf is
//     re-projected (HFTest.cpp:766-775), so this part is per entry.
__global__ void __launch_bounds__(VOTE_THREADS)
window_count_kernel(DevForest f, FrameGeom g, const __grid_constant__ ObjectSwitches sw, const int* __restrict__ locs,
                    const uint16_t* __restrict__ depth, const int* __restrict__ leaf_ord,
                    const int* __restrict__ counts, CentreTable ct, int half_win, int n_groups,
                    unsigned* __restrict__ cnt /*[S][n_groups]*/, unsigned long long* __restrict__ zacc /*[S][Z_BINS]*/) {
    __shared__ SharedCentres sc;
    extern __shared__ uint16_t s_cells[];  // [K][gy][gx]
    load_centres(sc, ct, f.K);
    const CellGrid cg = make_cell_grid(g.W, g.H, half_win);
    const int cells_per_class = cg.gx * cg.gy;
    for (int i = threadIdx.x; i < f.K * cells_per_class; i += blockDim.x) s_cells[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < f.K * HF6D_MAX_CENTRES; i += blockDim.x) {
        const int c = i / HF6D_MAX_CENTRES, k = i % HF6D_MAX_CENTRES;
        if (k >= sc.ctr[c].n || !sc.act[c][k]) continue;
        const int x_lo = sc.ctr[c].c[k].x - half_win, y_lo = sc.ctr[c].c[k].y - half_win;
        for (int y = y_lo; y < y_lo + 2 * half_win + cg.cell; y += cg.cell)
            for (int x = x_lo; x < x_lo + 2 * half_win + cg.cell; x += cg.cell) {
                const int yy = min(y, y_lo + 2 * half_win - 1), xx = min(x, x_lo + 2 * half_win - 1);
                const int cyi = (yy - cg.y0) / cg.cell, cxi = (xx - cg.x0) / cg.cell;
                if (cxi < 0 || cxi >= cg.gx || cyi < 0 || cyi >= cg.gy) continue;
                // 16-bit atomicOr through the containing 32-bit word
                const int idx = c * cells_per_class + cyi * cg.gx + cxi;
                atomicOr(reinterpret_cast<unsigned*>(s_cells) + (idx >> 1), (1u << k) << ((idx & 1) * 16));
            }
    }
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
                const int nctr = sc.ctr[c].n;
                if (nctr == 0) continue;
                const unsigned long long w = (unsigned)grp.y;
                for (int v0 = 0; v0 < grp.w; v0 += 32) {
                    const int v = v0 + lane;
                    int uu = INT_MIN, vv = INT_MIN;
                    unsigned mask = 0;
                    if (v < grp.w) {
                        const int vi = grp.z + v;
                        project(g, __fadd_rn(__ldg(f.ox + vi), tx), __fadd_rn(__ldg(f.oy + vi), ty),
                                __fadd_rn(__ldg(f.oz + vi), tz), uu, vv);
                        const int cxi = uu >= cg.x0 ? (uu - cg.x0) / cg.cell : -1, cyi = vv >= cg.y0 ? (vv - cg.y0) / cg.cell : -1;
                        unsigned cand = 0;
                        if (cxi >= 0 && cxi < cg.gx && cyi >= 0 && cyi < cg.gy)
                            cand = s_cells[c * cells_per_class + cyi * cg.gx + cxi];
                        while (cand) {  // exact test for the few centres whose window touches the cell
                            const int k = __ffs(cand) - 1;
                            cand &= cand - 1;
                            const int ccx = sc.ctr[c].c[k].x, ccy = sc.ctr[c].c[k].y;
                            const bool hit = vv >= ccy - half_win && vv < ccy + half_win && uu >= ccx - half_win && uu < ccx + half_win;
                            mask |= (unsigned)hit << k;
                        }
                    }
                    unsigned slots = __reduce_or_sync(0xffffffffu, mask);
                    while (slots) {  // warp-uniform: every centre window that received an entry from this chunk
                        const int k = __ffs(slots) - 1;
                        slots &= slots - 1;
                        unsigned hm = __ballot_sync(0xffffffffu, (mask >> k) & 1u);
                        const size_t s = (size_t)c * HF6D_MAX_CENTRES + k;
                        if (lane == 0) atomicAdd(cnt + s * n_groups + gi, (unsigned)__popc(hm));
                        unsigned long long* zs = zacc + s * HF6D_Z_BINS;
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
            }
        }
    }
}

// Pass A.2: yaw/pitch maps from the (slot, group) entry counts: every entry re-walks all votes of its leaf
// (HFTest.cpp:763-791), so a group adds  count * w  at each of its votes' bins (and their +-360 wrap copies).
// A warp reads 32 consecutive counters, then G lanes take one non-zero (slot, group) pair each.
constexpr int TABLE_THREADS = 256;
template <int G>
__global__ void __launch_bounds__(TABLE_THREADS)
yawpitch_from_counts_kernel(DevForest f, const unsigned* __restrict__ cnt, int n_groups, const uint8_t* __restrict__ active,
                            int S, PoseRegion reg, unsigned long long* __restrict__ ypacc /*[S][ny][np]*/) {
    const int lane = threadIdx.x & 31, sub = lane % G, part = lane / G;
    constexpr int PARTS = 32 / G;
    const size_t yp_slot = (size_t)reg.ny * reg.np;
    const int per_slot = (n_groups + 31) / 32;  // warp tasks per slot
    const long long n_tasks = (long long)S * per_slot;
    const long long warp_id = ((long long)blockIdx.x * TABLE_THREADS + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * TABLE_THREADS) >> 5;
    for (long long t = warp_id; t < n_tasks; t += n_warps) {
        const int s = (int)(t / per_slot);
        if (!active[s]) continue;
        const int g0 = (int)(t % per_slot) * 32;
        const int gi_l = g0 + lane;
        const unsigned my = gi_l < n_groups ? cnt[(size_t)s * n_groups + gi_l] : 0u;
        unsigned nz = __ballot_sync(0xffffffffu, my != 0);
        unsigned long long* yps = ypacc + (size_t)s * yp_slot;
        while (nz) {
            // the PARTS sub-groups of the warp each take one non-zero pair
            int src = -1;
            unsigned m = nz;
#pragma unroll
            for (int pp = 0; pp < PARTS; ++pp) {
                const int b = m ? __ffs(m) - 1 : -1;
                if (b >= 0) m &= m - 1;
                if (pp == part) src = b;
            }
            nz = m;
            const unsigned c_hits = __shfl_sync(0xffffffffu, my, src < 0 ? 0 : src);
            if (src < 0) continue;
            const int4 grp = __ldg(reinterpret_cast<const int4*>(f.groups) + g0 + src);
            const unsigned long long wk = (unsigned long long)(unsigned)grp.y * c_hits;
            for (int q = sub; q < grp.w; q += G) {
                const short4 bn = __ldg(f.bins + grp.z + q);
                const int yaw = bn.x, pit = bn.y;
                const int sy = yaw < 0 ? -1 : 1, sp = pit < 0 ? -1 : 1;  // copysign(1, (float)int): sign(0) = +1
#pragma unroll
                for (int k1 = 0; k1 < 2; ++k1)
#pragma unroll
                    for (int k2 = 0; k2 < 2; ++k2) {
                        // outside [0,720) the reference writes out of bounds; outside the region no kept peak sees the bin
                        const int ry = yaw - sy * k1 * 360 + 360 - reg.y0, rp = pit - sp * k2 * 360 + 360 - reg.p0;
                        if (ry < 0 || ry >= reg.ny || rp < 0 || rp >= reg.np) continue;
                        atomicAdd(yps + (size_t)ry * reg.np + rp, wk);
                    }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ pose pass B
// Roll histograms (HFTest.cpp:857-872): for a yaw/pitch peak (Yp, Pp) every roll_nodemap entry inside the +-17 box
// re-walks all votes of its leaf.  roll_nodemap holds one entry per (window entry, vote) pair (k1 = k2 = 0 copy), so a
// (slot, group) pair contributes  count * (#votes of the leaf in the box) * w  to the roll bins of every vote of the leaf.
struct PeakTable {
    const int* n_peaks;  // [S]
    const int* peak_yx;  // [S][max_peaks][2] = (row = yaw bin, col = pitch bin)
    int max_peaks;
};

template <int G>
__global__ void __launch_bounds__(TABLE_THREADS)
roll_from_counts_kernel(DevForest f, const unsigned* __restrict__ cnt, int n_groups, int S, PeakTable pk, int half_box,
                        unsigned long long* __restrict__ racc /*[S][max_peaks][POSE_BINS]*/) {
    const int lane = threadIdx.x & 31, sub = lane % G, part = lane / G;
    constexpr int PARTS = 32 / G;
    const unsigned gmask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (part * G));
    const int per_slot = (n_groups + 31) / 32;
    const long long n_tasks = (long long)S * per_slot;
    const long long warp_id = ((long long)blockIdx.x * TABLE_THREADS + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * TABLE_THREADS) >> 5;
    for (long long t = warp_id; t < n_tasks; t += n_warps) {
        const int s = (int)(t / per_slot);
        const int np = __ldg(pk.n_peaks + s);
        if (np == 0) continue;
        const int g0 = (int)(t % per_slot) * 32;
        const int gi_l = g0 + lane;
        const unsigned my = gi_l < n_groups ? cnt[(size_t)s * n_groups + gi_l] : 0u;
        unsigned nz = __ballot_sync(0xffffffffu, my != 0);
        while (nz) {
            int src = -1;
            unsigned m = nz;
#pragma unroll
            for (int pp = 0; pp < PARTS; ++pp) {
                const int b = m ? __ffs(m) - 1 : -1;
                if (b >= 0) m &= m - 1;
                if (pp == part) src = b;
            }
            nz = m;
            const unsigned c_hits = __shfl_sync(0xffffffffu, my, src < 0 ? 0 : src);
            if (src < 0) continue;
            const int4 grp = __ldg(reinterpret_cast<const int4*>(f.groups) + g0 + src);
            const int passes = (grp.w + G - 1) / G;
            for (int p = 0; p < np; ++p) {
                const int Yp = __ldg(pk.peak_yx + (s * pk.max_peaks + p) * 2), Pp = __ldg(pk.peak_yx + (s * pk.max_peaks + p) * 2 + 1);
                int inbox = 0;
                for (int ps = 0; ps < passes; ++ps) {
                    const int q = ps * G + sub;
                    bool in = false;
                    if (q < grp.w) {
                        const short4 bn = __ldg(f.bins + grp.z + q);
                        const int Y = (int)bn.x + 360, P0 = (int)bn.y + 360;
                        in = Y >= Yp - half_box && Y < Yp + half_box && P0 >= Pp - half_box && P0 < Pp + half_box;
                    }
                    inbox += __popc(__ballot_sync(gmask, in));
                }
                if (inbox == 0) continue;
                const unsigned long long wk = (unsigned long long)(unsigned)grp.y * c_hits * (unsigned long long)inbox;
                unsigned long long* rs = racc + ((size_t)s * pk.max_peaks + p) * HF6D_POSE_BINS;
                for (int q = sub; q < grp.w; q += G) {
                    const int r = __ldg(f.bins + grp.z + q).z;
                    const int b0 = r + 360, b1 = r < 0 ? r + 720 : r;
                    if (b0 >= 0 && b0 < HF6D_POSE_BINS) atomicAdd(rs + b0, wk);
                    if (b1 >= 0 && b1 < HF6D_POSE_BINS) atomicAdd(rs + b1, wk);
                }
            }
        }
    }
}

}  // namespace hf6d
