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

// Point3DToImage, HFTest.cpp:21-37
__device__ __forceinline__ void project(const FrameGeom& g, float x, float y, float z, int& u, int& v) {
    if (z == 0.f) { u = 0; v = 0; return; }
    u = f2i_x86(__fadd_rn(__fadd_rn(__fmul_rn(__fdiv_rn(x, z), g.fx), g.cx), 0.5f));
    v = f2i_x86(__fadd_rn(__fadd_rn(__fmul_rn(__fdiv_rn(y, z), g.fy), g.cy), 0.5f));
}

constexpr int VOTE_THREADS = 256;
constexpr int TABLE_THREADS = 256;  // the kernels that turn (slot, group) counters into yaw/pitch and roll votes
constexpr int VOTE_WARPS = VOTE_THREADS / 32;

// Per-warp staging of one batch of 32 (patch, tree) items: inclusive prefix of their vote counts and what a vote needs
// from its item -- the back-projected patch centre (HFTest.cpp:83-88) and the offset that turns a position in the
// batch's flattened vote list into a vote index.
struct WarpItems {
    int owner[32];   // item whose first vote is vote j0 + i of the current block of 32 votes
    float4 geo[32];  // tx, ty, tz, bits: vbeg - exclusive prefix
};

// Vote-parallel enumeration of every vote the reference casts (HFTest.cpp:177-214): a warp takes 32 (patch, tree) items,
// prefix-sums their leaves' vote counts, and then walks the flattened vote list 32 votes at a time -- one vote per lane
// whatever the votes-per-leaf distribution is.  A lane finds its item from the items' START positions: the items that begin
// inside the current block of 32 votes OR their bit into one mask (REDUX) and leave their index at that position; a vote
// belongs to the item at the highest set bit at or below its lane, or to the block's first owner if there is none.  (A
// 5-step binary search over the prefix sums in shared memory was 20 % of the vote kernel's stall samples:
// profiles/r02d_vote_hot_lines.txt.)
// body(valid, vote index, tx, ty, tz) is called by all 32 lanes (converged), so it may use warp collectives.
// Batches are handed out dynamically through a global counter (zeroed before the launch) when `next_batch` is given:
// the work per batch varies a lot in the pose pass (votes cluster where the objects are), and a static stride leaves
// most warps idle behind the few that own the busy batches.
// Leaf tables of the ranks of a tree-sharded group (peer exchange): tree t is read from the table of the rank that
// traversed it, base[t % world]; world == 1 means "the local table only".  Peer tables live in other GPUs' memory.
// patch_world > 1: the ranks share the PATCHES instead (PatchShard): patch p is read from the table of the rank whose range
// holds it, and `shard` restricts the enumeration to this rank's own patches (the vote kernel) or to everything (the pose
// stage, which needs every rank's votes for its classes).
struct LeafTables {
    const int* base[HF6D_MAX_PEERS];
    int world;
    int patch_world;
    PatchShard shard;  // items enumerated: this rank's patches (world > 1) or all (world == 1)
};

// The vote stream: one 8-byte record per cast vote, {vote index, (u + 64) | (v + 64) << 13 | class << 26}, written by the
// vote kernel in the order it enumerates the votes (a warp reserves the exact space of its batch with one atomic, its lanes
// write consecutive records) and read back by the pose stage (window_stream_kernel), which therefore does not enumerate,
// look up and project the votes a second time.  u, v are clamped to [-64, 8127]: a vote outside that range cannot lie in any
// centre window (windows are at most 128 wide and centred inside the image; frames wider than 8000 px use the old path).
constexpr int STREAM_BIAS = 64, STREAM_COORD_MAX = 8191;
struct VoteStream {
    uint2* rec;   // nullptr: no stream
    int* n;       // records reserved so far (zeroed before the vote kernel)
    int cap;
};
// The streams the pose stage reads: its own, or -- one stream of frames sharded over the GPUs of a box (peer exchange) -- the
// stream of every rank, the peers' in place over NVLink: together they hold every vote of the frame exactly once.
struct StreamSet {
    const uint2* rec[HF6D_MAX_PEERS];
    const int* n[HF6D_MAX_PEERS];
    int world, cap;
};
__device__ __forceinline__ unsigned stream_pack(int u, int v, int cls) {
    const unsigned uu = (unsigned)min(max(u + STREAM_BIAS, 0), STREAM_COORD_MAX);
    const unsigned vv = (unsigned)min(max(v + STREAM_BIAS, 0), STREAM_COORD_MAX);
    return uu | (vv << 13) | ((unsigned)cls << 26);
}

// body(valid, vote index, tx, ty, tz, pos): pos = position of the vote in the stream (stream_n given) or -1.
template <class Body>
__device__ __forceinline__ void for_each_cast_vote(const DevForest& f, const FrameGeom& g, const int* __restrict__ locs,
                                                   const uint16_t* __restrict__ depth, const LeafTables& lt,
                                                   int n_items, WarpItems& wi, int* next_batch, int* stream_n, Body&& body) {
    const int lane = threadIdx.x & 31;
    const int Pp_all = n_items / f.T;
    int item_lo = 0;
    if (lt.shard.world > 1) {  // patch sharding: only this rank's patches
        int p_lo, p_hi;
        patch_shard_range(n_items / f.T, lt.shard, p_lo, p_hi);
        item_lo = p_lo * f.T;
        n_items = p_hi * f.T;
    }
    const int warp0 = item_lo + (blockIdx.x * VOTE_WARPS + (threadIdx.x >> 5)) * 32;
    const int stride = gridDim.x * VOTE_THREADS;
    for (int base = warp0;; base += stride) {
        if (next_batch) {
            int b = 0;
            if (lane == 0) b = atomicAdd(next_batch, 1);
            base = item_lo + __shfl_sync(0xffffffffu, b, 0) * 32;
        }
        if (base >= n_items) break;
        int vbeg = 0, vcnt = 0;
        float tx = 0.f, ty = 0.f, tz = 0.f;
        const int item = base + lane;
        if (item < n_items) {
            const int p = item / f.T, t = item - p * f.T;
            const int ord = lt.base[lt.patch_world > 1 ? patch_shard_owner(Pp_all, lt.patch_world, p) : (lt.world > 1 ? t % lt.world : 0)][item];
            if (ord >= 0) {  // -1: tree owned by another rank (and no peer table given)
                const int2 lv = __ldg(f.leaf_votes + __ldg(f.leaf_base + t) + ord);
                vbeg = lv.x;
                vcnt = lv.y;
                if (vcnt > 0) {
                    const int2 pc = *reinterpret_cast<const int2*>(locs + 2 * p);
                    const float z = div_const<1000, 1>((float)depth[(size_t)pc.y * g.W + pc.x]);  // HFTest.cpp:628
                    tz = z;
                    tx = __fdiv_rn(__fmul_rn(__fsub_rn((float)pc.x, g.cx), z), g.fx);
                    ty = __fdiv_rn(__fmul_rn(__fsub_rn((float)pc.y, g.cy), z), g.fy);
                }
            }
        }
        int incl = vcnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        int sbase = -1;
        if (stream_n && total > 0) {
            if (lane == 0) sbase = atomicAdd(stream_n, total);
            sbase = __shfl_sync(0xffffffffu, sbase, 0);
        }
        __syncwarp();  // the previous batch's readers are done
        wi.geo[lane] = make_float4(tx, ty, tz, __int_as_float(vbeg - (incl - vcnt)));
        const int start = vcnt > 0 ? incl - vcnt : 0x3fffffff;  // this item's first vote (an item without votes owns none)
        for (int j0 = 0; j0 < total; j0 += 32) {
            const int j = j0 + lane;
            const bool valid = j < total;
            const unsigned sp = (unsigned)(start - j0);
            const bool starts_here = sp < 32u;
            const unsigned starts = __reduce_or_sync(0xffffffffu, starts_here ? (1u << sp) : 0u);
            // owner of the votes before the block's first start: the last item that began earlier (starts ascend with the index)
            const int own0 = (31 - __clz(__ballot_sync(0xffffffffu, start < j0))) & 31;
            if (starts_here) wi.owner[sp] = lane;
            __syncwarp();
            const unsigned below = starts & (0xffffffffu >> (31 - lane));
            const int lo = below ? wi.owner[31 - __clz(below)] : own0;  // smallest i with incl[i] > j
            const float4 ge = wi.geo[lo];
            __syncwarp();  // every lane has read its owner before the next block's items overwrite the slots
            body(valid, __float_as_int(ge.w) + j, ge.x, ge.y, ge.z, sbase < 0 ? -1 : sbase + j);
        }
    }
}

__global__ void __launch_bounds__(VOTE_THREADS, 8)  // the grid is 8 CTAs per SM: all of them resident (32 registers)
vote_kernel(DevForest f, FrameGeom g, const __grid_constant__ ObjectSwitches sw, const int* __restrict__ locs,
            const uint16_t* __restrict__ depth, const int* __restrict__ leaf_ord, const int* __restrict__ counts,
            unsigned long long* __restrict__ maps, VoteStream stream, PatchShard pshard) {
    LeafTables lt;  // votes are cast for this rank's own trees / patches only: the local table (foreign entries are -1)
    lt.base[0] = leaf_ord;
    lt.world = 1;
    lt.patch_world = 1;
    lt.shard = pshard;
    __shared__ WarpItems s_items[VOTE_WARPS];
    __shared__ uint8_t s_detect[HF6D_MAX_CLASSES];
    if (threadIdx.x < HF6D_MAX_CLASSES) s_detect[threadIdx.x] = sw.should_detect[threadIdx.x];
    __syncthreads();
    const size_t HW = (size_t)g.H * g.W;
    for_each_cast_vote(f, g, locs, depth, lt, counts[1] * f.T, s_items[threadIdx.x >> 5], nullptr, stream.rec ? stream.n : nullptr,
                       [&](bool valid, int vi, float tx, float ty, float tz, int pos) {
                           if (!valid) return;
                           const float4 v = __ldg(f.vote4 + vi);
                           const unsigned meta = __float_as_uint(v.w);
                           const int cls = (int)(meta & 31u);
                           int uu = -STREAM_BIAS, vv = -STREAM_BIAS;  // a class that is not detected: a record no window holds
                           if (s_detect[cls]) {
                               project(g, __fadd_rn(v.x, tx), __fadd_rn(v.y, ty), __fadd_rn(v.z, tz), uu, vv);
                               if (uu >= 0 && uu < g.W && vv >= 0 && vv < g.H)
                                   atomicAdd(maps + (size_t)cls * HW + (size_t)vv * g.W + uu, (unsigned long long)(meta >> 5));
                           }
                           if (pos >= 0 && pos < stream.cap) stream.rec[pos] = make_uint2((unsigned)vi, stream_pack(uu, vv, cls));
                       });
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

// Coarse lookup in shared memory: the image (plus a margin of two cells) is cut into square cells whose side is the
// smallest power of two >= half_win; a cell holds the OR of the centres whose window touches it.  Most votes hit an
// empty cell and skip the exact window tests.
struct CellGrid {
    int shift, gx, gy, x0, y0;  // log2(cell side), grid size, pixel of cell (0,0)
};
__host__ __device__ __forceinline__ CellGrid make_cell_grid(int W, int H, int half_win) {
    CellGrid cg;
    cg.shift = 0;
    while ((1 << cg.shift) < half_win) ++cg.shift;
    const int cell = 1 << cg.shift;
    cg.x0 = -2 * cell;
    cg.y0 = -2 * cell;
    cg.gx = (W + 4 * cell + cell - 1) >> cg.shift;
    cg.gy = (H + 4 * cell + cell - 1) >> cg.shift;
    return cg;
}
inline size_t cell_grid_bytes(int W, int H, int half_win, int K) {
    const CellGrid cg = make_cell_grid(W, H, half_win);
    return ((size_t)cg.gx * cg.gy * K * sizeof(uint16_t) + 15) / 16 * 16;
}

// One window entry of the reference's center_leaf_map: cast vote `vi` landed on a pixel that lies in the windows `mask`
// of class `c`; zz = depth of that pixel in metres, < 0 when the pixel is outside the image or has no depth.
//   * cnt[slot][group] += 1   -- everything the entry contributes to the yaw/pitch and roll maps depends only on its
//                                leaf, so those maps are built later from these counts;
//   * z histogram of the slot: the WINDOW pixel's depth stands in for the patch centre and every vote of the leaf is
//     re-projected (HFTest.cpp:766-775).  `add_z(centre rank, bin, weight)` receives the increments.
template <class AddZ>
__device__ __forceinline__ void accumulate_entry(const DevForest& f, int vi, int c, unsigned mask, float zz, int n_groups,
                                                 unsigned* __restrict__ cnt, AddZ&& add_z) {
    const int gi = __ldg(f.vgroup + vi);
    for (unsigned m = mask; m; m &= m - 1)
        atomicAdd(cnt + (size_t)(c * HF6D_MAX_CENTRES + __ffs(m) - 1) * n_groups + gi, 1u);
    if (zz < 0.f) return;
    const int4 grp = __ldg(reinterpret_cast<const int4*>(f.groups) + gi);  // cls, w, vbeg, vcnt
#pragma unroll 4
    for (int q = 0; q < grp.w; ++q) {
        const int zb = f2i_x86(div_const<1, 100>(__fadd_rn(__ldg(f.oz + grp.z + q), zz)));  // only the integer part is used
        if (zb < 0 || zb >= HF6D_Z_BINS) continue;
        for (unsigned m = mask; m; m &= m - 1) add_z(__ffs(m) - 1, zb, (unsigned)grp.y);
    }
}

constexpr int ENTRY_BLOCK = 64;               // entries a warp reserves at a time (one global atomic per block)
constexpr unsigned ENTRY_INVALID = 0xFFFFFFFFu;

// Pass A.1a: enumerate the cast votes again (same arithmetic as vote_kernel, so the same pixels) and append every vote
// that falls in the window of an active centre to the entry list (16 bytes: vote, class | window mask, pixel depth).
// Only ~10 % of the votes are entries, and what an entry costs (a walk over its leaf's votes with histogram updates that
// all land on a handful of hot bins) needs privatised histograms, so it is split off into its own fully parallel pass.
// A warp reserves list space ENTRY_BLOCK entries at a time and pads what it leaves unused with ENTRY_INVALID.  Entries
// that do not fit the list (capacity is a fixed budget, the worst case is every cast vote) are accumulated in place with
// global atomics, so the result never depends on the capacity.
__global__ void __launch_bounds__(VOTE_THREADS)
window_entries_kernel(DevForest f, FrameGeom g, const __grid_constant__ ObjectSwitches sw, const int* __restrict__ locs,
                      const uint16_t* __restrict__ depth, const __grid_constant__ LeafTables lt,
                      const int* __restrict__ counts, CentreTable ct, int half_win, int n_groups, int* next_batch,
                      uint4* __restrict__ entries, int entry_cap, int* __restrict__ n_reserved,
                      unsigned* __restrict__ cnt /*[S][n_groups]*/, unsigned long long* __restrict__ zacc /*[S][Z_BINS]*/) {
    __shared__ SharedCentres sc;
    __shared__ WarpItems s_items[VOTE_WARPS];
    __shared__ unsigned s_classes;  // classes that are detected and have at least one active centre
    extern __shared__ __align__(16) uint8_t wc_smem[];
    uint16_t* s_cells = reinterpret_cast<uint16_t*>(wc_smem);  // [K][gy][gx]
    load_centres(sc, ct, f.K);
    const CellGrid cg = make_cell_grid(g.W, g.H, half_win);
    const int cells_per_class = cg.gx * cg.gy;
    const int S = f.K * HF6D_MAX_CENTRES;
    for (int i = threadIdx.x; i < f.K * cells_per_class; i += blockDim.x) s_cells[i] = 0;
    if (threadIdx.x == 0) s_classes = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < S; i += blockDim.x) {
        const int c = i / HF6D_MAX_CENTRES, k = i % HF6D_MAX_CENTRES;
        if (k >= sc.ctr[c].n || !sc.act[c][k] || !sw.should_detect[c]) continue;
        atomicOr(&s_classes, 1u << c);
        const int x_lo = sc.ctr[c].c[k].x - half_win, y_lo = sc.ctr[c].c[k].y - half_win;
        const int cx0 = max(0, (x_lo - cg.x0) >> cg.shift), cx1 = min(cg.gx - 1, (x_lo + 2 * half_win - 1 - cg.x0) >> cg.shift);
        const int cy0 = max(0, (y_lo - cg.y0) >> cg.shift), cy1 = min(cg.gy - 1, (y_lo + 2 * half_win - 1 - cg.y0) >> cg.shift);
        for (int cyi = cy0; cyi <= cy1; ++cyi)
            for (int cxi = cx0; cxi <= cx1; ++cxi) {
                const int idx = c * cells_per_class + cyi * cg.gx + cxi;  // 16-bit atomicOr through the containing word
                atomicOr(reinterpret_cast<unsigned*>(s_cells) + (idx >> 1), (1u << k) << ((idx & 1) * 16));
            }
    }
    __syncthreads();
    const unsigned classes = s_classes;
    const int lane = threadIdx.x & 31;
    int blk_base = 0, blk_used = ENTRY_BLOCK;  // warp-uniform: the block this warp is filling (none yet)
    bool overflow = false;                     // the list is full: accumulate in place from now on
    for_each_cast_vote(f, g, locs, depth, lt, counts[1] * f.T, s_items[threadIdx.x >> 5], next_batch, nullptr,
                       [&](bool valid, int vi, float tx, float ty, float tz, int) {
        unsigned mask = 0;
        int c = 0, uu = 0, vv = 0;
        if (valid) {
            const float4 v = __ldg(f.vote4 + vi);
            c = (int)(__float_as_uint(v.w) & 31u);
            if ((classes >> c) & 1u) {
                project(g, __fadd_rn(v.x, tx), __fadd_rn(v.y, ty), __fadd_rn(v.z, tz), uu, vv);
                const int cxi = (uu - cg.x0) >> cg.shift, cyi = (vv - cg.y0) >> cg.shift;  // negative stays negative
                unsigned cand = 0;
                if (cxi >= 0 && cxi < cg.gx && cyi >= 0 && cyi < cg.gy) cand = s_cells[c * cells_per_class + cyi * cg.gx + cxi];
                while (cand) {  // exact test for the few centres whose window touches the cell
                    const int k = __ffs(cand) - 1;
                    cand &= cand - 1;
                    const int ccx = sc.ctr[c].c[k].x, ccy = sc.ctr[c].c[k].y;
                    const bool hit = vv >= ccy - half_win && vv < ccy + half_win && uu >= ccx - half_win && uu < ccx + half_win;
                    mask |= (unsigned)hit << k;
                }
            }
        }
        const unsigned hl = __ballot_sync(0xffffffffu, mask != 0);
        if (!hl) return;  // the common case: no lane of this chunk hit a window
        float zz = -1.f;
        if (mask && vv >= 0 && vv < g.H && uu >= 0 && uu < g.W) {  // the reference reads out of bounds here
            const unsigned d = depth[(size_t)vv * g.W + uu];
            if (d != 0) zz = div_const<1000, 1>((float)d);
        }
        const int n = __popc(hl);
        if (!overflow && blk_used + n > ENTRY_BLOCK) {  // pad the rest of the current block, reserve the next
            for (int i = blk_used + lane; i < ENTRY_BLOCK; i += 32) entries[blk_base + i] = make_uint4(ENTRY_INVALID, 0u, 0u, 0u);
            int b = 0;
            if (lane == 0) b = atomicAdd(n_reserved, ENTRY_BLOCK);
            blk_base = __shfl_sync(0xffffffffu, b, 0);
            blk_used = 0;
            if (blk_base + ENTRY_BLOCK > entry_cap) { overflow = true; blk_used = ENTRY_BLOCK; }
        }
        if (mask) {
            if (!overflow)
                entries[blk_base + blk_used + __popc(hl & ((1u << lane) - 1u))] =
                    make_uint4((unsigned)vi, ((unsigned)c << 16) | mask, __float_as_uint(zz), 0u);
            else
                accumulate_entry(f, vi, c, mask, zz, n_groups, cnt, [&](int k, int zb, unsigned w) {
                    atomicAdd(zacc + (size_t)(c * HF6D_MAX_CENTRES + k) * HF6D_Z_BINS + zb, (unsigned long long)w);
                });
        }
        if (!overflow) blk_used += n;
    });
    if (!overflow)
        for (int i = blk_used + lane; i < ENTRY_BLOCK; i += 32) entries[blk_base + i] = make_uint4(ENTRY_INVALID, 0u, 0u, 0u);
}

// Per-warp staging of 32 listed entries (see window_accumulate_kernel).
struct WarpEntries {
    float zz[32];        // depth of the window pixel [m]; < 0: no z contribution
    int vb[32];          // first vote of the entry's group
    int vn[32];          // votes of the group
    unsigned w[32];      // Q16 weight
    unsigned cm[32];     // class << 16 | mask of the centres whose window holds the entry
};

// Pass A.1b: a warp takes 32 listed entries at a time (one per lane: group lookup and the cnt increments), then G
// lanes walk one entry's leaf votes together (G = 16 when no vote group holds more than 16 votes: two entries per
// step, coalesced independent oz loads instead of one dependent load chain per entry).  SMEM_Z: every CTA keeps the z
// histograms of all centres that can be active in shared memory (uint32; a wrap of the 32-bit counter carries 2^32
// into the global 64-bit accumulator, so the sums stay exact) and flushes them once.  zoff[c] = first shared histogram
// of class c (centre rank k uses zoff[c] + k).
struct ZSlotTable {
    int16_t zoff[HF6D_MAX_CLASSES + 1];
};
constexpr int WA_THREADS = 1024;  // one CTA per SM when the histograms live in shared memory: 32 warps share them
template <bool SMEM_Z, int G>
__global__ void __launch_bounds__(WA_THREADS)
window_accumulate_kernel(DevForest f, const uint4* __restrict__ entries, int entry_cap, const int* __restrict__ n_reserved,
                         int n_groups, const __grid_constant__ ZSlotTable zt, unsigned* __restrict__ cnt,
                         unsigned long long* __restrict__ zacc, int* __restrict__ next_batch) {
    __shared__ WarpEntries s_we[WA_THREADS / 32];
    extern __shared__ unsigned s_z[];  // [zt.zoff[K]][Z_BINS]
    const int nz = zt.zoff[f.K];
    if (SMEM_Z) {
        for (int i = threadIdx.x; i < nz * HF6D_Z_BINS; i += WA_THREADS) s_z[i] = 0u;
        __syncthreads();
    }
    const int n = min(*n_reserved, entry_cap / ENTRY_BLOCK * ENTRY_BLOCK);
    const int lane = threadIdx.x & 31, sub = lane % G, part = lane / G;
    constexpr int PARTS = 32 / G;
    WarpEntries& we = s_we[threadIdx.x >> 5];
    // batches of 32 entries are handed out through a global counter (zeroed before the launch): the cost of a batch varies
    // with its share of padding entries and the vote counts of its leaves, and a static stride left 28 % of the warp
    // time waiting at the final barrier
    constexpr int WA_GRAB = 4;  // 32-entry batches per grab: one atomic round trip per 128 entries
    int grab = 0, left = 0;
    for (;;) {
        if (left == 0) {
            if (lane == 0) grab = atomicAdd(next_batch, 1) * (32 * WA_GRAB);
            grab = __shfl_sync(0xffffffffu, grab, 0);
            left = WA_GRAB;
        }
        const int base = grab + (WA_GRAB - left) * 32;
        --left;
        if (base >= n) break;
        const uint4 e = entries[base + lane];  // n is a multiple of ENTRY_BLOCK, so base + lane < n
        float zz = -1.f;
        int4 grp = make_int4(0, 0, 0, 0);
        if (e.x != ENTRY_INVALID) {
            const int c = (int)(e.y >> 16);
            const int gi = __ldg(f.vgroup + (int)e.x);
            grp = __ldg(reinterpret_cast<const int4*>(f.groups) + gi);  // cls, w, vbeg, vcnt
            for (unsigned m = e.y & 0xFFFFu; m; m &= m - 1)
                atomicAdd(cnt + (size_t)(c * HF6D_MAX_CENTRES + __ffs(m) - 1) * n_groups + gi, 1u);
            zz = __uint_as_float(e.z);
        }
        we.zz[lane] = zz;
        we.vb[lane] = grp.z;
        we.vn[lane] = grp.w;
        we.w[lane] = (unsigned)grp.y;
        we.cm[lane] = e.y;
        __syncwarp();
#pragma unroll 2
        for (int j = part; j < 32; j += PARTS) {
            const float zj = we.zz[j];
            if (zj < 0.f) continue;  // uniform within the G lanes of this entry
            const int vb = we.vb[j], vn = we.vn[j];
            const unsigned cm = we.cm[j], w = we.w[j];
            const int c = (int)(cm >> 16);
            for (int q = sub; q < vn; q += G) {
                const int zb = f2i_x86(div_const<1, 100>(__fadd_rn(__ldg(f.oz + vb + q), zj)));  // integer part only
                if (zb < 0 || zb >= HF6D_Z_BINS) continue;
                for (unsigned m = cm & 0xFFFFu; m; m &= m - 1) {
                    const int k = __ffs(m) - 1;
                    if (SMEM_Z) {
                        const unsigned old = atomicAdd(s_z + (zt.zoff[c] + k) * HF6D_Z_BINS + zb, w);
                        if (old + w < old) atomicAdd(zacc + (size_t)(c * HF6D_MAX_CENTRES + k) * HF6D_Z_BINS + zb, 1ull << 32);
                    } else {
                        atomicAdd(zacc + (size_t)(c * HF6D_MAX_CENTRES + k) * HF6D_Z_BINS + zb, (unsigned long long)w);
                    }
                }
            }
        }
        __syncwarp();
    }
    if (SMEM_Z) {
        __syncthreads();
        for (int c = 0; c < f.K; ++c) {
            const int n_k = zt.zoff[c + 1] - zt.zoff[c];
            for (int i = threadIdx.x; i < n_k * HF6D_Z_BINS; i += WA_THREADS) {
                const unsigned v = s_z[zt.zoff[c] * HF6D_Z_BINS + i];
                if (v) atomicAdd(zacc + (size_t)c * HF6D_MAX_CENTRES * HF6D_Z_BINS + i, (unsigned long long)v);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ pose pass A, fused
// window_stream_kernel does in one launch what window_entries_kernel + window_accumulate_kernel + z_mode_kernel do in three,
// and reads the votes from the vote stream instead of enumerating them again:
//   * every lane takes one stream record (vote, pixel, class), looks the pixel up in the cell grid and tests the few
//     candidate windows exactly; hits go to a per-warp ring in shared memory;
//   * whenever the ring holds 32 entries the warp processes them together: lane i owns entry i (group lookup, the
//     (slot, group) counter), then G lanes walk one entry's leaf votes for the z histogram (HFTest.cpp:766-775);
//   * a trained forest's votes are coherent -- the votes of one leaf land on the same few pixels and in the same z bin --
//     so both updates are aggregated inside the warp before they become atomics: entries with the same (group, windows)
//     are counted once (__match_any_sync), and a G-lane walk whose votes all fall into one z bin adds them in one atomic
//     (profiles/r02_pose_before_ncu.txt: 10 shared-memory wavefronts per atomic instruction without it);
//   * the first increment of a (slot, group) counter appends the pair to a list: the yaw/pitch and roll passes walk that
//     list instead of scanning the dense S x groups counter table (44 MB at configs[1]), and the roll pass, the last
//     reader, zeroes the counters it visits, so the table is never cleared wholesale either;
//   * the last CTA to finish runs the z mode seeking (HFTest.cpp:803-812) for every slot.
struct WarpRing {
    unsigned vi[64], cm[64];
    int pix[64];  // y * W + x of the window pixel the vote landed on, -1 outside the image (its depth is read in process())
};
struct PairList {
    uint2* pairs;   // (slot, group), each (slot, group) at most once per frame
    int* n;
    int cap;
};

// z mode of one slot by one warp: NMS (1 wide, z_nms tall) over the 300-bin histogram, highest score, earliest on ties.
__device__ __forceinline__ void z_mode_warp(const unsigned long long* __restrict__ zacc_s, int z_nms, float* zf /*[Z_BINS] smem*/,
                                            uint8_t* active_s, float* mode_z_s) {
    const int lane = threadIdx.x & 31;
    for (int i = lane; i < HF6D_Z_BINS; i += 32) zf[i] = (float)((double)__ldcg(zacc_s + i) / 65536.0);
    __syncwarp();
    unsigned long long best = 0;
    const int n_top = HF6D_Z_BINS - 2 * z_nms + 2;  // tops 0 .. rows-2*wy+1
    for (int top = lane; top < n_top; top += 32) {
        int brow = top;
        float bv = zf[top];
        for (int r = top + 1; r < top + z_nms; ++r)
            if (zf[r] > bv) { bv = zf[r]; brow = r; }
        if (bv != 0.f && brow == top + z_nms / 2)
            best = max(best, ((unsigned long long)__float_as_uint(bv) << 32) | (unsigned long long)(0xFFFFu - (unsigned)brow));
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (lane == 0) {
        if (best == 0) *active_s = 0;
        else *mode_z_s = __fmul_rn((float)(0xFFFF - (int)(best & 0xFFFFu)), 0.01f);
    }
    __syncwarp();
}

// Scheduling.  The kernel is a set of per-warp chains of dependent L2 round trips (chunk counter -> records -> window pixel's
// depth -> vote group -> group record -> counter atomic) at 32 warps per SM, and the work per record varies by orders of
// magnitude (background votes miss every window; a vote inside one drags its leaf's z offsets through a bisection), so
//   * a chunk is ROWS rows of 32 records; the records of the NEXT chunk are requested before the current one is examined,
//     the window pixel's depth is read together with the vote group, and the pair-list append waits for the counter's old
//     value until the lane's next batch (settle);
//   * chunks are handed out in ranges of clamp(remaining / (2 x warps), 1, k_max) chunks through one counter; a warp's first
//     range is its own index (no atomic), a request is issued at the first chunk of a range and looked at before its last.
//     Measured (profiles/r02_ws_tune.json): granularity decides -- k_max = 1 with 128-record chunks is fastest (pose stage
//     0.278 ms), ranges of 4 chunks cost 10 %, of 8 chunks 30 %: the last warp keeps its CTA, and with it the kernel, waiting
//     (12.9 % of the stall samples sit at the final barrier even so); the same-address atomics are not the limit.
//   * leaves whose sorted z offsets straddle several 1 cm bins are binned by bisection per occupied bin; one linear pass
//     over the list was measured slower at every list length (same file), as were 64-record chunks.
// HF6D_WS_TUNE="rows,k_max" (rows 2 or 4).
template <int G, int WS_ROWS>
__global__ void __launch_bounds__(WA_THREADS)
window_stream_kernel(DevForest f, FrameGeom g, const __grid_constant__ ObjectSwitches sw, const __grid_constant__ StreamSet ss,
                     const uint16_t* __restrict__ depth, CentreTable ct, int half_win, int n_groups,
                     const __grid_constant__ ZSlotTable zt, unsigned* __restrict__ cnt, PairList pl,
                     unsigned long long* __restrict__ zacc, int* __restrict__ next_chunk, int* __restrict__ done,
                     int z_nms, uint8_t* __restrict__ active, float* __restrict__ mode_z, int k_max) {
    __shared__ SharedCentres sc;
    __shared__ WarpRing s_ring[WA_THREADS / 32];
    __shared__ unsigned s_classes;
    __shared__ int s_last;
    __shared__ int s_sn[HF6D_MAX_PEERS], s_sbase[HF6D_MAX_PEERS + 1];  // records and first chunk of every stream
    extern __shared__ __align__(16) uint8_t ws_smem[];
    const int nz = zt.zoff[f.K];
    unsigned* s_z = reinterpret_cast<unsigned*>(ws_smem);                                       // [nz][Z_BINS]
    uint16_t* s_cells = reinterpret_cast<uint16_t*>(ws_smem + (((size_t)nz * HF6D_Z_BINS * 4 + 15) & ~(size_t)15));  // [K][gy][gx]
    load_centres(sc, ct, f.K);
    const CellGrid cg = make_cell_grid(g.W, g.H, half_win);
    const int cells_per_class = cg.gx * cg.gy;
    const int S = f.K * HF6D_MAX_CENTRES;
    for (int i = threadIdx.x; i < nz * HF6D_Z_BINS; i += WA_THREADS) s_z[i] = 0u;
    for (int i = threadIdx.x; i < f.K * cells_per_class; i += WA_THREADS) s_cells[i] = 0;
    if (threadIdx.x == 0) {
        s_classes = 0;
        int chunks = 0;
        for (int r = 0; r < ss.world; ++r) {
            s_sn[r] = min(*ss.n[r], ss.cap);
            s_sbase[r] = chunks;
            chunks += (s_sn[r] + 32 * WS_ROWS - 1) / (32 * WS_ROWS);
        }
        for (int r = ss.world; r <= HF6D_MAX_PEERS; ++r) s_sbase[r] = chunks;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < S; i += WA_THREADS) {
        const int c = i / HF6D_MAX_CENTRES, k = i % HF6D_MAX_CENTRES;
        if (k >= sc.ctr[c].n || !sc.act[c][k] || !sw.should_detect[c]) continue;
        atomicOr(&s_classes, 1u << c);
        const int x_lo = sc.ctr[c].c[k].x - half_win, y_lo = sc.ctr[c].c[k].y - half_win;
        const int cx0 = max(0, (x_lo - cg.x0) >> cg.shift), cx1 = min(cg.gx - 1, (x_lo + 2 * half_win - 1 - cg.x0) >> cg.shift);
        const int cy0 = max(0, (y_lo - cg.y0) >> cg.shift), cy1 = min(cg.gy - 1, (y_lo + 2 * half_win - 1 - cg.y0) >> cg.shift);
        for (int cyi = cy0; cyi <= cy1; ++cyi)
            for (int cxi = cx0; cxi <= cx1; ++cxi) {
                const int idx = c * cells_per_class + cyi * cg.gx + cxi;
                atomicOr(reinterpret_cast<unsigned*>(s_cells) + (idx >> 1), (1u << k) << ((idx & 1) * 16));
            }
    }
    __syncthreads();
    const unsigned classes = s_classes;
    const int lane = threadIdx.x & 31;
    WarpRing& ring = s_ring[threadIdx.x >> 5];
    const int total_chunks = s_sbase[HF6D_MAX_PEERS];

    // bin of one vote of a leaf for a window pixel at depth zj (HFTest.cpp:770-775), unclamped / -1 outside the histogram
    auto z_bin_raw = [&](float oz, float zj) { return f2i_x86(div_const<1, 100>(__fadd_rn(oz, zj))); };  // integer part only
    auto z_bin = [&](float oz, float zj) {
        const int zb = z_bin_raw(oz, zj);
        return (zb < 0 || zb >= HF6D_Z_BINS) ? -1 : zb;
    };

    // A (slot, group) counter that was zero before this lane's increment makes the lane append the pair to the list -- but the
    // atomic's return value is a full L2 round trip away, so it is looked at when the lane's NEXT batch starts (or after the
    // last one): pend_old is the value the counter held (non-zero = nothing to do), pend_slot / pend_gi the pair.
    unsigned pend_old = 1u;
    int pend_slot = 0, pend_gi = 0;
    auto settle = [&]() {  // called by the whole warp: one list reservation for all of its new pairs
        const unsigned fresh = __ballot_sync(0xffffffffu, pend_old == 0u);  // lanes whose increment was the pair's first
        if (fresh) {
            int at = 0;
            if (lane == __ffs(fresh) - 1) at = atomicAdd(pl.n, __popc(fresh));
            at = __shfl_sync(0xffffffffu, at, __ffs(fresh) - 1) + __popc(fresh & ((1u << lane) - 1u));
            if (pend_old == 0u && at < pl.cap) pl.pairs[at] = make_uint2((unsigned)pend_slot, (unsigned)pend_gi);
        }
        pend_old = 1u;
    };
    // the ring's entries [head, head + n_e), n_e <= 32, ONE LANE PER ENTRY: counters, pair list, z histograms
    auto process = [&](int head, int n_e) {
        settle();
        const bool have = lane < n_e;
        unsigned cm = 0;
        float zz = -1.f;
        int gi = 0;
        int4 grp = make_int4(0, 0, 0, 0);
        float2 ozr = make_float2(0.f, 0.f);
        if (have) {
            const int at = (head + lane) & 63;
            cm = ring.cm[at];
            const int pix = ring.pix[at];
            unsigned d = 0;
            if (pix >= 0) d = depth[pix];  // in flight together with the group lookup (the reference reads out of bounds here)
            gi = __ldg(f.vgroup + (int)ring.vi[at]);
            grp = __ldg(reinterpret_cast<const int4*>(f.groups) + gi);  // cls, w, vbeg, vcnt
            ozr = __ldg(f.oz_range + gi);
            if (d != 0) zz = div_const<1000, 1>((float)d);
        }
        // entries with the same group in the same windows (the votes of one leaf cast by one patch, typically): one update
        const unsigned long long key = have ? (((unsigned long long)(unsigned)gi << 16) | (cm & 0xFFFFu)) : (~0ull - (unsigned)lane);
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        if (have && lane == __ffs(peers) - 1) {
            const int c = (int)(cm >> 16);
            const unsigned mult = (unsigned)__popc(peers);
            unsigned m = cm & 0xFFFFu;
            // the first (usually the only) window of the entry: increment now, examine the old value later (settle)
            pend_slot = c * HF6D_MAX_CENTRES + __ffs(m) - 1;
            pend_gi = gi;
            pend_old = atomicAdd(cnt + (size_t)pend_slot * n_groups + gi, mult);
            for (m &= m - 1; m; m &= m - 1) {  // a pixel inside several overlapping windows: the others right away
                const int slot = c * HF6D_MAX_CENTRES + __ffs(m) - 1;
                if (atomicAdd(cnt + (size_t)slot * n_groups + gi, mult) == 0u) {  // first entry of this (slot, group)
                    const int at = atomicAdd(pl.n, 1);
                    if (at < pl.cap) pl.pairs[at] = make_uint2((unsigned)slot, (unsigned)gi);
                }
            }
        }
        // z histograms (HFTest.cpp:766-775): every vote of the entry's leaf lands in bin floor((oz + zz) / 0.01), zz = the window
        // pixel's depth standing in for the patch centre.  The bin is monotonic in oz, so with the group's oz values sorted the
        // votes of one bin are a contiguous run: the lane finds the run boundaries by bisection (the same float operations as
        // the reference, evaluated on the candidate vote) and adds weight x run length once per bin -- one bin for most
        // entries of a trained leaf, two when the leaf straddles a boundary -- instead of visiting every vote.
        const bool wants = have && zz >= 0.f && grp.w > 0;
        // (ordered: no NaN among the offsets; bounded: the bin stays inside the range where float -> int is monotonic)
        if (wants && !(ozr.x > ozr.y) && fabsf(ozr.x) < 1e6f && fabsf(ozr.y) < 1e6f) {
            const int c = (int)(cm >> 16);
            const float* ozs = f.oz_sorted + grp.z;
            const int hi = z_bin_raw(ozr.y, zz);
            int b = z_bin_raw(ozr.x, zz), prev = 0;
            auto add_run = [&](int bin, int votes) {  // `votes` votes of the leaf fall into `bin`
                if (bin < 0 || bin >= HF6D_Z_BINS) return;
                const unsigned add = (unsigned)grp.y * (unsigned)votes;
                for (unsigned m = cm & 0xFFFFu; m; m &= m - 1) {
                    const int k = __ffs(m) - 1;
                    const unsigned old = atomicAdd(s_z + (zt.zoff[c] + k) * HF6D_Z_BINS + bin, add);
                    if (old + add < old) atomicAdd(zacc + (size_t)(c * HF6D_MAX_CENTRES + k) * HF6D_Z_BINS + bin, 1ull << 32);
                }
            };
            if (b == hi) {  // the whole leaf in one bin
                add_run(b, grp.w);
                prev = grp.w;
            }
            while (prev < grp.w) {
                int idx = grp.w;  // first vote beyond bin b
                if (b < hi) {
                    int lo_i = prev, hi_i = grp.w;  // invariant: bin(lo_i - 1) <= b < bin(hi_i)
                    while (lo_i < hi_i) {
                        const int mid = (lo_i + hi_i) >> 1;
                        if (z_bin_raw(__ldg(ozs + mid), zz) > b) hi_i = mid; else lo_i = mid + 1;
                    }
                    idx = lo_i;
                }
                if (idx > prev && b >= 0 && b < HF6D_Z_BINS) {
                    const unsigned add = (unsigned)grp.y * (unsigned)(idx - prev);
                    for (unsigned m = cm & 0xFFFFu; m; m &= m - 1) {
                        const int k = __ffs(m) - 1;
                        const unsigned old = atomicAdd(s_z + (zt.zoff[c] + k) * HF6D_Z_BINS + b, add);
                        if (old + add < old) atomicAdd(zacc + (size_t)(c * HF6D_MAX_CENTRES + k) * HF6D_Z_BINS + b, 1ull << 32);
                    }
                }
                if (idx >= grp.w) break;
                prev = idx;
                b = z_bin_raw(__ldg(ozs + idx), zz);  // the next occupied bin
            }
        } else if (wants) {  // a group with a NaN or absurd offset (never in a sane forest): vote by vote
            const int c = (int)(cm >> 16);
            for (int q = 0; q < grp.w; ++q) {
                const int zb = z_bin(__ldg(f.oz + grp.z + q), zz);
                if (zb < 0) continue;
                for (unsigned m = cm & 0xFFFFu; m; m &= m - 1) {
                    const int k = __ffs(m) - 1;
                    const unsigned old = atomicAdd(s_z + (zt.zoff[c] + k) * HF6D_Z_BINS + zb, (unsigned)grp.y);
                    if (old + (unsigned)grp.y < old) atomicAdd(zacc + (size_t)(c * HF6D_MAX_CENTRES + k) * HF6D_Z_BINS + zb, 1ull << 32);
                }
            }
        }
        __syncwarp();
    };

    int head = 0, pending = 0;  // warp-uniform ring state
    // the records of chunk `chunk` (zero records beyond the end of its stream: class 0, pixel (-64, -64), never a hit)
    auto load_chunk = [&](int chunk, uint2 (&rec)[WS_ROWS], int& base, int& n) {
        int which = 0;  // the stream this chunk belongs to
        while (which + 1 < ss.world && chunk >= s_sbase[which + 1]) ++which;
        base = (chunk - s_sbase[which]) * (32 * WS_ROWS);
        n = s_sn[which];
        const uint2* __restrict__ srec = ss.rec[which];
#pragma unroll
        for (int r = 0; r < WS_ROWS; ++r) {
            const int i = base + r * 32 + lane;
            rec[r] = i < n ? __ldcs(srec + i) : make_uint2(0u, 0u);  // read once: streaming load
        }
    };
    const int n_warps_all = (int)gridDim.x * (WA_THREADS / 32);
    auto range_size = [&](int claimed) { return min(max((total_chunks - claimed) / (2 * n_warps_all), 1), k_max); };
    const int k0 = range_size(0);
    const int dyn0 = n_warps_all * k0;  // first chunk handed out by the counter
    int cur_c = ((int)blockIdx.x * (WA_THREADS / 32) + (threadIdx.x >> 5)) * k0, cur_end = min(cur_c + k0, total_chunks);
    int nxt_c = total_chunks, nxt_end = total_chunks, seen = dyn0, req = 0, raw = 0;
    bool first_of_range = true;
    uint2 nxt[WS_ROWS];
    int nbase = 0, nn = 0;
    if (cur_c < total_chunks) load_chunk(cur_c, nxt, nbase, nn);
    while (cur_c < total_chunks) {
        uint2 rec[WS_ROWS];
#pragma unroll
        for (int r = 0; r < WS_ROWS; ++r) rec[r] = nxt[r];
        const int base = nbase, n = nn;
        if (first_of_range) {  // ask for the next range now ...
            req = range_size(seen);
            if (lane == 0) raw = atomicAdd(next_chunk, req);
        }
        const bool last_of_range = cur_c + 1 >= cur_end;
        if (last_of_range) {   // ... and look at the answer before this range's last chunk
            nxt_c = min(dyn0 + __shfl_sync(0xffffffffu, raw, 0), total_chunks);
            nxt_end = min(nxt_c + req, total_chunks);
            seen = nxt_end;
        }
        {   // the chunk after this one: its records are in flight while this one is examined
            const int f = last_of_range ? nxt_c : cur_c + 1;
            nn = 0;
            if (f < total_chunks) load_chunk(f, nxt, nbase, nn);
        }
#pragma unroll
        for (int r = 0; r < WS_ROWS; ++r) {
            unsigned mask = 0;
            const int c = (int)(rec[r].y >> 26);
            const int uu = (int)(rec[r].y & 8191u) - STREAM_BIAS, vv = (int)((rec[r].y >> 13) & 8191u) - STREAM_BIAS;
            if (base + r * 32 + lane < n && ((classes >> c) & 1u)) {
                const int cxi = (uu - cg.x0) >> cg.shift, cyi = (vv - cg.y0) >> cg.shift;
                unsigned cand = 0;
                if (cxi >= 0 && cxi < cg.gx && cyi >= 0 && cyi < cg.gy) cand = s_cells[c * cells_per_class + cyi * cg.gx + cxi];
                while (cand) {  // exact test for the few centres whose window touches the cell
                    const int k = __ffs(cand) - 1;
                    cand &= cand - 1;
                    const int ccx = sc.ctr[c].c[k].x, ccy = sc.ctr[c].c[k].y;
                    const bool hit = vv >= ccy - half_win && vv < ccy + half_win && uu >= ccx - half_win && uu < ccx + half_win;
                    mask |= (unsigned)hit << k;
                }
            }
            const unsigned hl = __ballot_sync(0xffffffffu, mask != 0);
            if (!hl) continue;
            if (mask) {
                const int at = (head + pending + __popc(hl & ((1u << lane) - 1u))) & 63;
                ring.vi[at] = rec[r].x;
                ring.cm[at] = ((unsigned)c << 16) | mask;
                ring.pix[at] = (vv >= 0 && vv < g.H && uu >= 0 && uu < g.W) ? vv * g.W + uu : -1;
            }
            pending += __popc(hl);
            __syncwarp();
            if (pending >= 32) {
                process(head, 32);
                head = (head + 32) & 63;
                pending -= 32;
            }
        }
        first_of_range = last_of_range;
        if (last_of_range) {
            cur_c = nxt_c;
            cur_end = nxt_end;
        } else {
            ++cur_c;
        }
    }
    if (pending) process(head, pending);
    settle();

    __syncthreads();
    for (int c = 0; c < f.K; ++c) {
        const int n_k = zt.zoff[c + 1] - zt.zoff[c];
        for (int i = threadIdx.x; i < n_k * HF6D_Z_BINS; i += WA_THREADS) {
            const unsigned v = s_z[zt.zoff[c] * HF6D_Z_BINS + i];
            if (v) atomicAdd(zacc + (size_t)c * HF6D_MAX_CENTRES * HF6D_Z_BINS + i, (unsigned long long)v);
        }
    }
    // the last CTA to get here has every CTA's histogram flushes behind it: z mode seeking for all slots
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(done, 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // the shared histograms' memory is free now: one 300-float scratch row per warp, as many warps as there are rows
    const int warp = threadIdx.x >> 5, warps = min(WA_THREADS / 32, nz);
    float* zf = reinterpret_cast<float*>(ws_smem) + warp * HF6D_Z_BINS;
    if (warp < warps)
        for (int s = warp; s < S; s += warps)
            if (active[s]) z_mode_warp(zacc + (size_t)s * HF6D_Z_BINS, z_nms, zf, active + s, mode_z + s);
}

// Pass A.2 / B on the pair list: as yawpitch_from_counts_kernel / roll_from_counts_kernel below, but G lanes take one
// listed (slot, group) pair at a time instead of scanning the dense counter table.
template <int G>
__global__ void __launch_bounds__(TABLE_THREADS)
yawpitch_from_pairs_kernel(DevForest f, const unsigned* __restrict__ cnt, int n_groups, PairList pl,
                           const uint8_t* __restrict__ active, PoseRegion reg, unsigned long long* __restrict__ ypacc) {
    const int lane = threadIdx.x & 31, sub = lane % G;
    const size_t yp_slot = (size_t)reg.ny * reg.np;
    const int n = min(*pl.n, pl.cap);
    const int stride = gridDim.x * TABLE_THREADS / G;
    for (int e = (blockIdx.x * TABLE_THREADS + threadIdx.x) / G; e < n; e += stride) {
        const uint2 pr = __ldg(pl.pairs + e);
        const int s = (int)pr.x, gi = (int)pr.y;
        if (!active[s]) continue;
        const unsigned c_hits = __ldcg(cnt + (size_t)s * n_groups + gi);
        const int4 grp = __ldg(reinterpret_cast<const int4*>(f.groups) + gi);
        const unsigned long long wk = (unsigned long long)(unsigned)grp.y * c_hits;
        unsigned long long* yps = ypacc + (size_t)s * yp_slot;
        for (int q = sub; q < grp.w; q += G) {
            const short4 bn = __ldg(f.bins + grp.z + q);
            const int yaw = bn.x, pit = bn.y;
            const int sy = yaw < 0 ? -1 : 1, sp = pit < 0 ? -1 : 1;  // copysign(1, (float)int): sign(0) = +1
#pragma unroll
            for (int k1 = 0; k1 < 2; ++k1)
#pragma unroll
                for (int k2 = 0; k2 < 2; ++k2) {
                    const int ry = yaw - sy * k1 * 360 + 360 - reg.y0, rp = pit - sp * k2 * 360 + 360 - reg.p0;
                    if (ry < 0 || ry >= reg.ny || rp < 0 || rp >= reg.np) continue;
                    atomicAdd(yps + (size_t)ry * reg.np + rp, wk);
                }
        }
    }
}

// Pass A.2: yaw/pitch maps from the (slot, group) entry counts: every entry re-walks all votes of its leaf
// (HFTest.cpp:763-791), so a group adds  count * w  at each of its votes' bins (and their +-360 wrap copies).
// A warp reads 32 consecutive counters, then G lanes take one non-zero (slot, group) pair each.
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
            if (passes == 1) {
                // the usual case (a group's votes fit the G lanes): every lane keeps its vote's bins in registers for all
                // peaks instead of fetching them again per peak
                const bool have = sub < grp.w;
                const short4 bn = have ? __ldg(f.bins + grp.z + sub) : make_short4(0, 0, 0, 0);
                const int Y = (int)bn.x + 360, P0 = (int)bn.y + 360, r = bn.z;
                const int b0 = r + 360, b1 = r < 0 ? r + 720 : r;
                for (int p = 0; p < np; ++p) {
                    const int2 yp = __ldg(reinterpret_cast<const int2*>(pk.peak_yx) + s * pk.max_peaks + p);
                    const bool in = have && Y >= yp.x - half_box && Y < yp.x + half_box && P0 >= yp.y - half_box && P0 < yp.y + half_box;
                    const int inbox = __popc(__ballot_sync(gmask, in));
                    if (inbox == 0 || !have) continue;
                    const unsigned long long wk = (unsigned long long)(unsigned)grp.y * c_hits * (unsigned long long)inbox;
                    unsigned long long* rs = racc + ((size_t)s * pk.max_peaks + p) * HF6D_POSE_BINS;
                    if (b0 >= 0 && b0 < HF6D_POSE_BINS) atomicAdd(rs + b0, wk);
                    if (b1 >= 0 && b1 < HF6D_POSE_BINS) atomicAdd(rs + b1, wk);
                }
                continue;
            }
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

// The roll histograms from the pair list (see window_stream_kernel): G lanes per listed (slot, group) pair.  Last reader of
// the pair's counter: zeroes it for the next frame.
// One listed pair's contribution to the roll histograms of its slot's peaks; bn0 = the bins of vote `sub` (first pass).
template <int G>
__device__ __forceinline__ void roll_pair(const DevForest& f, int s, const int4& grp, unsigned c_hits, int np, const short4& bn0,
                                          const PeakTable& pk, int half_box, unsigned long long* __restrict__ racc, int sub,
                                          unsigned gmask) {
    const int passes = (grp.w + G - 1) / G;
    if (passes == 1) {
        // the usual case (a group's votes fit the G lanes): every lane keeps its vote's bins in registers for all peaks
        const bool have = sub < grp.w;
        const int Y = (int)bn0.x + 360, P0 = (int)bn0.y + 360, r = bn0.z;
        const int b0 = r + 360, b1 = r < 0 ? r + 720 : r;
        for (int p = 0; p < np; ++p) {
            const int2 yp = __ldg(reinterpret_cast<const int2*>(pk.peak_yx) + s * pk.max_peaks + p);
            const bool in = have && Y >= yp.x - half_box && Y < yp.x + half_box && P0 >= yp.y - half_box && P0 < yp.y + half_box;
            const int inbox = __popc(__ballot_sync(gmask, in));
            if (inbox == 0 || !have) continue;
            const unsigned long long wk = (unsigned long long)(unsigned)grp.y * c_hits * (unsigned long long)inbox;
            unsigned long long* rs = racc + ((size_t)s * pk.max_peaks + p) * HF6D_POSE_BINS;
            if (b0 >= 0 && b0 < HF6D_POSE_BINS) atomicAdd(rs + b0, wk);
            if (b1 >= 0 && b1 < HF6D_POSE_BINS) atomicAdd(rs + b1, wk);
        }
        return;
    }
    for (int p = 0; p < np; ++p) {
        const int Yp = __ldg(pk.peak_yx + (s * pk.max_peaks + p) * 2), Pp = __ldg(pk.peak_yx + (s * pk.max_peaks + p) * 2 + 1);
        int inbox = 0;
        for (int ps = 0; ps < passes; ++ps) {
            const int q = ps * G + sub;
            bool in = false;
            if (q < grp.w) {
                const short4 bn = ps == 0 ? bn0 : __ldg(f.bins + grp.z + q);
                const int Y = (int)bn.x + 360, P0 = (int)bn.y + 360;
                in = Y >= Yp - half_box && Y < Yp + half_box && P0 >= Pp - half_box && P0 < Pp + half_box;
            }
            inbox += __popc(__ballot_sync(gmask, in));
        }
        if (inbox == 0) continue;
        const unsigned long long wk = (unsigned long long)(unsigned)grp.y * c_hits * (unsigned long long)inbox;
        unsigned long long* rs = racc + ((size_t)s * pk.max_peaks + p) * HF6D_POSE_BINS;
        for (int q = sub; q < grp.w; q += G) {
            const int r = q == sub ? bn0.z : __ldg(f.bins + grp.z + q).z;
            const int b0 = r + 360, b1 = r < 0 ? r + 720 : r;
            if (b0 >= 0 && b0 < HF6D_POSE_BINS) atomicAdd(rs + b0, wk);
            if (b1 >= 0 && b1 < HF6D_POSE_BINS) atomicAdd(rs + b1, wk);
        }
    }
}

// The roll histograms from the pair list (see window_stream_kernel): G lanes per listed (slot, group) pair, two pairs per
// step -- a pair is a chain of three dependent loads (pair -> counter / group -> votes) and the kernel is bound by their
// latency (ncu: 12 % issue utilisation at full occupancy with one pair per step).  Last reader of a pair's counter: zeroes
// it for the next frame.
template <int G>
__global__ void __launch_bounds__(TABLE_THREADS)
roll_from_pairs_kernel(DevForest f, unsigned* __restrict__ cnt, int n_groups, PairList pl, PeakTable pk, int half_box,
                       unsigned long long* __restrict__ racc /*[S][max_peaks][POSE_BINS]*/) {
    const int lane = threadIdx.x & 31, sub = lane % G, part = lane / G;
    const unsigned gmask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (part * G));
    const int n = min(*pl.n, pl.cap);
    const int stride = gridDim.x * TABLE_THREADS / G;
    for (int e0 = (blockIdx.x * TABLE_THREADS + threadIdx.x) / G; e0 < n; e0 += 2 * stride) {
        const int e1 = e0 + stride;
        const bool two = e1 < n;  // uniform within the G lanes
        const uint2 pr0 = __ldg(pl.pairs + e0), pr1 = two ? __ldg(pl.pairs + e1) : make_uint2(0u, 0u);
        const int s0 = (int)pr0.x, g0 = (int)pr0.y, s1 = (int)pr1.x, g1 = (int)pr1.y;
        unsigned c0 = 0, c1 = 0;
        if (sub == 0) {
            unsigned* cp0 = cnt + (size_t)s0 * n_groups + g0;
            unsigned* cp1 = cnt + (size_t)s1 * n_groups + g1;
            c0 = __ldcg(cp0);
            if (two) c1 = __ldcg(cp1);
            *cp0 = 0u;  // last reader: the table is zero again for the next frame
            if (two) *cp1 = 0u;
        }
        const int np0 = __ldg(pk.n_peaks + s0), np1 = two ? __ldg(pk.n_peaks + s1) : 0;
        const int4 grp0 = __ldg(reinterpret_cast<const int4*>(f.groups) + g0);
        const int4 grp1 = two ? __ldg(reinterpret_cast<const int4*>(f.groups) + g1) : make_int4(0, 0, 0, 0);
        const short4 bn0 = (np0 && sub < grp0.w) ? __ldg(f.bins + grp0.z + sub) : make_short4(0, 0, 0, 0);
        const short4 bn1 = (np1 && sub < grp1.w) ? __ldg(f.bins + grp1.z + sub) : make_short4(0, 0, 0, 0);
        c0 = __shfl_sync(gmask, c0, part * G);
        c1 = __shfl_sync(gmask, c1, part * G);
        if (np0) roll_pair<G>(f, s0, grp0, c0, np0, bn0, pk, half_box, racc, sub, gmask);
        if (np1) roll_pair<G>(f, s1, grp1, c1, np1, bn1, pk, half_box, racc, sub, gmask);
    }
}

}  // namespace hf6d
