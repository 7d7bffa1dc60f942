// libhf6d.so: the C ABI of include/hf6d.h.  Owns the model on the device, the per-slot frame workspaces and streams,
// and launches the stage kernels (gather.cuh, encoder.cuh, forest.cuh, vote.cuh, modes.cuh).
// There is deliberately no CPU path in this file: without a usable sm_100 device every compute entry point fails.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <array>
#include <map>
#include <set>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <thread>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/hf6d.h"
#include "common.cuh"
#include "encoder.cuh"
#include "forest.cuh"
#include "gather.cuh"
#include "model.hpp"
#include "modes.cuh"
#include "refine.cuh"
#include "train.cuh"
#include "render.cuh"
#include "patchdb.hpp"
#include "texture_check.cuh"
#include "vote.cuh"

using namespace hf6d;

namespace {

thread_local std::string g_create_error;

constexpr int WA_MAX_DYN_SMEM = 200 * 1024;  // shared-memory z histograms of window_accumulate_kernel
constexpr int WS_MAX_DYN_SMEM = 192 * 1024;  // ... and of window_stream_kernel (31 KB of static shared memory beside them)
constexpr int MAX_YP = 16;    // upper bound for max_yaw_pitch_hypotheses
constexpr int MAX_ROLL = 8;   // upper bound for max_roll_hypotheses

struct DeviceModel {
    DevForest f{};
    // owned allocations
    std::vector<void*> allocs;
    // encoder
    __nv_bfloat16* W[3] = {nullptr, nullptr, nullptr};
    float* b[3] = {nullptr, nullptr, nullptr};
    int n_in[3], n_out[3], k_pad[3], n_pad[3], block_n[3];
    // split-bf16 mode (hf6d_set_encoder_mode 1), uploaded on first use: [n_pad][2 * k_pad] = (w_hi | w_lo), plain biases
    __nv_bfloat16* Ws[3] = {nullptr, nullptr, nullptr};
    float* bs[3] = {nullptr, nullptr, nullptr};
    // fp16 mode (hf6d_set_encoder_mode 2), uploaded on first use: fp16 weights [n_pad][k_pad]; layer 1 NOT divided by 255
    __half* Wh[3] = {nullptr, nullptr, nullptr};
    uint8_t* sep_ok = nullptr;
    uint8_t* class_mask = nullptr;  // [HF6D_MAX_CLASSES] classes whose centres / poses this context seeks
};

struct Slot {
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaEvent_t ev[HF6D_STAGE_COUNT + 1];
    bool ev_valid[HF6D_STAGE_COUNT + 1];
    int launches = 0;
    // frame
    uint8_t* bgr = nullptr;       // frame the kernels read: own_bgr or a caller-owned device frame (hf6d_bind_frame)
    uint16_t* depth = nullptr;
    uint2* tex = nullptr;         // packed texels {B | G<<8 | R<<16, depth mm} built from bgr + depth by the gather stage
    float4* normals = nullptr;    // (nx, ny, nz, 0) per pixel, patch_mode 1 only
    uint8_t* own_bgr = nullptr;
    uint16_t* own_depth = nullptr;
    cudaEvent_t ev_enc[4] = {nullptr, nullptr, nullptr, nullptr};
    bool ev_enc_valid = false;
    int* row_count = nullptr;
    int* row_off = nullptr;  // [gh] index of the first patch of every stride-grid row (gather_tile_kernel)
    int* counts = nullptr;  // [2] inside the result block
    int* locs = nullptr;
    __nv_bfloat16 *A0 = nullptr, *H1 = nullptr, *H2 = nullptr;
    uint8_t* q_u8 = nullptr;
    float* feat = nullptr;
    int* leaf_ord = nullptr;
    unsigned long long *maps = nullptr, *map_tmp = nullptr;
    float* blurred = nullptr;
    unsigned long long* list = nullptr;
    int* list_n = nullptr;
    unsigned long long *zacc = nullptr, *ypacc = nullptr, *yptmp = nullptr, *racc = nullptr;
    float* ypblur = nullptr;
    unsigned long long* bmax = nullptr;  // [maps][block rows][block cols] key-maxima of 8x8 blocks (NMS)
    uint4* entries = nullptr;     // window entries of the pose stage: {vote, class << 16 | window mask, pixel depth, -}
    unsigned* win_cnt = nullptr;  // [S][n_groups] window entries per (slot, vote group)
    // vote stream path of the pose stage (vote.cuh, window_stream_kernel)
    uint2* vstream = nullptr;     // one record per cast vote, written by the vote kernel
    int* stream_n = nullptr;      // records in the stream
    uint2* pairs = nullptr;       // distinct (slot, group) pairs of the frame
    bool stream_valid = false;    // the stream holds the votes of the slot's current leaf table
    bool cnt_dirty = true;        // win_cnt may hold non-zero counters (the stream path leaves it zeroed behind itself)
    // result block (one D2H)
    uint8_t* res_dev = nullptr;
    uint8_t* res_host = nullptr;  // pinned
    EncoderLayerLaunch enc[3];
    // split-bf16 mode: hidden activations as (hi | lo) halves [cap][2 * n_pad], allocated on first use
    __nv_bfloat16 *H1s = nullptr, *H2s = nullptr;
    EncoderLayerLaunch enc_split[3];
    bool split_ready = false;
    EncoderLayerLaunch enc_fp16[3];  // fp16 mode: the bf16 launches with the weight map and the operand format replaced
    bool fp16_ready = false;
    // feature storage 1 (hf6d_ctx::feat16): the feature layer writes fp16 rows, the traversal reads them; `feat` (fp32) is then
    // filled on demand only (hf6d_fetch, hf6d_device_ptr, hf6d_encode_patches) or by hf6d_inject
    __half* feat16 = nullptr;
    EncoderLayerLaunch enc16[2];     // feature-layer launches with fp16 output: [0] bf16 operands, [1] fp16 operands
    bool enc16_fp16_ready = false;
    bool feat_is16 = false;          // the slot's current features live in feat16
    // CUDA graph of one whole frame (HF6D_GRAPH=1): captured from the slot's own launches once a full eager frame has run with
    // the same frame pointers and context configuration, replayed while both stay the same (run_range)
    cudaGraphExec_t graph_exec = nullptr;
    const void *graph_bgr = nullptr, *graph_depth = nullptr;
    uint64_t graph_epoch = 0;
    int graph_launches = 0;
    bool capturing = false;          // run_stage / run_range_impl are being captured: no event records
    bool last_full_valid = false;    // the slot's previous run was a whole frame (eager or replayed) with the key below
    const void *last_bgr = nullptr, *last_depth = nullptr;
    uint64_t last_epoch = 0;
    uint32_t peer_seq = 0;  // frames this slot has pushed through the peer exchange (both flags carry it)
    bool busy = false;  // submit/wait bookkeeping
    int ticket = -1;
    std::vector<void*> allocs;
};

struct ResultLayout {
    size_t counts, centres, active, mode_z, n_peaks, peak_yx, peak_score, records, total;
};

}  // namespace

struct hf6d_refiner;  // stage REFINE (refine_api.inc), created on first use

struct hf6d_ctx {
    hf6d_params p{};
    hf6d_refiner* rf = nullptr;
    // object_options of the options file the context was created from (hf6d_load_option_models)
    std::vector<std::string> mesh_files;
    std::vector<float> obj_nn_search_radius;
    std::vector<int> obj_icp_iterations;
    hf6d_refine_params option_refine{};
    bool has_option_refine = false;
    int device = 0, n_slots = 1, sms = 148;
    HostForest hf;
    std::vector<HostLayer> layers;
    std::vector<hf6d_object> objects;
    DeviceModel dm;
    std::vector<Slot> slots;
    FrameGeom g{};
    int S = 0;  // accumulator slots = K * HF6D_MAX_CENTRES
    PoseRegion reg{};       // yaw/pitch accumulator rectangle
    MapRect yp_blur{};      // rectangle of the blurred yaw/pitch map that is computed
    int yp_left0 = 0, yp_nleft = 0, yp_top0 = 0, yp_ntop = 0;  // NMS window origins on the yaw/pitch map
    ResultLayout rl{};
    int lanes_per_hit = 32;      // lanes that share one (slot, group) pair: 16 when no vote group holds more than 16 votes
    int wc_ctas_per_sm = 1;      // resident CTAs of window_entries_kernel per SM
    int entry_cap = 0;           // capacity of a frame slot's window-entry list (entries beyond it are accumulated in place)
    bool gather_tiled = true;    // gather_tile_kernel (shared-memory staged texel tiles); HF6D_GATHER=direct: the per-tap loads
    int gather_tile_texels = 0, gather_smem = 0;
    bool use_stream = false;     // pose stage reads the vote stream instead of enumerating the votes again
    int stream_cap = 0, pair_cap = 0;
    int ws_rows = 4, ws_kmax = 1;  // window_stream_kernel: rows per chunk (2 / 4), largest chunk range (HF6D_WS_TUNE="rows,k_max")
    int shard_rank = 0, shard_world = 1;
    int class_rank = 0, class_world = 1;  // centres + pose only for classes k % class_world == class_rank
    PatchShard pshard{0, 1};              // gather .. vote only for this rank's patches (hf6d_set_patch_shard)
    int peer_split = 0;                   // what hf6d_peer_attach shards: 0 = trees, 1 = patches
    // peer exchange (tree-sharded mode over NVLink peer memory, hf6d_peer_attach)
    bool peer_on = false;
    int peer_rank = 0, peer_world = 1;
    uint32_t* peer_flags = nullptr;                                   // own flag block [n_slots][2][HF6D_MAX_PEERS]
    uint32_t* peer_flags_of[HF6D_MAX_PEERS] = {};                     // every rank's flag block (own entry = own block)
    std::vector<const unsigned long long*> peer_maps[HF6D_MAX_PEERS]; // [rank][slot]
    std::vector<const int*> peer_leaf[HF6D_MAX_PEERS];                // [rank][slot]
    std::vector<const uint2*> peer_vstream[HF6D_MAX_PEERS];           // [rank][slot] the peers' vote streams (pose stage)
    std::vector<const int*> peer_stream_n[HF6D_MAX_PEERS];            // [rank][slot]
    bool peer_streams = false;                                        // every rank of the group publishes its vote stream
    std::vector<void*> peer_opened;                                   // cudaIpcOpenMemHandle results
    int* peer_timeout = nullptr;                                      // set by the fallback wait kernel when it gives up
    bool peer_wait_expired = false;                                   // a bounded host-side wait ran out (sync_stream_bounded)
    bool use_graph = false;   // HF6D_GRAPH=1: whole frames are replayed from a CUDA graph per slot (one launch instead of 22 + memsets)
    uint64_t epoch = 1;       // bumped by every call that changes what a frame's launches look like
    int encoder_mode = 0;
    // Feature storage: 0 = fp32 rows (the reference's type), 1 = fp16 rows written by the feature layer of the bf16 / fp16
    // operand modes and read by the traversal (half the HBM bytes of both; the split mode always stores fp32).
    // HF6D_FEATURES=fp16 (tuning switch, read at create); needs F % 8 == 0 and a 160- or 256-wide feature tile.
    bool feat16 = false;
    int feat16_block_n = 0, feat16_variant = 0;
    // per encoder layer: row of HF6D_ENC_CONFIGS within the layer's shape class (0 = default: CTA pairs, cta_group::2;
    // 1 = stand-alone CTAs).  Tuning switch: HF6D_ENC_VARIANT="a,b,c".
    int enc_variant[3] = {0, 0, 0};
    int debug_capture = 0;
    int next_ticket = 0;
    std::string err;
    std::mutex mu;
};

namespace {

int fail(hf6d_ctx* c, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}

#define CU_TRY(c, expr)                                                                              \
    do {                                                                                             \
        cudaError_t e_ = (expr);                                                                     \
        if (e_ != cudaSuccess) return fail(c, HF6D_ECUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
    } while (0)

template <class T>
int dev_alloc(hf6d_ctx* c, std::vector<void*>& owner, T** out, size_t count) {
    void* p = nullptr;
    const size_t bytes = std::max<size_t>(count * sizeof(T), 16);
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return fail(c, HF6D_ENOMEM, "cudaMalloc(%zu bytes): %s", bytes, cudaGetErrorString(e));
    owner.push_back(p);
    *out = reinterpret_cast<T*>(p);
    return HF6D_OK;
}

template <class T>
int dev_upload(hf6d_ctx* c, std::vector<void*>& owner, const T** out, const std::vector<T>& v) {
    T* p = nullptr;
    int r = dev_alloc(c, owner, &p, v.size());
    if (r) return r;
    if (!v.empty()) CU_TRY(c, memcpy_h2d_done(p, v.data(), v.size() * sizeof(T)));
    *out = p;
    return HF6D_OK;
}

int round_up(int v, int m) { return (v + m - 1) / m * m; }

// The reference's roll-separation test (HFTest.cpp:918-921), tabulated with libm for every pair of kept roll bins.
void build_sep_table(std::vector<uint8_t>& t) {
    t.assign(361 * 361, 0);
    for (int a = 0; a <= 360; ++a)
        for (int b = 0; b <= 360; ++b) {
            const float prev = (float)(a + 180);
            const int ry = b + 180;
            const float dot = (float)(cos(prev / 180.0f * M_PI) * cos(ry / 180.0f * M_PI) +
                                      sin(prev / 180.0f * M_PI) * sin(ry / 180.0f * M_PI));
            t[a * 361 + b] = acos(dot) / M_PI * 180.0f > 7 ? 1 : 0;
        }
}

int upload_class_mask(hf6d_ctx* c);

int upload_model(hf6d_ctx* c) {
    DeviceModel& dm = c->dm;
    const HostForest& hf = c->hf;
    dm.f.T = hf.T; dm.f.K = hf.K; dm.f.F = hf.F;
    int r;
    if ((r = dev_upload(c, dm.allocs, &dm.f.recs, hf.recs))) return r;
    if ((r = dev_upload(c, dm.allocs, &dm.f.root, hf.rec_root))) return r;
    if ((r = dev_upload(c, dm.allocs, &dm.f.leaf_base, hf.leaf_base))) return r;
    if ((r = dev_upload(c, dm.allocs, &dm.f.group_off, hf.group_off))) return r;
    if ((r = dev_upload(c, dm.allocs, &dm.f.groups, hf.groups))) return r;
    if ((r = dev_upload(c, dm.allocs, &dm.f.oz, hf.oz))) return r;
    if ((r = dev_upload(c, dm.allocs, &dm.f.oz_sorted, hf.oz_sorted))) return r;
    if ((r = dev_upload(c, dm.allocs, &dm.f.vgroup, hf.vgroup))) return r;
    {
        std::vector<float4> v4(hf.ox.size());
        for (size_t i = 0; i < v4.size(); ++i) {
            const VoteGroup& vg = hf.groups[hf.vgroup[i]];
            const uint32_t meta = (uint32_t)vg.cls | (vg.w << 5);  // cls < 32, w <= 65536
            float mf;
            memcpy(&mf, &meta, 4);
            v4[i] = make_float4(hf.ox[i], hf.oy[i], hf.oz[i], mf);
        }
        if ((r = dev_upload(c, dm.allocs, &dm.f.vote4, v4))) return r;
        std::vector<int2> lv(hf.leaf_vbeg.size());
        for (size_t i = 0; i < lv.size(); ++i) lv[i] = make_int2(hf.leaf_vbeg[i], hf.leaf_vcnt[i]);
        if ((r = dev_upload(c, dm.allocs, &dm.f.leaf_votes, lv))) return r;
    }
    {
        std::vector<short4> bins(hf.yaw.size());
        for (size_t i = 0; i < bins.size(); ++i) bins[i] = make_short4(hf.yaw[i], hf.pitch[i], hf.roll[i], 0);
        if ((r = dev_upload(c, dm.allocs, &dm.f.bins, bins))) return r;
    }
    {
        std::vector<float2> rg(hf.groups.size());
        for (size_t i = 0; i < rg.size(); ++i) rg[i] = make_float2(hf.g_ozmin[i], hf.g_ozmax[i]);
        if ((r = dev_upload(c, dm.allocs, &dm.f.oz_range, rg))) return r;
    }
    std::vector<uint8_t> sep;
    build_sep_table(sep);
    const uint8_t* sp = nullptr;
    if ((r = dev_upload(c, dm.allocs, &sp, sep))) return r;
    dm.sep_ok = const_cast<uint8_t*>(sp);
    if ((r = dev_alloc(c, dm.allocs, &dm.class_mask, (size_t)HF6D_MAX_CLASSES))) return r;
    if ((r = upload_class_mask(c))) return r;

    // encoder: bf16 weights [n_pad][k_pad], zero padded; 1/255 folded into layer 1 (the A operand holds q itself)
    for (int l = 0; l < 3; ++l) {
        const HostLayer& L = c->layers[l];
        dm.n_in[l] = L.in;
        dm.n_out[l] = L.out;
        dm.k_pad[l] = l == 0 ? L.in : dm.n_pad[l - 1];
        const bool last = l == 2;
        if (last) {
            dm.block_n[l] = (L.out % 160 == 0) ? 160 : 256;
            // 208-wide tiles when the layer fits four of them (the reference's 800 features): 4 % padded flops, but the A
            // tile is re-read 4 times instead of 5 and the layer is bound by operand delivery (98 -> 94 us).  CTA pairs
            // only, so the choice is made after probing them; HF6D_ENC_N3=160 keeps the narrow tiles.
            const char* e3 = getenv("HF6D_ENC_N3");
            if (L.out > 3 * 208 && L.out <= 4 * 208 && !(e3 && atoi(e3) != 208)) {
                EncoderLayerLaunch probe{};
                probe.block_n = 208;
                probe.last = true;
                probe.variant = 0;
                if (launch_encoder_layer(probe, nullptr, c->sms, nullptr, true) == cudaSuccess) dm.block_n[l] = 208;
                else cudaGetLastError();
            }
            dm.n_pad[l] = round_up(L.out, dm.block_n[l]);
        } else {
            dm.block_n[l] = 256;
            dm.n_pad[l] = round_up(L.out, 256);
        }
        if (dm.k_pad[l] % ENC_BLOCK_K) return fail(c, HF6D_EINVAL, "encoder layer %d: K=%d is not a multiple of 64", l, dm.k_pad[l]);
        if (dm.n_pad[l] > ENC_MAX_N) return fail(c, HF6D_EINVAL, "encoder layer %d wider than %d", l, ENC_MAX_N);
        std::vector<__nv_bfloat16> w((size_t)dm.n_pad[l] * dm.k_pad[l], __float2bfloat16(0.f));
        for (int n = 0; n < L.out; ++n)
            for (int k = 0; k < L.in; ++k) {
                float v = L.W[(size_t)n * L.in + k];
                if (l == 0) v = v / 255.0f;
                w[(size_t)n * dm.k_pad[l] + k] = __float2bfloat16(v);
            }
        std::vector<float> b(dm.n_pad[l], 0.f);
        for (int n = 0; n < L.out; ++n) b[n] = last ? L.b[n] : 0.5f * L.b[n];  // hidden layers evaluate tanh(x/2)
        const __nv_bfloat16* wp = nullptr;
        const float* bp = nullptr;
        if ((r = dev_upload(c, dm.allocs, &wp, w))) return r;
        if ((r = dev_upload(c, dm.allocs, &bp, b))) return r;
        dm.W[l] = const_cast<__nv_bfloat16*>(wp);
        dm.b[l] = const_cast<float*>(bp);
    }
    return HF6D_OK;
}

int alloc_slot(hf6d_ctx* c, Slot& s) {
    const FrameGeom& g = c->g;
    const int K = c->hf.K, T = c->hf.T, F = c->hf.F, S = c->S;
    const size_t HW = (size_t)g.W * g.H;
    int r;
    CU_TRY(c, cudaStreamCreateWithFlags(&s.own_stream, cudaStreamNonBlocking));
    s.stream = s.own_stream;
    for (int i = 0; i <= HF6D_STAGE_COUNT; ++i) {
        CU_TRY(c, cudaEventCreate(&s.ev[i]));
        s.ev_valid[i] = false;
    }
    for (int i = 0; i < 4; ++i) CU_TRY(c, cudaEventCreate(&s.ev_enc[i]));
    if ((r = dev_alloc(c, s.allocs, &s.own_bgr, HW * 3))) return r;
    if ((r = dev_alloc(c, s.allocs, &s.own_depth, HW))) return r;
    s.bgr = s.own_bgr;
    s.depth = s.own_depth;
    if ((r = dev_alloc(c, s.allocs, &s.tex, HW))) return r;
    if (c->p.patch_mode == 1 && (r = dev_alloc(c, s.allocs, &s.normals, HW))) return r;
    if ((r = dev_alloc(c, s.allocs, &s.row_count, (size_t)g.gh))) return r;
    if ((r = dev_alloc(c, s.allocs, &s.row_off, (size_t)g.gh))) return r;
    if ((r = dev_alloc(c, s.allocs, &s.locs, (size_t)g.cap * 2))) return r;
    if ((r = dev_alloc(c, s.allocs, &s.A0, (size_t)g.cap * c->dm.k_pad[0]))) return r;
    if ((r = dev_alloc(c, s.allocs, &s.q_u8, (size_t)g.cap * c->dm.n_in[0]))) return r;
    if ((r = dev_alloc(c, s.allocs, &s.H1, (size_t)g.cap * c->dm.n_pad[0]))) return r;
    if ((r = dev_alloc(c, s.allocs, &s.H2, (size_t)g.cap * c->dm.n_pad[1]))) return r;
    if ((r = dev_alloc(c, s.allocs, &s.feat, (size_t)g.cap * F))) return r;
    if ((r = dev_alloc(c, s.allocs, &s.leaf_ord, (size_t)g.cap * T))) return r;
    if ((r = dev_alloc(c, s.allocs, &s.maps, HW * K))) return r;
    if ((r = dev_alloc(c, s.allocs, &s.map_tmp, HW * K))) return r;
    if ((r = dev_alloc(c, s.allocs, &s.blurred, HW * K))) return r;
    const int n_lists = std::max(K, S);
    if ((r = dev_alloc(c, s.allocs, &s.list, (size_t)n_lists * NMS_LIST_CAP))) return r;
    if ((r = dev_alloc(c, s.allocs, &s.list_n, (size_t)n_lists + 8))) return r;  // +0..2: batch counter, reserved entries, accumulate-pass batch counter; +3..5: stream chunk counter, finished CTAs, (slot, group) pairs (pose)
    if (c->use_stream) {
        if ((r = dev_alloc(c, s.allocs, &s.vstream, (size_t)c->stream_cap))) return r;
        if ((r = dev_alloc(c, s.allocs, &s.stream_n, (size_t)4))) return r;
        if ((r = dev_alloc(c, s.allocs, &s.pairs, (size_t)c->pair_cap))) return r;
    }
    if ((r = dev_alloc(c, s.allocs, &s.entries, (size_t)c->entry_cap))) return r;
    const size_t yp = (size_t)c->reg.ny * c->reg.np;
    const size_t yp_tmp = (size_t)c->reg.ny * c->yp_blur.nc, yp_out = (size_t)c->yp_blur.nr * c->yp_blur.nc;
    if ((r = dev_alloc(c, s.allocs, &s.zacc, (size_t)S * HF6D_Z_BINS))) return r;
    if ((r = dev_alloc(c, s.allocs, &s.ypacc, (size_t)S * yp))) return r;
    if ((r = dev_alloc(c, s.allocs, &s.yptmp, (size_t)S * yp_tmp))) return r;
    if ((r = dev_alloc(c, s.allocs, &s.ypblur, (size_t)S * yp_out))) return r;
    {
        const BlockGrid b1 = make_block_grid(MapRect{0, 0, g.H, g.W}), b2 = make_block_grid(c->yp_blur);
        if ((r = dev_alloc(c, s.allocs, &s.bmax, std::max((size_t)K * b1.by * b1.bx, (size_t)S * b2.by * b2.bx)))) return r;
    }
    if ((r = dev_alloc(c, s.allocs, &s.win_cnt, (size_t)S * c->hf.groups.size()))) return r;
    if ((r = dev_alloc(c, s.allocs, &s.racc, (size_t)S * MAX_YP * HF6D_POSE_BINS))) return r;
    if ((r = dev_alloc(c, s.allocs, &s.res_dev, c->rl.total))) return r;
    CU_TRY(c, cudaMemset(s.res_dev, 0, c->rl.total));
    CU_TRY(c, cudaMemset(s.win_cnt, 0, (size_t)S * c->hf.groups.size() * 4));
    s.cnt_dirty = false;
    CU_TRY(c, cudaMemset(s.leaf_ord, 0xFF, (size_t)g.cap * T * 4));
    CU_TRY(c, cudaMemset(s.A0, 0, (size_t)g.cap * c->dm.k_pad[0] * 2));
    CU_TRY(c, cudaMallocHost(reinterpret_cast<void**>(&s.res_host), c->rl.total));
    memset(s.res_host, 0, c->rl.total);
    s.counts = reinterpret_cast<int*>(s.res_dev + c->rl.counts);

    // encoder launches: tensor maps over this slot's activation buffers
    const DeviceModel& dm = c->dm;
    const void* a_in[3] = {s.A0, s.H1, s.H2};
    void* outs[3] = {s.H1, s.H2, s.feat};
    for (int l = 0; l < 3; ++l) {
        EncoderLayerLaunch& L = s.enc[l];
        const bool short_k = dm.k_pad[l] / ENC_BLOCK_K <= 6;
        const EncoderConfig cfg = encoder_config(dm.block_n[l], l == 2, short_k, c->enc_variant[l]);
        if (!cfg.pair) return fail(c, HF6D_EINVAL, "encoder layer %d has no kernel variant %d", l, c->enc_variant[l]);
        if (!make_bf16_kmajor_map(&L.tmA, a_in[l], (uint64_t)g.cap, (uint64_t)dm.k_pad[l], ENC_BLOCK_M) ||
            !make_bf16_kmajor_map(&L.tmB, dm.W[l], (uint64_t)dm.n_pad[l], (uint64_t)dm.k_pad[l],
                                  (uint32_t)(dm.block_n[l] / cfg.pair)))
            return fail(c, HF6D_ECUDA, "cuTensorMapEncodeTiled failed for encoder layer %d", l);
        L.last = l == 2;
        if (!make_out_map(&L.tmC, outs[l], (uint64_t)g.cap, (uint64_t)(L.last ? F : dm.n_pad[l]), L.last ? 4 : 2,
                          cfg.chunk_bytes))
            return fail(c, HF6D_ECUDA, "cuTensorMapEncodeTiled failed for the output of encoder layer %d", l);
        L.bias = dm.b[l];
        L.K = dm.k_pad[l];
        L.n_pad = dm.n_pad[l];
        L.block_n = dm.block_n[l];
        L.short_k = short_k;
        L.variant = c->enc_variant[l];
        // gather writes A0 ascending (36 MB: all of it stays in L2); layer 1 ascending -> layer 2 descending -> layer 3
        // ascending -> traverse descending: every consumer starts where its producer stopped
        static const bool snake = !(getenv("HF6D_SNAKE") && atoi(getenv("HF6D_SNAKE")) == 0);  // tuning override
        L.reverse_m = snake && l == 1;
    }
    return HF6D_OK;
}

void free_all(hf6d_ctx* c) {
    for (Slot& s : c->slots) {
        for (void* p : s.allocs) cudaFree(p);
        if (s.res_host) cudaFreeHost(s.res_host);
        for (int i = 0; i <= HF6D_STAGE_COUNT; ++i) if (s.ev[i]) cudaEventDestroy(s.ev[i]);
        for (int i = 0; i < 4; ++i) if (s.ev_enc[i]) cudaEventDestroy(s.ev_enc[i]);
        if (s.graph_exec) cudaGraphExecDestroy(s.graph_exec);
        if (s.own_stream) cudaStreamDestroy(s.own_stream);
    }
    for (void* p : c->dm.allocs) cudaFree(p);
}

void rf_free(hf6d_refiner* r);  // refine_api.inc

// Classes whose centres and poses this context seeks: detected AND owned under the class shard (votes are cast for every
// detected class regardless, because the summed maps of a class need the votes of all ranks' trees).
bool seeks_class(const hf6d_ctx* c, int k) {
    return c->objects[k].should_detect && (k % c->class_world) == c->class_rank;
}

ObjectSwitches switches_of(const hf6d_ctx* c, bool pose_stage = false) {
    ObjectSwitches sw;
    memset(&sw, 0, sizeof sw);
    for (size_t i = 0; i < c->objects.size(); ++i)
        sw.should_detect[i] = (pose_stage ? seeks_class(c, (int)i) : (bool)c->objects[i].should_detect) ? 1 : 0;
    return sw;
}

int upload_class_mask(hf6d_ctx* c) {
    uint8_t m[HF6D_MAX_CLASSES];
    memset(m, 0, sizeof m);
    for (int k = 0; k < c->hf.K && k < (int)c->objects.size(); ++k) m[k] = seeks_class(c, k) ? 1 : 0;
    if (c->dm.class_mask) {
        cudaError_t e = memcpy_h2d_done(c->dm.class_mask, m, sizeof m);
        if (e != cudaSuccess) { c->err = std::string("cudaMemcpy(class mask): ") + cudaGetErrorString(e); return HF6D_ECUDA; }
    }
    return HF6D_OK;
}

struct ResView {
    int* counts;
    hf6d_centre_list* centres;
    uint8_t* active;
    float* mode_z;
    int* n_peaks;
    int* peak_yx;
    float* peak_score;
    HypRecord* records;
};
ResView view_of(const hf6d_ctx* c, uint8_t* base) {
    ResView v;
    v.counts = reinterpret_cast<int*>(base + c->rl.counts);
    v.centres = reinterpret_cast<hf6d_centre_list*>(base + c->rl.centres);
    v.active = base + c->rl.active;
    v.mode_z = reinterpret_cast<float*>(base + c->rl.mode_z);
    v.n_peaks = reinterpret_cast<int*>(base + c->rl.n_peaks);
    v.peak_yx = reinterpret_cast<int*>(base + c->rl.peak_yx);
    v.peak_score = reinterpret_cast<float*>(base + c->rl.peak_score);
    v.records = reinterpret_cast<HypRecord*>(base + c->rl.records);
    return v;
}

#define LAUNCH_CHECK(c, s)                                                                                  \
    do {                                                                                                    \
        cudaError_t e_ = cudaGetLastError();                                                                \
        if (e_ != cudaSuccess) return fail(c, HF6D_ECUDA, "kernel launch (%s:%d): %s", __FILE__, __LINE__,  \
                                           cudaGetErrorString(e_));                                         \
        ++(s).launches;                                                                                     \
    } while (0)

// ------------------------------------------------------------------------------------------------ peer exchange
// Tree-sharded mode without a collective library.  Every rank maps its peers' vote maps, leaf tables and a small flag
// block (CUDA IPC); the first kernels after the exchange point read the peers' buffers in place over NVLink
// (box_rows_kernel sums the ranks' maps of this rank's classes as it loads them, window_entries_kernel reads a tree's leaf
// ordinals from the rank that traversed it).  What is left of the "collective" is its synchronisation, two flags per
// (slot, rank), each written REMOTELY by a one-warp kernel and waited for LOCALLY by a stream memory operation
// (cuStreamWaitValue32: no SM, no spinning kernel):
//   ready[slot][r]    = n : rank r has finished traverse + vote of its n-th frame on this slot      (r -> everyone)
//   consumed[slot][r] = n : rank r has finished every kernel that reads its peers' n-th frame       (r -> everyone)
// Order on a slot's stream:  wait consumed >= n-1 | traverse, vote | write ready = n | wait ready >= n | centres, pose |
// write consumed = n.  Every wait depends only on work the peer enqueued BEFORE its own next wait, so the ranks cannot
// wait for each other in a cycle as long as every rank runs the same frames on the same slots.
constexpr int PEER_FLAG_READY = 0, PEER_FLAG_CONSUMED = 1;
inline size_t peer_flag_index(int slot, int kind, int rank) { return ((size_t)slot * 2 + kind) * HF6D_MAX_PEERS + rank; }

struct PeerTargets {
    uint32_t* p[HF6D_MAX_PEERS];
    int n;
};
__global__ void peer_signal_kernel(const __grid_constant__ PeerTargets t, uint32_t value) {
    if ((int)threadIdx.x < t.n) {
        __threadfence_system();  // everything this stream wrote before is visible to whoever sees the flag
        *reinterpret_cast<volatile uint32_t*>(t.p[threadIdx.x]) = value;
    }
}
// Fallback when the driver offers no stream memory operations: a one-warp spinning kernel with a bounded wait.
__global__ void peer_wait_kernel(const uint32_t* flags, int n, uint32_t value, int* timed_out) {
    if ((int)threadIdx.x >= n) return;
    const volatile uint32_t* f = flags + threadIdx.x;
    for (long long spin = 0; spin < (1ll << 24); ++spin) {
        if ((int32_t)(*f - value) >= 0) { __threadfence_system(); return; }
        __nanosleep(200);
    }
    *timed_out = 1;
}

typedef CUresult (*PFN_streamWaitValue32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
inline PFN_streamWaitValue32 get_stream_wait_value() {
    static PFN_streamWaitValue32 fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        const char* mode = getenv("HF6D_PEER_WAIT");
        if (mode && !strcmp(mode, "kernel")) return nullptr;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_streamWaitValue32>(p);
    }
    return fn;
}

int peer_signal(hf6d_ctx* c, Slot& s, int slot, int kind, uint32_t value) {
    PeerTargets t;
    memset(&t, 0, sizeof t);
    for (int r = 0; r < c->peer_world; ++r)
        if (r != c->peer_rank) t.p[t.n++] = c->peer_flags_of[r] + peer_flag_index(slot, kind, c->peer_rank);
    if (!t.n) return HF6D_OK;
    peer_signal_kernel<<<1, 32, 0, s.stream>>>(t, value);
    CU_TRY(c, cudaGetLastError());
    ++s.launches;
    return HF6D_OK;
}

int peer_wait(hf6d_ctx* c, Slot& s, int slot, int kind, uint32_t value) {
    PFN_streamWaitValue32 wait = get_stream_wait_value();
    for (int r = 0; r < c->peer_world; ++r) {
        if (r == c->peer_rank) continue;
        uint32_t* flag = c->peer_flags + peer_flag_index(slot, kind, r);
        if (wait) {
            const CUresult e = wait(reinterpret_cast<CUstream>(s.stream), reinterpret_cast<CUdeviceptr>(flag), value,
                                    CU_STREAM_WAIT_VALUE_GEQ);
            if (e != CUDA_SUCCESS) return fail(c, HF6D_ECUDA, "cuStreamWaitValue32 failed (%d)", (int)e);
        } else {
            peer_wait_kernel<<<1, 32, 0, s.stream>>>(flag, 1, value, c->peer_timeout);
            CU_TRY(c, cudaGetLastError());
            ++s.launches;
        }
    }
    return HF6D_OK;
}

// cudaStreamSynchronize with a deadline while the peer exchange is on: a stream that waits for a peer's flag (stream memory
// operation: no kernel, hence no in-kernel timeout) would otherwise block the host for ever if that peer died or skipped a
// frame.  HF6D_PEER_TIMEOUT_MS (default 20000).  On expiry the stream is still blocked: the context must be destroyed.
int sync_stream_bounded(hf6d_ctx* c, cudaStream_t st) {
    if (!c->peer_on) {
        CU_TRY(c, cudaStreamSynchronize(st));
        return HF6D_OK;
    }
    static const long timeout_ms = getenv("HF6D_PEER_TIMEOUT_MS") ? atol(getenv("HF6D_PEER_TIMEOUT_MS")) : 20000;
    const auto t0 = std::chrono::steady_clock::now();
    for (;;) {
        const cudaError_t e = cudaStreamQuery(st);
        if (e == cudaSuccess) return HF6D_OK;
        if (e != cudaErrorNotReady) return fail(c, HF6D_ECUDA, "cudaStreamQuery: %s", cudaGetErrorString(e));
        if (std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count() > timeout_ms) {
            c->peer_wait_expired = true;
            return fail(c, HF6D_ECUDA, "peer exchange: no answer from a peer within %ld ms (the slot's stream is still waiting; destroy the context)", timeout_ms);
        }
        std::this_thread::yield();
    }
}

inline int slot_index(const hf6d_ctx* c, const Slot& s) { return (int)(&s - c->slots.data()); }
PeerMaps peer_maps_of(const hf6d_ctx* c, int slot) {
    PeerMaps pm;
    memset(&pm, 0, sizeof pm);
    if (c->peer_on)
        for (int r = 0; r < c->peer_world; ++r)
            if (r != c->peer_rank) pm.base[pm.n++] = c->peer_maps[r][slot];
    return pm;
}
LeafTables leaf_tables_of(const hf6d_ctx* c, const Slot& s, int slot) {
    LeafTables lt;
    memset(&lt, 0, sizeof lt);
    lt.base[0] = s.leaf_ord;
    lt.world = 1;
    lt.patch_world = 1;
    lt.shard = PatchShard{0, 1};  // the pose stage enumerates every patch
    if (c->peer_on) {
        if (c->peer_split == 1) lt.patch_world = c->peer_world;
        else lt.world = c->peer_world;
        for (int r = 0; r < c->peer_world; ++r) lt.base[r] = r == c->peer_rank ? s.leaf_ord : c->peer_leaf[r][slot];
    }
    return lt;
}

int run_stage(hf6d_ctx* c, Slot& s, int stage) {
    const FrameGeom& g = c->g;
    const DevForest& f = c->dm.f;
    const int K = f.K, S = c->S;
    cudaStream_t st = s.stream;
    const hf6d_params& p = c->p;
    ResView rv = view_of(c, s.res_dev);
    switch (stage) {
        case HF6D_STAGE_SCAN: {
            const int blocks = (g.gh + 7) / 8;
            scan_count_kernel<<<blocks, 256, 0, st>>>(s.depth, g, s.row_count);
            LAUNCH_CHECK(c, s);
            scan_compact_kernel<<<blocks, 256, 0, st>>>(s.depth, g, s.row_count, s.locs, s.counts, s.row_off);
            LAUNCH_CHECK(c, s);
            break;
        }
        case HF6D_STAGE_GATHER: {
            pack_frame_kernel<<<(g.W * g.H + 255) / 256, 256, 0, st>>>(s.bgr, s.depth, g.W * g.H, s.tex);
            LAUNCH_CHECK(c, s);
            if (c->p.patch_mode == 1 && c->pshard.world > 1) return fail(c, HF6D_EINVAL, "patch sharding is not available in patch_mode 1");
            if (c->p.patch_mode == 1) {
                normals_kernel<<<(g.W * g.H + 255) / 256, 256, 0, st>>>(s.depth, g.W, g.H, c->p.normals_focal, s.normals);
                LAUNCH_CHECK(c, s);
                gather_normals_kernel<<<g.cap / GATHER_PATCHES_PER_CTA, GATHER_THREADS, 0, st>>>(
                    s.tex, s.normals, g, s.locs, s.counts, s.A0, c->debug_capture ? s.q_u8 : nullptr);
            } else if (c->gather_tiled)
                gather_tile_kernel<<<dim3((g.gw + GT_PATCHES - 1) / GT_PATCHES, g.gh), GT_THREADS, c->gather_smem, st>>>(
                    s.tex, g, s.locs, s.counts, s.row_count, s.row_off, c->gather_tile_texels, c->pshard, s.A0, c->debug_capture ? s.q_u8 : nullptr,
                    c->encoder_mode == 2);
            else
                gather_normalise_kernel<<<g.cap / GATHER_PATCHES_PER_CTA, GATHER_THREADS, 0, st>>>(
                    s.tex, g, s.locs, s.counts, c->pshard, s.A0, c->debug_capture ? s.q_u8 : nullptr, c->encoder_mode == 2);
            LAUNCH_CHECK(c, s);
            break;
        }
        case HF6D_STAGE_ENCODE: {
            if (!s.capturing) CU_TRY(c, cudaEventRecord(s.ev_enc[0], st));
            if (c->encoder_mode == 1 && !s.split_ready) return fail(c, HF6D_ESTATE, "split encoder buffers missing");
            if (c->encoder_mode == 2 && !s.fp16_ready) return fail(c, HF6D_ESTATE, "fp16 encoder state missing");
            for (int l = 0; l < 3; ++l) {
                const bool to16 = l == 2 && c->feat16 && c->encoder_mode != 1;
                EncoderLayerLaunch& L = to16 ? s.enc16[c->encoder_mode == 2] : c->encoder_mode == 1 ? s.enc_split[l]
                                        : c->encoder_mode == 2 ? s.enc_fp16[l] : s.enc[l];
                L.shard_rank = c->pshard.rank;
                L.shard_world = c->pshard.world;
                cudaError_t e = launch_encoder_layer(L, s.counts + 1, c->sms, st);
                if (e != cudaSuccess) return fail(c, HF6D_ECUDA, "encoder layer %d launch: %s", l, cudaGetErrorString(e));
                ++s.launches;
                if (!s.capturing) CU_TRY(c, cudaEventRecord(s.ev_enc[l + 1], st));
            }
            s.ev_enc_valid = !s.capturing;
            s.feat_is16 = c->feat16 && c->encoder_mode != 1;
            break;
        }
        case HF6D_STAGE_TRAVERSE: {
            const TraversePlan tp = traverse_plan(f.F, f.T, s.feat_is16 ? 2 : 4);
            const int n_owned = (f.T - c->shard_rank + c->shard_world - 1) / c->shard_world;
            const int per_slot = (n_owned + TRV_SLOTS - 1) / TRV_SLOTS;
            const int threads = (tp.n_bufs + 1) * 32, n_recs = (int)c->hf.recs.size();
            if (c->pshard.world > 1)  // rows of other ranks' patches: "not traversed here" (-1), as for trees that are not owned
                CU_TRY(c, cudaMemsetAsync(s.leaf_ord, 0xFF, (size_t)g.cap * f.T * 4, st));
            static const int trv_reverse = !(getenv("HF6D_SNAKE") && atoi(getenv("HF6D_SNAKE")) == 0);  // see alloc_slot
#define HF6D_TRV(NCH, FT, FEAT)                                                                                              \
    traverse_kernel<NCH, FT><<<c->sms, threads, tp.smem, st>>>(FEAT, f, s.counts, s.leaf_ord, c->shard_rank, c->shard_world, \
                                                               tp.n_bufs, tp.n_cache, n_recs, trv_reverse, c->pshard)
            if (s.feat_is16) {
                if (per_slot <= 1) HF6D_TRV(1, __half, s.feat16);
                else if (per_slot <= 2) HF6D_TRV(2, __half, s.feat16);
                else HF6D_TRV(4, __half, s.feat16);
            } else {
                if (per_slot <= 1) HF6D_TRV(1, float, s.feat);
                else if (per_slot <= 2) HF6D_TRV(2, float, s.feat);
                else HF6D_TRV(4, float, s.feat);
            }
#undef HF6D_TRV
            LAUNCH_CHECK(c, s);
            break;
        }
        case HF6D_STAGE_VOTE: {
            CU_TRY(c, cudaMemsetAsync(s.maps, 0, (size_t)K * g.W * g.H * 8, st));
            const long long items = (long long)g.cap * f.T;
            const int blocks = (int)std::min<long long>((items + VOTE_THREADS - 1) / VOTE_THREADS, (long long)c->sms * 8);
            // with a peer group the windows need the other ranks' votes too: they are read from the peers' streams in place
            const bool streaming = c->use_stream && ((c->shard_world == 1 && c->pshard.world == 1 && !c->peer_on) ||
                                                     (c->peer_on && c->peer_streams));
            VoteStream vs{nullptr, nullptr, 0};
            if (streaming) {
                CU_TRY(c, cudaMemsetAsync(s.stream_n, 0, 4, st));
                vs = VoteStream{s.vstream, s.stream_n, c->stream_cap};
            }
            vote_kernel<<<blocks, VOTE_THREADS, 0, st>>>(f, g, switches_of(c), s.locs, s.depth, s.leaf_ord, s.counts, s.maps, vs, c->pshard);
            LAUNCH_CHECK(c, s);
            s.stream_valid = streaming;
            break;
        }
        case HF6D_STAGE_CENTRES: {
            const MapDims md{g.H, g.W};
            const MapRect full{0, 0, g.H, g.W};
            const int kb = p.centers_blur_size, w = p.centers_nms_wsize;
            box_rows_kernel<<<dim3((g.H + BLUR_WARPS - 1) / BLUR_WARPS, K), BLUR_WARPS * 32,
                              box_rows_smem_bytes(g.W), st>>>(s.maps, s.map_tmp, md, full, full, kb, c->dm.class_mask,
                                                                        peer_maps_of(c, slot_index(c, s)));
            LAUNCH_CHECK(c, s);
            box_cols_kernel<<<dim3((g.W + 127) / 128, (g.H + BLUR_COL_CHUNK - 1) / BLUR_COL_CHUNK, K), 128, 0, st>>>(
                s.map_tmp, s.blurred, md, full, full, kb, 1.0 / ((double)kb * kb), c->dm.class_mask, s.bmax);
            LAUNCH_CHECK(c, s);
            CU_TRY(c, cudaMemsetAsync(s.list_n, 0, (size_t)std::max(K, S) * 4, st));
            // reference loop bounds: lefts 0..cols-wx, tops 0..rows-2*wy+1 (HFTest.cpp:246-249)
            const int n_left = g.W - w + 1, n_top = g.H - 2 * w + 2;
            if (n_left > 0 && n_top > 0) {
                const BlockGrid bg = make_block_grid(full);
                nms_select_kernel<<<dim3((bg.by * bg.bx + NMS_SELECT_THREADS - 1) / NMS_SELECT_THREADS, 1, K), NMS_SELECT_THREADS, 0, st>>>(s.blurred, s.bmax, full, w, w, 0, n_left, 0,
                                                                                         n_top, s.list, s.list_n, c->dm.class_mask);
                LAUNCH_CHECK(c, s);
            }
            ObjectLimits lim;
            memset(&lim, 0, sizeof lim);
            for (int k = 0; k < K; ++k) {
                lim.max_loc[k] = c->objects[k].max_location_hypotheses;
                lim.should_detect[k] = seeks_class(c, k) ? 1 : 0;
            }
            select_centres_kernel<<<K, 256, 0, st>>>(s.list, s.list_n, lim, p.min_location_score_ratio, rv.centres, rv.active);
            LAUNCH_CHECK(c, s);
            break;
        }
        case HF6D_STAGE_POSE: {
            const size_t yp = (size_t)c->reg.ny * c->reg.np;
            const int max_yp = p.max_yaw_pitch_hypotheses, max_roll = p.max_roll_hypotheses;
            const int n_groups = (int)c->hf.groups.size();
            // Two implementations of pass A (HFTest.cpp:742-802).  The stream path needs the vote stream the vote kernel wrote
            // for this slot's current leaf table; sharded contexts, huge forests (stream over budget) and stage-isolated runs
            // that replaced the leaf table after voting use the enumeration path.  Both give identical accumulators.
            const bool stream_path = c->use_stream && s.stream_valid &&
                                     ((c->shard_world == 1 && c->pshard.world == 1 && !c->peer_on) || (c->peer_on && c->peer_streams));
            // clear only the accumulator slots a class can use (slot = class * HF6D_MAX_CENTRES + centre rank): one launch.
            // The (slot, group) counters are left zeroed by the stream path itself (its roll pass zeroes what it visited).
            {
                ClearPlan cp;
                memset(&cp, 0, sizeof cp);
                cp.base[0] = s.zacc;    cp.slot_bytes[0] = (unsigned long long)HF6D_Z_BINS * 8;
                cp.base[1] = s.ypacc;   cp.slot_bytes[1] = (unsigned long long)yp * 8;
                cp.base[2] = s.racc;    cp.slot_bytes[2] = (unsigned long long)max_yp * HF6D_POSE_BINS * 8;
                cp.base[3] = stream_path ? nullptr : s.win_cnt; cp.slot_bytes[3] = (unsigned long long)n_groups * 4;
                bool any = false;
                for (int k = 0; k < K; ++k) {
                    cp.n_slots[k] = seeks_class(c, k) ? c->objects[k].max_location_hypotheses : 0;
                    any |= cp.n_slots[k] > 0;
                }
                if (any) {
                    clear_accumulators_kernel<<<dim3(c->sms, K, stream_path ? 3 : 4), 256, 0, st>>>(cp);
                    LAUNCH_CHECK(c, s);
                }
                if (stream_path && s.cnt_dirty) {  // the enumeration path (or nothing yet) used the table last: all of it, once
                    CU_TRY(c, cudaMemsetAsync(s.win_cnt, 0, (size_t)S * n_groups * 4, st));
                    s.cnt_dirty = false;
                }
                if (!stream_path) s.cnt_dirty = true;
            }
            CU_TRY(c, cudaMemsetAsync(s.list_n, 0, ((size_t)std::max(K, S) + 8) * 4, st));
            const long long items = (long long)g.cap * f.T;
            CentreTable ct{rv.centres, rv.active};
            const int half_win = p.centers_nms_wsize / 2;
            const size_t cell_bytes = cell_grid_bytes(g.W, g.H, half_win, K);
            const int table_blocks = c->sms * 8;
            int* ctr = s.list_n + std::max(K, S);
            ZSlotTable zt;
            memset(&zt, 0, sizeof zt);
            for (int k = 0; k < K; ++k)
                zt.zoff[k + 1] = (int16_t)(zt.zoff[k] + (seeks_class(c, k) ? c->objects[k].max_location_hypotheses : 0));
            const size_t zbytes = (size_t)zt.zoff[K] * HF6D_Z_BINS * 4;
            const PairList pl{s.pairs, ctr + 5, c->pair_cap};
            if (stream_path) {
                StreamSet vs;
                memset(&vs, 0, sizeof vs);
                vs.cap = c->stream_cap;
                if (c->peer_on) {  // every rank's stream: the votes of the whole frame, each once
                    const int si = slot_index(c, s);
                    vs.world = c->peer_world;
                    for (int r = 0; r < c->peer_world; ++r) {
                        vs.rec[r] = r == c->peer_rank ? s.vstream : c->peer_vstream[r][si];
                        vs.n[r] = r == c->peer_rank ? s.stream_n : c->peer_stream_n[r][si];
                    }
                } else {
                    vs.world = 1;
                    vs.rec[0] = s.vstream;
                    vs.n[0] = s.stream_n;
                }
                const size_t dyn = ((zbytes + 15) & ~(size_t)15) + cell_bytes;
#define HF6D_WS(GG, RR)                                                                                                         \
    window_stream_kernel<GG, RR><<<c->sms, WA_THREADS, dyn, st>>>(f, g, switches_of(c, true), vs, s.depth, ct, half_win, n_groups, zt, \
                                                             s.win_cnt, pl, s.zacc, ctr + 3, ctr + 4, 20 /* HFTest.cpp:745 */,     \
                                                             rv.active, rv.mode_z, c->ws_kmax)
                if (c->ws_rows == 2) { if (c->lanes_per_hit == 16) HF6D_WS(16, 2); else HF6D_WS(32, 2); }
                else { if (c->lanes_per_hit == 16) HF6D_WS(16, 4); else HF6D_WS(32, 4); }
#undef HF6D_WS
                LAUNCH_CHECK(c, s);
                if (c->lanes_per_hit == 16)
                    yawpitch_from_pairs_kernel<16><<<table_blocks, TABLE_THREADS, 0, st>>>(f, s.win_cnt, n_groups, pl, rv.active, c->reg, s.ypacc);
                else
                    yawpitch_from_pairs_kernel<32><<<table_blocks, TABLE_THREADS, 0, st>>>(f, s.win_cnt, n_groups, pl, rv.active, c->reg, s.ypacc);
                LAUNCH_CHECK(c, s);
            } else {
                // one resident wave; batches of 32 items are handed out through list_n[n_lists], list space is reserved
                // through list_n[n_lists + 1] (both zeroed above)
                const int wc_blocks = (int)std::min<long long>((items + VOTE_THREADS - 1) / VOTE_THREADS, (long long)c->sms * c->wc_ctas_per_sm);
                window_entries_kernel<<<wc_blocks, VOTE_THREADS, cell_bytes, st>>>(f, g, switches_of(c, true), s.locs, s.depth,
                                                                                 leaf_tables_of(c, s, slot_index(c, s)),
                                                                                 s.counts, ct, half_win, n_groups, ctr, s.entries,
                                                                                 c->entry_cap, ctr + 1, s.win_cnt, s.zacc);
                LAUNCH_CHECK(c, s);
                if (c->entry_cap >= ENTRY_BLOCK) {
                    static const bool smem_z_ok = !(getenv("HF6D_WA_SMEM_Z") && atoi(getenv("HF6D_WA_SMEM_Z")) == 0);  // tuning override
                    const bool zs = smem_z_ok && zbytes <= (size_t)WA_MAX_DYN_SMEM;
                    const int wa_grid = c->sms * (zs ? (int)std::max<size_t>(1, std::min<size_t>(2, (200 * 1024) / (zbytes + 24 * 1024))) : 2);
#define HF6D_WA(SZ, GG)                                                                                                     \
    window_accumulate_kernel<SZ, GG><<<wa_grid, WA_THREADS, (SZ) ? zbytes : 0, st>>>(f, s.entries, c->entry_cap, ctr + 1, n_groups, \
                                                                                   zt, s.win_cnt, s.zacc, ctr + 2)
                    if (zs) { if (c->lanes_per_hit == 16) HF6D_WA(true, 16); else HF6D_WA(true, 32); }
                    else { if (c->lanes_per_hit == 16) HF6D_WA(false, 16); else HF6D_WA(false, 32); }
#undef HF6D_WA
                    LAUNCH_CHECK(c, s);
                }
                if (c->lanes_per_hit == 16)
                    yawpitch_from_counts_kernel<16><<<table_blocks, TABLE_THREADS, 0, st>>>(f, s.win_cnt, n_groups, rv.active, S, c->reg, s.ypacc);
                else
                    yawpitch_from_counts_kernel<32><<<table_blocks, TABLE_THREADS, 0, st>>>(f, s.win_cnt, n_groups, rv.active, S, c->reg, s.ypacc);
                LAUNCH_CHECK(c, s);
                z_mode_kernel<<<S, 128, 0, st>>>(s.zacc, 20 /* HFTest.cpp:745 */, rv.active, rv.mode_z);
                LAUNCH_CHECK(c, s);
            }
            const MapDims md{HF6D_POSE_BINS, HF6D_POSE_BINS};
            const MapRect rin{c->reg.y0, c->reg.p0, c->reg.ny, c->reg.np};
            const MapRect rout = c->yp_blur;
            const int kb = p.pose_blur_size, w = p.pose_nms_wsize;
            box_rows_kernel<<<dim3((rin.nr + BLUR_WARPS - 1) / BLUR_WARPS, S), BLUR_WARPS * 32,
                              box_rows_smem_bytes(rin.nc), st>>>(s.ypacc, s.yptmp, md, rin, rout, kb, rv.active,
                                                                           PeerMaps{{}, 0});
            LAUNCH_CHECK(c, s);
            box_cols_kernel<<<dim3((rout.nc + 127) / 128, (rout.nr + BLUR_COL_CHUNK - 1) / BLUR_COL_CHUNK, S), 128, 0, st>>>(
                s.yptmp, s.ypblur, md, rin, rout, kb, 1.0 / ((double)kb * kb), rv.active, s.bmax);
            LAUNCH_CHECK(c, s);
            if (c->yp_nleft > 0 && c->yp_ntop > 0) {
                const BlockGrid bg = make_block_grid(rout);
                nms_select_kernel<<<dim3((bg.by * bg.bx + NMS_SELECT_THREADS - 1) / NMS_SELECT_THREADS, 1, S), NMS_SELECT_THREADS, 0, st>>>(
                    s.ypblur, s.bmax, rout, w, w, c->yp_left0, c->yp_nleft, c->yp_top0, c->yp_ntop, s.list, s.list_n, rv.active);
                LAUNCH_CHECK(c, s);
            }
            select_peaks_kernel<<<S, 256, 0, st>>>(s.list, s.list_n, rv.active, max_yp, p.min_yaw_pitch_drop_ratio,
                                                   rv.n_peaks, rv.peak_yx, rv.peak_score);
            LAUNCH_CHECK(c, s);
            PeakTable pk{rv.n_peaks, rv.peak_yx, max_yp};
            if (stream_path) {
                if (c->lanes_per_hit == 16)
                    roll_from_pairs_kernel<16><<<table_blocks, TABLE_THREADS, 0, st>>>(f, s.win_cnt, n_groups, pl, pk, p.pose_blur_size / 2, s.racc);
                else
                    roll_from_pairs_kernel<32><<<table_blocks, TABLE_THREADS, 0, st>>>(f, s.win_cnt, n_groups, pl, pk, p.pose_blur_size / 2, s.racc);
            } else if (c->lanes_per_hit == 16)
                roll_from_counts_kernel<16><<<table_blocks, TABLE_THREADS, 0, st>>>(f, s.win_cnt, n_groups, S, pk, p.pose_blur_size / 2, s.racc);
            else
                roll_from_counts_kernel<32><<<table_blocks, TABLE_THREADS, 0, st>>>(f, s.win_cnt, n_groups, S, pk, p.pose_blur_size / 2, s.racc);
            LAUNCH_CHECK(c, s);
            roll_modes_kernel<<<dim3(S, max_yp), 128, 0, st>>>(s.racc, rv.n_peaks, max_yp, p.pose_blur_size,
                                                               p.pose_nms_wsize, max_roll, c->dm.sep_ok, rv.records);
            LAUNCH_CHECK(c, s);
            break;
        }
        default:
            return fail(c, HF6D_EINVAL, "unknown stage %d", stage);
    }
    return HF6D_OK;
}

int run_range_impl(hf6d_ctx* c, Slot& s, int first, int last);

// One whole frame as a CUDA graph.  Capture needs nothing but the launches themselves: every kernel reads its sizes from
// device memory (patch count, centre lists, pair list), so the graph of one frame is the graph of every frame with the same
// frame buffers and the same context configuration.  Event records are left out (hf6d_stage_ms / hf6d_encoder_layer_ms
// report zeros for replayed frames).
int capture_frame(hf6d_ctx* c, Slot& s) {
    if (s.graph_exec) {
        cudaGraphExecDestroy(s.graph_exec);
        s.graph_exec = nullptr;
    }
    CU_TRY(c, cudaStreamBeginCapture(s.stream, cudaStreamCaptureModeThreadLocal));
    s.capturing = true;
    const int r = run_range_impl(c, s, 0, HF6D_STAGE_COUNT - 1);
    s.capturing = false;
    cudaGraph_t graph = nullptr;
    const cudaError_t e = cudaStreamEndCapture(s.stream, &graph);
    if (r || e != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        if (r) return r;
        return fail(c, HF6D_ECUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
    }
    const cudaError_t ei = cudaGraphInstantiate(&s.graph_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ei != cudaSuccess) {
        s.graph_exec = nullptr;
        return fail(c, HF6D_ECUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(ei));
    }
    s.graph_bgr = s.bgr;
    s.graph_depth = s.depth;
    s.graph_epoch = c->epoch;
    s.graph_launches = s.launches;
    return HF6D_OK;
}

int run_range(hf6d_ctx* c, Slot& s, int first, int last) {
    const bool whole = first == 0 && last == HF6D_STAGE_COUNT - 1;
    if (whole && c->use_graph && !c->peer_on && c->shard_world == 1 && c->pshard.world == 1) {
        // replay only what an eager frame has just established: same buffers, same configuration, and the host-side state the
        // launches depend on (vote stream valid, counter table clean, feature storage) left by a whole frame
        if (s.last_full_valid && s.last_bgr == s.bgr && s.last_depth == s.depth && s.last_epoch == c->epoch && !s.cnt_dirty) {
            if (!s.graph_exec || s.graph_bgr != s.bgr || s.graph_depth != s.depth || s.graph_epoch != c->epoch) {
                if (capture_frame(c, s) != HF6D_OK) {  // this context cannot capture: stay eager from here on
                    c->use_graph = false;
                    s.last_full_valid = false;
                    return run_range_impl(c, s, first, last);
                }
            }
            CU_TRY(c, cudaGraphLaunch(s.graph_exec, s.stream));
            s.launches = s.graph_launches;
            for (int i = 0; i <= HF6D_STAGE_COUNT; ++i) s.ev_valid[i] = false;
            s.ev_enc_valid = false;
            return HF6D_OK;
        }
        const int r = run_range_impl(c, s, first, last);
        s.last_full_valid = r == HF6D_OK;
        s.last_bgr = s.bgr;
        s.last_depth = s.depth;
        s.last_epoch = c->epoch;
        return r;
    }
    s.last_full_valid = false;
    const uint32_t seq0 = s.peer_seq;
    const int r = run_range_impl(c, s, first, last);
    if (r && c->peer_on) {
        // a frame that was not enqueued completely must not count: drain what was enqueued (bounded: a peer may never answer)
        // and put the slot's sequence number back, so that this rank stays in step with its peers
        const std::string err = c->err;
        sync_stream_bounded(c, s.stream);
        s.peer_seq = seq0;
        c->err = err;
    }
    return r;
}

int run_range_impl(hf6d_ctx* c, Slot& s, int first, int last) {
    if (first < 0 || last >= HF6D_STAGE_COUNT || first > last) return fail(c, HF6D_EINVAL, "bad stage range %d..%d", first, last);
    s.launches = 0;
    for (int i = 0; i <= HF6D_STAGE_COUNT; ++i) s.ev_valid[i] = false;
    if (!s.capturing) {
        CU_TRY(c, cudaEventRecord(s.ev[first], s.stream));
        s.ev_valid[first] = true;
    }
    const int slot = slot_index(c, s);
    for (int st = first; st <= last; ++st) {
        int r;
        if (c->peer_on && st == HF6D_STAGE_TRAVERSE) {  // the peers must be done with this slot's previous frame
            if ((r = peer_wait(c, s, slot, PEER_FLAG_CONSUMED, s.peer_seq))) return r;
            ++s.peer_seq;
        }
        if (c->peer_on && st == HF6D_STAGE_CENTRES && (r = peer_wait(c, s, slot, PEER_FLAG_READY, s.peer_seq))) return r;
        if ((r = run_stage(c, s, st))) return r;
        if (c->peer_on && st == HF6D_STAGE_VOTE && (r = peer_signal(c, s, slot, PEER_FLAG_READY, s.peer_seq))) return r;
        if (c->peer_on && st == HF6D_STAGE_POSE && (r = peer_signal(c, s, slot, PEER_FLAG_CONSUMED, s.peer_seq))) return r;
        if (!s.capturing) {
            CU_TRY(c, cudaEventRecord(s.ev[st + 1], s.stream));
            s.ev_valid[st + 1] = true;
        }
    }
    return HF6D_OK;
}

int collect_host(hf6d_ctx* c, Slot& s, hf6d_hypothesis* out, int cap, int* n_out) {
    CU_TRY(c, cudaMemcpyAsync(s.res_host, s.res_dev, c->rl.total, cudaMemcpyDeviceToHost, s.stream));
    {
        const int r = sync_stream_bounded(c, s.stream);
        if (r) return r;
    }
    ResView rv = view_of(c, s.res_host);
    const int K = c->hf.K, max_yp = c->p.max_yaw_pitch_hypotheses, max_roll = c->p.max_roll_hypotheses;
    int n = 0;
    for (int cls = 0; cls < K; ++cls) {
        const hf6d_centre_list& cl = rv.centres[cls];
        for (int k = 0; k < cl.n; ++k) {
            const int s_ = cls * HF6D_MAX_CENTRES + k;
            if (!rv.active[s_]) continue;
            int per_centre = 0;
            for (int h = 0; h < rv.n_peaks[s_]; ++h)
                for (int r = 0; r < max_roll; ++r) {
                    const HypRecord& rec = rv.records[((size_t)s_ * max_yp + h) * max_roll + r];
                    if (!rec.valid) continue;
                    if (per_centre >= HF6D_MAX_HYPOTHESES_PER_CENTRE) continue;
                    ++per_centre;
                    if (n < cap && out) {
                        hf6d_hypothesis& o = out[n];
                        o.cls = cls;
                        o.cx = cl.c[k].x;
                        o.cy = cl.c[k].y;
                        o.z = rv.mode_z[s_];
                        o.yaw_deg = rv.peak_yx[(s_ * max_yp + h) * 2] - 360;
                        o.pitch_deg = rv.peak_yx[(s_ * max_yp + h) * 2 + 1] - 360;
                        o.roll_deg = rec.roll_bin - 360;
                        o.loc_score = cl.c[k].score;
                        o.yawpitch_score = rv.peak_score[s_ * max_yp + h];
                        o.roll_score = rec.roll_score;
                        hf6d_pose_from_tuple(&c->p, o.cx, o.cy, o.z, o.yaw_deg, o.pitch_deg, o.roll_deg, o.pose);
                    }
                    ++n;
                }
        }
    }
    if (n_out) *n_out = n;
    return HF6D_OK;
}

int check_slot(hf6d_ctx* c, int slot) {
    if (!c) return HF6D_EINVAL;
    if (slot < 0 || slot >= c->n_slots) return fail(c, HF6D_EINVAL, "slot %d out of range [0,%d)", slot, c->n_slots);
    return HF6D_OK;
}

// Feature storage 1: the fp32 view of the slot's features (what hf6d_fetch / hf6d_device_ptr / hf6d_encode_patches hand out)
// is produced on demand from the fp16 rows the traversal reads -- an exact widening, so the caller sees the values the leaf
// tests compared.
__global__ void widen_features_kernel(const __half* __restrict__ src, const int* __restrict__ counts, int cap, int F,
                                      float* __restrict__ dst) {
    const long long n = (long long)min(counts[1], cap) * F / 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float2 v = __half22float2(reinterpret_cast<const __half2*>(src)[i]);
        reinterpret_cast<float2*>(dst)[i] = v;
    }
}
int features_fp32(hf6d_ctx* c, Slot& s) {
    if (!s.feat_is16) return HF6D_OK;
    widen_features_kernel<<<c->sms * 8, 256, 0, s.stream>>>(s.feat16, s.counts, c->g.cap, c->hf.F, s.feat);
    CU_TRY(c, cudaGetLastError());
    CU_TRY(c, cudaStreamSynchronize(s.stream));
    return HF6D_OK;
}

struct BufInfo {
    void* ptr;
    size_t bytes;  // capacity
};
int buffer_of(hf6d_ctx* c, Slot& s, int what, BufInfo& b) {
    const FrameGeom& g = c->g;
    const size_t HW = (size_t)g.W * g.H;
    switch (what) {
        case HF6D_BUF_COUNTS: b = {s.counts, 8}; break;
        case HF6D_BUF_LOCS: b = {s.locs, (size_t)g.cap * 8}; break;
        case HF6D_BUF_PATCH_U8: b = {s.q_u8, (size_t)g.cap * c->dm.n_in[0]}; break;
        case HF6D_BUF_NORMALS:
            if (!s.normals) return fail(c, HF6D_ESTATE, "surface normals exist only in patch_mode 1");
            b = {s.normals, HW * 16};
            break;
        case HF6D_BUF_FEATURES: b = {s.feat, (size_t)g.cap * c->hf.F * 4}; break;
        case HF6D_BUF_LEAF_ORD: b = {s.leaf_ord, (size_t)g.cap * c->hf.T * 4}; break;
        case HF6D_BUF_MAPS: b = {s.maps, HW * c->hf.K * 8}; break;
        case HF6D_BUF_BLURRED: b = {s.blurred, HW * c->hf.K * 4}; break;
        case HF6D_BUF_CENTRES: b = {s.res_dev + c->rl.centres, (size_t)c->hf.K * sizeof(hf6d_centre_list)}; break;
        case HF6D_BUF_FRAME_BGR: b = {s.bgr, HW * 3}; break;
        case HF6D_BUF_FRAME_DEPTH: b = {s.depth, HW * 2}; break;
        default: return fail(c, HF6D_EINVAL, "unknown buffer %d", what);
    }
    return HF6D_OK;
}

int ensure_feat16(hf6d_ctx* c);

int finish_create(hf6d_ctx* c, int device, int n_slots) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        return fail(c, HF6D_ECUDA, "no CUDA device available (libhf6d has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(c, HF6D_EINVAL, "device %d out of range (%d devices)", device, ndev);
    cudaDeviceProp prop;
    CU_TRY(c, cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(c, HF6D_ECUDA, "device %d is sm_%d%d; libhf6d is built for sm_100a only", device, prop.major, prop.minor);
    CU_TRY(c, cudaSetDevice(device));
    c->device = device;
    c->sms = prop.multiProcessorCount;
    c->n_slots = std::max(1, n_slots);

    hf6d_params& p = c->p;
    p.patch_vox = c->hf.ps;  // forest.txt wins (HFBase.cpp:118-119)
    p.voxel_m = c->hf.vox;
    if (p.W <= 0 || p.H <= 0 || p.W > 65535 || p.H > 65535) return fail(c, HF6D_EINVAL, "bad frame size %dx%d", p.W, p.H);
    if (p.stride <= 0) return fail(c, HF6D_EINVAL, "Stride should be more than 0");
    if (p.batch_size <= 0) return fail(c, HF6D_EINVAL, "batch_size must be positive");
    if (p.patch_vox != 8) return fail(c, HF6D_EINVAL, "patch_size_in_voxels = %d: only 8 (the 256-input encoder) is supported", p.patch_vox);
    if (c->layers.size() != 3) return fail(c, HF6D_EINVAL, "encoder must have 3 layers");
    if (p.patch_mode != 0 && p.patch_mode != 1) return fail(c, HF6D_EINVAL, "patch_mode must be 0 (RGB-D) or 1 (RGB + normals)");
    if (p.patch_mode == 1 && !(p.normals_focal > 0.f)) return fail(c, HF6D_EINVAL, "normals_focal must be positive");
    {
        const int nch = p.patch_mode == 1 ? 6 : 4;
        if (c->layers[0].in != nch * p.patch_vox * p.patch_vox)
            return fail(c, HF6D_EINVAL, "encoder input %d != %d*patch^2 = %d", c->layers[0].in, nch, nch * p.patch_vox * p.patch_vox);
    }
    if (c->layers[1].in != c->layers[0].out || c->layers[2].in != c->layers[1].out)
        return fail(c, HF6D_EINVAL, "encoder layer shapes do not chain");
    if (c->layers[2].out != c->hf.F)  // HFTest.cpp:595-596
        return fail(c, HF6D_EINVAL, "Output vector size of caffe net (%d) and Input vector size of forest (%d) do not match", c->layers[2].out, c->hf.F);
    if (c->hf.F % 4) return fail(c, HF6D_EINVAL, "feature length must be a multiple of 4");
    if (p.max_yaw_pitch_hypotheses < 0 || p.max_yaw_pitch_hypotheses > MAX_YP || p.max_roll_hypotheses < 0 || p.max_roll_hypotheses > MAX_ROLL)
        return fail(c, HF6D_EINVAL, "max_yaw_pitch_hypotheses / max_roll_hypotheses out of range");
    if (p.centers_blur_size < 1 || p.pose_blur_size < 1 || p.centers_nms_wsize < 1 || p.pose_nms_wsize < 1 ||
        p.centers_nms_wsize > 128 || p.pose_nms_wsize > 128)
        return fail(c, HF6D_EINVAL, "blur / NMS window sizes out of range");

    FrameGeom& g = c->g;
    g.W = p.W; g.H = p.H; g.stride = p.stride;
    g.gw = (p.W + p.stride - 1) / p.stride;
    g.gh = (p.H + p.stride - 1) / p.stride;
    g.fx = p.fx; g.fy = p.fy; g.cx = p.cx; g.cy = p.cy;
    g.focal = p.patch_mode == 1 ? p.normals_focal : p.fx;  // HFTest.cpp:394 vs :356
    g.ps = p.patch_vox; g.vox = p.voxel_m; g.range = p.max_depth_range_m; g.dist_thr = p.distance_threshold_m;
    g.fill_random = p.fill_random; g.fill_seed = p.fill_seed; g.batch = p.batch_size;
    g.cap = round_up(g.gw * g.gh, 128);

    const int K = c->hf.K;
    c->S = K * HF6D_MAX_CENTRES;
    // Yaw/pitch accumulator rectangle: bins that (a) can influence a kept peak -- [180,540] +- nms/2 +- blur/2 -- and
    // (b) can receive a vote at all: the bounding box of every vote copy of the loaded forest (HFTest.cpp:778-791).
    {
        const int bh = p.pose_blur_size / 2 + 1, nh = p.pose_nms_wsize / 2 + 1;
        const int need_lo = std::max(0, 180 - nh - bh), need_hi = std::min(HF6D_POSE_BINS - 1, 540 + nh + bh);
        int ylo = INT_MAX, yhi = INT_MIN, plo = INT_MAX, phi = INT_MIN;
        const HostForest& hf = c->hf;
        for (size_t i = 0; i < hf.yaw.size(); ++i) {
            const int yaw = hf.yaw[i], pit = hf.pitch[i];
            const int sy = yaw < 0 ? -1 : 1, sp = pit < 0 ? -1 : 1;
            for (int k1 = 0; k1 < 2; ++k1) {
                const int Y = yaw - sy * k1 * 360 + 360;
                if (Y >= need_lo && Y <= need_hi) { ylo = std::min(ylo, Y); yhi = std::max(yhi, Y); }
                const int P = pit - sp * k1 * 360 + 360;
                if (P >= need_lo && P <= need_hi) { plo = std::min(plo, P); phi = std::max(phi, P); }
            }
        }
        if (ylo > yhi) { ylo = yhi = need_lo; }
        if (plo > phi) { plo = phi = need_lo; }
        c->reg = PoseRegion{ylo, yhi - ylo + 1, plo, phi - plo + 1};
        // blurred values can be non-zero up to blur/2 away from the data; peaks are kept only in [180,540] and the
        // NMS window reads nms/2 further
        const int b_lo = std::max(0, 180 - nh), b_hi = std::min(HF6D_POSE_BINS - 1, 540 + nh);
        const int by0 = std::max(b_lo, ylo - bh), by1 = std::min(b_hi, yhi + bh);
        const int bp0 = std::max(b_lo, plo - bh), bp1 = std::min(b_hi, phi + bh);
        c->yp_blur = MapRect{by0, bp0, std::max(1, by1 - by0 + 1), std::max(1, bp1 - bp0 + 1)};
        // window origins: centres in [180,540] (HFTest.cpp:825-834) that can be non-zero, within the reference's loop
        // bounds lefts 0..720-w, tops 0..720-2w+1
        const int w = p.pose_nms_wsize;
        const int cx0 = std::max(180, bp0), cx1 = std::min(540, bp1), cy0 = std::max(180, by0), cy1 = std::min(540, by1);
        const int l0 = std::max(cx0 - w / 2, 0), l1 = std::min(cx1 - w / 2, HF6D_POSE_BINS - w);
        const int t0 = std::max(cy0 - w / 2, 0), t1 = std::min(cy1 - w / 2, HF6D_POSE_BINS - 2 * w + 1);
        c->yp_left0 = l0; c->yp_nleft = std::max(0, l1 - l0 + 1);
        c->yp_top0 = t0; c->yp_ntop = std::max(0, t1 - t0 + 1);
    }

    // (slot, vote group) entry counters of the pose stage
    {
        int max_group_votes = 1;
        for (const VoteGroup& vg : c->hf.groups) max_group_votes = std::max(max_group_votes, vg.vcnt);
        c->lanes_per_hit = max_group_votes <= 16 ? 16 : 32;
        if ((long long)c->S * (long long)c->hf.groups.size() > (1LL << 30))
            return fail(c, HF6D_ENOMEM, "forest too large: %zu vote groups x %d centre slots exceeds the counter table budget",
                        c->hf.groups.size(), c->S);
    }

    if ((int)c->objects.size() != K) {
        c->objects.assign(K, hf6d_object{});
        for (int k = 0; k < K; ++k) {
            snprintf(c->objects[k].name, sizeof c->objects[k].name, "object%d", k);
            c->objects[k].should_detect = 1;
            c->objects[k].max_location_hypotheses = 12;
            c->objects[k].instances = 1;
        }
    }
    for (int k = 0; k < K; ++k)
        if (c->objects[k].max_location_hypotheses < 0 || c->objects[k].max_location_hypotheses > HF6D_MAX_CENTRES)
            return fail(c, HF6D_EINVAL, "max_location_hypotheses of object %d must be in [0, %d]", k, HF6D_MAX_CENTRES);

    // result block layout
    ResultLayout& rl = c->rl;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o += (bytes + 15) / 16 * 16; return at; };
    rl.counts = take(16);
    rl.centres = take((size_t)K * sizeof(hf6d_centre_list));
    rl.active = take((size_t)c->S);
    rl.mode_z = take((size_t)c->S * 4);
    rl.n_peaks = take((size_t)c->S * 4);
    const size_t n_yp = (size_t)std::max(1, p.max_yaw_pitch_hypotheses), n_roll = (size_t)std::max(1, p.max_roll_hypotheses);
    rl.peak_yx = take((size_t)c->S * n_yp * 2 * 4);
    rl.peak_score = take((size_t)c->S * n_yp * 4);
    rl.records = take((size_t)c->S * n_yp * n_roll * sizeof(HypRecord));
    rl.total = o;

    int r = upload_model(c);
    if (r) return r;

    // opt in to large dynamic shared memory once
    CU_TRY(c, cudaFuncSetAttribute(traverse_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)traverse_plan(c->hf.F, c->hf.T).smem));
    CU_TRY(c, cudaFuncSetAttribute(traverse_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)traverse_plan(c->hf.F, c->hf.T).smem));
    CU_TRY(c, cudaFuncSetAttribute(traverse_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)traverse_plan(c->hf.F, c->hf.T).smem));
    if (c->hf.F % 8 == 0) {  // the fp16-row instances (feature storage 1)
        const int sm16 = (int)traverse_plan(c->hf.F, c->hf.T, 2).smem;
        CU_TRY(c, cudaFuncSetAttribute(traverse_kernel<1, __half>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm16));
        CU_TRY(c, cudaFuncSetAttribute(traverse_kernel<2, __half>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm16));
        CU_TRY(c, cudaFuncSetAttribute(traverse_kernel<4, __half>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm16));
    }
    if (cell_grid_bytes(p.W, p.H, p.centers_nms_wsize / 2, K) > 96 * 1024)
        return fail(c, HF6D_EINVAL, "frame too large for the centre-window lookup grid");
    if ((long long)g.cap * c->hf.T >= (1LL << 31)) return fail(c, HF6D_EINVAL, "patches x trees exceeds 2^31");
    CU_TRY(c, cudaFuncSetAttribute(window_entries_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)cell_grid_bytes(p.W, p.H, p.centers_nms_wsize / 2, K)));
    CU_TRY(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->wc_ctas_per_sm, window_entries_kernel, VOTE_THREADS,
                                                             cell_grid_bytes(p.W, p.H, p.centers_nms_wsize / 2, K)));
    c->wc_ctas_per_sm = std::max(1, c->wc_ctas_per_sm);
    CU_TRY(c, cudaFuncSetAttribute(box_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)box_rows_smem_bytes(std::max(p.W, HF6D_POSE_BINS))));

    {
        // worst case: every cast vote is an entry; budget 64 MB per frame slot (4 M entries), the overflow path stays exact
        long long max_leaf_votes = 1;
        for (size_t i = 0; i < c->hf.leaf_vcnt.size(); ++i) max_leaf_votes = std::max<long long>(max_leaf_votes, c->hf.leaf_vcnt[i]);
        const long long worst = ((long long)g.cap * c->hf.T * max_leaf_votes + ENTRY_BLOCK - 1) / ENTRY_BLOCK * ENTRY_BLOCK +
                                (long long)c->sms * 64 * ENTRY_BLOCK;  // + one partly filled block per resident warp
        long long cap_entries = std::min<long long>(worst, 4LL << 20);
        if (const char* e = getenv("HF6D_ENTRY_CAP")) cap_entries = std::max(0LL, atoll(e));  // tests force the overflow path
        c->entry_cap = (int)cap_entries;
    }
    {
        // Texel tile of the gather: the footprints of GT_PATCHES consecutive centres of a row at the closest range planned
        // for (0.5 m); segments that need more (closer surfaces, rows broken by holes) read their taps from global memory.
        const int amax = (int)((float)p.patch_vox * p.voxel_m / 0.5f * g.focal);
        const long long w = (long long)(GT_PATCHES - 1) * p.stride + amax + 2, h = amax + 3;
        long long bytes = std::max<long long>((w | 1) * h * 8, GT_VAL_BYTES);
        bytes = std::min<long long>(bytes, 100 * 1024);
        c->gather_smem = (int)bytes;
        c->gather_tile_texels = (int)(bytes / 8);
        const char* e = getenv("HF6D_GATHER");
        c->gather_tiled = !(e && !strcmp(e, "direct"));
        CU_TRY(c, cudaFuncSetAttribute(gather_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, c->gather_smem));
    }
    {
        // Vote stream for the pose stage: worst case one record per (patch, tree, vote of the largest leaf), one pair per (slot,
        // group).  Budget 512 MB / 256 MB per frame slot; frames beyond the 13-bit pixel fields, forests beyond the budget
        // and HF6D_POSE_STREAM=0 (tuning switch; the tests run both paths) use the enumeration path.
        long long max_leaf_votes = 1;
        for (size_t i = 0; i < c->hf.leaf_vcnt.size(); ++i) max_leaf_votes = std::max<long long>(max_leaf_votes, c->hf.leaf_vcnt[i]);
        const long long records = (long long)g.cap * c->hf.T * max_leaf_votes;
        const long long pairs = (long long)c->S * (long long)c->hf.groups.size();
        size_t zmax = 0;
        for (int k = 0; k < K; ++k) zmax += (size_t)HF6D_MAX_CENTRES * HF6D_Z_BINS * 4;
        const size_t dyn = zmax + cell_grid_bytes(p.W, p.H, p.centers_nms_wsize / 2, K) + 16;
        const char* e = getenv("HF6D_POSE_STREAM");
        c->use_stream = !(e && atoi(e) == 0) && p.W + 2 * STREAM_BIAS <= STREAM_COORD_MAX && p.H + 2 * STREAM_BIAS <= STREAM_COORD_MAX &&
                        p.centers_nms_wsize / 2 <= STREAM_BIAS && records <= (64LL << 20) && pairs <= (32LL << 20) &&
                        dyn <= (size_t)WS_MAX_DYN_SMEM;
        if (const char* t = getenv("HF6D_WS_TUNE")) {
            int a = 0, b = 0;
            if (sscanf(t, "%d,%d", &a, &b) == 2 && (a == 2 || a == 4) && b >= 1 && b <= 64) {
                c->ws_rows = a;
                c->ws_kmax = b;
            }
        }
        c->stream_cap = c->use_stream ? (int)std::max<long long>(records, 1) : 0;
        c->pair_cap = c->use_stream ? (int)std::max<long long>(pairs, 1) : 0;
        if (c->use_stream) {
            CU_TRY(c, (cudaFuncSetAttribute(window_stream_kernel<16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_MAX_DYN_SMEM)));
            CU_TRY(c, (cudaFuncSetAttribute(window_stream_kernel<32, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_MAX_DYN_SMEM)));
            CU_TRY(c, (cudaFuncSetAttribute(window_stream_kernel<16, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_MAX_DYN_SMEM)));
            CU_TRY(c, (cudaFuncSetAttribute(window_stream_kernel<32, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_MAX_DYN_SMEM)));
        }
    }
    CU_TRY(c, cudaFuncSetAttribute(window_accumulate_kernel<true, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, WA_MAX_DYN_SMEM));
    CU_TRY(c, cudaFuncSetAttribute(window_accumulate_kernel<true, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, WA_MAX_DYN_SMEM));
    // Two sets of encoder kernels (HF6D_ENC_CONFIGS): a context with several frame slots is built for pipelining, where the
    // smaller shared-memory footprints (variant 0) let other frames' CTAs run beside the persistent encoder CTA; a
    // one-slot context runs one frame at a time, where the deeper operand rings (variant 2) are simply faster.
    for (int l = 0; l < 3; ++l) {
        const int cls = encoder_shape_class(c->dm.block_n[l], l == 2, c->dm.k_pad[l] / ENC_BLOCK_K <= 6);
        c->enc_variant[l] = (c->n_slots == 1 && (cls == 0 || cls == 1 || cls == 4)) ? 2 : 0;
    }
    if (const char* e = getenv("HF6D_ENC_VARIANT")) {
        int v[3];
        if (sscanf(e, "%d,%d,%d", &v[0], &v[1], &v[2]) == 3)
            for (int l = 0; l < 3; ++l) c->enc_variant[l] = v[l];
    }
    for (int l = 0; l < 3; ++l) {  // CTA pairs need co-schedulable clusters: fall back to stand-alone CTAs if they are not
        EncoderLayerLaunch probe{};
        probe.block_n = c->dm.block_n[l];
        probe.last = l == 2;
        probe.short_k = c->dm.k_pad[l] / ENC_BLOCK_K <= 6;
        probe.variant = c->enc_variant[l];
        if (launch_encoder_layer(probe, nullptr, c->sms, nullptr, true) != cudaSuccess) {
            cudaGetLastError();
            if (c->enc_variant[l] == 1) return fail(c, HF6D_ECUDA, "encoder layer %d: no kernel variant can run on this device", l);
            c->enc_variant[l] = 1;
            probe.variant = 1;
            if (launch_encoder_layer(probe, nullptr, c->sms, nullptr, true) != cudaSuccess)
                return fail(c, HF6D_ECUDA, "encoder layer %d: no kernel variant can run on this device", l);
        }
    }
    c->slots.resize(c->n_slots);
    for (Slot& s : c->slots) {
        memset(s.ev, 0, sizeof s.ev);
        if ((r = alloc_slot(c, s))) return r;
    }
    c->use_graph = getenv("HF6D_GRAPH") && atoi(getenv("HF6D_GRAPH")) != 0;
    {   // feature storage: fp16 rows wherever the feature layer has a kernel for them; HF6D_FEATURES=fp32 keeps the fp32 rows
        const char* e = getenv("HF6D_FEATURES");
        if (!(e && !strcmp(e, "fp32"))) {
            if (ensure_feat16(c) == HF6D_OK) c->feat16 = true;
            else if (e && !strcmp(e, "fp16")) return HF6D_EINVAL;  // asked for explicitly: say why not (c->err)
            else c->err.clear();
        }
    }
    CU_TRY(c, cudaDeviceSynchronize());
    return HF6D_OK;
}

void params_from_options(const HostOptions& o, hf6d_params& p) {
    p.stride = o.stride;
    p.fx = o.fx; p.fy = o.fy; p.cx = o.cx; p.cy = o.cy;
    p.max_depth_range_m = o.max_depth_range;
    p.distance_threshold_m = o.distance_threshold;
    p.fill_random = o.are_objects_segmented ? 0 : 1;  // HFTest.cpp:1235
    p.batch_size = o.batch_size;
}

}  // namespace

// ==================================================================================================== C ABI
extern "C" {

void hf6d_default_params(hf6d_params* p) {
    memset(p, 0, sizeof *p);
    p->W = 640; p->H = 480;
    p->stride = 2;  // generate_scripts.sh:53; HFTest.h:98
    p->fx = 575.f; p->fy = 575.f; p->cx = 319.5f; p->cy = 239.5f;
    p->patch_vox = 8; p->voxel_m = 0.005f;
    p->max_depth_range_m = 0.25f;
    p->distance_threshold_m = 1.5f;
    p->fill_random = 1; p->fill_seed = 0;
    p->batch_size = 100;
    p->max_yaw_pitch_hypotheses = 7;  // HFTest.h:175-183
    p->max_roll_hypotheses = 3;
    p->min_location_score_ratio = 1.0f / 1000.0f;
    p->min_yaw_pitch_drop_ratio = 1.0f / 1000.0f;
    p->centers_blur_size = 13; p->centers_nms_wsize = 40;
    p->pose_blur_size = 35; p->pose_nms_wsize = 35;
    p->patch_mode = 0;
    p->normals_focal = 575.0f;  // HFTest.cpp:329, :356
}

const char* hf6d_last_error(const hf6d_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int hf6d_create(const hf6d_params* p, const char* forest_dir, const char* weights_path, int device, int n_slots,
                hf6d_ctx** out) {
    if (!p || !forest_dir || !weights_path || !out) return fail(nullptr, HF6D_EINVAL, "null argument");
    *out = nullptr;
    hf6d_ctx* c = new hf6d_ctx();
    c->p = *p;
    std::string err;
    int r = HF6D_OK;
    if (!load_forest(forest_dir, c->hf, err)) r = fail(nullptr, HF6D_EIO, "%s", err.c_str());
    else if (!load_weights(weights_path, c->layers, err)) r = fail(nullptr, HF6D_EIO, "%s", err.c_str());
    else {
        r = finish_create(c, device, n_slots);
        if (r) g_create_error = c->err;
    }
    if (r) { free_all(c); delete c; return r; }
    *out = c;
    return HF6D_OK;
}

int hf6d_create_from_options(const char* options_path, int W, int H, int device, int n_slots, hf6d_ctx** out) {
    if (!options_path || !out) return fail(nullptr, HF6D_EINVAL, "null argument");
    *out = nullptr;
    HostOptions o;
    std::string err;
    if (!load_options(options_path, o, err)) return fail(nullptr, HF6D_EIO, "%s", err.c_str());
    hf6d_params p;
    hf6d_default_params(&p);
    p.W = W; p.H = H;
    params_from_options(o, p);
    hf6d_ctx* c = new hf6d_ctx();
    c->p = p;
    c->objects = o.objects;
    c->mesh_files = o.mesh_files;
    c->obj_nn_search_radius = o.nn_search_radius;
    c->obj_icp_iterations = o.icp_iterations;
    c->option_refine = o.refine;
    c->has_option_refine = true;
    int r = HF6D_OK;
    if (!load_forest(o.forest_folder, c->hf, err)) r = fail(nullptr, HF6D_EIO, "%s", err.c_str());
    else if (!load_weights(o.caffe_weights, c->layers, err)) r = fail(nullptr, HF6D_EIO, "%s", err.c_str());
    else if ((int)o.objects.size() != c->hf.K)  // HFTest.cpp:1186
        r = fail(nullptr, HF6D_EINVAL, "Number of objects provided in the options file (%zu) does not match the number of classes in the forest (%d)", o.objects.size(), c->hf.K);
    else {
        r = finish_create(c, device >= 0 ? device : std::max(o.gpu, 0), n_slots);
        if (r) g_create_error = c->err;
    }
    if (r) { free_all(c); delete c; return r; }
    *out = c;
    return HF6D_OK;
}

int hf6d_parse_options(const char* options_path, hf6d_options* out, hf6d_object* objs, int cap) {
    if (!options_path || !out) return fail(nullptr, HF6D_EINVAL, "null argument");
    HostOptions o;
    std::string err;
    if (!load_options(options_path, o, err)) return fail(nullptr, HF6D_EIO, "%s", err.c_str());
    memset(out, 0, sizeof *out);
    hf6d_default_params(&out->params);
    params_from_options(o, out->params);
    out->gpu = o.gpu;
    out->n_objects = (int32_t)o.objects.size();
    out->location_score_coeff = o.location_score_coeff;
    out->pose_score_coeff = o.pose_score_coeff;
    snprintf(out->forest_folder, sizeof out->forest_folder, "%s", o.forest_folder.c_str());
    snprintf(out->caffe_weights, sizeof out->caffe_weights, "%s", o.caffe_weights.c_str());
    snprintf(out->caffe_definition, sizeof out->caffe_definition, "%s", o.caffe_definition.c_str());
    for (int k = 0; k < (int)o.objects.size() && k < cap && objs; ++k) objs[k] = o.objects[k];
    return HF6D_OK;
}

int hf6d_inspect_forest(const char* forest_dir, hf6d_model_info* out) {
    if (!forest_dir || !out) return fail(nullptr, HF6D_EINVAL, "null argument");
    HostForest hf;
    std::string err;
    if (!load_forest(forest_dir, hf, err)) return fail(nullptr, HF6D_EIO, "%s", err.c_str());
    memset(out, 0, sizeof *out);
    out->T = hf.T; out->K = hf.K; out->F = hf.F; out->patch_vox = hf.ps;
    out->voxel_m = hf.vox;
    out->n_leaves = (int64_t)hf.leaf_id.size();
    out->n_internal = hf.n_internal;
    out->n_votes = (int64_t)hf.ox.size();
    out->max_depth = hf.max_depth;
    return HF6D_OK;
}

int hf6d_inspect_weights(const char* weights_path, int32_t dims[4]) {
    if (!weights_path || !dims) return fail(nullptr, HF6D_EINVAL, "null argument");
    std::vector<HostLayer> layers;
    std::string err;
    if (!load_weights(weights_path, layers, err)) return fail(nullptr, HF6D_EIO, "%s", err.c_str());
    if (layers.size() != 3) return fail(nullptr, HF6D_EINVAL, "encoder must have 3 layers, file has %zu", layers.size());
    dims[0] = layers[0].in;
    for (int l = 0; l < 3; ++l) dims[l + 1] = layers[l].out;
    return HF6D_OK;
}

void hf6d_destroy(hf6d_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    if (c->peer_on || !c->peer_opened.empty()) hf6d_peer_detach(c);
    if (c->peer_flags) cudaFree(c->peer_flags);
    if (c->peer_timeout) cudaFree(c->peer_timeout);
    rf_free(c->rf);
    free_all(c);
    delete c;
}

int hf6d_get_params(const hf6d_ctx* c, hf6d_params* out) {
    if (!c || !out) return HF6D_EINVAL;
    *out = c->p;
    return HF6D_OK;
}

int hf6d_model(const hf6d_ctx* c, hf6d_model_info* out) {
    if (!c || !out) return HF6D_EINVAL;
    out->T = c->hf.T; out->K = c->hf.K; out->F = c->hf.F; out->patch_vox = c->hf.ps;
    out->voxel_m = c->hf.vox;
    out->n_leaves = (int64_t)c->hf.leaf_id.size();
    out->n_internal = c->hf.n_internal;
    out->n_votes = (int64_t)c->hf.ox.size();
    out->max_depth = c->hf.max_depth;
    out->dims[0] = c->layers[0].in;
    for (int l = 0; l < 3; ++l) out->dims[l + 1] = c->layers[l].out;
    return HF6D_OK;
}

int hf6d_set_objects(hf6d_ctx* c, const hf6d_object* objs, int n) {
    if (!c || !objs) return HF6D_EINVAL;
    ++c->epoch;  // a captured frame (HF6D_GRAPH) no longer describes this context
    if (n != c->hf.K) return fail(c, HF6D_EINVAL, "Number of objects (%d) does not match the number of classes in the forest (%d)", n, c->hf.K);
    for (int k = 0; k < n; ++k)
        if (objs[k].max_location_hypotheses < 0 || objs[k].max_location_hypotheses > HF6D_MAX_CENTRES)
            return fail(c, HF6D_EINVAL, "max_location_hypotheses of object %d must be in [0, %d]", k, HF6D_MAX_CENTRES);
    c->objects.assign(objs, objs + n);
    return upload_class_mask(c);
}

int hf6d_get_objects(const hf6d_ctx* c, hf6d_object* objs, int cap) {
    if (!c) return HF6D_EINVAL;
    for (int k = 0; k < (int)c->objects.size() && k < cap && objs; ++k) objs[k] = c->objects[k];
    return (int)c->objects.size();
}

int hf6d_set_fill_seed(hf6d_ctx* c, uint64_t seed) {
    if (!c) return HF6D_EINVAL;
    ++c->epoch;  // a captured frame (HF6D_GRAPH) no longer describes this context
    c->p.fill_seed = seed;
    c->g.fill_seed = seed;
    return HF6D_OK;
}

int hf6d_set_tree_shard(hf6d_ctx* c, int rank, int world) {
    if (!c) return HF6D_EINVAL;
    ++c->epoch;
    if (world < 1 || rank < 0 || rank >= world) return fail(c, HF6D_EINVAL, "bad tree shard %d/%d", rank, world);
    c->shard_rank = rank;
    c->shard_world = world;
    for (Slot& s : c->slots) s.stream_valid = false;
    return HF6D_OK;
}

int hf6d_set_patch_shard(hf6d_ctx* c, int rank, int world) {
    if (!c) return HF6D_EINVAL;
    ++c->epoch;
    if (world < 1 || rank < 0 || rank >= world) return fail(c, HF6D_EINVAL, "bad patch shard %d/%d", rank, world);
    c->pshard = PatchShard{rank, world};
    for (Slot& s : c->slots) s.stream_valid = false;
    return HF6D_OK;
}

int hf6d_set_peer_split(hf6d_ctx* c, int split) {
    if (!c) return HF6D_EINVAL;
    ++c->epoch;
    if (split != 0 && split != 1) return fail(c, HF6D_EINVAL, "peer split %d: 0 = trees, 1 = patches", split);
    if (c->peer_on) return fail(c, HF6D_ESTATE, "hf6d_set_peer_split must be called before hf6d_peer_attach");
    c->peer_split = split;
    return HF6D_OK;
}

int hf6d_set_class_shard(hf6d_ctx* c, int rank, int world) {
    if (!c) return HF6D_EINVAL;
    ++c->epoch;
    if (world < 1 || rank < 0 || rank >= world) return fail(c, HF6D_EINVAL, "bad class shard %d/%d", rank, world);
    c->class_rank = rank;
    c->class_world = world;
    return upload_class_mask(c);
}

namespace {
struct PeerBlob {  // what a rank publishes: POD, exchanged by the caller (torch.distributed all_gather, MPI, a file ...)
    uint32_t magic;
    int32_t n_slots, W, H, K, T, cap, device;
    cudaIpcMemHandle_t flags;
    cudaIpcMemHandle_t maps[HF6D_MAX_SLOTS];
    cudaIpcMemHandle_t leaf[HF6D_MAX_SLOTS];
    int32_t stream_cap;  // 0: this rank has no vote stream (then nobody uses the streams)
    cudaIpcMemHandle_t vstream[HF6D_MAX_SLOTS];
    cudaIpcMemHandle_t stream_n[HF6D_MAX_SLOTS];
};
constexpr uint32_t PEER_MAGIC = 0x36644650u;  // "PFd6"
}  // namespace

size_t hf6d_peer_blob_bytes(void) { return sizeof(PeerBlob); }

int hf6d_peer_export(hf6d_ctx* c, void* blob, size_t cap_bytes) {
    if (!c || !blob) return HF6D_EINVAL;
    if (cap_bytes < sizeof(PeerBlob)) return fail(c, HF6D_EINVAL, "peer blob needs %zu bytes", sizeof(PeerBlob));
    if (c->n_slots > HF6D_MAX_SLOTS) return fail(c, HF6D_EINVAL, "peer exchange supports at most %d slots", HF6D_MAX_SLOTS);
    CU_TRY(c, cudaSetDevice(c->device));
    if (!c->peer_flags) {
        const size_t n = (size_t)c->n_slots * 2 * HF6D_MAX_PEERS;
        CU_TRY(c, cudaMalloc(reinterpret_cast<void**>(&c->peer_flags), n * 4));
        CU_TRY(c, cudaMemset(c->peer_flags, 0, n * 4));
        CU_TRY(c, cudaMalloc(reinterpret_cast<void**>(&c->peer_timeout), 4));
        CU_TRY(c, cudaMemset(c->peer_timeout, 0, 4));
    }
    PeerBlob b;
    memset(&b, 0, sizeof b);
    b.magic = PEER_MAGIC;
    b.n_slots = c->n_slots; b.W = c->g.W; b.H = c->g.H; b.K = c->hf.K; b.T = c->hf.T; b.cap = c->g.cap; b.device = c->device;
    CU_TRY(c, cudaIpcGetMemHandle(&b.flags, c->peer_flags));
    for (int i = 0; i < c->n_slots; ++i) {
        CU_TRY(c, cudaIpcGetMemHandle(&b.maps[i], c->slots[i].maps));
        CU_TRY(c, cudaIpcGetMemHandle(&b.leaf[i], c->slots[i].leaf_ord));
        if (c->use_stream) {
            CU_TRY(c, cudaIpcGetMemHandle(&b.vstream[i], c->slots[i].vstream));
            CU_TRY(c, cudaIpcGetMemHandle(&b.stream_n[i], c->slots[i].stream_n));
        }
    }
    b.stream_cap = c->use_stream ? c->stream_cap : 0;
    memcpy(blob, &b, sizeof b);
    return HF6D_OK;
}

int hf6d_peer_detach(hf6d_ctx* c) {
    if (!c) return HF6D_EINVAL;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (void* p : c->peer_opened) cudaIpcCloseMemHandle(p);
    c->peer_opened.clear();
    for (int r = 0; r < HF6D_MAX_PEERS; ++r) {
        c->peer_maps[r].clear();
        c->peer_leaf[r].clear();
        c->peer_flags_of[r] = nullptr;
    }
    c->peer_on = false;
    c->peer_world = 1;
    c->peer_rank = 0;
    return HF6D_OK;
}

int hf6d_peer_attach(hf6d_ctx* c, int rank, int world, const void* blobs, size_t bytes_each) {
    if (!c || !blobs) return HF6D_EINVAL;
    if (world < 2 || world > HF6D_MAX_PEERS || rank < 0 || rank >= world) return fail(c, HF6D_EINVAL, "bad peer group %d/%d", rank, world);
    if (bytes_each != sizeof(PeerBlob)) return fail(c, HF6D_EINVAL, "peer blob size %zu, expected %zu", bytes_each, sizeof(PeerBlob));
    if (!c->peer_flags) return fail(c, HF6D_ESTATE, "hf6d_peer_export must be called before hf6d_peer_attach");
    if (c->peer_on) hf6d_peer_detach(c);
    CU_TRY(c, cudaSetDevice(c->device));
    {
        // A slot that waits for a peer's flag blocks its hardware queue; if two slots' streams alias one queue, the frame the
        // peer is waiting for can be stuck behind that wait (include/hf6d.h, peer contract).  The number of queues is fixed when
        // the driver initialises, from CUDA_DEVICE_MAX_CONNECTIONS (default 8): with more slots than queues the exchange is
        // refused here instead of deadlocking later.
        const char* e = getenv("CUDA_DEVICE_MAX_CONNECTIONS");
        const int queues = e && atoi(e) > 0 ? atoi(e) : 8;
        if (c->n_slots > 1 && c->n_slots + 1 > queues)
            return fail(c, HF6D_ESTATE, "peer exchange with %d frame slots needs CUDA_DEVICE_MAX_CONNECTIONS >= %d in the environment "
                        "before CUDA initialises (it is %d): streams that share a hardware queue can deadlock on the peer flags",
                        c->n_slots, c->n_slots + 1, queues);
    }
    const PeerBlob* all = static_cast<const PeerBlob*>(blobs);
    for (int r = 0; r < world; ++r) {
        const PeerBlob& b = all[r];
        if (b.magic != PEER_MAGIC || b.n_slots != c->n_slots || b.W != c->g.W || b.H != c->g.H || b.K != c->hf.K ||
            b.T != c->hf.T || b.cap != c->g.cap)
            return fail(c, HF6D_EINVAL, "rank %d runs a different configuration (slots / frame size / forest)", r);
    }
    // the pose stage reads every rank's vote stream in place when all of them keep one of the same capacity; otherwise it
    // enumerates the votes again from the ranks' leaf tables (HF6D_PEER_STREAMS=0 forces that, for the tests)
    c->peer_streams = c->use_stream && !(getenv("HF6D_PEER_STREAMS") && atoi(getenv("HF6D_PEER_STREAMS")) == 0);
    for (int r = 0; r < world; ++r) c->peer_streams = c->peer_streams && all[r].stream_cap == c->stream_cap;
    for (int r = 0; r < world; ++r) {
        c->peer_maps[r].assign(c->n_slots, nullptr);
        c->peer_leaf[r].assign(c->n_slots, nullptr);
        c->peer_vstream[r].assign(c->n_slots, nullptr);
        c->peer_stream_n[r].assign(c->n_slots, nullptr);
        if (r == rank) {
            c->peer_flags_of[r] = c->peer_flags;
            continue;
        }
        const PeerBlob& b = all[r];
        if (b.device != c->device) {
            int can = 0;
            CU_TRY(c, cudaDeviceCanAccessPeer(&can, c->device, b.device));
            if (!can) return fail(c, HF6D_ECUDA, "device %d cannot access device %d (no NVLink / P2P path)", c->device, b.device);
        }
        void* p = nullptr;
        CU_TRY(c, cudaIpcOpenMemHandle(&p, b.flags, cudaIpcMemLazyEnablePeerAccess));
        c->peer_opened.push_back(p);
        c->peer_flags_of[r] = static_cast<uint32_t*>(p);
        for (int i = 0; i < c->n_slots; ++i) {
            CU_TRY(c, cudaIpcOpenMemHandle(&p, b.maps[i], cudaIpcMemLazyEnablePeerAccess));
            c->peer_opened.push_back(p);
            c->peer_maps[r][i] = static_cast<const unsigned long long*>(p);
            CU_TRY(c, cudaIpcOpenMemHandle(&p, b.leaf[i], cudaIpcMemLazyEnablePeerAccess));
            c->peer_opened.push_back(p);
            c->peer_leaf[r][i] = static_cast<const int*>(p);
            if (c->peer_streams) {
                CU_TRY(c, cudaIpcOpenMemHandle(&p, b.vstream[i], cudaIpcMemLazyEnablePeerAccess));
                c->peer_opened.push_back(p);
                c->peer_vstream[r][i] = static_cast<const uint2*>(p);
                CU_TRY(c, cudaIpcOpenMemHandle(&p, b.stream_n[i], cudaIpcMemLazyEnablePeerAccess));
                c->peer_opened.push_back(p);
                c->peer_stream_n[r][i] = static_cast<const int*>(p);
            }
        }
    }
    c->peer_rank = rank;
    c->peer_world = world;
    c->peer_on = true;
    // The flag block is NOT cleared here and the slots' sequence numbers keep counting: a peer that attached earlier may
    // already have signalled into this rank's flags, and erasing that would leave this rank waiting for ever.  The flags were
    // zeroed once when the block was created (hf6d_peer_export); every rank of a group counts the same frames, so the numbers
    // stay in step across detach / attach cycles of the whole group.
    int r;
    if (c->peer_split == 1) {
        if ((r = hf6d_set_tree_shard(c, 0, 1))) return r;
        if ((r = hf6d_set_patch_shard(c, rank, world))) return r;
    } else {
        if ((r = hf6d_set_patch_shard(c, 0, 1))) return r;
        if ((r = hf6d_set_tree_shard(c, rank, world))) return r;
    }
    return hf6d_set_class_shard(c, rank, world);
}

int hf6d_peer_timed_out(hf6d_ctx* c) {
    if (!c || !c->peer_timeout) return 0;
    if (c->peer_wait_expired) return 1;
    int v = 0;
    cudaSetDevice(c->device);
    if (cudaMemcpy(&v, c->peer_timeout, 4, cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
    return v;
}

namespace {
// Split-bf16 encoder state: weights as (hi | lo) bf16 halves, per-slot (hi | lo) hidden activations and tensor maps.
int ensure_split_encoder(hf6d_ctx* c) {
    DeviceModel& dm = c->dm;
    int r;
    if (!dm.Ws[0]) {
        for (int l = 0; l < 3; ++l) {
            const HostLayer& L = c->layers[l];
            const size_t kp = (size_t)dm.k_pad[l];
            std::vector<__nv_bfloat16> w((size_t)dm.n_pad[l] * 2 * kp, __float2bfloat16(0.f));
            for (int n = 0; n < L.out; ++n)
                for (int k = 0; k < L.in; ++k) {
                    float v = L.W[(size_t)n * L.in + k];
                    if (l == 0) v = v / 255.0f;  // the A operand of the first layer holds q itself
                    const __nv_bfloat16 hi = __float2bfloat16(v);
                    w[(size_t)n * 2 * kp + k] = hi;
                    w[(size_t)n * 2 * kp + kp + k] = __float2bfloat16(v - __bfloat162float(hi));
                }
            std::vector<float> b(dm.n_pad[l], 0.f);
            for (int n = 0; n < L.out; ++n) b[n] = L.b[n];
            const __nv_bfloat16* wp = nullptr;
            const float* bp = nullptr;
            if ((r = dev_upload(c, dm.allocs, &wp, w))) return r;
            if ((r = dev_upload(c, dm.allocs, &bp, b))) return r;
            dm.Ws[l] = const_cast<__nv_bfloat16*>(wp);
            dm.bs[l] = const_cast<float*>(bp);
        }
    }
    const FrameGeom& g = c->g;
    for (Slot& s : c->slots) {
        if (s.split_ready) continue;
        if ((r = dev_alloc(c, s.allocs, &s.H1s, (size_t)g.cap * 2 * dm.n_pad[0]))) return r;
        if ((r = dev_alloc(c, s.allocs, &s.H2s, (size_t)g.cap * 2 * dm.n_pad[1]))) return r;
        const void* a_in[3] = {s.A0, s.H1s, s.H2s};
        void* outs[3] = {s.H1s, s.H2s, s.feat};
        for (int l = 0; l < 3; ++l) {
            EncoderLayerLaunch& L = s.enc_split[l];
            memset(&L, 0, sizeof L);
            L.split = true;
            L.last = l == 2;
            L.variant = c->enc_variant[l] == 1 ? 1 : 0;  // stand-alone CTAs only where pairs cannot be scheduled
            L.block_n = dm.block_n[l];
            const EncoderConfig cfg = encoder_config(L.block_n, L.last, false, L.variant, true);
            if (!cfg.pair) return fail(c, HF6D_EINVAL, "split encoder: no kernel for layer %d (N tile %d, variant %d)", l, L.block_n, L.variant);
            const uint64_t a_cols = l == 0 ? (uint64_t)dm.k_pad[0] : 2ull * dm.k_pad[l];
            if (!make_bf16_kmajor_map(&L.tmA, a_in[l], (uint64_t)g.cap, a_cols, ENC_BLOCK_M) ||
                !make_bf16_kmajor_map(&L.tmB, dm.Ws[l], (uint64_t)dm.n_pad[l], 2ull * dm.k_pad[l], (uint32_t)(L.block_n / cfg.pair)) ||
                !make_out_map(&L.tmC, outs[l], (uint64_t)g.cap, (uint64_t)(L.last ? c->hf.F : 2 * dm.n_pad[l]), L.last ? 4 : 2, cfg.chunk_bytes))
                return fail(c, HF6D_ECUDA, "cuTensorMapEncodeTiled failed for split encoder layer %d", l);
            L.bias = dm.bs[l];
            L.K = dm.k_pad[l];
            L.n_pad = dm.n_pad[l];
            L.short_k = false;
            L.reverse_m = s.enc[l].reverse_m;
            L.n_seg = l == 0 ? 2 : 3;
            L.lo_off = L.last ? 0 : dm.n_pad[l];
            if (launch_encoder_layer(L, nullptr, c->sms, nullptr, true) != cudaSuccess) {
                cudaGetLastError();
                return fail(c, HF6D_ECUDA, "split encoder layer %d cannot run on this device", l);
            }
        }
        s.split_ready = true;
    }
    return HF6D_OK;
}
// fp16 encoder state: fp16 weights (same padded shapes as the bf16 ones) and per-slot launches that differ from the bf16 ones
// in the weight tensor map, the operand format and -- first layer -- the 1/255 applied to the accumulator instead of W.
int ensure_fp16_encoder(hf6d_ctx* c) {
    DeviceModel& dm = c->dm;
    int r;
    if (c->p.patch_mode == 1) return fail(c, HF6D_EINVAL, "encoder mode 2 is not available in patch_mode 1");
    if (!dm.Wh[0]) {
        for (int l = 0; l < 3; ++l) {
            const HostLayer& L = c->layers[l];
            std::vector<__half> w((size_t)dm.n_pad[l] * dm.k_pad[l], __float2half(0.f));
            for (int n = 0; n < L.out; ++n)
                for (int k = 0; k < L.in; ++k) w[(size_t)n * dm.k_pad[l] + k] = __float2half(L.W[(size_t)n * L.in + k]);
            const __half* wp = nullptr;
            if ((r = dev_upload(c, dm.allocs, &wp, w))) return r;
            dm.Wh[l] = const_cast<__half*>(wp);
        }
    }
    for (Slot& s : c->slots) {
        if (s.fp16_ready) continue;
        for (int l = 0; l < 3; ++l) {
            EncoderLayerLaunch& L = s.enc_fp16[l];
            L = s.enc[l];
            const EncoderConfig cfg = encoder_config(L.block_n, L.last, L.short_k, L.variant);
            if (!make_bf16_kmajor_map(&L.tmB, dm.Wh[l], (uint64_t)dm.n_pad[l], (uint64_t)dm.k_pad[l], (uint32_t)(L.block_n / cfg.pair)))
                return fail(c, HF6D_ECUDA, "cuTensorMapEncodeTiled failed for fp16 encoder layer %d", l);
            L.fp16 = 1;
            L.in_scale = l == 0 ? 1.0f / 255.0f : 1.0f;
        }
        s.fp16_ready = true;
    }
    return c->feat16_block_n ? ensure_feat16(c) : HF6D_OK;  // the fp16-operand feature layer with fp16 output, where that exists
}
}  // namespace

extern "C++" {
namespace {
// Feature storage 1: per-slot fp16 feature rows and the feature-layer launches that write them (bf16 operands; fp16 operands
// once encoder mode 2 has been set up).  Idempotent.
int ensure_feat16(hf6d_ctx* c) {
    DeviceModel& dm = c->dm;
    const FrameGeom& g = c->g;
    const int F = c->hf.F;
    int r;
    if (!c->feat16_block_n) {
        if (F % 8 || !(F % 160 == 0 || dm.block_n[2] == 256))
            return fail(c, HF6D_EINVAL, "fp16 feature rows need a feature length that is a multiple of 8 and of 160, or 256-wide feature tiles (F = %d)", F);
        const int block_n = F % 160 == 0 ? 160 : 256;
        EncoderLayerLaunch probe{};
        probe.block_n = block_n;
        probe.last = true;
        probe.out16 = true;
        probe.variant = c->enc_variant[2] == 1 ? 1 : (c->n_slots == 1 ? 2 : 0);
        if (const char* e = getenv("HF6D_FEAT16_VARIANT")) probe.variant = atoi(e);  // tuning override (row of HF6D_ENC_OUT16_CONFIGS)
        if (launch_encoder_layer(probe, nullptr, c->sms, nullptr, true) != cudaSuccess) {
            cudaGetLastError();
            probe.variant = 1;
            if (launch_encoder_layer(probe, nullptr, c->sms, nullptr, true) != cudaSuccess) {
                cudaGetLastError();
                return fail(c, HF6D_ECUDA, "fp16 feature rows: no kernel variant can run on this device");
            }
        }
        c->feat16_variant = probe.variant;
        c->feat16_block_n = block_n;
    }
    for (Slot& s : c->slots) {
        if (!s.feat16) {
            if ((r = dev_alloc(c, s.allocs, &s.feat16, (size_t)g.cap * F))) return r;
            EncoderLayerLaunch& L = s.enc16[0];
            L = s.enc[2];
            L.out16 = true;
            L.block_n = c->feat16_block_n;
            L.variant = c->feat16_variant;
            L.n_pad = round_up(F, L.block_n);
            const EncoderConfig cfg = encoder_config(L.block_n, true, false, L.variant, false, true);
            if (!cfg.pair) return fail(c, HF6D_EINVAL, "fp16 feature rows: no kernel variant %d for %d-wide tiles", L.variant, L.block_n);
            if (!make_bf16_kmajor_map(&L.tmB, dm.W[2], (uint64_t)dm.n_pad[2], (uint64_t)dm.k_pad[2], (uint32_t)(L.block_n / cfg.pair)) ||
                !make_out_map(&L.tmC, s.feat16, (uint64_t)g.cap, (uint64_t)F, 2, cfg.chunk_bytes))
                return fail(c, HF6D_ECUDA, "cuTensorMapEncodeTiled failed for the fp16 feature layer");
        }
        if (s.fp16_ready && !s.enc16_fp16_ready) {
            EncoderLayerLaunch& L = s.enc16[1];
            L = s.enc16[0];
            const EncoderConfig cfg = encoder_config(L.block_n, true, false, L.variant, false, true);
            if (!make_bf16_kmajor_map(&L.tmB, dm.Wh[2], (uint64_t)dm.n_pad[2], (uint64_t)dm.k_pad[2], (uint32_t)(L.block_n / cfg.pair)))
                return fail(c, HF6D_ECUDA, "cuTensorMapEncodeTiled failed for the fp16 feature layer (fp16 operands)");
            L.fp16 = 1;
            L.in_scale = 1.0f;
            s.enc16_fp16_ready = true;
        }
    }
    return HF6D_OK;
}
}  // namespace
}  // extern "C++" (declared before finish_create)

int hf6d_set_feature_storage(hf6d_ctx* c, int storage) {
    if (!c) return HF6D_EINVAL;
    ++c->epoch;  // a captured frame (HF6D_GRAPH) no longer describes this context
    if (storage != 0 && storage != 1) return fail(c, HF6D_EINVAL, "feature storage %d: 0 = fp32 rows, 1 = fp16 rows", storage);
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, cudaDeviceSynchronize());
    if (storage == 1) {
        const int r = ensure_feat16(c);
        if (r) return r;
    }
    c->feat16 = storage == 1;
    for (Slot& s : c->slots) s.stream_valid = false;
    return HF6D_OK;
}

int hf6d_get_feature_storage(const hf6d_ctx* c) { return c ? (c->feat16 ? 1 : 0) : HF6D_EINVAL; }

int hf6d_set_encoder_mode(hf6d_ctx* c, int mode) {
    if (!c) return HF6D_EINVAL;
    ++c->epoch;  // a captured frame (HF6D_GRAPH) no longer describes this context
    if (mode < 0 || mode > 2)
        return fail(c, HF6D_EINVAL, "encoder mode %d: 0 = bf16 operands, 1 = split bf16 (hi + lo), 2 = fp16 operands", mode);
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, cudaDeviceSynchronize());
    if (mode == 1) {
        const int r = ensure_split_encoder(c);
        if (r) return r;
    }
    if (mode == 2) {
        const int r = ensure_fp16_encoder(c);
        if (r) return r;
    }
    c->encoder_mode = mode;
    for (Slot& s : c->slots) s.stream_valid = false;
    return HF6D_OK;
}

int hf6d_get_encoder_mode(const hf6d_ctx* c) { return c ? c->encoder_mode : HF6D_EINVAL; }

int hf6d_set_debug_capture(hf6d_ctx* c, int on) {
    if (!c) return HF6D_EINVAL;
    ++c->epoch;  // a captured frame (HF6D_GRAPH) no longer describes this context
    c->debug_capture = on ? 1 : 0;
    return HF6D_OK;
}

void* hf6d_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
    return p;
}
void hf6d_host_free(void* p) { if (p) cudaFreeHost(p); }

int hf6d_upload(hf6d_ctx* c, int slot, const uint8_t* bgr, const uint16_t* depth_mm) {
    int r = check_slot(c, slot);
    if (r) return r;
    if (!bgr || !depth_mm) return fail(c, HF6D_EINVAL, "null frame");
    Slot& s = c->slots[slot];
    const size_t HW = (size_t)c->g.W * c->g.H;
    CU_TRY(c, cudaSetDevice(c->device));
    s.bgr = s.own_bgr;
    s.depth = s.own_depth;
    s.stream_valid = false;
    CU_TRY(c, cudaMemcpyAsync(s.bgr, bgr, HW * 3, cudaMemcpyHostToDevice, s.stream));
    CU_TRY(c, cudaMemcpyAsync(s.depth, depth_mm, HW * 2, cudaMemcpyHostToDevice, s.stream));
    return HF6D_OK;
}

int hf6d_bind_frame(hf6d_ctx* c, int slot, const void* d_bgr, const void* d_depth_mm) {
    int r = check_slot(c, slot);
    if (r) return r;
    Slot& s = c->slots[slot];
    if ((d_bgr == nullptr) != (d_depth_mm == nullptr)) return fail(c, HF6D_EINVAL, "bind both planes or neither");
    s.stream_valid = false;
    s.bgr = d_bgr ? const_cast<uint8_t*>(static_cast<const uint8_t*>(d_bgr)) : s.own_bgr;
    s.depth = d_depth_mm ? const_cast<uint16_t*>(static_cast<const uint16_t*>(d_depth_mm)) : s.own_depth;
    return HF6D_OK;
}

int hf6d_encoder_layer_ms(hf6d_ctx* c, int slot, float* ms) {
    int r = check_slot(c, slot);
    if (r) return r;
    Slot& s = c->slots[slot];
    CU_TRY(c, cudaStreamSynchronize(s.stream));
    for (int l = 0; l < 3; ++l) {
        ms[l] = 0.f;
        if (s.ev_enc_valid) cudaEventElapsedTime(&ms[l], s.ev_enc[l], s.ev_enc[l + 1]);
    }
    return HF6D_OK;
}

int hf6d_run(hf6d_ctx* c, int slot, int first_stage, int last_stage) {
    int r = check_slot(c, slot);
    if (r) return r;
    CU_TRY(c, cudaSetDevice(c->device));
    return run_range(c, c->slots[slot], first_stage, last_stage);
}

int hf6d_sync(hf6d_ctx* c, int slot) {
    int r = check_slot(c, slot);
    if (r) return r;
    return sync_stream_bounded(c, c->slots[slot].stream);
}

int hf6d_collect(hf6d_ctx* c, int slot, hf6d_hypothesis* out, int cap, int* n_out) {
    int r = check_slot(c, slot);
    if (r) return r;
    CU_TRY(c, cudaSetDevice(c->device));
    return collect_host(c, c->slots[slot], out, cap, n_out);
}

int hf6d_detect(hf6d_ctx* c, const uint8_t* bgr, const uint16_t* depth_mm, hf6d_hypothesis* out, int cap, int* n_out) {
    int r = hf6d_upload(c, 0, bgr, depth_mm);
    if (r) return r;
    if ((r = hf6d_run(c, 0, 0, HF6D_STAGE_COUNT - 1))) return r;
    return hf6d_collect(c, 0, out, cap, n_out);
}

int hf6d_submit(hf6d_ctx* c, const uint8_t* bgr, const uint16_t* depth_mm, int* ticket) {
    if (!c || !ticket) return HF6D_EINVAL;
    const int t = c->next_ticket;
    const int slot = t % c->n_slots;
    Slot& s = c->slots[slot];
    if (s.busy) return fail(c, HF6D_ESTATE, "all %d slots are in flight: hf6d_wait ticket %d first", c->n_slots, s.ticket);
    int r = hf6d_upload(c, slot, bgr, depth_mm);
    if (r) return r;
    if ((r = hf6d_run(c, slot, 0, HF6D_STAGE_COUNT - 1))) return r;
    CU_TRY(c, cudaMemcpyAsync(s.res_host, s.res_dev, c->rl.total, cudaMemcpyDeviceToHost, s.stream));
    s.busy = true;
    s.ticket = t;
    ++c->next_ticket;
    *ticket = t;
    return HF6D_OK;
}

int hf6d_wait(hf6d_ctx* c, int ticket, hf6d_hypothesis* out, int cap, int* n_out) {
    if (!c) return HF6D_EINVAL;
    if (ticket < 0) return fail(c, HF6D_ESTATE, "unknown ticket %d", ticket);
    Slot& s = c->slots[ticket % c->n_slots];
    if (!s.busy || s.ticket != ticket) return fail(c, HF6D_ESTATE, "unknown ticket %d", ticket);
    int r = collect_host(c, s, out, cap, n_out);
    s.busy = false;
    return r;
}

int64_t hf6d_fetch(hf6d_ctx* c, int slot, int what, void* dst, size_t cap_bytes) {
    int r = check_slot(c, slot);
    if (r) return r;
    Slot& s = c->slots[slot];
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, cudaStreamSynchronize(s.stream));
    BufInfo b;
    if ((r = buffer_of(c, s, what, b))) return r;
    if (what == HF6D_BUF_FEATURES && (r = features_fp32(c, s))) return r;
    size_t bytes = b.bytes;
    if (what == HF6D_BUF_LOCS || what == HF6D_BUF_PATCH_U8 || what == HF6D_BUF_FEATURES || what == HF6D_BUF_LEAF_ORD) {
        int counts[2];
        CU_TRY(c, cudaMemcpy(counts, s.counts, 8, cudaMemcpyDeviceToHost));
        const size_t P = (size_t)std::min(counts[0], c->g.cap), Pp = (size_t)std::min(counts[1], c->g.cap);
        if (what == HF6D_BUF_LOCS) bytes = P * 8;
        else if (what == HF6D_BUF_PATCH_U8) {
            if (!c->debug_capture) return fail(c, HF6D_ESTATE, "HF6D_BUF_PATCH_U8 needs hf6d_set_debug_capture(1) before the run");
            bytes = Pp * (size_t)c->dm.n_in[0];
        } else if (what == HF6D_BUF_FEATURES) bytes = Pp * c->hf.F * 4;
        else bytes = Pp * c->hf.T * 4;
    }
    if (bytes > cap_bytes) return fail(c, HF6D_EINVAL, "destination too small: need %zu bytes, have %zu", bytes, cap_bytes);
    if (bytes) CU_TRY(c, cudaMemcpy(dst, b.ptr, bytes, cudaMemcpyDeviceToHost));
    return (int64_t)bytes;
}

int hf6d_inject(hf6d_ctx* c, int slot, int what, const void* src, size_t bytes) {
    int r = check_slot(c, slot);
    if (r) return r;
    Slot& s = c->slots[slot];
    CU_TRY(c, cudaSetDevice(c->device));
    BufInfo b;
    if ((r = buffer_of(c, s, what, b))) return r;
    if (bytes > b.bytes) return fail(c, HF6D_EINVAL, "buffer %d holds %zu bytes, got %zu", what, b.bytes, bytes);
    CU_TRY(c, cudaStreamSynchronize(s.stream));
    CU_TRY(c, memcpy_h2d_done(b.ptr, src, bytes));
    if (what == HF6D_BUF_FEATURES) s.feat_is16 = false;  // injected features are fp32 rows: the traversal reads those
    s.last_full_valid = false;
    s.stream_valid = false;  // whatever was injected, the vote stream no longer describes the slot
    return HF6D_OK;
}

int hf6d_device_ptr(hf6d_ctx* c, int slot, int what, void** ptr, size_t* bytes) {
    int r = check_slot(c, slot);
    if (r) return r;
    BufInfo b;
    if ((r = buffer_of(c, c->slots[slot], what, b))) return r;
    if (what == HF6D_BUF_FEATURES) {  // the caller gets (and may overwrite) the fp32 rows: they are the slot's features from here on
        Slot& s = c->slots[slot];
        CU_TRY(c, cudaSetDevice(c->device));
        if ((r = features_fp32(c, s))) return r;
        s.feat_is16 = false;
    }
    c->slots[slot].last_full_valid = false;  // the caller may write through the pointer
    if (ptr) *ptr = b.ptr;
    if (bytes) *bytes = b.bytes;
    return HF6D_OK;
}

int hf6d_set_stream(hf6d_ctx* c, int slot, void* cuda_stream) {
    int r = check_slot(c, slot);
    if (r) return r;
    Slot& s = c->slots[slot];
    cudaStreamSynchronize(s.stream);
    s.stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : s.own_stream;
    return HF6D_OK;
}

int hf6d_stage_ms(hf6d_ctx* c, int slot, float* ms) {
    int r = check_slot(c, slot);
    if (r) return r;
    Slot& s = c->slots[slot];
    CU_TRY(c, cudaStreamSynchronize(s.stream));
    for (int st = 0; st < HF6D_STAGE_COUNT; ++st) {
        ms[st] = 0.f;
        if (s.ev_valid[st] && s.ev_valid[st + 1]) cudaEventElapsedTime(&ms[st], s.ev[st], s.ev[st + 1]);
    }
    return HF6D_OK;
}

int64_t hf6d_debug_texture_gather(hf6d_ctx* c, int slot, float* dst, size_t cap_bytes) {
    int r = check_slot(c, slot);
    if (r) return r;
    Slot& s = c->slots[slot];
    const FrameGeom& g = c->g;
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, cudaStreamSynchronize(s.stream));
    int counts[2];
    CU_TRY(c, cudaMemcpy(counts, s.counts, 8, cudaMemcpyDeviceToHost));
    const size_t Pp = (size_t)std::min(counts[1], g.cap);
    const size_t bytes = Pp * g.ps * g.ps * 4 * sizeof(float);
    if (bytes > cap_bytes) return fail(c, HF6D_EINVAL, "destination too small: need %zu bytes, have %zu", bytes, cap_bytes);
    if (!Pp) return 0;
    float* vol = nullptr;
    float* out = nullptr;
    cudaArray_t arr = nullptr;
    cudaTextureObject_t tex = 0;
    int rc = HF6D_OK;
    auto cleanup = [&]() {
        if (tex) cudaDestroyTextureObject(tex);
        if (arr) cudaFreeArray(arr);
        if (vol) cudaFree(vol);
        if (out) cudaFree(out);
    };
#define TX_TRY(expr)                                                                                   \
    do {                                                                                               \
        cudaError_t e_ = (expr);                                                                       \
        if (e_ != cudaSuccess) { rc = fail(c, HF6D_ECUDA, "%s: %s", #expr, cudaGetErrorString(e_)); cleanup(); return rc; } \
    } while (0)
    const size_t HW = (size_t)g.W * g.H;
    TX_TRY(cudaMalloc(&vol, HW * 4 * sizeof(float)));
    TX_TRY(cudaMalloc(&out, bytes));
    texture_build_kernel<<<(unsigned)((HW + 255) / 256), 256, 0, s.stream>>>(s.bgr, s.depth, g.W, g.H, vol);
    TX_TRY(cudaGetLastError());
    // extent (width = 4 channels, height = W, depth = H): patch_extractor.cu:322-337
    cudaChannelFormatDesc fd = cudaCreateChannelDesc<float>();
    const cudaExtent ext = make_cudaExtent(4, g.W, g.H);
    TX_TRY(cudaMalloc3DArray(&arr, &fd, ext));
    cudaMemcpy3DParms cp;
    memset(&cp, 0, sizeof cp);
    cp.srcPtr = make_cudaPitchedPtr(vol, 4 * sizeof(float), 4, g.W);
    cp.dstArray = arr;
    cp.extent = ext;
    cp.kind = cudaMemcpyDeviceToDevice;
    TX_TRY(cudaStreamSynchronize(s.stream));
    TX_TRY(cudaMemcpy3D(&cp));
    cudaResourceDesc rd;
    memset(&rd, 0, sizeof rd);
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = arr;
    cudaTextureDesc td;
    memset(&td, 0, sizeof td);
    td.normalizedCoords = 0;
    td.filterMode = cudaFilterModeLinear;
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeBorder;
    td.readMode = cudaReadModeElementType;
    TX_TRY(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    texture_gather_kernel<<<(unsigned)Pp, 64, 0, s.stream>>>(tex, g, s.locs, s.counts, out);
    TX_TRY(cudaGetLastError());
    TX_TRY(cudaStreamSynchronize(s.stream));
    TX_TRY(cudaMemcpy(dst, out, bytes, cudaMemcpyDeviceToHost));
#undef TX_TRY
    cleanup();
    return (int64_t)bytes;
}

int64_t hf6d_result_bytes(const hf6d_ctx* c) { return c ? (int64_t)c->rl.total : HF6D_EINVAL; }

int hf6d_launch_count(const hf6d_ctx* c, int slot) {
    if (!c || slot < 0 || slot >= c->n_slots) return HF6D_EINVAL;
    return c->slots[slot].launches;
}

int64_t hf6d_count_cast_votes(hf6d_ctx* c, int slot) {
    int r = check_slot(c, slot);
    if (r) return r;
    Slot& s = c->slots[slot];
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, cudaStreamSynchronize(s.stream));
    int counts[2];
    CU_TRY(c, cudaMemcpy(counts, s.counts, 8, cudaMemcpyDeviceToHost));
    const size_t Pp = (size_t)std::min(counts[1], c->g.cap), T = (size_t)c->hf.T;
    std::vector<int> ord(Pp * T);
    if (Pp) CU_TRY(c, cudaMemcpy(ord.data(), s.leaf_ord, Pp * T * 4, cudaMemcpyDeviceToHost));
    int64_t n = 0;
    for (size_t i = 0; i < Pp; ++i)
        for (size_t t = 0; t < T; ++t) {
            const int o = ord[i * T + t];
            if (o >= 0) n += c->hf.leaf_vcnt[(size_t)c->hf.leaf_base[t] + o];
        }
    return n;
}

void hf6d_pose_from_tuple(const hf6d_params* p, int cx, int cy, float z, int yaw_deg, int pitch_deg, int roll_deg,
                          float pose[16]) {
    const float yaw = (float)((float)yaw_deg / 180.0f * M_PI);
    const float pitch = (float)((float)pitch_deg / 180.0f * M_PI);
    const float roll = (float)((float)roll_deg / 180.0f * M_PI);
    float R[9];
    rot_from_ypr(yaw, pitch, roll, R);
    const float x = ((float)cx - p->cx) * z / p->fx;
    const float y = ((float)cy - p->cy) * z / p->fy;
    pose[0] = R[0]; pose[1] = R[1]; pose[2] = R[2]; pose[3] = x;
    pose[4] = R[3]; pose[5] = R[4]; pose[6] = R[5]; pose[7] = y;
    pose[8] = R[6]; pose[9] = R[7]; pose[10] = R[8]; pose[11] = z;
    pose[12] = 0; pose[13] = 0; pose[14] = 0; pose[15] = 1;
}

}  // extern "C"

#include "refine_api.inc"
#include "train_api.inc"
#include "render_api.inc"
#include "patchdb_api.inc"
