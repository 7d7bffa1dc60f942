// Host-side model loading for libhf6d: the reference's on-disk artefacts -> flat tables ready for the device.
//   forest.txt / tree<N>.dat        HoughForest/src/HFBase.cpp:58-145 (pre-order, left first, sizeof(bool)==1)
//   .caffemodel (V1 layers)         weights of generate_scripts.sh:424-524 (encode1..3), or the raw HF6DW001 container
//   text-format DetectorOptions     HoughForest/include/proto/detector_options.proto, HFTest.cpp:1155-1235
// Plain C++ (no CUDA), compiled with -ffp-contract=off so the vote pre-rotation below rounds exactly as written.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <climits>
#include <string>
#include <vector>

#include "../../include/hf6d.h"

namespace hf6d {

// ------------------------------------------------------------------------------------------------ forest
// Internal nodes only, breadth-first per tree, all trees concatenated.  A child entry e is either the global index of
// an internal node (e >= 0) or ~global_leaf (e < 0), so a descent never touches leaf payload.
struct PackedNode {
    uint32_t f1f2;   // feature1 | feature2 << 16 ; index F addresses a constant 0.0f slot (measure_mode 1 -> f2 = F)
    float thr;
    int32_t left, right;
};
static_assert(sizeof(PackedNode) == 16, "one node = one 16-byte load");

// Two tree levels per load: a record holds the test of an internal node at even depth and the tests of its two
// children, plus the four grandchild entries.  A missing child test (the child is a leaf) is "0 < +inf" -> left, with
// both entries of that side pointing at the leaf.  An entry is a record index (>= 0) or ~global_leaf (< 0).
struct PackedRecord {
    uint32_t f1f2[3];  // node, left child, right child
    float thr[3];
    int32_t next[4];   // [2 * (went right at the node) + (went right at the child)]
    uint32_t pad[2];
};
static_assert(sizeof(PackedRecord) == 48, "one record = three 16-byte loads");

struct VoteGroup {   // votes of one (leaf, class) pair that passes the class_prob >= 0.5 gate (HFTest.cpp:191)
    int32_t cls;
    uint32_t w;      // Q16 weight
    int32_t vbeg, vcnt;
};

struct HostForest {
    int T = 0, K = 0, F = 0, ps = 0;
    float vox = 0;
    int max_depth = 0;
    std::vector<PackedNode> nodes;
    std::vector<PackedRecord> recs;  // the same trees, two levels per record (what the traversal kernel walks)
    std::vector<int32_t> root;       // [T] entry of each tree's root (node numbering)
    std::vector<int32_t> rec_root;   // [T] entry of each tree's root (record numbering)
    std::vector<int32_t> leaf_base;  // [T+1] first global leaf of tree t (leaves in file order)
    std::vector<int32_t> leaf_id;    // [L] leaf_id field of the file
    std::vector<float> class_prob;   // [L][K]
    std::vector<int32_t> group_off;  // [L+1]
    std::vector<VoteGroup> groups;
    // per gated vote, SoA
    std::vector<float> ox, oy, oz;          // R(yaw,pitch,roll) * (-x,-y,-z): the patch-independent part of HFTest.cpp:41-102
    std::vector<int16_t> yaw, pitch, roll;  // integer-degree bins (HFTest.cpp:779-780, :863), saturated to +-30000
    std::vector<int32_t> vgroup;            // [votes] vote -> its group
    std::vector<int32_t> leaf_vbeg, leaf_vcnt;  // [L] all gated votes of a leaf are contiguous: first vote, count
    std::vector<float> g_ozmin, g_ozmax;        // [groups] smallest / largest oz of the group's votes (z-histogram shortcut)
    std::vector<float> oz_sorted;               // [votes] oz again, ascending within every group (counting votes per z bin)
    int64_t n_internal = 0;
};

inline int32_t d2i_x86(double y) {  // x86 cvttsd2si: NaN / overflow -> INT_MIN
    if (!(y == y) || y >= 2147483648.0 || y < -2147483648.0) return INT_MIN;
    return (int32_t)y;
}
inline int16_t sat16(int32_t v) { return (int16_t)(v > 30000 ? 30000 : (v < -30000 ? -30000 : v)); }

// corr * Rz(yaw) * Ry(pitch) * Rx(roll), 3x3 part; cos/sin in double on float arguments, narrowed to float
// (HFTest.cpp:45-80 / MeshUtils.cpp:29-59); products in the order a 4x4 Eigen product accumulates them.
inline void rot_from_ypr(float yaw, float pitch, float roll, float R[9]) {
    float cyw = (float)cos((double)yaw), syw = (float)sin((double)yaw);
    float cp = (float)cos((double)pitch), sp = (float)sin((double)pitch);
    float cr = (float)cos((double)roll), sr = (float)sin((double)roll);
    float a00 = cyw * cp, a01 = -syw, a02 = cyw * sp;
    float a10 = syw * cp, a11 = cyw, a12 = syw * sp;
    float a20 = -sp, a21 = 0.0f, a22 = cp;
    float nsr = -sr;
    R[0] = a00; R[1] = a01 * cr + a02 * sr; R[2] = a01 * nsr + a02 * cr;
    float b10 = a10, b11 = a11 * cr + a12 * sr, b12 = a11 * nsr + a12 * cr;
    float b20 = a20, b21 = a21 * cr + a22 * sr, b22 = a21 * nsr + a22 * cr;
    R[3] = -b10; R[4] = -b11; R[5] = -b12;
    R[6] = -b20; R[7] = -b21; R[8] = -b22;
}

struct FileBuf {
    std::vector<uint8_t> d;
    size_t pos = 0;
    bool load(const std::string& path) {
        FILE* fp = fopen(path.c_str(), "rb");
        if (!fp) return false;
        fseek(fp, 0, SEEK_END);
        long n = ftell(fp);
        fseek(fp, 0, SEEK_SET);
        d.resize(n > 0 ? (size_t)n : 0);
        bool ok = n <= 0 || fread(d.data(), 1, (size_t)n, fp) == (size_t)n;
        fclose(fp);
        pos = 0;
        return ok;
    }
    bool get(void* dst, size_t n) {
        if (pos + n > d.size()) return false;
        memcpy(dst, d.data() + pos, n);
        pos += n;
        return true;
    }
};

// Loads and flattens a forest.  On failure returns false with a message in err.
inline bool load_forest(const std::string& dir, HostForest& hf, std::string& err) {
    {
        FILE* fp = fopen((dir + "/forest.txt").c_str(), "r");
        if (!fp) { err = "cannot open " + dir + "/forest.txt"; return false; }
        int n = fscanf(fp, "%d %d %d %d %f", &hf.T, &hf.K, &hf.F, &hf.ps, &hf.vox);
        fclose(fp);
        if (n != 5 || hf.T <= 0 || hf.K <= 0 || hf.F <= 0 || hf.ps <= 0) { err = "malformed forest.txt"; return false; }
        if (hf.K > HF6D_MAX_CLASSES) { err = "too many classes"; return false; }
        if (hf.F >= 65535) { err = "feature vector too long for 16-bit feature indices"; return false; }
    }
    struct TmpNode { int32_t mode, f1, f2; float thr; int32_t child[2]; int32_t leaf; /* -1 or global leaf */ };
    hf.root.assign(hf.T, 0);
    hf.leaf_base.assign(hf.T + 1, 0);
    hf.group_off.assign(1, 0);
    for (int t = 0; t < hf.T; ++t) {
        FileBuf fb;
        std::string path = dir + "/tree" + std::to_string(t) + ".dat";
        if (!fb.load(path)) { err = "cannot read " + path; return false; }
        std::vector<TmpNode> tn;
        // iterative pre-order parse: stack of (parent, which child) slots waiting for a node
        std::vector<std::pair<int32_t, int>> pending;
        pending.push_back({-1, 0});
        int32_t root_idx = -1;
        std::vector<int> depth_of;
        while (!pending.empty()) {
            auto slot = pending.back();
            pending.pop_back();
            uint8_t leaf;
            if (!fb.get(&leaf, 1)) { err = "truncated " + path; return false; }
            TmpNode n{};
            n.child[0] = n.child[1] = -1;
            n.leaf = -1;
            int32_t idx = (int32_t)tn.size();
            int d = slot.first < 0 ? 0 : depth_of[slot.first] + 1;
            if (leaf) {
                int32_t lid;
                if (!fb.get(&lid, 4)) { err = "truncated " + path; return false; }
                int32_t g = (int32_t)hf.leaf_id.size();
                n.leaf = g;
                hf.leaf_id.push_back(lid);
                size_t cp0 = hf.class_prob.size();
                hf.class_prob.resize(cp0 + hf.K);
                if (!fb.get(&hf.class_prob[cp0], 4 * (size_t)hf.K)) { err = "truncated " + path; return false; }
                for (int c = 0; c < hf.K; ++c) {
                    int32_t nm;
                    if (!fb.get(&nm, 4) || nm < 0) { err = "bad vote count in " + path; return false; }
                    if ((size_t)nm * 24 > fb.d.size() - fb.pos) { err = "truncated " + path; return false; }
                    const float prob = hf.class_prob[cp0 + c];
                    const bool gated = prob >= 0.5f && nm > 0;
                    if (gated) {
                        VoteGroup g2;
                        g2.cls = c;
                        g2.w = (uint32_t)(prob * 65536.0f + 0.5f);
                        g2.vbeg = (int32_t)hf.ox.size();
                        g2.vcnt = nm;
                        hf.groups.push_back(g2);
                    }
                    for (int i = 0; i < nm; ++i) {
                        float v[6];
                        fb.get(v, 24);
                        if (!gated) continue;
                        float R[9];
                        rot_from_ypr(v[0], v[1], v[2], R);
                        const float vx = -v[3], vy = -v[4], vz = -v[5];
                        hf.vgroup.push_back((int32_t)hf.groups.size() - 1);
                        hf.ox.push_back((R[0] * vx + R[1] * vy) + R[2] * vz);
                        hf.oy.push_back((R[3] * vx + R[4] * vy) + R[5] * vz);
                        hf.oz.push_back((R[6] * vx + R[7] * vy) + R[8] * vz);
                        hf.yaw.push_back(sat16(d2i_x86((double)v[0] / M_PI * 180.0)));
                        hf.pitch.push_back(sat16(d2i_x86((double)v[1] / M_PI * 180.0)));
                        hf.roll.push_back(sat16(d2i_x86((double)(v[2] * 180.0f) / M_PI)));
                    }
                }
                hf.group_off.push_back((int32_t)hf.groups.size());
                {
                    const int32_t g0 = hf.group_off[hf.group_off.size() - 2], g1 = hf.group_off.back();
                    const int32_t vb = g1 > g0 ? hf.groups[g0].vbeg : (int32_t)hf.ox.size();
                    hf.leaf_vbeg.push_back(vb);
                    hf.leaf_vcnt.push_back((int32_t)hf.ox.size() - vb);
                }
            } else {
                int32_t hdr[3];
                if (!fb.get(hdr, 12) || !fb.get(&n.thr, 4)) { err = "truncated " + path; return false; }
                n.mode = hdr[0]; n.f1 = hdr[1]; n.f2 = hdr[2];
                if ((n.mode == 0 || n.mode == 1) && (n.f1 < 0 || n.f1 >= hf.F)) { err = "feature1 out of range in " + path; return false; }
                if (n.mode == 0 && (n.f2 < 0 || n.f2 >= hf.F)) { err = "feature2 out of range in " + path; return false; }
                pending.push_back({idx, 1});  // right is parsed after the whole left subtree
                pending.push_back({idx, 0});
            }
            tn.push_back(n);
            depth_of.push_back(d);
            if (d > hf.max_depth) hf.max_depth = d;
            if (slot.first < 0) root_idx = idx; else tn[slot.first].child[slot.second] = idx;
        }
        hf.leaf_base[t + 1] = (int32_t)hf.leaf_id.size();
        // breadth-first renumbering of the internal nodes
        const int32_t base = (int32_t)hf.nodes.size();
        std::vector<int32_t> order, newidx(tn.size(), -1);
        if (tn[root_idx].leaf < 0) order.push_back(root_idx);
        for (size_t h = 0; h < order.size(); ++h) {
            const TmpNode& n = tn[order[h]];
            newidx[order[h]] = base + (int32_t)h;
            for (int s = 0; s < 2; ++s)
                if (tn[n.child[s]].leaf < 0) order.push_back(n.child[s]);
        }
        auto entry = [&](int32_t i) { return tn[i].leaf >= 0 ? ~tn[i].leaf : newidx[i]; };
        hf.root[t] = entry(root_idx);
        for (size_t h = 0; h < order.size(); ++h) {
            const TmpNode& n = tn[order[h]];
            PackedNode pn;
            uint32_t f1 = (uint32_t)hf.F, f2 = (uint32_t)hf.F;  // unknown mode: val = 0 (the reference leaves it uninitialised)
            if (n.mode == 0) { f1 = (uint32_t)n.f1; f2 = (uint32_t)n.f2; }
            else if (n.mode == 1) { f1 = (uint32_t)n.f1; }
            pn.f1f2 = f1 | (f2 << 16);
            pn.thr = n.thr;
            pn.left = entry(n.child[0]);
            pn.right = entry(n.child[1]);
            hf.nodes.push_back(pn);
        }
        hf.n_internal += (int64_t)order.size();
        // two-level records: breadth-first over the internal nodes at even depth
        {
            const int32_t rbase = (int32_t)hf.recs.size();
            std::vector<int32_t> rorder, ridx(tn.size(), -1);
            if (tn[root_idx].leaf < 0) { rorder.push_back(root_idx); ridx[root_idx] = rbase; }
            for (size_t h = 0; h < rorder.size(); ++h) {
                const TmpNode& n = tn[rorder[h]];
                for (int s = 0; s < 2; ++s) {
                    const TmpNode& ch = tn[n.child[s]];
                    if (ch.leaf >= 0) continue;
                    for (int b = 0; b < 2; ++b) {
                        const int32_t gc = ch.child[b];
                        if (tn[gc].leaf < 0) { ridx[gc] = rbase + (int32_t)rorder.size(); rorder.push_back(gc); }
                    }
                }
            }
            auto test_of = [&](const TmpNode& n, uint32_t& f1f2, float& thr) {
                uint32_t f1 = (uint32_t)hf.F, f2 = (uint32_t)hf.F;
                if (n.mode == 0) { f1 = (uint32_t)n.f1; f2 = (uint32_t)n.f2; }
                else if (n.mode == 1) { f1 = (uint32_t)n.f1; }
                f1f2 = f1 | (f2 << 16);
                thr = n.thr;
            };
            hf.rec_root.push_back(tn[root_idx].leaf >= 0 ? ~tn[root_idx].leaf : ridx[root_idx]);
            for (size_t h = 0; h < rorder.size(); ++h) {
                const TmpNode& n = tn[rorder[h]];
                PackedRecord r;
                memset(&r, 0, sizeof r);
                test_of(n, r.f1f2[0], r.thr[0]);
                for (int s = 0; s < 2; ++s) {
                    const TmpNode& ch = tn[n.child[s]];
                    if (ch.leaf >= 0) {
                        r.f1f2[1 + s] = (uint32_t)hf.F | ((uint32_t)hf.F << 16);  // 0 - 0 < +inf: always "left"
                        r.thr[1 + s] = INFINITY;
                        r.next[2 * s] = r.next[2 * s + 1] = ~ch.leaf;
                    } else {
                        test_of(ch, r.f1f2[1 + s], r.thr[1 + s]);
                        for (int b = 0; b < 2; ++b) {
                            const int32_t gc = ch.child[b];
                            r.next[2 * s + b] = tn[gc].leaf >= 0 ? ~tn[gc].leaf : ridx[gc];
                        }
                    }
                }
                hf.recs.push_back(r);
            }
        }
    }
    hf.g_ozmin.resize(hf.groups.size());
    hf.g_ozmax.resize(hf.groups.size());
    for (size_t gi = 0; gi < hf.groups.size(); ++gi) {
        const VoteGroup& vg = hf.groups[gi];
        float lo = INFINITY, hi = -INFINITY;
        bool ordered = true;  // a NaN offset compares false everywhere: such a group never takes the shortcut
        for (int q = 0; q < vg.vcnt; ++q) {
            const float z = hf.oz[(size_t)vg.vbeg + q];
            if (!(z == z)) ordered = false;
            lo = z < lo ? z : lo;
            hi = z > hi ? z : hi;
        }
        hf.g_ozmin[gi] = ordered ? lo : INFINITY;
        hf.g_ozmax[gi] = ordered ? hi : -INFINITY;
    }
    hf.oz_sorted = hf.oz;
    for (size_t gi = 0; gi < hf.groups.size(); ++gi) {
        const VoteGroup& vg = hf.groups[gi];
        if (hf.g_ozmin[gi] <= hf.g_ozmax[gi])  // ordered groups only (the others are walked vote by vote)
            std::sort(hf.oz_sorted.begin() + vg.vbeg, hf.oz_sorted.begin() + vg.vbeg + vg.vcnt);
    }
    if (hf.nodes.empty()) hf.nodes.push_back(PackedNode{0, 0.f, -1, -1});  // keep device arrays non-empty
    if (hf.recs.empty()) { PackedRecord r; memset(&r, 0, sizeof r); hf.recs.push_back(r); }
    return true;
}

// ------------------------------------------------------------------------------------------------ encoder weights
struct HostLayer {
    int out = 0, in = 0;
    std::vector<float> W, b;  // W [out][in] row-major (Caffe InnerProduct), b [out]
};

namespace pb {
inline bool varint(const uint8_t*& p, const uint8_t* e, uint64_t& v) {
    v = 0;
    for (int s = 0; p < e && s < 64; s += 7) {
        uint8_t b = *p++;
        v |= (uint64_t)(b & 0x7F) << s;
        if (!(b & 0x80)) return true;
    }
    return false;
}
// Iterates the fields of one message; cb(field, wire_type, payload begin, payload end, varint value) -> keep going
template <class CB>
inline bool fields(const uint8_t* p, const uint8_t* e, CB cb) {
    while (p < e) {
        uint64_t key, v = 0;
        if (!varint(p, e, key)) return false;
        const int wt = (int)(key & 7), fn = (int)(key >> 3);
        const uint8_t* b = p;
        const uint8_t* q = p;
        if (wt == 0) { if (!varint(p, e, v)) return false; q = p; }
        else if (wt == 1) { if (e - p < 8) return false; p += 8; q = p; }
        else if (wt == 5) { if (e - p < 4) return false; p += 4; q = p; }
        else if (wt == 2) { uint64_t n; if (!varint(p, e, n) || (uint64_t)(e - p) < n) return false; b = p; p += n; q = p; }
        else return false;
        cb(fn, wt, b, q, v);
    }
    return true;
}
}  // namespace pb

// BVLC caffe.proto: NetParameter{ layers = 2 (V1LayerParameter), layer = 100 (LayerParameter) };
// V1LayerParameter{ name = 4, blobs = 6 }; LayerParameter{ name = 1, blobs = 7 };
// BlobProto{ num = 1, channels = 2, height = 3, width = 4, data = 5 (packed or repeated float), shape = 7 { dim = 1 } }.
inline bool parse_blob(const uint8_t* b, const uint8_t* e, std::vector<float>& data, std::vector<int64_t>& shape) {
    int64_t legacy[4] = {0, 0, 0, 0};
    bool has_legacy = false;
    std::vector<int64_t> dims;
    bool ok = pb::fields(b, e, [&](int fn, int wt, const uint8_t* p, const uint8_t* q, uint64_t v) {
        if (fn >= 1 && fn <= 4 && wt == 0) { legacy[fn - 1] = (int64_t)v; has_legacy = true; }
        else if (fn == 5 && wt == 2) { size_t n = (q - p) / 4; size_t o = data.size(); data.resize(o + n); memcpy(&data[o], p, n * 4); }
        else if (fn == 5 && wt == 5) { float f; memcpy(&f, q - 4, 4); data.push_back(f); }
        else if (fn == 7 && wt == 2) {
            pb::fields(p, q, [&](int f2, int w2, const uint8_t* p2, const uint8_t* q2, uint64_t v2) {
                if (f2 == 1 && w2 == 0) dims.push_back((int64_t)v2);
                else if (f2 == 1 && w2 == 2) { const uint8_t* r = p2; uint64_t x; while (r < q2 && pb::varint(r, q2, x)) dims.push_back((int64_t)x); }
            });
        }
    });
    if (!ok) return false;
    if (!dims.empty()) shape = dims;
    else if (has_legacy) shape.assign(legacy, legacy + 4);
    return true;
}

inline bool load_weights(const std::string& path, std::vector<HostLayer>& layers, std::string& err) {
    FileBuf fb;
    if (!fb.load(path)) { err = "cannot read " + path; return false; }
    layers.clear();
    if (fb.d.size() >= 12 && memcmp(fb.d.data(), "HF6DW001", 8) == 0) {
        fb.pos = 8;
        int32_t n;
        fb.get(&n, 4);
        if (n <= 0 || n > 16) { err = "bad layer count in " + path; return false; }
        for (int i = 0; i < n; ++i) {
            HostLayer L;
            int32_t oi[2];
            if (!fb.get(oi, 8) || oi[0] <= 0 || oi[1] <= 0) { err = "truncated " + path; return false; }
            L.out = oi[0]; L.in = oi[1];
            L.W.resize((size_t)L.out * L.in);
            L.b.resize(L.out);
            if (!fb.get(L.W.data(), L.W.size() * 4) || !fb.get(L.b.data(), L.b.size() * 4)) { err = "truncated " + path; return false; }
            layers.push_back(std::move(L));
        }
        return true;
    }
    // protobuf wire format
    struct Named { std::string name; std::vector<std::vector<float>> blobs; std::vector<std::vector<int64_t>> shapes; };
    std::vector<Named> found;
    const uint8_t* b = fb.d.data();
    bool ok = pb::fields(b, b + fb.d.size(), [&](int fn, int wt, const uint8_t* p, const uint8_t* q, uint64_t) {
        if (wt != 2 || (fn != 2 && fn != 100)) return;
        const int name_field = fn == 2 ? 4 : 1, blob_field = fn == 2 ? 6 : 7;
        Named nm;
        pb::fields(p, q, [&](int f2, int w2, const uint8_t* p2, const uint8_t* q2, uint64_t) {
            if (w2 != 2) return;
            if (f2 == name_field) nm.name.assign((const char*)p2, (size_t)(q2 - p2));
            else if (f2 == blob_field) {
                std::vector<float> d; std::vector<int64_t> s;
                if (parse_blob(p2, q2, d, s)) { nm.blobs.push_back(std::move(d)); nm.shapes.push_back(std::move(s)); }
            }
        });
        if (nm.blobs.size() >= 2) found.push_back(std::move(nm));
    });
    if (!ok) { err = "not a caffemodel or HF6DW001 file: " + path; return false; }
    const char* want[3] = {"encode1", "encode2", "encode3"};
    for (int i = 0; i < 3; ++i) {
        const Named* nm = nullptr;
        for (auto& f : found) if (f.name == want[i]) nm = &f;
        if (!nm) { err = std::string("layer ") + want[i] + " with weights not found in " + path; return false; }
        HostLayer L;
        L.out = (int)nm->blobs[1].size();
        if (L.out <= 0 || nm->blobs[0].size() % (size_t)L.out) { err = std::string("bad blob sizes in layer ") + want[i]; return false; }
        L.in = (int)(nm->blobs[0].size() / (size_t)L.out);
        L.W = nm->blobs[0];
        L.b = nm->blobs[1];
        layers.push_back(std::move(L));
    }
    return true;
}

// ------------------------------------------------------------------------------------------------ options file
struct HostOptions {
    std::vector<hf6d_object> objects;
    std::vector<std::string> mesh_files;
    std::vector<float> nn_search_radius;   // per object, -1 = not given (MeshUtils.h:239-245)
    std::vector<int> icp_iterations;       // per object, -1 = not given
    hf6d_refine_params refine;             // the MeshUtils settings of HFTest.cpp:1203-1225 (proto defaults until a key sets them)
    bool refine_defaults_set = false;
    std::string forest_folder, caffe_definition, caffe_weights;
    int stride = 4, gpu = -1, num_threads = 4, batch_size = 100;  // proto defaults (detector_options.proto:23-27)
    float max_depth_range = 0.25f, fx = 575.f, fy = 575.f, cx = 319.5f, cy = 239.5f, distance_threshold = 1.5f;
    bool are_objects_segmented = false;
    float location_score_coeff = 1.0f, pose_score_coeff = 0.7f;  // detector_options.proto:46-47 (used by the CLI's ranking)
};

// Minimal protobuf text-format reader for DetectorOptions.Options: `key: value`, `key { ... }`, `key: { ... }`,
// '#' comments, quoted strings with \" \\ \n escapes, optional ',' or ';' separators.  Unknown keys are an error
// (the reference ignores parse failures silently, HFTest.cpp:1164).
class OptionsParser {
  public:
    explicit OptionsParser(const std::string& text) : s_(text) {}
    bool parse(HostOptions& o, std::string& err) {
        if (!o.refine_defaults_set) { refine_defaults(o.refine); o.refine_defaults_set = true; }
        while (true) {
            skip();
            if (p_ >= s_.size()) break;
            std::string key;
            if (!ident(key)) return fail(err, "expected a field name");
            skip();
            if (key == "object_options") {
                if (peek() == ':') { ++p_; skip(); }
                char close = 0;
                if (peek() == '{') close = '}'; else if (peek() == '<') close = '>';
                if (!close) return fail(err, "expected '{' after object_options");
                ++p_;
                hf6d_object ob;
                memset(&ob, 0, sizeof ob);
                ob.should_detect = 1; ob.max_location_hypotheses = 12; ob.instances = 1;
                std::string mesh;
                bool has_name = false;
                float nn_radius = 0.01f;  // proto defaults (detector_options.proto:10-11): an object always carries both
                int icp_iters = 60;
                while (true) {
                    skip();
                    if (p_ >= s_.size()) return fail(err, "unterminated object_options block");
                    if (peek() == close) { ++p_; break; }
                    std::string k2, v;
                    if (!ident(k2)) return fail(err, "expected a field name in object_options");
                    skip();
                    if (peek() != ':') return fail(err, "expected ':' after " + k2);
                    ++p_;
                    if (!value(v)) return fail(err, "bad value for " + k2);
                    if (k2 == "name") { snprintf(ob.name, sizeof ob.name, "%s", v.c_str()); has_name = true; }
                    else if (k2 == "mesh_file") mesh = v;
                    else if (k2 == "instances") ob.instances = atoi(v.c_str());
                    else if (k2 == "max_location_hypotheses") ob.max_location_hypotheses = atoi(v.c_str());
                    else if (k2 == "should_detect") { if (!boolean(v, ob.should_detect)) return fail(err, "bad bool for should_detect"); }
                    else if (k2 == "nn_search_radius") nn_radius = strtof(v.c_str(), nullptr);
                    else if (k2 == "icp_iterations") icp_iters = atoi(v.c_str());
                    else if (k2 == "align_z_axis") {}
                    else return fail(err, "unknown field object_options." + k2);
                    sep();
                }
                if (!has_name) return fail(err, "object_options without a name");
                o.objects.push_back(ob);
                o.mesh_files.push_back(mesh);
                o.nn_search_radius.push_back(nn_radius);
                o.icp_iterations.push_back(icp_iters);
            } else {
                if (peek() != ':') return fail(err, "expected ':' after " + key);
                ++p_;
                std::string v;
                if (!value(v)) return fail(err, "bad value for " + key);
                int b = 0;
                if (key == "forest_folder") o.forest_folder = v;
                else if (key == "caffe_definition") o.caffe_definition = v;
                else if (key == "caffe_weights") o.caffe_weights = v;
                else if (key == "stride") o.stride = atoi(v.c_str());
                else if (key == "gpu") o.gpu = atoi(v.c_str());
                else if (key == "num_threads") o.num_threads = atoi(v.c_str());
                else if (key == "batch_size") o.batch_size = atoi(v.c_str());
                else if (key == "max_depth_range_in_patch_in_m") o.max_depth_range = strtof(v.c_str(), nullptr);
                else if (key == "fx") o.fx = strtof(v.c_str(), nullptr);
                else if (key == "fy") o.fy = strtof(v.c_str(), nullptr);
                else if (key == "cx") o.cx = strtof(v.c_str(), nullptr);
                else if (key == "cy") o.cy = strtof(v.c_str(), nullptr);
                else if (key == "distance_threshold") o.distance_threshold = strtof(v.c_str(), nullptr);
                else if (key == "are_objects_segmented") { if (!boolean(v, b)) return fail(err, "bad bool for " + key); o.are_objects_segmented = b != 0; }
                else if (key == "location_score_coeff") o.refine.location_score_coeff = o.location_score_coeff = strtof(v.c_str(), nullptr);
                else if (key == "pose_score_coeff") o.refine.pose_score_coeff = o.pose_score_coeff = strtof(v.c_str(), nullptr);
                else if (key == "search_single_object_instance" || key == "search_single_object_in_group" || key == "use_color_similarity") {
                    if (!boolean(v, b)) return fail(err, "bad bool for " + key);
                    (key == "search_single_object_instance" ? o.refine.search_single_object_instance
                     : key == "search_single_object_in_group" ? o.refine.search_single_object_in_group
                                                              : o.refine.use_color_similarity) = b;
                } else if (float* f = refine_float(o.refine, key)) *f = strtof(v.c_str(), nullptr);
                else if (key == "cluster_min_points") o.refine.cluster_min_points = atoi(v.c_str());
                else return fail(err, "unknown field " + key);
            }
            sep();
        }
        if (o.forest_folder.empty()) { err = "No forest folder specified"; return false; }     // HFTest.cpp:1166
        if (o.caffe_weights.empty()) { err = "No caffe weights model defined."; return false; }  // HFTest.cpp:1170
        if (o.stride <= 0) { err = "Stride should be more than 0"; return false; }               // HFTest.cpp:1173
        if (!(o.max_depth_range > 0)) { err = "max_depth_range_in_patch_in_m must be greater than 0"; return false; }
        return true;
    }

  private:
    // MeshUtils' members as DetectObjects sets them from the options (HFTest.cpp:1203-1225; detector_options.proto:33-66)
    static void refine_defaults(hf6d_refine_params& r) {
        memset(&r, 0, sizeof r);
        r.scene_leaf_m = 0.005f; r.object_leaf_m = 0.005f; r.normals_radius_m = 0.03f; r.nn_search_radius_m = 0.01f;
        r.occlusion_threshold_m = 0.02f;
        r.similarity_coeff = 10.f; r.inliers_coeff = 2.5f; r.clutter_coeff = 1.4f; r.location_score_coeff = 1.f; r.pose_score_coeff = 0.7f;
        r.group_total_explain_coeff = 0.5f; r.group_common_explain_coeff = 0.3f;
        r.inliers_threshold = 0.6f; r.clutter_threshold = 0.6f; r.final_score_threshold = 10.f;
        r.cluster_eps_angle_threshold = 0.05f; r.cluster_curvature_threshold = 0.1f; r.cluster_tolerance_near = 0.03f;
        r.cluster_tolerance_far = 0.05f; r.cluster_min_points = 5;
        r.use_color_similarity = 1; r.use_normal_similarity = 1; r.default_icp_iterations = 60;
    }
    static float* refine_float(hf6d_refine_params& r, const std::string& k) {
        if (k == "similarity_coeff") return &r.similarity_coeff;
        if (k == "inliers_coeff") return &r.inliers_coeff;
        if (k == "clutter_coeff") return &r.clutter_coeff;
        if (k == "group_total_explain_coeff") return &r.group_total_explain_coeff;
        if (k == "group_common_explain_coeff") return &r.group_common_explain_coeff;
        if (k == "inliers_threshold") return &r.inliers_threshold;
        if (k == "clutter_threshold") return &r.clutter_threshold;
        if (k == "final_score_threshold") return &r.final_score_threshold;
        if (k == "cluster_eps_angle_threshold") return &r.cluster_eps_angle_threshold;
        if (k == "cluster_curvature_threshold") return &r.cluster_curvature_threshold;
        if (k == "cluster_tolerance_near") return &r.cluster_tolerance_near;
        if (k == "cluster_tolerance_far") return &r.cluster_tolerance_far;
        return nullptr;
    }
    char peek() const { return p_ < s_.size() ? s_[p_] : '\0'; }
    void skip() {
        while (p_ < s_.size()) {
            char c = s_[p_];
            if (c == '#') { while (p_ < s_.size() && s_[p_] != '\n') ++p_; }
            else if (c == ' ' || c == '\t' || c == '\n' || c == '\r') ++p_;
            else break;
        }
    }
    void sep() { skip(); if (peek() == ',' || peek() == ';') ++p_; }
    bool ident(std::string& out) {
        size_t b = p_;
        while (p_ < s_.size() && (isalnum((unsigned char)s_[p_]) || s_[p_] == '_')) ++p_;
        out = s_.substr(b, p_ - b);
        return p_ > b;
    }
    bool value(std::string& out) {
        skip();
        out.clear();
        if (peek() == '"' || peek() == '\'') {
            const char q = s_[p_++];
            while (p_ < s_.size() && s_[p_] != q) {
                if (s_[p_] == '\\' && p_ + 1 < s_.size()) {
                    char c = s_[p_ + 1];
                    out.push_back(c == 'n' ? '\n' : c == 't' ? '\t' : c);
                    p_ += 2;
                } else out.push_back(s_[p_++]);
            }
            if (p_ >= s_.size()) return false;
            ++p_;
            return true;
        }
        size_t b = p_;
        while (p_ < s_.size() && !isspace((unsigned char)s_[p_]) && s_[p_] != ',' && s_[p_] != ';' && s_[p_] != '}' && s_[p_] != '>' && s_[p_] != '#') ++p_;
        out = s_.substr(b, p_ - b);
        return p_ > b;
    }
    static bool boolean(const std::string& v, int& out) {
        if (v == "true" || v == "True" || v == "t" || v == "1") { out = 1; return true; }
        if (v == "false" || v == "False" || v == "f" || v == "0") { out = 0; return true; }
        return false;
    }
    bool fail(std::string& err, const std::string& what) {
        int line = 1;
        for (size_t i = 0; i < p_ && i < s_.size(); ++i) line += s_[i] == '\n';
        err = "options file, line " + std::to_string(line) + ": " + what;
        return false;
    }
    std::string s_;
    size_t p_ = 0;
};

inline bool load_options(const std::string& path, HostOptions& o, std::string& err) {
    FileBuf fb;
    if (!fb.load(path)) { err = "Detector options file not found! (" + path + ")"; return false; }
    OptionsParser ps(std::string(fb.d.begin(), fb.d.end()));
    return ps.parse(o, err);
}

}  // namespace hf6d
