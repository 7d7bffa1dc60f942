// oracle/_ref/libhf6d_refsrc.so -- the reference's OWN sources for the detection path, compiled from where they lie under
// /root/reference against the stand-in headers of oracle/ref_shim/ (Eigen, OpenCV, Caffe, boost, glog/gflags, protobuf,
// CUDA runtime, MeshUtils).  TEST INFRASTRUCTURE ONLY: it pins the C oracle (hf6d_oracle.c) to the reference's code;
// nothing in the product loads it.  Built by oracle/build_ref.py; this file is the C ABI around it (C++98 like the rest).
//
//   HoughForest/src/HFBase.cpp          unmodified   forest loader (A5)
//   HoughForest/src/HFTest.cpp          unmodified   get_leaf, detect, non_max_suppression, test_image (A1, A3, A6-A11)
//   PatchGen/src/cuda/patch_extractor.cu   `<<<>>>` -> HF6D_SHIM_LAUNCH   centre scan + gather kernels on the host (A2)
//   PatchGen/src/cuda/surface_normals.cu   `<<<>>>` -> HF6D_SHIM_LAUNCH   normals kernel on the host (A2c)
//   HoughForest/src/MeshUtils.cpp:29-59, 423-440     get_rotmat_from_yaw_pitch_roll + the pre-ICP head of icp (A12)
//
// What the stand-ins decide instead of the absent libraries is listed in their headers (matrix product order, box-filter
// accumulation, texture filter, fill RNG, encoder arithmetic).
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include <omp.h>

// every header HFTest.h pulls in is included first (its include guards then make it a no-op below), so the access hack
// touches the two reference class definitions only: members of `class HFTest` before its first access specifier are
// private by default, hence `class` -> `struct` as well.  Layout and name mangling do not depend on either.
#include <Eigen/Dense>
#include <boost/unordered_map.hpp>
#include <caffe/caffe.hpp>
#include <cv.h>
#include <MeshUtils.h>
#include <detector_options.pb.h>
#include <google/protobuf/text_format.h>
#define private public
#define protected public
#define class struct
#include <HFTest.h>
#undef class
#undef private
#undef protected
#include <cuda/patch_extractor.h>
#include <cuda/surface_normals.h>

// ------------------------------------------------------------------------------------------------ gflags / CUDA globals
bool FLAGS_visualize_hypotheses = false;
std::string FLAGS_detector_options_file;
std::string FLAGS_output_folder;
Hf6dShimDim threadIdx, blockIdx, blockDim, gridDim;
const char* hf6d_shim_kernel_name = "";
unsigned long long hf6d_shim_fill_seed = 0;
int hf6d_shim_clock_calls = 0;

// ------------------------------------------------------------------------------------------------ capture
static bool g_capture = false;
static int g_map_rows = 0, g_map_cols = 0;
static std::vector<float> g_net_input;   // every batch handed to the net, concatenated
static std::vector<float> g_maps;        // every rows x cols map handed to cv::blur (the pre-blur centre maps, class order)
static std::vector<float> g_blurred;     // ... and what came out

// ------------------------------------------------------------------------------------------------ OpenCV stand-ins
namespace cv {
static inline int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}
void blur(const Mat& src_in, Mat& dst, Size k) {
    Mat src;
    src_in.copyTo(src);
    const int rows = src.rows, cols = src.cols;
    const bool cap = g_capture && rows == g_map_rows && cols == g_map_cols;
    if (cap) g_maps.insert(g_maps.end(), (const float*)src.data, (const float*)src.data + (size_t)rows * cols);
    std::vector<double> tmp((size_t)rows * cols);
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) {
            double s = 0;
            for (int i = 0; i < k.width; ++i) s += (double)src.at<float>(r, reflect101(c - k.width / 2 + i, cols));
            tmp[(size_t)r * cols + c] = s;
        }
    Mat out(rows, cols, CV_32FC1);
    const double scale = 1.0 / ((double)k.width * (double)k.height);
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) {
            double s = 0;
            for (int i = 0; i < k.height; ++i) s += tmp[(size_t)reflect101(r - k.height / 2 + i, rows) * cols + c];
            out.at<float>(r, c) = (float)(s * scale);
        }
    if (cap) g_blurred.insert(g_blurred.end(), (const float*)out.data, (const float*)out.data + (size_t)rows * cols);
    dst = out;
}
Mat imread(const std::string&, int) { return Mat(); }
bool imwrite(const std::string&, const Mat&) { return false; }
}  // namespace cv

// ------------------------------------------------------------------------------------------------ Caffe stand-in
typedef void (*hf6d_encode_fn)(const float* x, int P, int n0, const float* W1, const float* b1, int n1, const float* W2,
                               const float* b2, int n2, const float* W3, const float* b3, int n3, float* out);
static hf6d_encode_fn g_encoder = 0;
static int g_batch = 100;

namespace caffe {
template <typename T>
Net<T>::Net(const std::string&, Phase) {
    in_v_.push_back(&in_);
    out_v_.push_back(&out_);
}
template <typename T>
void Net<T>::CopyTrainedLayersFrom(const std::string& path) {  // raw "HF6DW001" container: n, then per layer out, in, W, b
    std::ifstream f(path.c_str(), std::ios::in | std::ios::binary);
    char magic[8];
    int n = 0;
    f.read(magic, 8);
    f.read((char*)&n, 4);
    CHECK(f && std::memcmp(magic, "HF6DW001", 8) == 0 && n == 3) << "stand-in Net: cannot read " << path;
    W_.resize(n);
    b_.resize(n);
    dims_.assign(n + 1, 0);
    for (int l = 0; l < n; ++l) {
        int oi[2];
        f.read((char*)oi, 8);
        if (l == 0) dims_[0] = oi[1];
        dims_[l + 1] = oi[0];
        W_[l].resize((size_t)oi[0] * oi[1]);
        b_[l].resize(oi[0]);
        f.read((char*)&W_[l][0], sizeof(T) * W_[l].size());
        f.read((char*)&b_[l][0], sizeof(T) * b_[l].size());
    }
    CHECK(f) << "stand-in Net: truncated " << path;
    in_.num_ = g_batch; in_.chan_ = dims_[0];
    in_.d.assign((size_t)g_batch * dims_[0], 0);
    out_.num_ = g_batch; out_.chan_ = dims_[3];
    out_.d.assign((size_t)g_batch * dims_[3], 0);
}
template <typename T>
const std::vector<Blob<T>*>& Net<T>::ForwardPrefilled() {
    CHECK(g_encoder != 0) << "stand-in Net: no encoder installed (hf6d_refsrc_set_encoder)";
    if (g_capture) g_net_input.insert(g_net_input.end(), in_.d.begin(), in_.d.end());
    g_encoder(&in_.d[0], g_batch, dims_[0], &W_[0][0], &b_[0][0], dims_[1], &W_[1][0], &b_[1][0], dims_[2], &W_[2][0],
              &b_[2][0], dims_[3], &out_.d[0]);
    return out_v_;
}
template class Net<float>;
}  // namespace caffe

// ------------------------------------------------------------------------------------------------ MeshUtils stand-ins
static std::vector<Hf6dShimHypothesis> g_hyps;
std::vector<Hf6dShimHypothesis>& hf6d_shim_hypotheses() { return g_hyps; }
static Hf6dShimHypothesis g_pending;
#pragma omp threadprivate(g_pending)

void MeshUtils::record_icp(int obj_id, int row, int col, float z, float yaw, float pitch, float roll,
                           const Eigen::Matrix4f& rotmat) {
    g_pending.obj_id = obj_id; g_pending.row = row; g_pending.col = col; g_pending.z = z;
    g_pending.yaw = yaw; g_pending.pitch = pitch; g_pending.roll = roll;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) g_pending.rotmat[i * 4 + j] = rotmat(i, j);
}
bool MeshUtils::evaluate_hypothesis(ObjectHypothesis& h, float location_score, float pose_score) {
    g_pending.location_score = location_score;
    g_pending.pose_score = pose_score;
    h.eval.location_score = location_score;
    h.eval.pose_score = pose_score;
#pragma omp critical(hf6d_shim_hyps)
    {
        h.eval.final_score = -(float)g_hyps.size();  // keeps the reference's final sort in call order
        g_hyps.push_back(g_pending);
    }
    return true;
}
std::vector<int> MeshUtils::optimize_hypotheses(std::vector<ObjectHypothesis>& h) {
    std::vector<int> all(h.size());
    for (size_t i = 0; i < h.size(); ++i) all[i] = (int)i;
    return all;
}

// ------------------------------------------------------------------------------------------------ C ABI
struct Hf6dRefSrc {
    HFTest* t;
    std::string forest_dir, weights;
};

struct Hf6dRefSrcParams {  // same layout as hf6d_ref_params / hf6d_params
    int W, H, stride;
    float fx, fy, cx, cy;
    int patch_vox;
    float voxel_m, max_depth_range_m, distance_threshold_m;
    int fill_random;
    unsigned long long fill_seed;
    int batch_size, max_yaw_pitch_hypotheses, max_roll_hypotheses;
    float min_location_score_ratio, min_yaw_pitch_drop_ratio;
    int centers_blur_size, centers_nms_wsize, pose_blur_size, pose_nms_wsize;
    int patch_mode;
    float normals_focal;
};

namespace {
struct CoutSilencer {
    std::streambuf* old;
    std::ostringstream sink;
    CoutSilencer() : old(std::cout.rdbuf(sink.rdbuf())) {}
    ~CoutSilencer() { std::cout.rdbuf(old); }
};
void texture_rgbd(const unsigned char* bgr, const unsigned short* depth, int W, int H, std::vector<float>& tex) {
    // HFTest.cpp:370-379 (inside test_image; the function-level gather entry point needs the same texture)
    tex.resize((size_t)W * H * 4);
    size_t pos = 0;
    for (int row = 0; row < H; ++row)
        for (int col = 0; col < W; ++col) {
            const size_t i = (size_t)row * W + col;
            tex[pos++] = (float)bgr[i * 3 + 0] / 255.0f;
            tex[pos++] = (float)bgr[i * 3 + 1] / 255.0f;
            tex[pos++] = (float)bgr[i * 3 + 2] / 255.0f;
            tex[pos++] = (float)depth[i];
        }
}
}  // namespace

extern "C" {

void hf6d_refsrc_set_encoder(hf6d_encode_fn fn) { g_encoder = fn; }

void* hf6d_refsrc_create(const char* forest_dir, const char* weights_path) {
    CoutSilencer quiet;
    Hf6dRefSrc* h = new Hf6dRefSrc;
    h->t = new HFTest();
    h->forest_dir = forest_dir;
    h->weights = weights_path ? weights_path : "";
    if (!h->t->setInputForest(forest_dir)) {  // HFBase::loadForestFromFolder
        delete h->t;
        delete h;
        return 0;
    }
    h->t->setCaffeModel("deploy.prototxt (unused by the stand-in)", h->weights);
    return h;
}
void hf6d_refsrc_destroy(void* hv) {
    Hf6dRefSrc* h = (Hf6dRefSrc*)hv;
    if (!h) return;
    delete h->t;
    delete h;
}
float hf6d_refsrc_forest_info(void* hv, int* info /*T, K, F, patch_size_in_voxels*/) {
    HFTest* t = ((Hf6dRefSrc*)hv)->t;
    info[0] = t->number_of_trees_; info[1] = t->number_of_classes_; info[2] = t->feature_vector_length_;
    info[3] = t->patch_size_in_voxels_;
    return t->voxel_size_in_m_;
}

// HFTest::get_leaf (HFTest.cpp:144-163) for P feature vectors x all trees -> leaf_id[P][T]
void hf6d_refsrc_get_leaves(void* hv, const float* features, int P, int* leaf_id) {
    HFTest* t = ((Hf6dRefSrc*)hv)->t;
    const int F = t->feature_vector_length_, T = (int)t->trees_.size();
    for (int i = 0; i < P; ++i) {
        std::vector<float> v(features + (size_t)i * F, features + (size_t)(i + 1) * F);
        for (int k = 0; k < T; ++k) leaf_id[(size_t)i * T + k] = t->get_leaf(t->trees_[k], v)->leaf_id;
    }
}

// Loader dump (HFBase::loadNodeFromFile, HFBase.cpp:58-108): leaves of tree `tree` in file (pre-)order.  Returns the leaf
// count; fills leaf_id[n], class_prob[n][K], vote_count[n][K] and, if votes != 0, the votes concatenated ([..][6]).
static void walk_leaves(HFBase::TreeNode* n, int K, std::vector<int>& ids, std::vector<float>& probs, std::vector<int>& counts,
                        std::vector<float>& votes) {
    if (n->leaf) {
        ids.push_back(n->leaf_id);
        for (int c = 0; c < K; ++c) {
            probs.push_back(n->class_prob[c]);
            counts.push_back((int)n->hough_votes[c].size());
            for (size_t i = 0; i < n->hough_votes[c].size(); ++i)
                for (int j = 0; j < 6; ++j) votes.push_back(n->hough_votes[c][i](j));
        }
        return;
    }
    walk_leaves(n->left, K, ids, probs, counts, votes);
    walk_leaves(n->right, K, ids, probs, counts, votes);
}
static void walk_internal(HFBase::TreeNode* n, std::vector<int>& tests, std::vector<float>& thr) {
    if (n->leaf) return;
    tests.push_back(n->test.measure_mode); tests.push_back(n->test.feature1); tests.push_back(n->test.feature2);
    thr.push_back(n->test.threshold);
    walk_internal(n->left, tests, thr);
    walk_internal(n->right, tests, thr);
}
long long hf6d_refsrc_tree_dump(void* hv, int tree, int* leaf_id, float* class_prob, int* vote_count, float* votes,
                                long long votes_cap, int* tests /*[n_internal][3]*/, float* thresholds, int* n_internal) {
    HFTest* t = ((Hf6dRefSrc*)hv)->t;
    std::vector<int> ids, counts, tests_v;
    std::vector<float> probs, vv, thr;
    walk_leaves(t->trees_[tree], t->number_of_classes_, ids, probs, counts, vv);
    walk_internal(t->trees_[tree], tests_v, thr);
    if (leaf_id) std::memcpy(leaf_id, ids.empty() ? 0 : &ids[0], ids.size() * 4);
    if (class_prob) std::memcpy(class_prob, probs.empty() ? 0 : &probs[0], probs.size() * 4);
    if (vote_count) std::memcpy(vote_count, counts.empty() ? 0 : &counts[0], counts.size() * 4);
    if (votes && (long long)vv.size() <= votes_cap && !vv.empty()) std::memcpy(votes, &vv[0], vv.size() * 4);
    if (tests && !tests_v.empty()) std::memcpy(tests, &tests_v[0], tests_v.size() * 4);
    if (thresholds && !thr.empty()) std::memcpy(thresholds, &thr[0], thr.size() * 4);
    if (n_internal) *n_internal = (int)thr.size();
    return (long long)ids.size();
}

// get_obj_center_vote_from_6dof + Point3DToImage (HFTest.cpp:41-102, 21-37) for n votes cast from patch (px, py, depth)
void hf6d_refsrc_vote_pixels(const float* dof6, int n, const int* px, const int* py, const unsigned short* depth_mm,
                             const float* intr /*fx, fy, cx, cy*/, float* c3 /*[n][3]*/, int* uv /*[n][2]*/) {
    HFTest t;
    t.setCameraIntrinsics(intr[0], intr[1], intr[2], intr[3]);
    for (int i = 0; i < n; ++i) {
        Eigen::VectorXf dof(6);
        for (int j = 0; j < 6; ++j) dof(j) = dof6[(size_t)i * 6 + j];
        const Eigen::Vector3f c = t.get_obj_center_vote_from_6dof(dof, px[i], py[i], depth_mm[i]);
        const Eigen::Vector2i p = t.Point3DToImage(c);
        c3[i * 3] = c(0); c3[i * 3 + 1] = c(1); c3[i * 3 + 2] = c(2);
        uv[i * 2] = p(0); uv[i * 2 + 1] = p(1);
    }
}

// HFTest::non_max_suppression(cv::Mat, ...) (HFTest.cpp:219-268); results in the reference's order (std::sort by score)
int hf6d_refsrc_nms(const float* img, int rows, int cols, int wx, int wy, float* score, int* xs, int* ys, int cap) {
    HFTest t;
    cv::Mat m(rows, cols, CV_32FC1, (void*)img);
    std::vector<HFTest::MapHypothesis> out;
    t.non_max_suppression(m, out, cv::Size2i(wx, wy));
    for (size_t i = 0; i < out.size() && (int)i < cap; ++i) {
        score[i] = out[i].first; xs[i] = out[i].second.x; ys[i] = out[i].second.y;
    }
    return (int)out.size();
}

// patch_extractor_gpu::extract_patches_rgbd (patch_extractor.cu:318-433 + kernel :230-309).  Returns the patch count.
int hf6d_refsrc_extract_rgbd(const unsigned char* bgr, const unsigned short* depth, const Hf6dRefSrcParams* p,
                             float* patches /*[cap][ps][ps][4]*/, int* locs /*[cap][2]*/, int cap) {
    std::vector<float> tex, host_patches;
    std::vector<int> loc;
    texture_rgbd(bgr, depth, p->W, p->H, tex);
    hf6d_shim_fill_seed = p->fill_seed;
    patch_extractor_gpu::extract_patches_rgbd(tex, p->W, p->H, p->patch_vox, p->voxel_m, p->max_depth_range_m, p->stride, 1.0f,
                                              p->fx, host_patches, loc, p->fill_random != 0, p->distance_threshold_m);
    const int n = (int)loc.size() / 2, per = p->patch_vox * p->patch_vox * 4;
    for (int i = 0; i < n && i < cap; ++i) {
        locs[2 * i] = loc[2 * i]; locs[2 * i + 1] = loc[2 * i + 1];
        std::memcpy(patches + (size_t)i * per, &host_patches[(size_t)i * per], sizeof(float) * per);
    }
    return n;
}

// surface_normals_gpu::generate_normals (surface_normals.cu:11-123) -> normals[H][W][3]
void hf6d_refsrc_normals(const unsigned short* depth, int W, int H, float focal, float* normals) {
    std::vector<unsigned short> d(depth, depth + (size_t)W * H);
    std::vector<float> out;
    surface_normals_gpu::generate_normals(d, W, H, focal, out);
    std::memcpy(normals, &out[0], sizeof(float) * (size_t)W * H * 3);
}

// patch_extractor_gpu::extract_patches (the normals variant, patch_extractor.cu:12-221) on the 7-channel texture the
// reference builds at HFTest.cpp:333-347.  Returns the patch count.
int hf6d_refsrc_extract_normals(const unsigned char* bgr, const unsigned short* depth, const float* normals,
                                const Hf6dRefSrcParams* p, float* patches /*[cap][ps][ps][6]*/, int* locs, int cap) {
    const int W = p->W, H = p->H;
    std::vector<float> tex((size_t)W * H * 7), host_patches;
    size_t pos = 0, np = 0;
    for (int row = 0; row < H; ++row)
        for (int col = 0; col < W; ++col) {
            const size_t i = (size_t)row * W + col;
            tex[pos++] = (float)bgr[i * 3 + 0] / 255.0f;
            tex[pos++] = (float)bgr[i * 3 + 1] / 255.0f;
            tex[pos++] = (float)bgr[i * 3 + 2] / 255.0f;
            tex[pos++] = (float)depth[i];
            tex[pos++] = normals[np++];
            tex[pos++] = normals[np++];
            tex[pos++] = normals[np++];
        }
    std::vector<int> loc;
    hf6d_shim_fill_seed = p->fill_seed;
    patch_extractor_gpu::extract_patches(tex, W, H, p->patch_vox, p->voxel_m, p->stride, p->normals_focal, host_patches, loc,
                                         p->fill_random != 0, p->distance_threshold_m);
    const int n = (int)loc.size() / 2, per = p->patch_vox * p->patch_vox * 6;
    for (int i = 0; i < n && i < cap; ++i) {
        locs[2 * i] = loc[2 * i]; locs[2 * i + 1] = loc[2 * i + 1];
        std::memcpy(patches + (size_t)i * per, &host_patches[(size_t)i * per], sizeof(float) * per);
    }
    return n;
}

// HFTest::test_image (HFTest.cpp:296-1024), the whole per-frame path.  Hypotheses in the order the reference hands them to
// MeshUtils (class, centre rank, yaw/pitch rank, roll rank; n_threads = 1 keeps that order deterministic).
int hf6d_refsrc_test_image(void* hv, const unsigned char* bgr, const unsigned short* depth, const Hf6dRefSrcParams* p,
                           const unsigned char* should_detect, const int* max_loc, int n_threads, int capture,
                           Hf6dShimHypothesis* out, int cap) {
    Hf6dRefSrc* h = (Hf6dRefSrc*)hv;
    HFTest* t = h->t;
    t->setStrideInPixels(p->stride);
    t->setCameraIntrinsics(p->fx, p->fy, p->cx, p->cy);
    t->setNumThreads(n_threads);
    t->setMaxDepthRange(p->max_depth_range_m);
    t->setBatchSizeCaffe(p->batch_size);
    t->setHypothesesCalculationOption(p->max_yaw_pitch_hypotheses, p->max_roll_hypotheses, p->min_location_score_ratio,
                                      p->min_yaw_pitch_drop_ratio, p->centers_blur_size, p->centers_nms_wsize,
                                      p->pose_blur_size, p->pose_nms_wsize);
    g_batch = p->batch_size;
    hf6d_shim_fill_seed = p->fill_seed;
    DetectorOptions::Options opt;
    for (int k = 0; k < t->number_of_classes_; ++k) {
        DetectorOptions::ObjectOptions o;
        std::ostringstream nm;
        nm << "object" << k;
        o.name_ = nm.str();
        o.should_detect_ = should_detect ? should_detect[k] != 0 : true;
        o.max_location_hypotheses_ = max_loc ? max_loc[k] : 12;
        opt.objects_.push_back(o);
    }
    MeshUtils mu;
    mu.setIntrinsics(p->fx, p->fy, p->cx, p->cy);
    mu.setNumThreads(n_threads);
    cv::Mat rgb(p->H, p->W, CV_8UC3, (void*)bgr), dep(p->H, p->W, CV_16UC1, (void*)depth);
    g_hyps.clear();
    g_capture = capture != 0;
    g_map_rows = p->H; g_map_cols = p->W;
    g_net_input.clear(); g_maps.clear(); g_blurred.clear();
    {
        CoutSilencer quiet;
        t->test_image(rgb, dep, opt, mu, p->fill_random != 0, p->distance_threshold_m);
    }
    g_capture = false;
    for (size_t i = 0; i < g_hyps.size() && (int)i < cap; ++i) out[i] = g_hyps[i];
    return (int)g_hyps.size();
}

static long long copy_out(const std::vector<float>& v, float* dst, long long cap) {
    if (dst && (long long)v.size() <= cap && !v.empty()) std::memcpy(dst, &v[0], v.size() * sizeof(float));
    return (long long)v.size();
}
long long hf6d_refsrc_captured_net_input(float* dst, long long cap) { return copy_out(g_net_input, dst, cap); }
long long hf6d_refsrc_captured_maps(float* dst, long long cap) { return copy_out(g_maps, dst, cap); }
long long hf6d_refsrc_captured_blurred(float* dst, long long cap) { return copy_out(g_blurred, dst, cap); }

}  // extern "C"
