/* hf6d CPU oracle -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference's `HoughForest --test` per-frame detection path, written from the reference
 * sources (it cannot be built here: SURVEY.md F5).  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
 * legs may load this; the product (libhf6d.so) never links or calls it.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors, forests, weights or images and the
 * arithmetic of three third-party pieces it leans on (Caffe InnerProduct/Sigmoid, OpenCV 2.4 cv::blur, the CUDA
 * texture unit) is not under /root/reference.  This file is therefore the definition of "correct" for the repo;
 * every place it had to choose a semantics is marked CHOICE.
 */
#ifndef HF6D_ORACLE_H_
#define HF6D_ORACLE_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HF6D_REF_WEIGHT_SHIFT 16 /* vote weights are Q16 fixed point: w = (uint32)(p*65536+0.5) */
#define HF6D_REF_Z_BINS 300
#define HF6D_REF_POSE_BINS 720

typedef struct hf6d_ref_forest hf6d_ref_forest;

typedef struct {
    int32_t W, H;
    int32_t stride;
    float fx, fy, cx, cy;
    int32_t patch_vox; /* patch_size_in_voxels (forest.txt) */
    float voxel_m;     /* voxel_size_in_m      (forest.txt) */
    float max_depth_range_m;
    float distance_threshold_m;
    int32_t fill_random; /* = !are_objects_segmented */
    uint64_t fill_seed;
    int32_t batch_size;
    /* HFTest.h:175-194 */
    int32_t max_yaw_pitch_hypotheses, max_roll_hypotheses;
    float min_location_score_ratio, min_yaw_pitch_drop_ratio;
    int32_t centers_blur_size, centers_nms_wsize, pose_blur_size, pose_nms_wsize;
    /* A2c: 0 = RGB-D patches (4 channels, the live path), 1 = RGB + surface normals (6 channels, the variant the
     * reference keeps commented out at HFTest.cpp:322-363 / :443-470); normals_focal is its hard-wired 575.0f */
    int32_t patch_mode;
    float normals_focal;
} hf6d_ref_params;

typedef struct {
    int32_t cls, cx, cy; /* class, centre pixel */
    float z;             /* mode_z [m] */
    int32_t yaw_deg, pitch_deg, roll_deg;
    float loc_score, yawpitch_score, roll_score;
    float pose[16]; /* pre-ICP 4x4, row-major */
} hf6d_ref_hypothesis;

void hf6d_ref_default_params(hf6d_ref_params* p);

hf6d_ref_forest* hf6d_ref_forest_load(const char* dir);
void hf6d_ref_forest_free(hf6d_ref_forest* f);
/* info[0..5] = T, K, F, patch_vox, total leaves, total internal nodes ; returns voxel size */
float hf6d_ref_forest_info(const hf6d_ref_forest* f, int32_t* info);
/* leaves of tree t in file (pre-order) order */
int32_t hf6d_ref_tree_leaf_count(const hf6d_ref_forest* f, int32_t t);

/* A2a: valid patch centres, row-major scan order.  locs = [x0,y0,x1,y1..]; returns P (may exceed cap: then truncated) */
int32_t hf6d_ref_scan_centres(const uint16_t* depth_mm, const hf6d_ref_params* p, int32_t* locs, int32_t cap);
/* A1+A2b: patches [P][ps][ps][4] f32 HWC */
void hf6d_ref_gather(const uint8_t* bgr, const uint16_t* depth_mm, const hf6d_ref_params* p, const int32_t* locs,
                     int32_t P, float* patches);
/* A2c: surface normals [H][W][3] from the depth map (surface_normals.cu:11-73) */
void hf6d_ref_normals(const uint16_t* depth_mm, int32_t W, int32_t H, float focal, float* normals);
/* A2c: patches [P][ps][ps][6] f32 HWC = (B, G, R, nx, ny, nz)  (patch_extractor.cu:12-111) */
void hf6d_ref_gather_normals(const uint8_t* bgr, const uint16_t* depth_mm, const float* normals, const hf6d_ref_params* p,
                             const int32_t* locs, int32_t P, float* patches);
/* A2c: q [P][6*ps*ps] u8, CHW -- plain quantisation, no local normalisation (HFTest.cpp:443-470) */
void hf6d_ref_quantise_normals(const float* patches, int32_t P, int32_t ps, uint8_t* q);
/* A3: q [P][4*ps*ps] u8, CHW */
void hf6d_ref_normalise(const float* patches, int32_t P, int32_t ps, uint8_t* q);
/* A4: features [P][n3].  W* are [out][in] row-major fp32 */
void hf6d_ref_encode(const uint8_t* q, int32_t P, int32_t n0, const float* W1, const float* b1, int32_t n1,
                     const float* W2, const float* b2, int32_t n2, const float* W3, const float* b3, int32_t n3,
                     float* features);
/* A4 on fp32 inputs x [P][n0] (the net input k/255.0f); same arithmetic as hf6d_ref_encode */
void hf6d_ref_encode_f32(const float* x, int32_t P, int32_t n0, const float* W1, const float* b1, int32_t n1,
                         const float* W2, const float* b2, int32_t n2, const float* W3, const float* b3, int32_t n3,
                         float* features);
/* A6: leaf_id [P][T] = leaf_id field of the file; leaf_ord [P][T] = file-order ordinal of the leaf inside its tree */
void hf6d_ref_traverse(const hf6d_ref_forest* f, const float* features, int32_t P, int32_t* leaf_id, int32_t* leaf_ord);
/* A7/A8: maps [K][H][W] u64 Q16 (zeroed here).  Returns number of votes cast (in or out of bounds). */
int64_t hf6d_ref_vote(const hf6d_ref_forest* f, const int32_t* leaf_ord, const int32_t* locs, const uint16_t* depth_mm,
                      int32_t P, const hf6d_ref_params* p, const uint8_t* should_detect, uint64_t* maps);
/* A9: box blur (BORDER_REFLECT_101) of a Q16 map -> float */
void hf6d_ref_blur(const uint64_t* acc, int32_t rows, int32_t cols, int32_t kx, int32_t ky, float* out);
/* A9: sliding-window NMS; outputs sorted by score desc (stable); returns count (truncated to cap) */
int32_t hf6d_ref_nms(const float* in, int32_t rows, int32_t cols, int32_t wx, int32_t wy, float* score, int32_t* xs,
                     int32_t* ys, int32_t cap);
/* A9-A12 for all classes from leaf assignments; hyps sorted by (class, centre rank, yaw/pitch rank, roll rank) */
int32_t hf6d_ref_hypotheses(const hf6d_ref_forest* f, const int32_t* leaf_ord, const int32_t* locs,
                            const uint16_t* depth_mm, int32_t P, const hf6d_ref_params* p,
                            const uint8_t* should_detect, const int32_t* max_location_hypotheses,
                            const uint64_t* maps, hf6d_ref_hypothesis* hyps, int32_t cap);
/* Whole frame.  features_override (may be NULL) replaces the encoder output (rows = P'); dbg_* may be NULL. */
int32_t hf6d_ref_detect(const hf6d_ref_forest* f, const uint8_t* bgr, const uint16_t* depth_mm,
                        const hf6d_ref_params* p, const float* const* weights /*W1,b1,W2,b2,W3,b3*/,
                        const int32_t* dims /*n0,n1,n2,n3*/, const uint8_t* should_detect,
                        const int32_t* max_location_hypotheses, const float* features_override,
                        hf6d_ref_hypothesis* hyps, int32_t cap, int32_t* n_patches_out, double* stage_seconds /*[6]*/);

#ifdef __cplusplus
}
#endif
#endif
