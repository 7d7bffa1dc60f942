/* libhf6d.so -- C ABI of the B200-native Hough-forest detection path.
 *
 * Drop-in boundary for the per-frame hot path of the reference's `HoughForest --test`:
 *
 *   reference interface                                          replaced by
 *   ------------------------------------------------------------ --------------------------------------------------
 *   HFTest::DetectObjects()        HoughForest/src/HFTest.cpp:1152  hf6d_create_from_options + hf6d_detect loop
 *   HFTest::setInputForest()       HoughForest/include/HFTest.h:131 hf6d_create (forest_dir)
 *   HFTest::setCaffeModel()        HoughForest/include/HFTest.h:122 hf6d_create (weights_path)
 *   HFTest::test_image()           HoughForest/include/HFTest.h:114 hf6d_detect / hf6d_submit + hf6d_wait
 *   patch_extractor_gpu::extract_patches_rgbd()
 *                                  PatchGen/include/cuda/patch_extractor.h:29
 *                                                                   stages HF6D_STAGE_SCAN + HF6D_STAGE_GATHER
 *   caffe::Net::ForwardPrefilled() HoughForest/src/HFTest.cpp:593   stage HF6D_STAGE_ENCODE
 *   HFTest::detect()/get_leaf()    HoughForest/src/HFTest.cpp:144-217 stages HF6D_STAGE_TRAVERSE + HF6D_STAGE_VOTE
 *   cv::blur + non_max_suppression HoughForest/src/HFTest.cpp:702-707 stage HF6D_STAGE_CENTRES
 *   z / yaw-pitch / roll seeking   HoughForest/src/HFTest.cpp:742-925 stage HF6D_STAGE_POSE
 *   MeshUtils::icp() pre-ICP pose  HoughForest/src/MeshUtils.cpp:423-440 hf6d_hypothesis.pose
 *
 * Everything is plain C: opaque handle, pointers and sizes.  Every entry point returns 0 on success or a negative
 * HF6D_E* code; hf6d_last_error() gives the message.  Nothing here aborts, prints, or reads stdin (the reference's
 * CUDA_SAFE_CALL blocks on getchar(), PatchGen/include/cuda/cuda_utils.h:18-72).
 *
 * There is no CPU fallback: every entry point that computes fails with HF6D_ECUDA when no sm_100 device is usable.
 */
#ifndef HF6D_H_
#define HF6D_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HF6D_OK 0
#define HF6D_EINVAL (-1)  /* bad argument */
#define HF6D_EIO (-2)     /* file missing / malformed (forest, weights, options) */
#define HF6D_ECUDA (-3)   /* CUDA runtime error or no usable device */
#define HF6D_ENOMEM (-4)
#define HF6D_ESTATE (-5)  /* call order (e.g. wait on an unknown ticket) */

#define HF6D_WEIGHT_SHIFT 16 /* vote weights are Q16 fixed point: w = (uint32)(class_prob * 65536 + 0.5) */
#define HF6D_Z_BINS 300      /* HFTest.cpp:743-745: 3 m / 1 cm */
#define HF6D_POSE_BINS 720   /* HFTest.cpp:756: degrees in [-360, 359] */
#define HF6D_MAX_CLASSES 32
#define HF6D_MAX_CENTRES 16  /* max_location_hypotheses upper bound per object */
#define HF6D_MAX_HYPOTHESES_PER_CENTRE 32

typedef struct hf6d_ctx hf6d_ctx;

/* Frame geometry + the reference's options (detector_options.proto:18-70) + hard-coded constants (HFTest.h:175-194).
 * Layout is shared with the oracle's parameter block so that parity tests pass one struct to both. */
typedef struct {
    int32_t W, H;
    int32_t stride;                /* Options.stride */
    float fx, fy, cx, cy;          /* Options.fx .. cy */
    int32_t patch_vox;             /* forest.txt: patch_size_in_voxels (overwritten from the forest at create) */
    float voxel_m;                 /* forest.txt: voxel_size_in_m      (overwritten from the forest at create) */
    float max_depth_range_m;       /* Options.max_depth_range_in_patch_in_m */
    float distance_threshold_m;    /* Options.distance_threshold */
    int32_t fill_random;           /* = !Options.are_objects_segmented */
    uint64_t fill_seed;            /* counter-based RNG key for the border fill (reference: clock64()) */
    int32_t batch_size;            /* Options.batch_size: P' = floor(P / batch_size) * batch_size patches are processed */
    int32_t max_yaw_pitch_hypotheses, max_roll_hypotheses;
    float min_location_score_ratio, min_yaw_pitch_drop_ratio;
    int32_t centers_blur_size, centers_nms_wsize, pose_blur_size, pose_nms_wsize;
    /* Patch channels: 0 = B,G,R,D with local normalisation (the reference's live path, patch_extractor.cu:230-309 +
     * HFTest.cpp:500-570; encoder input 4*ps*ps); 1 = B,G,R + interpolated surface normals with plain quantisation (the
     * variant the reference keeps next to it: surface_normals.cu:11-73, patch_extractor.cu:12-111, HFTest.cpp:322-363 and
     * :443-470; encoder input 6*ps*ps).  normals_focal is that variant's hard-wired focal length (575.0f). */
    int32_t patch_mode;
    float normals_focal;
} hf6d_params;

typedef struct {
    int32_t cls, cx, cy;                  /* class, centre pixel (column, row) */
    float z;                              /* mode_z [m] */
    int32_t yaw_deg, pitch_deg, roll_deg; /* quantised pose mode */
    float loc_score, yawpitch_score, roll_score;
    float pose[16];                       /* pre-ICP 4x4, row-major, camera (xtion) frame */
} hf6d_hypothesis;

/* Per-object switches (detector_options.proto:3-16). */
typedef struct {
    char name[64];
    int32_t should_detect;
    int32_t max_location_hypotheses;
    int32_t instances;
} hf6d_object;

typedef enum {
    HF6D_STAGE_SCAN = 0,     /* valid patch centres, row-major order             patch_extractor.cu:372-391 */
    HF6D_STAGE_GATHER = 1,   /* bilinear RGB-D gather + local normalise + quantise  patch_extractor.cu:230-309, HFTest.cpp:500-570 */
    HF6D_STAGE_ENCODE = 2,   /* 256-1500-1000-800 sigmoid encoder (tcgen05)      HFTest.cpp:585-596 */
    HF6D_STAGE_TRAVERSE = 3, /* every patch through every (owned) tree            HFTest.cpp:144-163 */
    HF6D_STAGE_VOTE = 4,     /* centre votes into per-class accumulators         HFTest.cpp:166-217 */
    HF6D_STAGE_CENTRES = 5,  /* box blur + sliding-window NMS -> centre list     HFTest.cpp:702-707 */
    HF6D_STAGE_POSE = 6,     /* z / yaw-pitch / roll mode seeking -> hypotheses  HFTest.cpp:742-925 */
    HF6D_STAGE_COUNT = 7
} hf6d_stage;

typedef enum {
    HF6D_BUF_COUNTS = 0,   /* int32[2]  = P (valid centres), P' (processed) */
    HF6D_BUF_LOCS = 1,     /* int32[P][2] = (x, y) */
    HF6D_BUF_PATCH_U8 = 2, /* uint8[P'][C*ps*ps] quantised CHW patches, C = 4 or 6 (only when debug capture is on) */
    HF6D_BUF_FEATURES = 3, /* float[P'][F] (feature storage 1: the fp32 widening of the stored fp16 rows) */
    HF6D_BUF_LEAF_ORD = 4, /* int32[P'][T]: file-order ordinal of the leaf inside its tree; -1 for trees not owned */
    HF6D_BUF_MAPS = 5,     /* uint64[K][H][W] Q16 vote sums */
    HF6D_BUF_BLURRED = 6,  /* float[K][H][W] */
    HF6D_BUF_CENTRES = 7,  /* per class: int32 n, then HF6D_MAX_CENTRES x {float score, int32 x, int32 y} (see hf6d_centre) */
    HF6D_BUF_FRAME_BGR = 8,
    HF6D_BUF_FRAME_DEPTH = 9,
    HF6D_BUF_NORMALS = 10, /* float[H][W][4] = (nx, ny, nz, 0): surface normals (patch_mode 1 only) */
    HF6D_BUF_COUNT = 11
} hf6d_buffer;

typedef struct {
    float score;
    int32_t x, y;
} hf6d_centre;

typedef struct {
    int32_t n;
    hf6d_centre c[HF6D_MAX_CENTRES];
} hf6d_centre_list;

typedef struct {
    int32_t T, K, F, patch_vox;
    float voxel_m;
    int64_t n_leaves, n_internal, n_votes; /* n_votes: votes of (leaf, class) groups that pass the 0.5 gate */
    int32_t max_depth;
    int32_t dims[4]; /* encoder 256, 1500, 1000, 800 */
} hf6d_model_info;

/* ---------------------------------------------------------------------------------------------- lifecycle */
void hf6d_default_params(hf6d_params* p);

/* forest_dir: forest.txt + tree<N>.dat (HFBase.cpp:110-145).  weights_path: a V1 .caffemodel with layers encode1..3
 * (generate_scripts.sh:424-524) or the raw "HF6DW001" container.  n_slots frames may be in flight (>=1): n_slots == 1 is
 * the latency configuration (one frame at a time, the encoder kernels that are fastest alone), n_slots > 1 the throughput
 * configuration (4 is enough to saturate a B200; encoder kernels with a smaller shared-memory footprint, so that the
 * other frames' kernels run beside them).  Results are bit-identical in both. */
int hf6d_create(const hf6d_params* p, const char* forest_dir, const char* weights_path, int device, int n_slots,
                hf6d_ctx** out);
/* Text-format DetectorOptions.Options file, as HoughForest --test --detector_options_file takes (HFTest.cpp:1155-1235).
 * W, H: frame size the context is sized for. */
int hf6d_create_from_options(const char* options_path, int W, int H, int device, int n_slots, hf6d_ctx** out);
void hf6d_destroy(hf6d_ctx* c);
const char* hf6d_last_error(const hf6d_ctx* c); /* c may be NULL: error of the last failed create on this thread */

int hf6d_get_params(const hf6d_ctx* c, hf6d_params* out);
int hf6d_model(const hf6d_ctx* c, hf6d_model_info* out);
int hf6d_set_objects(hf6d_ctx* c, const hf6d_object* objs, int n);   /* n must equal K */
int hf6d_get_objects(const hf6d_ctx* c, hf6d_object* objs, int cap); /* returns K */
int hf6d_set_fill_seed(hf6d_ctx* c, uint64_t seed);
/* Tree sharding (one process per GPU): this context traverses and votes only trees t with t % world == rank.
 * The caller sums HF6D_BUF_MAPS across ranks (NCCL all-reduce, uint64 sum) and max-reduces HF6D_BUF_LEAF_ORD between
 * hf6d_run(.., VOTE) and hf6d_run(CENTRES, ..). */
int hf6d_set_tree_shard(hf6d_ctx* c, int rank, int world);
/* Patch sharding (one stream of frames over several GPUs, the split that scales: the encoder is 0.3 of a frame and shards with
 * the patches, not with the trees): this context gathers, encodes, traverses (every tree) and votes only its share of the
 * frame's patches -- rank r of `world` takes the 128-patch row blocks [B*r/world, B*(r+1)/world) of the frame's B blocks, in
 * the reference's patch order (its batches of 100 run in an OpenMP loop, HFTest.cpp:612).  The scan is replicated.  Between
 * hf6d_run(.., VOTE) and hf6d_run(CENTRES, ..) the caller sums HF6D_BUF_MAPS and max-reduces HF6D_BUF_LEAF_ORD across the
 * ranks, exactly as for hf6d_set_tree_shard (rows of other ranks' patches are -1), or uses the peer exchange below. */
int hf6d_set_patch_shard(hf6d_ctx* c, int rank, int world);
/* What hf6d_peer_attach shards: 0 = trees (default), 1 = patches.  Call before hf6d_peer_attach. */
int hf6d_set_peer_split(hf6d_ctx* c, int split);
/* Class sharding of the stages after the exchange: this context seeks centres and poses (HF6D_STAGE_CENTRES, _POSE)
 * only for classes k with k % world == rank; votes are still cast for every detected class.  The hypothesis lists of
 * the ranks, concatenated in class order, equal the unsharded list. */
int hf6d_set_class_shard(hf6d_ctx* c, int rank, int world);
/* Peer exchange: the tree-sharded mode without a collective library, for the GPUs of one NVLink / NVSwitch box (one
 * process per GPU).  Every rank publishes a small blob (CUDA IPC handles of its vote maps, leaf tables and flag block),
 * the caller hands every rank all blobs (any transport: torch.distributed all_gather, MPI, a file), and from then on
 * hf6d_run exchanges nothing through the host: the kernels after the exchange point read the peers' maps, vote streams (or
 * leaf tables, when a rank keeps no stream) in place over NVLink, and the ranks synchronise through flags in each other's memory (see "peer exchange" in
 * csrc/hf6d_api.cu).  hf6d_peer_attach also sets the tree shard and the class shard to rank/world.  Contract:
 *  - the ranks of a group export, attach (and later detach) TOGETHER, with a barrier of the caller's transport between
 *    hf6d_peer_attach on every rank and the first hf6d_run on any (a rank must not signal into memory a peer has not mapped
 *    yet) and another one before hf6d_peer_detach / hf6d_destroy (nobody unmaps while a peer may still read);
 *  - every rank then runs the same frames on the same slots, in the same split: whole frames (SCAN..POSE), or SCAN..VOTE
 *    followed by CENTRES..POSE.  Any other sub-range (e.g. POSE alone on one rank) breaks the flag sequence;
 *  - the slots' streams must not share a hardware queue (CUDA_DEVICE_MAX_CONNECTIONS >= number of streams the process
 *    uses): a slot that waits for a peer's flag blocks its queue, and the peer may be waiting for a frame queued behind it.
 *    hf6d_peer_attach checks the variable (the driver reads it once, at initialisation) and fails with HF6D_ESTATE when it
 *    allows fewer queues than n_slots + 1, instead of leaving a deadlock for later;
 *  - waits are bounded: hf6d_sync / hf6d_collect / hf6d_wait give up after HF6D_PEER_TIMEOUT_MS (default 20000) with
 *    HF6D_ECUDA and hf6d_peer_timed_out() == 1 when a peer never answers; the context must then be destroyed.  A frame
 *    whose enqueue fails half way does not count towards the slot's sequence number.
 * Replaces the reference's merge of per-thread vote maps and leaf lists, HoughForest/src/HFTest.cpp:645-654, across GPUs. */
#define HF6D_MAX_SLOTS 16
size_t hf6d_peer_blob_bytes(void);
int hf6d_peer_export(hf6d_ctx* c, void* blob, size_t cap_bytes);
int hf6d_peer_attach(hf6d_ctx* c, int rank, int world, const void* blobs /* world blobs, rank order */, size_t bytes_each);
int hf6d_peer_detach(hf6d_ctx* c);
int hf6d_peer_timed_out(hf6d_ctx* c); /* 1 if a flag wait gave up (host-side deadline or the fallback wait kernel); results are then invalid */
/* Encoder arithmetic (replaces Caffe's fp32 sgemm, HoughForest/src/HFTest.cpp:585-596):
 *   0 = bf16 operands, fp32 accumulation (default: the throughput mode; features within 3e-2 of fp32);
 *   1 = split bf16: every operand as hi + lo bf16 halves, a product as a_hi*w_hi + a_lo*w_hi + a_hi*w_lo on the same
 *       tensor-core kernel (three passes over K), fp32 sigmoid -- features within ~3e-5 of an fp32 evaluation, i.e. as
 *       close to the reference's as one fp32 summation order is to another; about 3x the encoder time.
 *   2 = fp16 operands, fp32 accumulation: the same kernel at the same tensor-core rate with 11-bit instead of 8-bit
 *       significands (every operand of this net is bounded: patch values 0..255, sigmoid outputs, weights of a few units;
 *       the first layer's 1/255 moves from the weights to the accumulator so that small weights stay normal numbers).
 *       Features within ~2e-3 of fp32.  Not what BASELINE's north star names (bf16), hence not the default.
 * Synchronises the device; the first switch to mode 1 allocates the hi/lo activation buffers of every slot. */
int hf6d_set_encoder_mode(hf6d_ctx* c, int mode);
int hf6d_get_encoder_mode(const hf6d_ctx* c);
/* Feature storage between the encoder and the forest (the reference keeps the Caffe output blob in fp32 and reads it in
 * HFTest::get_leaf, HoughForest/src/HFTest.cpp:144-163, 595-628):
 *   0 = fp32 rows, float[P'][F];
 *   1 = fp16 rows: the feature layer of encoder modes 0 / 2 rounds its fp32 sigmoid once to fp16 and the traversal reads
 *       those rows (half the HBM bytes of both kernels; sigmoid outputs lie in (0,1), the rounding is <= 2.5e-4, two orders
 *       below the bf16 operand error of mode 0).  Every leaf test compares the exact fp32 widening of the stored half, and
 *       that widening is what HF6D_BUF_FEATURES returns (hf6d_fetch / hf6d_device_ptr / hf6d_encode_patches), so everything
 *       downstream stays bit-exact on the handed-out values.  Features injected with hf6d_inject are fp32 rows and are
 *       traversed as such.  Encoder mode 1 (the near-fp32 mode) always stores fp32.
 * Default: 1 where the feature layer has a kernel for it (F a multiple of 160, or 256-wide feature tiles), else 0;
 * HF6D_FEATURES=fp32 in the environment makes 0 the default.  Synchronises the device; the first switch to 1 allocates the
 * fp16 rows of every slot. */
int hf6d_set_feature_storage(hf6d_ctx* c, int storage);
int hf6d_get_feature_storage(const hf6d_ctx* c);
int hf6d_set_debug_capture(hf6d_ctx* c, int on); /* keep HF6D_BUF_PATCH_U8 */

/* ---------------------------------------------------------------------------------------------- host-only helpers */
/* These three touch no GPU: they parse and validate the model artefacts exactly as hf6d_create* does (same loaders), so
 * the CLI can report a bad options file / forest / weights file before it selects a device, and the CPU test-suite can
 * cover the format readers.  On failure the message is available from hf6d_last_error(NULL). */
typedef struct {
    hf6d_params params;         /* Options fields mapped onto hf6d_params (W, H left at the defaults 640 x 480) */
    int32_t gpu;                /* Options.gpu (-1 = unset) */
    int32_t n_objects;          /* number of object_options blocks */
    char forest_folder[1024];   /* Options.forest_folder */
    char caffe_weights[1024];   /* Options.caffe_weights */
    char caffe_definition[1024];
    float location_score_coeff; /* Options.location_score_coeff / pose_score_coeff: the Hough terms of the reference's */
    float pose_score_coeff;     /* final score (MeshUtils.cpp:780-784); the CLI ranks pre-ICP hypotheses with them    */
} hf6d_options;
int hf6d_parse_options(const char* options_path, hf6d_options* out, hf6d_object* objs, int cap);
int hf6d_inspect_forest(const char* forest_dir, hf6d_model_info* out); /* dims[] left zero */
int hf6d_inspect_weights(const char* weights_path, int32_t dims[4]);

/* ---------------------------------------------------------------------------------------------- whole frame */
/* Host buffers: bgr uint8[H][W][3] (OpenCV imread order), depth_mm uint16[H][W].  Synchronous. */
int hf6d_detect(hf6d_ctx* c, const uint8_t* bgr, const uint16_t* depth_mm, hf6d_hypothesis* out, int cap, int* n_out);
/* Pipelined form: up to n_slots frames in flight; H2D copy, kernels and D2H of frame i overlap those of frame i+1.
 * Host buffers must stay valid until the matching hf6d_wait; pinned memory (hf6d_host_alloc) makes the copy async. */
int hf6d_submit(hf6d_ctx* c, const uint8_t* bgr, const uint16_t* depth_mm, int* ticket);
int hf6d_wait(hf6d_ctx* c, int ticket, hf6d_hypothesis* out, int cap, int* n_out);
void* hf6d_host_alloc(size_t bytes); /* pinned */
void hf6d_host_free(void* p);

/* ---------------------------------------------------------------------------------------------- stage level */
/* Used by the parity tests, the bench (device-resident timing) and the multi-GPU driver. slot in [0, n_slots). */
int hf6d_upload(hf6d_ctx* c, int slot, const uint8_t* bgr, const uint16_t* depth_mm); /* async H2D on the slot stream */
/* Make the slot read a frame that already lives in device memory (caller-owned, same layouts); NULL, NULL restores
 * the slot's own frame buffer.  hf6d_upload also restores it. */
int hf6d_bind_frame(hf6d_ctx* c, int slot, const void* d_bgr, const void* d_depth_mm);
int hf6d_run(hf6d_ctx* c, int slot, int first_stage, int last_stage);                 /* async, inclusive range */
int hf6d_sync(hf6d_ctx* c, int slot);
int hf6d_collect(hf6d_ctx* c, int slot, hf6d_hypothesis* out, int cap, int* n_out);   /* D2H + host pose finalise; syncs */
/* Copy a device buffer to the host (syncs the slot).  Returns bytes written, or <0. */
int64_t hf6d_fetch(hf6d_ctx* c, int slot, int what, void* dst, size_t cap_bytes);
/* Overwrite a device buffer from the host: FEATURES (rows = P'), LEAF_ORD, MAPS, COUNTS/LOCS (stage-isolated parity). */
int hf6d_inject(hf6d_ctx* c, int slot, int what, const void* src, size_t bytes);
/* Raw device pointer of a buffer (for NCCL collectives issued by the caller). */
int hf6d_device_ptr(hf6d_ctx* c, int slot, int what, void** ptr, size_t* bytes);
/* Run the slot on a caller-owned CUDA stream (cudaStream_t as void*), e.g. torch's current stream; NULL restores. */
int hf6d_set_stream(hf6d_ctx* c, int slot, void* cuda_stream);
/* Milliseconds per stage of the last hf6d_run on this slot (CUDA events on the slot stream); ms[HF6D_STAGE_COUNT]. */
int hf6d_stage_ms(hf6d_ctx* c, int slot, float* ms);
/* Milliseconds of the three encoder layer launches of the last run that included HF6D_STAGE_ENCODE; ms[3]. */
int hf6d_encoder_layer_ms(hf6d_ctx* c, int slot, float* ms);
/* Bytes of the per-frame result block that hf6d_collect / hf6d_wait copy device -> host. */
int64_t hf6d_result_bytes(const hf6d_ctx* c);
/* Number of kernels the last hf6d_run on this slot launched. */
int hf6d_launch_count(const hf6d_ctx* c, int slot);
/* Diagnostic (syncs the slot, host arithmetic): how many votes the slot's current leaf table casts -- the sum over its
 * (patch, owned tree) pairs of the gated votes of the leaf reached (HFTest.cpp:191-214).  The bench derives the vote and
 * pose stages' algorithmic bytes from it.  Returns the count, or < 0. */
int64_t hf6d_count_cast_votes(hf6d_ctx* c, int slot);

/* Diagnostic (not on the hot path): gathers the slot's P' patches through a real CUDA texture object built exactly as
 * the reference builds its texture (patch_extractor.cu:339-343, fill = 0) and copies them to the host as
 * float[P'][ps][ps][4] (HWC).  Needs a prior hf6d_run(.., SCAN, ..) on the slot.  Returns bytes written, or <0. */
int64_t hf6d_debug_texture_gather(hf6d_ctx* c, int slot, float* dst, size_t cap_bytes);

/* Pre-ICP pose of a hypothesis tuple (HFTest.cpp:922-924 + MeshUtils.cpp:423-440); host arithmetic. */
void hf6d_pose_from_tuple(const hf6d_params* p, int cx, int cy, float z, int yaw_deg, int pitch_deg, int roll_deg,
                          float pose[16]);

/* ---------------------------------------------------------------------------------------------- refinement
 * What HFTest::test_image does with every hypothesis tuple after the Hough stage (HoughForest/src/HFTest.cpp:922-994) and what
 * HFTest::DetectObjects does with the result (HFTest.cpp:1261-1303):
 *
 *   reference interface                                              replaced by
 *   MeshUtils::setIntrinsics / setReg / setGroupReg / set*Threshold /
 *     setClusteringOptions / useColorSimilarity / searchSingle*      hf6d_set_refine_params   (HFTest.cpp:1203-1225)
 *   MeshUtils::insertObjectFromPLY   HoughForest/include/MeshUtils.h:213  hf6d_load_object_ply / hf6d_set_object_model
 *   MeshUtils::setScene              HoughForest/src/MeshUtils.cpp:340     hf6d_refine (first step, on the slot's frame)
 *   MeshUtils::icp                   MeshUtils.cpp:423                     hf6d_refine (pose refinement of every tuple)
 *   MeshUtils::evaluate_hypothesis   MeshUtils.cpp:629                     hf6d_refine (scores + acceptance)
 *   MeshUtils::optimize_hypotheses   MeshUtils.cpp:1160                    hf6d_refine (selection)
 *   the instance cap of the output loop  HFTest.cpp:1269-1273              hf6d_detection.rank
 *
 * The arithmetic the reference leaves to PCL 1.7 (VoxelGrid, NormalEstimation, KdTree, IterativeClosestPoint) is restated
 * from PCL's published algorithms; oracle/refine.py lists the choices.  All of it runs on the GPU; the host enumerates the
 * solution vectors of a hypothesis group (MeshUtils::get_next_solution_vector) and picks the best from the GPU's counts. */
typedef struct {
    float scene_leaf_m, object_leaf_m;   /* VoxelGrid leaf sizes (MeshUtils.h:140-141: 0.005) */
    float normals_radius_m;              /* MeshUtils.cpp:198: 0.03 */
    float nn_search_radius_m;            /* MeshUtils.h:127: 0.01 -- search radius of objects without their own, and the divisor of the depth score */
    float occlusion_threshold_m;         /* MeshUtils.h:128: 0.02 */
    float similarity_coeff, inliers_coeff, clutter_coeff, location_score_coeff, pose_score_coeff;  /* Options.*_coeff */
    float group_total_explain_coeff, group_common_explain_coeff;
    float inliers_threshold, clutter_threshold, final_score_threshold;
    float cluster_eps_angle_threshold, cluster_curvature_threshold, cluster_tolerance_near, cluster_tolerance_far;
    int32_t cluster_min_points;
    int32_t use_color_similarity, use_normal_similarity;
    int32_t search_single_object_instance, search_single_object_in_group;
    int32_t default_icp_iterations;      /* detector_options.proto:11: 60 */
} hf6d_refine_params;

typedef struct {
    int32_t hypothesis;                  /* index into the hypothesis list handed to hf6d_refine */
    int32_t cls;
    float pose[16];                      /* refined 4x4, row-major, camera frame (the Hough pose when ICP did not converge) */
    float similarity, inliers_ratio, clutter, location_score, pose_score, final_score;  /* MeshUtils::HypothesisEvaluation */
    int32_t icp_converged, icp_iterations;
    int32_t visible, inliers, explained; /* visible model points, points with a scene neighbour, scene points explained */
    int32_t accepted;                    /* evaluate_hypothesis returned true */
    int32_t selected;                    /* chosen by optimize_hypotheses */
    int32_t rank;                        /* position in the output of DetectObjects (final score order, at most `instances`
                                            per object), -1 when not written */
} hf6d_detection;

typedef enum {
    HF6D_RBUF_SCENE_POINTS = 0,   /* float[S][4] = x, y, z, bits r | g << 8 | b << 16: the down-sampled scene after normals_not_nan */
    HF6D_RBUF_SCENE_NORMALS = 1,  /* float[S][4] = nx, ny, nz, curvature */
    HF6D_RBUF_SCENE_LABELS = 2,   /* int32[S] smooth-cluster id, -1 = none */
    HF6D_RBUF_CLUSTER_SIZES = 3,  /* int32[n_clusters] */
    HF6D_RBUF_MODEL_POINTS = 4,   /* float[M][4] of object `arg` (down-sampled, normals_not_nan applied) */
    HF6D_RBUF_MODEL_NORMALS = 5,  /* float[M][4] of object `arg`: normals re-estimated on that cloud (NaN rows are dropped by the scoring) */
    HF6D_RBUF_MODEL_VERTICES = 6  /* float[n][3] of object `arg`: the mesh vertices as read (what MeshUtils::renderObject projects) */
} hf6d_refine_buffer;

void hf6d_default_refine_params(hf6d_refine_params* p);
int hf6d_set_refine_params(hf6d_ctx* c, const hf6d_refine_params* p); /* before the models are set: they depend on the leaf size */
int hf6d_get_refine_params(const hf6d_ctx* c, hf6d_refine_params* out);
/* Vertices of the object's mesh as MeshUtils::getPointCloudFromPLY reads them: xyz float[n][3] metres (object frame),
 * rgb uint8[n][3].  nn_search_radius / icp_iterations: ObjectOptions values, -1 = not given (MeshUtils.h:239-245). */
int hf6d_set_object_model(hf6d_ctx* c, int cls, const float* xyz, const uint8_t* rgb, int n, float nn_search_radius,
                          int icp_iterations);
int hf6d_load_object_ply(hf6d_ctx* c, int cls, const char* ply_path, float nn_search_radius, int icp_iterations);
/* For a context made by hf6d_create_from_options: inserts the mesh_file of every object with should_detect, with the object's
 * nn_search_radius and icp_iterations, as DetectObjects does (HFTest.cpp:1227-1233); the MeshUtils settings of the options file
 * (HFTest.cpp:1203-1225) are already the context's refine parameters.  HF6D_EIO names the first mesh that cannot be read. */
int hf6d_load_option_models(hf6d_ctx* c);
/* Refines and scores the n hypotheses (as hf6d_collect / hf6d_wait / hf6d_detect returned them) against the frame the slot
 * still holds (hf6d_detect: slot 0; ticket t: slot t % n_slots, until the next submit reuses it), then selects.  Writes one
 * hf6d_detection per hypothesis, in input order.  Synchronous.  Every detected class needs a model. */
int hf6d_refine(hf6d_ctx* c, int slot, const hf6d_hypothesis* hyps, int n, hf6d_detection* out, int cap, int* n_out);
/* Milliseconds of the last hf6d_refine: ms[0] scene (cloud, VoxelGrid, normals, clusters), ms[1] ICP, ms[2] scoring,
 * ms[3] joint optimisation (GPU kernels + host enumeration). */
int hf6d_refine_ms(hf6d_ctx* c, float ms[4]);
/* Copies a buffer of the last hf6d_refine (or of a model) to the host; returns bytes written, or < 0.  dst == NULL: the size. */
int64_t hf6d_refine_fetch(hf6d_ctx* c, int what, int arg, void* dst, size_t cap_bytes);

/* ---------------------------------------------------------------------------------------------- training
 * `HoughForest --train` (HoughForest/src/main.cpp:41-66): HFTrain::train (HoughForest/src/HFTrain.cpp:1199-1265) on the GPU.
 *
 *   reference interface                                       replaced by
 *   HFTrain::setNumTrees / setMinSamples / setTestsPerNode /
 *     setThresPerTest / setStartTreeNo / setPatchSizeInVoxels /
 *     setVoxelSizeInM              HFTrain.h:66-118           hf6d_train_params
 *   HFTrain::setInputPatchesFilename + getTrainSet  HFTrain.cpp:14-67   hf6d_train_forest (the training-vector file
 *                                                             train_patch_generator writes: int32 K, int32 F, then records
 *                                                             {int32 class, float yaw pitch roll x y z, float[F]})
 *   HFTrain::train                 HFTrain.cpp:1199           hf6d_train_forest / hf6d_train_forest_mem
 *   forest.txt + tree<N>.dat       HFTrain.cpp:1225-1231, HFBase.cpp:4-38   written to output_folder, read back by hf6d_create
 *
 * The reference draws every random number from rand() seeded with the clock, inside OpenMP threads: its forests are not
 * reproducible.  Here the draws come from a counter-based generator keyed by `seed`; for a given seed the forest is the same
 * bit for bit on every run (oracle/train.py restates the algorithm with the same draws and is the checker).  No CPU path. */
typedef struct {
    int32_t trees;                /* --trees (3) */
    int32_t min_samples;          /* --min_samples (30): a child with at most this many samples is a leaf */
    int32_t tests_per_node;       /* --tests_per_node (30) */
    int32_t thresholds_per_test;  /* --thresholds_per_test (10) */
    int32_t start_tree_no;        /* --start_tree_no (0): files tree<start> .. tree<start + trees - 1> */
    int32_t patch_size_in_voxels; /* --patch_size_in_voxels: only copied into forest.txt */
    float voxel_size_in_m;        /* --voxel_size_in_m: only copied into forest.txt */
    uint64_t seed;
    int32_t device;
} hf6d_train_params;

typedef struct {
    int64_t nodes, leaves;        /* over all trees */
    int32_t max_depth;
    int32_t training_samples;     /* per tree: int(2/3 * samples), HFTrain.cpp:1143 */
    float train_ms;               /* GPU time of all trees (CUDA events), file writing included */
} hf6d_train_stats;

void hf6d_default_train_params(hf6d_train_params* p);
/* input_file: the training vectors (see above); output_folder must exist.  stats may be NULL.  Errors: hf6d_last_error(NULL). */
int hf6d_train_forest(const hf6d_train_params* p, const char* input_file, const char* output_folder, hf6d_train_stats* stats);
/* The same from memory: cls int32[n] in [0, K), dof float[n][6] = yaw, pitch, roll, x, y, z, features float[n][F]. */
int hf6d_train_forest_mem(const hf6d_train_params* p, int n, int K, int F, const int32_t* cls, const float* dof,
                          const float* features, const char* output_folder, hf6d_train_stats* stats);

/* ---------------------------------------------------------------------------------------------- view renderer
 * `PatchGen --render` (PatchGen/src/main.cpp:62-81): RenderViewsTesselatedSphere::generateViews
 * (PatchGen/src/render_views_tesselated_sphere_mod.cpp:140-342) without VTK / OpenGL -- the reference's camera geometry
 * (tessellated-sphere directions, heights, in-plane rotations, view-up rule, area-weighted focal point, 45.3105 degree view
 * angle) and output contract (8-bit colour on white, uint16 millimetre depth with 0 = no surface, the 4 x 4 view transform
 * of pose<N>.txt), pixels from a z-buffer triangle rasteriser on the GPU.
 *
 *   reference interface                                            replaced by
 *   setPlyFileName / setTesselationLevel / setInPlaceCamRotations /
 *     setLightings / setHeight / setStartHeight / setAboveZ / setBelowZ /
 *     setRenderAround0 / setObjectRadius / setResolution   .h:85-150   hf6d_render_params + hf6d_renderer_create*
 *   generateViews' camera loop                              .cpp:255-340   hf6d_renderer_view_count / hf6d_renderer_view
 *   render_win->Render() + save_rendering's buffers         .cpp:60-104    hf6d_render
 *
 * Shading, fill rule and depth precision are OpenGL's in the reference and therefore choices here (oracle/render.py V1-V4). */
typedef struct {
    int32_t W, H;                 /* 640 x 480 */
    float view_angle_deg;         /* vertical view angle, 45.3105 -> f = 575 px at H = 480 */
    int32_t tesselation_level;    /* --tessel_level (1) */
    int32_t use_vertices;         /* camera on the sphere's vertices (1, the class default) or face centres */
    int32_t in_place_rotations;   /* --inPlaceCamRot (24) */
    int32_t lightings;            /* --lightings (3): every view is rendered with ambient = 0, 0.1, .. */
    int32_t heights;              /* --numHeights (4) */
    float height_step;            /* --heightStep (0.25 m) */
    float start_height;           /* --startHeight (0.3 m) */
    int32_t above_z, below_z, render_around_0;
    float object_radius;          /* --object_radius (-1: the mesh's largest extent) */
    int32_t device;
} hf6d_render_params;

typedef struct hf6d_renderer hf6d_renderer;
void hf6d_default_render_params(hf6d_render_params* p);
/* ASCII PLY with coloured vertices and faces (polygons are fanned), or the mesh from memory: xyz float[n][3], rgb uint8[n][3],
 * faces int32[m][3].  Errors: hf6d_last_error(NULL).  No CPU path. */
int hf6d_renderer_create_ply(const hf6d_render_params* p, const char* ply_path, hf6d_renderer** out);
int hf6d_renderer_create(const hf6d_render_params* p, const float* xyz, const uint8_t* rgb, int n_vertices, const int32_t* faces,
                         int n_faces, hf6d_renderer** out);
void hf6d_renderer_destroy(hf6d_renderer* r);
/* Camera poses of generateViews in its order (direction, height, in-plane rotation); each is rendered `lightings` times by the
 * reference.  pose: row-major 4 x 4 world -> camera, camera looking down -z, y up (what pose<N>.txt holds). */
int hf6d_renderer_view_count(const hf6d_renderer* r);
int hf6d_renderer_view(const hf6d_renderer* r, int view, double pose[16]);
/* One view with any pose: bgr uint8[H][W][3] (row 0 = top), depth_mm uint16[H][W].  ambient = lighting index * 0.1. */
int hf6d_render(hf6d_renderer* r, const double pose[16], float ambient, uint8_t* bgr, uint16_t* depth_mm);

/* ---------------------------------------------------------------------------------------------- patch database + training vectors
 * The files between PatchGen and HoughForest --train, in the reference's own formats, so that either side of a reference
 * training run can be swapped for this library:
 *
 *   reference                                                                   replaced by
 *   mdb_env_open / mdb_put / mdb_txn_commit of "%04d_%08d" -> caffe::Datum       hf6d_patchdb_create / _put / _close
 *     patch_generator.cpp:479-493, 562-580 (read back by Caffe's DATA layer)
 *   mdb_cursor_get(MDB_FIRST / MDB_NEXT) + Datum::ParseFromArray                 hf6d_patchdb_open / _next
 *     train_patch_generator.cpp:33-52, 79-101
 *   patch_generator::get_yaw_pitch_roll_from_rot_mat + get_object_coords        hf6d_patch_annotation
 *     patch_generator.cpp:20-56 (one line of patch_annotation_lmdb.txt)
 *   patch_extractor_gpu::extract_patches_rgbd + the normalisation of            hf6d_create_extractor, then hf6d_upload +
 *     insert_patches_to_db_rgbd (patch_generator.cpp:382-470: the same            hf6d_run(SCAN..GATHER) + hf6d_fetch(LOCS /
 *     arithmetic as HFTest.cpp:500-570)                                           PATCH_U8): the Datum bytes
 *   train_patch_generator::generate_train_patches                               hf6d_generate_train_vectors (hf6d_encode_patches
 *     train_patch_generator.cpp:22-200                                            per batch of database entries)
 *
 * data.mdb is LMDB's format version 1 with 4096-byte pages (lmdb 0.9.x on x86-64), written as one compact committed tree;
 * the reader follows the current meta page of any such file.  Keys must be put in ascending order (patch_generator's are).
 * Host code, except hf6d_encode_patches / hf6d_generate_train_vectors, whose encoder has no CPU path. */
typedef struct hf6d_patchdb hf6d_patchdb;
int hf6d_patchdb_create(const char* folder, hf6d_patchdb** out); /* folder must exist and hold no data.mdb */
int hf6d_patchdb_open(const char* folder, hf6d_patchdb** out);
/* One patch: data = uint8[channels][height][width]; label = object index (Datum.label). */
int hf6d_patchdb_put(hf6d_patchdb* db, const char* key, int channels, int height, int width, int label, const uint8_t* data);
int64_t hf6d_patchdb_entries(const hf6d_patchdb* db);
/* Next entry in key order: 1 = filled, 0 = end of the database, < 0 = error.  dims = {channels, height, width, label};
 * *n_data = bytes of Datum.data (copied when they fit cap_bytes, HF6D_EINVAL when they do not). */
int hf6d_patchdb_next(hf6d_patchdb* db, char key[64], int32_t dims[4], uint8_t* data, size_t cap_bytes, size_t* n_data);
int hf6d_patchdb_close(hf6d_patchdb* db); /* a writer commits here; frees the handle either way */

/* (yaw, pitch, roll, x, y, z) of a patch centred on pixel (x, y) with depth_mm, seen under the view transform `pose`
 * (row-major 4 x 4, what pose<N>.txt holds): the vote the forest is trained on.  Focal length from the vertical view angle. */
void hf6d_patch_annotation(int W, int H, float view_angle_deg, int x, int y, uint16_t depth_mm, const float pose[16],
                           float out[6]);

/* A context that only extracts patches and encodes them: no forest is read (TRAVERSE and later stages vote for nothing);
 * weights_path may be NULL when only SCAN..GATHER will run (the auto-encoder does not exist yet when patches are generated).
 * p->patch_vox and p->voxel_m are taken from p (there is no forest.txt to overrule them). */
int hf6d_create_extractor(const hf6d_params* p, const char* weights_path, int device, hf6d_ctx** out);
int hf6d_patch_capacity(const hf6d_ctx* c); /* most patches one frame / one hf6d_encode_patches call can hold */
/* The encoder over caller-held quantised patches (uint8[n][C*ps*ps], CHW: Datum.data): features float[n][F]. */
int hf6d_encode_patches(hf6d_ctx* c, int slot, const uint8_t* patches, int n, float* features);

typedef struct {
    int64_t entries;   /* in the database */
    int64_t written;   /* training vectors written: batch_size * floor((entries - 1) / batch_size) -- the reference stops
                          at the batch in which the cursor reaches the end (train_patch_generator.cpp:98-106) */
    int32_t classes, feature_length;
    float encode_ms;
} hf6d_trainvec_stats;
/* lmdb_folder: data.mdb + patch_annotation_lmdb.txt.  output_file: int32 K, int32 F, then per patch int32 object, float
 * yaw pitch roll x y z, float[F] (what hf6d_train_forest reads).  stats may be NULL.  Errors: hf6d_last_error(NULL). */
int hf6d_generate_train_vectors(const char* weights_path, const char* lmdb_folder, const char* output_file, int batch_size,
                                int device, int encoder_mode, hf6d_trainvec_stats* stats);

#ifdef __cplusplus
}
#endif
#endif
