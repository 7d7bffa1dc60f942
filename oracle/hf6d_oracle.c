/* hf6d CPU oracle -- TEST INFRASTRUCTURE ONLY (see hf6d_oracle.h).  PINNED to the reference's own sources: HFBase.cpp,
 * HFTest.cpp, patch_extractor.cu, surface_normals.cu and the head of MeshUtils::icp, compiled where they lie under
 * /root/reference against stand-in headers (oracle/build_ref.py -> oracle/_ref/), agree with this file bit for bit on seeded
 * inputs, the whole per-frame path included (tests/test_ref_pins.py).  Still UNPINNED, because the arithmetic lives in absent
 * third-party code: the encoder (Caffe / BLAS summation order, choice C5), the texture unit's filter (C1, checked on the B200),
 * cv::blur's accumulation (C9, checked against cv2), the clock-seeded fill (C2).  The reference holds no golden vectors.
 *
 * Restates, stage by stage, what `HoughForest --test` computes for one RGB-D frame:
 *   A1  texture build                 HoughForest/src/HFTest.cpp:370-379
 *   A2a valid-centre scan             PatchGen/src/cuda/patch_extractor.cu:372-391
 *   A2b RGB-D patch gather            PatchGen/src/cuda/patch_extractor.cu:230-309
 *   A3  local normalise + quantise    HoughForest/src/HFTest.cpp:500-570
 *   A4  auto-encoder forward          HoughForest/src/HFTest.cpp:585-596, generate_scripts.sh:424-524
 *   A5  forest file format            HoughForest/src/HFBase.cpp:58-145, HoughForest/include/HFBase.h:23-60
 *   A6  tree descent                  HoughForest/src/HFTest.cpp:144-163
 *   A7  vote casting                  HoughForest/src/HFTest.cpp:21-102, 166-217
 *   A9  box blur + sliding-window NMS HoughForest/src/HFTest.cpp:219-268, 702-707
 *   A10 z / yaw-pitch accumulation    HoughForest/src/HFTest.cpp:742-802
 *   A11 z / pose mode seeking         HoughForest/src/HFTest.cpp:803-925
 *   A12 pre-ICP pose                  HoughForest/src/HFTest.cpp:922-928, HoughForest/src/MeshUtils.cpp:29-59, 423-440
 *
 * Build: gcc -O3 -fopenmp -ffp-contract=off -mavx2 (no FMA contraction anywhere: the reference was an SSE2 build).
 *
 * CHOICES this file makes where the reference leaves semantics to a library, hardware or UB:
 *  C1 texture filter: software bilinear, fraction rounded to 8 bits (round-to-nearest), border texel = 0,
 *     S = ((w00*T00 + w10*T10) + w01*T01) + w11*T11 in fp32.  (For patch_vox = 8 the fraction is k/8: exact.)
 *  C2 border fill values: counter-based hash of (fill_seed, patch index) instead of clock64()-seeded cuRAND.
 *  C3 the variance terms are evaluated in double: `std += pow(x - mean, 2) / N` (HFTest.cpp:527-534) is an unqualified call,
 *     and on the reference's tested toolchain (README.md:16, Ubuntu 14.04 = gcc 4.8 / glibc 2.19, whose <math.h> puts only
 *     the C functions in the global namespace) it binds to pow(double, double): the float deviation is squared in double
 *     (exact), divided by (double)N, added to (double)std, and the sum narrowed back to the float accumulator -- per
 *     element.  A present-day gcc in its default dialect evaluates the same expression in double as well (C++11
 *     std::pow(float, int) promotes).  The means (`mean += x / N`, float / int) stay in float.  Pinned by
 *     tests/test_ref_pins.py against the reference's own source compiled with a C-only <math.h>.
 *  C4 (unsigned char)(NaN) == 0 (x86 cvttss2si low byte).
 *  C5 encoder: fp32, fixed accumulation order (8 interleaved partial sums, pairwise combine), bias added last,
 *     sigmoid = 1/(1+expf(-x)).  Caffe/BLAS order is unknowable.
 *  C6 cos/sin of vote angles evaluated in double, narrowed to float (HFTest.cpp:45-50 with <math.h>).
 *  C7 Eigen 4x4 * 4-vector accumulates k = 0..3 left to right.
 *  C8 vote weights are Q16 fixed point accumulated in integers (the reference's float maps are summed in a
 *     thread-schedule dependent order, HFTest.cpp:645-654, so they are only defined up to rounding anyway);
 *     integer sums make every later stage order-independent and bit-reproducible across GPUs.
 *  C9 box filter = exact integer window sum, then (float)((double)S/65536 * (1.0/(kx*ky)))  (cv::blur on CV_32F sums
 *     in double and scales once).
 *  C10 NMS result order for equal scores: emission order (std::sort is unstable in the reference).
 *  C11 float->int conversions that overflow or are NaN give INT_MIN (x86).
 *  C12 (A2c, the RGB + surface-normals patch mode) the normal computation of surface_normals.cu and the length
 *     normalisations of patch_extractor.cu:70-84 are evaluated in fp32 without FMA contraction; pixels the reference's
 *     launch leaves unwritten are zero; the random normal used as fill comes from the counter-based generator of C2.
 */
#include "hf6d_oracle.h"

#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------------ forest (A5) */
typedef struct ref_node {
    uint8_t leaf;
    int32_t leaf_id;
    int32_t mode, f1, f2;
    float thr;
    struct ref_node *left, *right;
    float* class_prob; /* [K] */
    int32_t* nvotes;   /* [K] */
    float** votes;     /* [K][n*6] yaw,pitch,roll,x,y,z */
    int32_t ordinal;   /* file-order index among the leaves of this tree */
    int64_t gordinal;  /* ordinal + leaves of previous trees */
} ref_node;

struct hf6d_ref_forest {
    int32_t T, K, F, ps;
    float vox;
    ref_node** roots;
    int32_t* n_leaves;
    int32_t* n_internal;
    ref_node*** leaves; /* [T][n_leaves[t]] */
    int32_t* leaf_cap;
};

static void free_node(ref_node* n, int K) {
    if (!n) return;
    free_node(n->left, K);
    free_node(n->right, K);
    if (n->votes)
        for (int c = 0; c < K; ++c) free(n->votes[c]);
    free(n->votes);
    free(n->nvotes);
    free(n->class_prob);
    free(n);
}

static ref_node* load_node(FILE* fp, hf6d_ref_forest* f, int t, int* ok) {
    ref_node* n = (ref_node*)calloc(1, sizeof(ref_node));
    uint8_t leaf;
    if (fread(&leaf, 1, 1, fp) != 1) { *ok = 0; return n; }
    n->leaf = leaf;
    n->leaf_id = -1;
    if (leaf) {
        const int K = f->K;
        if (fread(&n->leaf_id, 4, 1, fp) != 1) { *ok = 0; return n; }
        n->class_prob = (float*)malloc(sizeof(float) * K);
        if (fread(n->class_prob, 4, K, fp) != (size_t)K) { *ok = 0; return n; }
        n->nvotes = (int32_t*)calloc(K, sizeof(int32_t));
        n->votes = (float**)calloc(K, sizeof(float*));
        for (int c = 0; c < K; ++c) {
            int32_t nm;
            if (fread(&nm, 4, 1, fp) != 1 || nm < 0) { *ok = 0; return n; }
            n->nvotes[c] = nm;
            n->votes[c] = (float*)malloc(sizeof(float) * 6 * (nm > 0 ? nm : 1));
            if (nm > 0 && fread(n->votes[c], 4, (size_t)nm * 6, fp) != (size_t)nm * 6) { *ok = 0; return n; }
        }
        if (f->n_leaves[t] == f->leaf_cap[t]) {
            f->leaf_cap[t] = f->leaf_cap[t] ? f->leaf_cap[t] * 2 : 1024;
            f->leaves[t] = (ref_node**)realloc(f->leaves[t], sizeof(ref_node*) * f->leaf_cap[t]);
        }
        n->ordinal = f->n_leaves[t];
        f->leaves[t][f->n_leaves[t]++] = n;
    } else {
        int32_t hdr[3];
        if (fread(hdr, 4, 3, fp) != 3 || fread(&n->thr, 4, 1, fp) != 1) { *ok = 0; return n; }
        n->mode = hdr[0];
        n->f1 = hdr[1];
        n->f2 = hdr[2];
        f->n_internal[t]++;
        n->left = load_node(fp, f, t, ok);
        if (!*ok) return n;
        n->right = load_node(fp, f, t, ok);
    }
    return n;
}

hf6d_ref_forest* hf6d_ref_forest_load(const char* dir) {
    char path[4096];
    snprintf(path, sizeof path, "%s/forest.txt", dir);
    FILE* fp = fopen(path, "r");
    if (!fp) return NULL;
    hf6d_ref_forest* f = (hf6d_ref_forest*)calloc(1, sizeof *f);
    if (fscanf(fp, "%d %d %d %d %f", &f->T, &f->K, &f->F, &f->ps, &f->vox) != 5 || f->T <= 0 || f->K <= 0) {
        fclose(fp);
        free(f);
        return NULL;
    }
    fclose(fp);
    f->roots = (ref_node**)calloc(f->T, sizeof(ref_node*));
    f->n_leaves = (int32_t*)calloc(f->T, sizeof(int32_t));
    f->n_internal = (int32_t*)calloc(f->T, sizeof(int32_t));
    f->leaves = (ref_node***)calloc(f->T, sizeof(ref_node**));
    f->leaf_cap = (int32_t*)calloc(f->T, sizeof(int32_t));
    int64_t base = 0;
    for (int t = 0; t < f->T; ++t) {
        snprintf(path, sizeof path, "%s/tree%d.dat", dir, t);
        fp = fopen(path, "rb");
        int ok = fp != NULL;
        if (ok) {
            f->roots[t] = load_node(fp, f, t, &ok);
            fclose(fp);
        }
        if (!ok) {
            hf6d_ref_forest_free(f);
            return NULL;
        }
        for (int i = 0; i < f->n_leaves[t]; ++i) f->leaves[t][i]->gordinal = base + i;
        base += f->n_leaves[t];
    }
    return f;
}

void hf6d_ref_forest_free(hf6d_ref_forest* f) {
    if (!f) return;
    for (int t = 0; t < f->T; ++t) {
        free_node(f->roots[t], f->K);
        free(f->leaves[t]);
    }
    free(f->roots);
    free(f->n_leaves);
    free(f->n_internal);
    free(f->leaves);
    free(f->leaf_cap);
    free(f);
}

float hf6d_ref_forest_info(const hf6d_ref_forest* f, int32_t* info) {
    int32_t nl = 0, ni = 0;
    for (int t = 0; t < f->T; ++t) {
        nl += f->n_leaves[t];
        ni += f->n_internal[t];
    }
    info[0] = f->T;
    info[1] = f->K;
    info[2] = f->F;
    info[3] = f->ps;
    info[4] = nl;
    info[5] = ni;
    return f->vox;
}

int32_t hf6d_ref_tree_leaf_count(const hf6d_ref_forest* f, int32_t t) { return f->n_leaves[t]; }

void hf6d_ref_default_params(hf6d_ref_params* p) {
    memset(p, 0, sizeof *p);
    p->W = 640;
    p->H = 480;
    p->stride = 2; /* generate_scripts.sh:53; HFTest.h:98 */
    p->fx = 575.f;
    p->fy = 575.f;
    p->cx = 319.5f;
    p->cy = 239.5f;
    p->patch_vox = 8;
    p->voxel_m = 0.005f;
    p->max_depth_range_m = 0.25f;
    p->distance_threshold_m = 1.5f;
    p->fill_random = 1;
    p->fill_seed = 0;
    p->batch_size = 100;
    p->max_yaw_pitch_hypotheses = 7; /* HFTest.h:175-183 */
    p->max_roll_hypotheses = 3;
    p->min_location_score_ratio = 1.0f / 1000.0f;
    p->min_yaw_pitch_drop_ratio = 1.0f / 1000.0f;
    p->centers_blur_size = 13;
    p->centers_nms_wsize = 40;
    p->pose_blur_size = 35;
    p->pose_nms_wsize = 35;
    p->patch_mode = 0;
    p->normals_focal = 575.0f; /* HFTest.cpp:329, :356 */
}

static inline int32_t d2i_x86(double y) { /* C11 */
    if (!(y == y) || y >= 2147483648.0 || y < -2147483648.0) return INT_MIN;
    return (int32_t)y;
}
static inline int32_t f2i_x86(float y) { /* C11 */
    if (!(y == y) || y >= 2147483648.0f || y < -2147483648.0f) return INT_MIN;
    return (int32_t)y;
}

/* ------------------------------------------------------------------------------------------------ A2a */
static int adaptive_size(const hf6d_ref_params* p, float depth_m) {
    /* patch_extractor.cu:257 / :378 -- ((ps*vox)/d)*f, float, truncated */
    /* the RGB-D extractor is handed fx (HFTest.cpp:394), the normals variant the literal 575.0f (HFTest.cpp:356) */
    float v = (float)p->patch_vox * p->voxel_m / depth_m * (p->patch_mode ? p->normals_focal : p->fx);
    return (int)v;
}

int32_t hf6d_ref_scan_centres(const uint16_t* depth, const hf6d_ref_params* p, int32_t* locs, int32_t cap) {
    int32_t n = 0;
    for (int h = 0; h < p->H; h += p->stride)
        for (int w = 0; w < p->W; w += p->stride) {
            float d = (float)depth[(size_t)h * p->W + w];
            if (d != 0 && d / 1000.0f < p->distance_threshold_m) {
                int a = adaptive_size(p, d / 1000.0f);
                int x0 = w - a / 2, x1 = x0 + a - 1;
                int y0 = h - a / 2, y1 = y0 + a - 1;
                if (x0 >= 0 && y0 >= 0 && x1 < p->W && y1 < p->H) {
                    if (n < cap) {
                        locs[2 * n] = w;
                        locs[2 * n + 1] = h;
                    }
                    ++n;
                }
            }
        }
    return n;
}

/* ------------------------------------------------------------------------------------------------ A1 + A2b */
static inline float texel(const uint8_t* bgr, const uint16_t* depth, int W, int H, int x, int y, int ch) {
    if (x < 0 || y < 0 || x >= W || y >= H) return 0.0f; /* cudaAddressModeBorder */
    if (ch == 3) return (float)depth[(size_t)y * W + x];
    return (float)bgr[((size_t)y * W + x) * 3 + ch] / 255.0f; /* HFTest.cpp:374-376 */
}

static inline float frac8(float a) { /* C1: 1.8 fixed-point filter weight */
    return floorf(a * 256.0f + 0.5f) / 256.0f;
}

static inline float bilinear(const uint8_t* bgr, const uint16_t* depth, int W, int H, float u, float v, int ch) {
    float fu = floorf(u), fv = floorf(v);
    int i = (int)fu, j = (int)fv;
    float a = frac8(u - fu), b = frac8(v - fv);
    float w00 = (1.0f - a) * (1.0f - b), w10 = a * (1.0f - b), w01 = (1.0f - a) * b, w11 = a * b;
    float t00 = texel(bgr, depth, W, H, i, j, ch), t10 = texel(bgr, depth, W, H, i + 1, j, ch);
    float t01 = texel(bgr, depth, W, H, i, j + 1, ch), t11 = texel(bgr, depth, W, H, i + 1, j + 1, ch);
    float s = w00 * t00;
    s = s + w10 * t10;
    s = s + w01 * t01;
    s = s + w11 * t11;
    return s;
}

static inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

void hf6d_ref_gather(const uint8_t* bgr, const uint16_t* depth, const hf6d_ref_params* p, const int32_t* locs,
                     int32_t P, float* patches) {
    const int ps = p->patch_vox, W = p->W, H = p->H;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; ++i) {
        const int cx = locs[2 * i], cy = locs[2 * i + 1];
        const float dc = (float)depth[(size_t)cy * W + cx] / 1000.0f; /* exact texel fetch :253 */
        const int a = adaptive_size(p, dc);
        const int x0 = cx - a / 2, y0 = cy - a / 2;
        const float step = (float)a / (float)ps;
        float fill[4] = {0, 0, 0, 0};
        if (p->fill_random) { /* C2 */
            uint64_t z = mix64(p->fill_seed + 0x9E3779B97F4A7C15ULL * (uint64_t)(i + 1));
            float r = (float)((z & 0xFFFF) % 255) / 255.0f;
            float g = (float)(((z >> 16) & 0xFFFF) % 255) / 255.0f;
            float b = (float)(((z >> 32) & 0xFFFF) % 255) / 255.0f;
            float d = (float)(((z >> 48) & 0xFFFF) % 255) / 255.0f;
            fill[0] = b; fill[1] = g; fill[2] = r; fill[3] = d; /* :295-298 */
        }
        float* out = patches + (size_t)i * ps * ps * 4;
        for (int ty = 0; ty < ps; ++ty)
            for (int tx = 0; tx < ps; ++tx) {
                float u = (float)x0 + (float)tx * step;
                float v = (float)y0 + (float)ty * step;
                float d = bilinear(bgr, depth, W, H, u, v, 3) / 1000.0f;
                float* o = out + (ty * ps + tx) * 4;
                if (d > 0) {
                    o[0] = bilinear(bgr, depth, W, H, u, v, 0);
                    o[1] = bilinear(bgr, depth, W, H, u, v, 1);
                    o[2] = bilinear(bgr, depth, W, H, u, v, 2);
                    float td = (d - dc) / p->max_depth_range_m + 0.5f;
                    if (td > 1.0f) td = 1.0f;
                    if (td < 0.0f) td = 0.0f;
                    o[3] = td;
                } else {
                    o[0] = fill[0]; o[1] = fill[1]; o[2] = fill[2]; o[3] = fill[3];
                }
            }
    }
}

/* ------------------------------------------------------------------------------------------------ A2c */
/* surface_normals.cu:11-73.  The texture fetches sit on texel centres, so they return the texel itself.
 * CHOICE (C12): products and sums are evaluated without FMA contraction, like everything else here (the reference's
 * kernel is compiled by nvcc, whose contraction choices are not recoverable); pixels the reference's launch
 * configuration leaves unwritten when W*H is not a multiple of 64 (:99-100) are zero. */
void hf6d_ref_normals(const uint16_t* depth, int32_t W, int32_t H, float focal, float* normals) {
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            float* o = normals + ((size_t)y * W + x) * 3;
            o[0] = o[1] = o[2] = 0.0f;
            if (!(x > 0 && x < W - 1 && y > 0 && y < H - 1)) continue;
            const float z = (float)depth[(size_t)y * W + x] / 1000.0f;
            const float z_left = (float)depth[(size_t)y * W + x - 1] / 1000.0f;
            const float z_right = (float)depth[(size_t)y * W + x + 1] / 1000.0f;
            const float z_up = (float)depth[(size_t)(y - 1) * W + x] / 1000.0f;
            const float z_down = (float)depth[(size_t)(y + 1) * W + x] / 1000.0f;
            if (!(z != 0 && z_left != 0 && z_right != 0 && z_up != 0 && z_down != 0)) continue;
            const float hw = (float)W / 2.0f, hh = (float)H / 2.0f;
            const float x_left = ((float)x - 1 - hw) * z_left / focal;
            const float x_right = ((float)x + 1 - hw) * z_right / focal;
            const float x_up = ((float)x - hw) * z_up / focal;
            const float x_down = ((float)x - hw) * z_down / focal;
            const float y_left = ((float)y - hh) * z_left / focal;
            const float y_right = ((float)y - hh) * z_right / focal;
            const float y_up = ((float)y - 1 - hh) * z_up / focal;
            const float y_down = ((float)y + 1 - hh) * z_down / focal;
            const float ax = (x_left - x_right) / 2.0f, ay = (y_left - y_right) / 2.0f, az = (z_left - z_right) / 2.0f;
            const float bx = (x_down - x_up) / 2.0f, by = (y_down - y_up) / 2.0f, bz = (z_down - z_up) / 2.0f;
            float nx = -(ay * bz - az * by);
            float ny = -(az * bx - ax * bz);
            float nz = -(ax * by - ay * bx);
            const float mag = sqrtf(nx * nx + ny * ny + nz * nz);
            o[0] = nx / mag; /* mag == 0 gives NaN, as in the reference */
            o[1] = ny / mag;
            o[2] = nz / mag;
        }
}

static inline float texel7(const uint8_t* bgr, const uint16_t* depth, const float* normals, int W, int H, int x, int y,
                           int ch) {
    if (x < 0 || y < 0 || x >= W || y >= H) return 0.0f; /* cudaAddressModeBorder */
    if (ch < 3) return (float)bgr[((size_t)y * W + x) * 3 + ch] / 255.0f;
    if (ch == 3) return (float)depth[(size_t)y * W + x];
    return normals[((size_t)y * W + x) * 3 + (ch - 4)];
}

static inline float bilinear7(const uint8_t* bgr, const uint16_t* depth, const float* normals, int W, int H, float u,
                              float v, int ch) {
    float fu = floorf(u), fv = floorf(v);
    int i = (int)fu, j = (int)fv;
    float a = frac8(u - fu), b = frac8(v - fv);
    float w00 = (1.0f - a) * (1.0f - b), w10 = a * (1.0f - b), w01 = (1.0f - a) * b, w11 = a * b;
    float s = w00 * texel7(bgr, depth, normals, W, H, i, j, ch);
    s = s + w10 * texel7(bgr, depth, normals, W, H, i + 1, j, ch);
    s = s + w01 * texel7(bgr, depth, normals, W, H, i, j + 1, ch);
    s = s + w11 * texel7(bgr, depth, normals, W, H, i + 1, j + 1, ch);
    return s;
}

/* patch_extractor.cu:12-111.  CHOICE (C2 again): the per-patch fill values come from a counter-based generator keyed
 * on (fill_seed, patch index) instead of clock64()-seeded cuRAND; a zero-length draw is retried (:21-40). */
void hf6d_ref_gather_normals(const uint8_t* bgr, const uint16_t* depth, const float* normals, const hf6d_ref_params* p,
                             const int32_t* locs, int32_t P, float* patches) {
    const int ps = p->patch_vox, W = p->W, H = p->H;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; ++i) {
        const int cx = locs[2 * i], cy = locs[2 * i + 1];
        const float dc = (float)depth[(size_t)cy * W + cx] / 1000.0f;
        const int a = adaptive_size(p, dc);
        const int x0 = cx - a / 2, y0 = cy - a / 2;
        const float step = (float)a / (float)ps;
        float fill[6] = {0, 0, 0, 0, 0, 0};
        if (p->fill_random) {
            uint64_t z = mix64(p->fill_seed + 0x9E3779B97F4A7C15ULL * (uint64_t)(i + 1));
            fill[2] = (float)((z & 0xFFFF) % 255) / 255.0f;         /* r */
            fill[1] = (float)(((z >> 16) & 0xFFFF) % 255) / 255.0f; /* g */
            fill[0] = (float)(((z >> 32) & 0xFFFF) % 255) / 255.0f; /* b */
            fill[3] = 0.0f; fill[4] = 0.0f; fill[5] = 1.0f;
            for (int attempt = 0; attempt < 4; ++attempt) {
                const uint64_t z2 = mix64(z + 0x9E3779B97F4A7C15ULL), z3 = mix64(z2 + 0x9E3779B97F4A7C15ULL);
                z = z3;
                const float xr = (float)((uint32_t)z2 % 100000u) - 50000.0f;
                const float yr = (float)((uint32_t)(z2 >> 32) % 100000u) - 50000.0f;
                const float zr = (float)((uint32_t)z3 % 50000u); /* z >= 0: the normal faces the camera */
                const float norm = sqrtf(xr * xr + yr * yr + zr * zr);
                if (norm != 0) { fill[3] = xr / norm; fill[4] = yr / norm; fill[5] = zr / norm; break; }
            }
        }
        float* out = patches + (size_t)i * ps * ps * 6;
        for (int ty = 0; ty < ps; ++ty)
            for (int tx = 0; tx < ps; ++tx) {
                const float u = (float)x0 + (float)tx * step, v = (float)y0 + (float)ty * step;
                float* o = out + (ty * ps + tx) * 6;
                int in_object = 0;
                const float d = bilinear7(bgr, depth, normals, W, H, u, v, 3) / 1000.0f;
                if (d > 0) {
                    const float x = bilinear7(bgr, depth, normals, W, H, u, v, 4);
                    const float y = bilinear7(bgr, depth, normals, W, H, u, v, 5);
                    const float z = bilinear7(bgr, depth, normals, W, H, u, v, 6);
                    const float norm = sqrtf(x * x + y * y + z * z);
                    if (norm > 0) {
                        o[0] = bilinear7(bgr, depth, normals, W, H, u, v, 0);
                        o[1] = bilinear7(bgr, depth, normals, W, H, u, v, 1);
                        o[2] = bilinear7(bgr, depth, normals, W, H, u, v, 2);
                        o[3] = x / norm; o[4] = y / norm; o[5] = z / norm;
                        in_object = 1;
                    }
                }
                if (!in_object)
                    for (int c = 0; c < 6; ++c) o[c] = fill[c];
            }
    }
}

/* HFTest.cpp:443-470: colour channels (uchar)(v*255.0f); normals (uchar)((v/2.0 + 0.5f)*255.0f) -- the literal 2.0
 * makes that expression double.  HWC -> CHW. */
void hf6d_ref_quantise_normals(const float* patches, int32_t P, int32_t ps, uint8_t* q) {
    const int n6 = ps * ps * 6;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; ++i) {
        const float* src = patches + (size_t)i * n6;
        int pos = 0;
        for (int c = 0; c < 6; ++c)
            for (int row = 0; row < ps; ++row)
                for (int col = 0; col < ps; ++col) {
                    const float v = src[row * ps * 6 + col * 6 + c];
                    int32_t k;
                    if (c < 3) k = f2i_x86(v * 255.0f);
                    else k = d2i_x86(((double)v / 2.0 + (double)0.5f) * (double)255.0f);
                    q[(size_t)i * n6 + pos++] = (uint8_t)(k & 0xFF); /* C4 */
                }
    }
}

/* ------------------------------------------------------------------------------------------------ A3 */

void hf6d_ref_normalise(const float* patches, int32_t P, int32_t ps, uint8_t* q) {
    const int n1 = ps * ps, n3 = ps * ps * 3, n4 = ps * ps * 4;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; ++i) {
        float buf[4 * 32 * 32];
        const float* src = patches + (size_t)i * n4;
        float mean_rgb = 0, mean_d = 0;
        int pos = 0;
        for (int c = 0; c < 4; ++c)
            for (int row = 0; row < ps; ++row)
                for (int col = 0; col < ps; ++col) {
                    buf[pos] = src[row * ps * 4 + col * 4 + c];
                    if (c < 3) mean_rgb += buf[pos] / (float)n3;
                    else mean_d += buf[pos] / (float)n1;
                    pos++;
                }
        float var_rgb = 0, var_d = 0; /* "std" in the source, never sqrt'ed */
        for (int j = 0; j < n4; ++j) { /* C3: pow(double, double), double division and addition, narrowed per element */
            if (j < n3) { const double d = (double)(buf[j] - mean_rgb); var_rgb = (float)((double)var_rgb + (d * d) / (double)n3); }
            else { const double d = (double)(buf[j] - mean_d); var_d = (float)((double)var_d + (d * d) / (double)n1); }
        }
        for (int j = 0; j < n4; ++j) {
            const float m = j < n3 ? mean_rgb : mean_d;
            const float lim = 3 * (j < n3 ? var_rgb : var_d);
            float x = buf[j] - m;
            if (x > lim) x = lim;
            if (x < -lim) x = -lim;
            x = x / lim;
            x = (x + 1) * 0.4f + 0.1f;
            q[(size_t)i * n4 + j] = (uint8_t)(f2i_x86(x * 255.0f) & 0xFF); /* C4 */
        }
    }
}

/* ------------------------------------------------------------------------------------------------ A4 */
static void dense_sigmoid(const float* X, int M, int Kd, const float* Wt, const float* b, int N, float* Y) {
    /* C5: Y[m][n] = sigmoid( dot8(X[m], W[n]) + b[n] ), dot8 = 8 interleaved partial sums combined pairwise */
    enum { MB = 16 };
#pragma omp parallel for schedule(dynamic, 4)
    for (int m0 = 0; m0 < M; m0 += MB) {
        int mb = M - m0 < MB ? M - m0 : MB;
        for (int n = 0; n < N; ++n) {
            const float* w = Wt + (size_t)n * Kd;
            for (int mi = 0; mi < mb; ++mi) {
                const float* x = X + (size_t)(m0 + mi) * Kd;
                float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                int k = 0;
                for (; k + 8 <= Kd; k += 8)
                    for (int l = 0; l < 8; ++l) acc[l] += w[k + l] * x[k + l];
                for (int l = 0; k < Kd; ++k, ++l) acc[l] += w[k] * x[k];
                float s = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
                s = s + b[n];
                Y[(size_t)(m0 + mi) * N + n] = 1.0f / (1.0f + expf(-s));
            }
        }
    }
}

void hf6d_ref_encode(const uint8_t* q, int32_t P, int32_t n0, const float* W1, const float* b1, int32_t n1,
                     const float* W2, const float* b2, int32_t n2, const float* W3, const float* b3, int32_t n3,
                     float* features) {
    float* x = (float*)malloc(sizeof(float) * (size_t)P * n0);
    float* h1 = (float*)malloc(sizeof(float) * (size_t)P * n1);
    float* h2 = (float*)malloc(sizeof(float) * (size_t)P * n2);
    for (size_t i = 0; i < (size_t)P * n0; ++i) x[i] = (float)q[i] / 255.0f; /* HFTest.cpp:565 */
    dense_sigmoid(x, P, n0, W1, b1, n1, h1);
    dense_sigmoid(h1, P, n1, W2, b2, n2, h2);
    dense_sigmoid(h2, P, n2, W3, b3, n3, features);
    free(x);
    free(h1);
    free(h2);
}

/* The same three layers on fp32 inputs (the net input k/255.0f the reference builds at HFTest.cpp:565).  The
 * reference-source pin library (oracle/ref_driver.cpp) installs this as its stand-in Caffe net's forward pass. */
void hf6d_ref_encode_f32(const float* x, int32_t P, int32_t n0, const float* W1, const float* b1, int32_t n1,
                         const float* W2, const float* b2, int32_t n2, const float* W3, const float* b3, int32_t n3,
                         float* features) {
    float* h1 = (float*)malloc(sizeof(float) * (size_t)P * n1);
    float* h2 = (float*)malloc(sizeof(float) * (size_t)P * n2);
    dense_sigmoid(x, P, n0, W1, b1, n1, h1);
    dense_sigmoid(h1, P, n1, W2, b2, n2, h2);
    dense_sigmoid(h2, P, n2, W3, b3, n3, features);
    free(h1);
    free(h2);
}

/* ------------------------------------------------------------------------------------------------ A6 */
static const ref_node* descend(const ref_node* n, const float* fv) {
    while (!n->leaf) {
        float val = 0.0f;
        if (n->mode == 0) val = fv[n->f1] - fv[n->f2];
        else if (n->mode == 1) val = fv[n->f1];
        n = (val < n->thr) ? n->left : n->right;
    }
    return n;
}

void hf6d_ref_traverse(const hf6d_ref_forest* f, const float* features, int32_t P, int32_t* leaf_id,
                       int32_t* leaf_ord) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; ++i)
        for (int t = 0; t < f->T; ++t) {
            const ref_node* l = descend(f->roots[t], features + (size_t)i * f->F);
            if (leaf_id) leaf_id[(size_t)i * f->T + t] = l->leaf_id;
            if (leaf_ord) leaf_ord[(size_t)i * f->T + t] = l->ordinal;
        }
}

/* ------------------------------------------------------------------------------------------------ A7 geometry */
static void rot_from_ypr(float yaw, float pitch, float roll, float R[9]) {
    /* corr * Rz*Ry*Rx, HFTest.cpp:45-80 (C6) */
    float cyw = (float)cos((double)yaw), syw = (float)sin((double)yaw);
    float cp = (float)cos((double)pitch), sp = (float)sin((double)pitch);
    float cr = (float)cos((double)roll), sr = (float)sin((double)roll);
    /* A = Rz*Ry */
    float a00 = cyw * cp, a01 = -syw, a02 = cyw * sp;
    float a10 = syw * cp, a11 = cyw, a12 = syw * sp;
    float a20 = -sp, a21 = 0.0f, a22 = cp;
    /* B = A*Rx */
    float nsr = -sr;
    R[0] = a00; R[1] = a01 * cr + a02 * sr; R[2] = a01 * nsr + a02 * cr;
    float b10 = a10, b11 = a11 * cr + a12 * sr, b12 = a11 * nsr + a12 * cr;
    float b20 = a20, b21 = a21 * cr + a22 * sr, b22 = a21 * nsr + a22 * cr;
    R[3] = -b10; R[4] = -b11; R[5] = -b12;
    R[6] = -b20; R[7] = -b21; R[8] = -b22;
}

static void centre3d(const float* vote, int px, int py, uint16_t depth_mm, const hf6d_ref_params* p, float out[3]) {
    float R[9];
    rot_from_ypr(vote[0], vote[1], vote[2], R);
    float z = (float)depth_mm / 1000.0f;
    float x = ((float)px - p->cx) * z / p->fx;
    float y = ((float)py - p->cy) * z / p->fy;
    float vx = -vote[3], vy = -vote[4], vz = -vote[5];
    out[0] = ((R[0] * vx + R[1] * vy) + R[2] * vz) + x; /* C7 */
    out[1] = ((R[3] * vx + R[4] * vy) + R[5] * vz) + y;
    out[2] = ((R[6] * vx + R[7] * vy) + R[8] * vz) + z;
}

static void project(const float c3[3], const hf6d_ref_params* p, int* u, int* v) {
    if (c3[2] == 0) { *u = 0; *v = 0; return; } /* HFTest.cpp:25-28 */
    *u = f2i_x86(c3[0] / c3[2] * p->fx + p->cx + 0.5f);
    *v = f2i_x86(c3[1] / c3[2] * p->fy + p->cy + 0.5f);
}

static inline uint32_t qweight(float prob) { return (uint32_t)(prob * 65536.0f + 0.5f); } /* C8 */

typedef struct { int32_t u, v; const ref_node* leaf; } vote_entry;
typedef struct { vote_entry* e; int64_t n, cap; } entry_list;

static void entry_push(entry_list* l, int u, int v, const ref_node* leaf) {
    if (l->n == l->cap) {
        l->cap = l->cap ? l->cap * 2 : (1 << 16);
        l->e = (vote_entry*)realloc(l->e, sizeof(vote_entry) * (size_t)l->cap);
    }
    l->e[l->n].u = u; l->e[l->n].v = v; l->e[l->n].leaf = leaf;
    l->n++;
}

/* Casts the votes of every processed patch; entries (the reference's center_leaf_map, HFTest.cpp:208-211) optional.
 * Threads take contiguous patch ranges with their own maps and entry lists, merged in thread order, as the reference votes
 * inside an OpenMP loop and merges per-thread maps (HFTest.cpp:601-656).  Integer sums: the result does not depend on the
 * number of threads, and the merged entry lists are in patch order whatever it is. */
static int64_t cast_votes_range(const hf6d_ref_forest* f, const int32_t* leaf_ord, const int32_t* locs, const uint16_t* depth,
                                int32_t i0, int32_t i1, const hf6d_ref_params* p, const uint8_t* should_detect, uint64_t* maps,
                                entry_list* entries /*[K] or NULL*/) {
    const int W = p->W, H = p->H, K = f->K, T = f->T;
    int64_t cast = 0;
    for (int i = i0; i < i1; ++i) {
        const int px = locs[2 * i], py = locs[2 * i + 1];
        const uint16_t d = depth[(size_t)py * W + px]; /* HFTest.cpp:628 */
        for (int t = 0; t < T; ++t) {
            const int32_t ord = leaf_ord[(size_t)i * T + t];
            if (ord < 0) continue; /* tree not owned by this shard (multi-GPU tests); never in the reference */
            const ref_node* leaf = f->leaves[t][ord];
            for (int c = 0; c < K; ++c) {
                if (should_detect && !should_detect[c]) continue;
                if (!(leaf->class_prob[c] >= 0.5f)) continue; /* HFTest.cpp:191 */
                const uint32_t w = qweight(leaf->class_prob[c]);
                for (int v = 0; v < leaf->nvotes[c]; ++v) {
                    float c3[3];
                    int uu, vv;
                    centre3d(leaf->votes[c] + 6 * v, px, py, d, p, c3);
                    project(c3, p, &uu, &vv);
                    if (maps && uu >= 0 && uu < W && vv >= 0 && vv < H) maps[((size_t)c * H + vv) * W + uu] += w;
                    if (entries) entry_push(&entries[c], uu, vv, leaf);
                    ++cast;
                }
            }
        }
    }
    return cast;
}

static int64_t cast_votes(const hf6d_ref_forest* f, const int32_t* leaf_ord, const int32_t* locs,
                          const uint16_t* depth, int32_t P, const hf6d_ref_params* p, const uint8_t* should_detect,
                          uint64_t* maps, entry_list* entries /*[K] or NULL*/) {
    const int K = f->K;
    const size_t map_n = (size_t)K * p->W * p->H;
    if (maps) memset(maps, 0, sizeof(uint64_t) * map_n);
    int nt = 1;
#ifdef _OPENMP
    nt = omp_get_max_threads();
#endif
    if (nt > P / 256) nt = P / 256 > 0 ? P / 256 : 1;
    if (nt <= 1) return cast_votes_range(f, leaf_ord, locs, depth, 0, P, p, should_detect, maps, entries);
    uint64_t** tmaps = (uint64_t**)calloc(nt, sizeof(uint64_t*));
    entry_list* tent = (entry_list*)calloc((size_t)nt * K, sizeof(entry_list));
    int64_t* tcast = (int64_t*)calloc(nt, sizeof(int64_t));
#pragma omp parallel num_threads(nt)
    {
        int th = 0;
#ifdef _OPENMP
        th = omp_get_thread_num();
#endif
        const int32_t i0 = (int32_t)((int64_t)P * th / nt), i1 = (int32_t)((int64_t)P * (th + 1) / nt);
        if (maps) tmaps[th] = (uint64_t*)calloc(map_n, sizeof(uint64_t));
        tcast[th] = cast_votes_range(f, leaf_ord, locs, depth, i0, i1, p, should_detect, maps ? tmaps[th] : NULL,
                                     entries ? tent + (size_t)th * K : NULL);
    }
    int64_t cast = 0;
    for (int th = 0; th < nt; ++th) cast += tcast[th];
    if (maps) {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < (int64_t)map_n; ++i) {
            uint64_t sum = 0;
            for (int th = 0; th < nt; ++th) sum += tmaps[th][i];
            maps[i] = sum;
        }
        for (int th = 0; th < nt; ++th) free(tmaps[th]);
    }
    if (entries)
        for (int c = 0; c < K; ++c) {
            int64_t total = entries[c].n;
            for (int th = 0; th < nt; ++th) total += tent[(size_t)th * K + c].n;
            if (total > entries[c].cap) {
                entries[c].cap = total;
                entries[c].e = (vote_entry*)realloc(entries[c].e, sizeof(vote_entry) * (size_t)(total > 0 ? total : 1));
            }
            for (int th = 0; th < nt; ++th) {  /* thread order = patch order */
                entry_list* l = &tent[(size_t)th * K + c];
                if (l->n) memcpy(entries[c].e + entries[c].n, l->e, sizeof(vote_entry) * (size_t)l->n);
                entries[c].n += l->n;
                free(l->e);
            }
        }
    free(tmaps);
    free(tent);
    free(tcast);
    return cast;
}

int64_t hf6d_ref_vote(const hf6d_ref_forest* f, const int32_t* leaf_ord, const int32_t* locs, const uint16_t* depth,
                      int32_t P, const hf6d_ref_params* p, const uint8_t* should_detect, uint64_t* maps) {
    return cast_votes(f, leaf_ord, locs, depth, P, p, should_detect, maps, NULL);
}

/* ------------------------------------------------------------------------------------------------ A9 blur */
static inline int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        else i = 2 * n - 2 - i;
    }
    return i;
}

void hf6d_ref_blur(const uint64_t* acc, int32_t rows, int32_t cols, int32_t kx, int32_t ky, float* out) {
    /* C9.  Anchor = kernel centre (k/2), BORDER_REFLECT_101 (cv::blur defaults).  Window sums as running sums (what
     * cv::blur's RowSum / ColumnSum do); the sums are integers, so they equal the plain k-term sums exactly. */
    const double scale = 1.0 / (double)(kx * ky);
    uint64_t* tmp = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)rows * cols);
#pragma omp parallel for schedule(static)
    for (int r = 0; r < rows; ++r) {
        const uint64_t* a = acc + (size_t)r * cols;
        uint64_t s = 0;
        for (int k = 0; k < kx; ++k) s += a[reflect101(-(kx / 2) + k, cols)];
        for (int c = 0; c < cols; ++c) {
            tmp[(size_t)r * cols + c] = s;
            s += a[reflect101(c + 1 - kx / 2 + kx - 1, cols)];
            s -= a[reflect101(c - kx / 2, cols)];
        }
    }
#pragma omp parallel for schedule(static)
    for (int c0 = 0; c0 < cols; c0 += 64) { /* a strip of columns per task: rows are walked contiguously */
        const int c1 = c0 + 64 < cols ? c0 + 64 : cols;
        uint64_t s[64];
        for (int c = c0; c < c1; ++c) {
            s[c - c0] = 0;
            for (int k = 0; k < ky; ++k) s[c - c0] += tmp[(size_t)reflect101(-(ky / 2) + k, rows) * cols + c];
        }
        for (int r = 0; r < rows; ++r) {
            const uint64_t* add = tmp + (size_t)reflect101(r + 1 - ky / 2 + ky - 1, rows) * cols;
            const uint64_t* sub = tmp + (size_t)reflect101(r - ky / 2, rows) * cols;
            for (int c = c0; c < c1; ++c) {
                out[(size_t)r * cols + c] = (float)(((double)s[c - c0] / 65536.0) * scale);
                s[c - c0] += add[c];
                s[c - c0] -= sub[c];
            }
        }
    }
    free(tmp);
}

/* ------------------------------------------------------------------------------------------------ A9 NMS */
typedef struct { float score; int32_t x, y, seq; } nms_hit;

static int nms_cmp(const void* a, const void* b) {
    const nms_hit* x = (const nms_hit*)a;
    const nms_hit* y = (const nms_hit*)b;
    if (x->score > y->score) return -1;
    if (x->score < y->score) return 1;
    return x->seq < y->seq ? -1 : (x->seq > y->seq); /* C10 */
}

/* Returns a malloc'ed, sorted list of hits (HFTest.cpp:219-268).  Two passes of a "first maximum" sliding window:
 * rows first (window wx, ties -> leftmost), then columns of the row results (window wy, ties -> topmost); a window
 * emits its maximum only if that maximum sits at the window centre and is non-zero.  The vertical pass only visits
 * rows 0 .. rows-wy (the reference's loop bound), so centres in the bottom wy-1 rows of windows are never produced. */
static nms_hit* nms_run(const float* in, int rows, int cols, int wx, int wy, int* count) {
    *count = 0;
    if (cols - wx + 1 <= 0 || rows - wy + 1 <= 0) return NULL;
    const int ncol = cols - wx + 1;
    /* pass 1: position of the first maximum of every horizontal window, by a monotonic queue like the reference's deque
     * (HFTest.cpp:232-243: an element is dropped from the back only by a strictly larger one, so the front is the
     * leftmost maximum) */
    int32_t* argx = (int32_t*)malloc(sizeof(int32_t) * (size_t)rows * ncol);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < rows; ++i) {
        int32_t* q = (int32_t*)malloc(sizeof(int32_t) * (size_t)cols);
        int qh = 0, qt = 0;
        const float* row = in + (size_t)i * cols;
        for (int j = 0; j < cols; ++j) {
            if (qt > qh && q[qh] == j - wx) ++qh;
            while (qt > qh && row[q[qt - 1]] < row[j]) --qt;
            q[qt++] = j;
            if (j >= wx - 1) argx[(size_t)i * ncol + j - wx + 1] = q[qh];
        }
        free(q);
    }
    int cap = 256, n = 0;
    nms_hit* hits = (nms_hit*)malloc(sizeof(nms_hit) * cap);
    const int last_i = rows - wy; /* inclusive: the reference's vertical pass stops here (HFTest.cpp:247) */
    int32_t* q = (int32_t*)malloc(sizeof(int32_t) * (size_t)rows);
    for (int j = 0; j < ncol; ++j) {
        int qh = 0, qt = 0; /* rows of the column's window, values in[r][argx[r][j]]: the front is the topmost maximum */
        for (int i = 0; i <= last_i; ++i) {
            const float v = in[(size_t)i * cols + argx[(size_t)i * ncol + j]];
            if (qt > qh && q[qh] == i - wy) ++qh;
            while (qt > qh && in[(size_t)q[qt - 1] * cols + argx[(size_t)q[qt - 1] * ncol + j]] < v) --qt;
            q[qt++] = i;
            if (i < wy - 1) continue;
            const int top = i - wy + 1, brow = q[qh];
            const float bv = in[(size_t)brow * cols + argx[(size_t)brow * ncol + j]];
            const int bcol = argx[(size_t)brow * ncol + j];
            const int ccx = j + wx / 2, ccy = top + wy / 2;
            if (bv != 0 && brow == ccy && bcol == ccx) {
                if (n == cap) { cap *= 2; hits = (nms_hit*)realloc(hits, sizeof(nms_hit) * cap); }
                hits[n].score = bv; hits[n].x = bcol; hits[n].y = brow; hits[n].seq = n;
                ++n;
            }
        }
    }
    free(q);
    free(argx);
    qsort(hits, n, sizeof(nms_hit), nms_cmp);
    *count = n;
    return hits;
}

int32_t hf6d_ref_nms(const float* in, int32_t rows, int32_t cols, int32_t wx, int32_t wy, float* score, int32_t* xs,
                     int32_t* ys, int32_t cap) {
    int n = 0;
    nms_hit* h = nms_run(in, rows, cols, wx, wy, &n);
    for (int i = 0; i < n && i < cap; ++i) {
        score[i] = h[i].score; xs[i] = h[i].x; ys[i] = h[i].y;
    }
    free(h);
    return n;
}

/* ------------------------------------------------------------------------------------------------ A10-A12 */
static void pose_from_tuple(const hf6d_ref_params* p, int cx, int cy, float z, int yaw_deg, int pitch_deg,
                            int roll_deg, float pose[16]) {
    /* HFTest.cpp:922-924 (float deg / 180.0f * M_PI, narrowed to float) then MeshUtils.cpp:423-440 */
    float yaw = (float)((float)yaw_deg / 180.0f * M_PI);
    float pitch = (float)((float)pitch_deg / 180.0f * M_PI);
    float roll = (float)((float)roll_deg / 180.0f * M_PI);
    float R[9];
    rot_from_ypr(yaw, pitch, roll, R);
    float x = ((float)cx - p->cx) * z / p->fx;
    float y = ((float)cy - p->cy) * z / p->fy;
    pose[0] = R[0]; pose[1] = R[1]; pose[2] = R[2]; pose[3] = x;
    pose[4] = R[3]; pose[5] = R[4]; pose[6] = R[5]; pose[7] = y;
    pose[8] = R[6]; pose[9] = R[7]; pose[10] = R[8]; pose[11] = z;
    pose[12] = 0; pose[13] = 0; pose[14] = 0; pose[15] = 1;
}

typedef struct { int32_t Y, Pp; const ref_node* leaf; } roll_entry;

/* The reference's center_leaf_map is a hash map from a pixel to the leaves that voted for it (HFTest.cpp:208-211), read
 * back pixel by pixel over a centre's window (:757-762).  Here: the entries of one class sorted by pixel (stable counting
 * sort, so a pixel keeps its insertion order) over the image plus a margin of half a window -- votes may land outside the
 * image and still inside a window. */
typedef struct {
    int32_t x0, y0, gw, gh; /* pixel of cell (0,0), grid size */
    int64_t* start;         /* [gw*gh + 1] */
    vote_entry* sorted;
} entry_index;

static void entry_index_build(entry_index* ix, const entry_list* l, int W, int H, int half) {
    ix->x0 = -half; ix->y0 = -half; ix->gw = W + 2 * half; ix->gh = H + 2 * half;
    const int64_t cells = (int64_t)ix->gw * ix->gh;
    ix->start = (int64_t*)calloc((size_t)cells + 1, sizeof(int64_t));
    ix->sorted = (vote_entry*)malloc(sizeof(vote_entry) * (size_t)(l->n > 0 ? l->n : 1));
    for (int64_t e = 0; e < l->n; ++e) {
        const int64_t cx = (int64_t)l->e[e].u - ix->x0, cy = (int64_t)l->e[e].v - ix->y0;
        if (cx >= 0 && cx < ix->gw && cy >= 0 && cy < ix->gh) ix->start[cy * ix->gw + cx + 1]++;
    }
    for (int64_t i = 0; i < cells; ++i) ix->start[i + 1] += ix->start[i];
    int64_t* fill = (int64_t*)malloc(sizeof(int64_t) * (size_t)cells);
    memcpy(fill, ix->start, sizeof(int64_t) * (size_t)cells);
    for (int64_t e = 0; e < l->n; ++e) {
        const int64_t cx = (int64_t)l->e[e].u - ix->x0, cy = (int64_t)l->e[e].v - ix->y0;
        if (cx >= 0 && cx < ix->gw && cy >= 0 && cy < ix->gh) ix->sorted[fill[cy * ix->gw + cx]++] = l->e[e];
    }
    free(fill);
}
static void entry_index_free(entry_index* ix) { free(ix->start); free(ix->sorted); }

static int hypotheses_for_centre(const hf6d_ref_params* p, int c, const uint16_t* depth,
                                 const entry_index* ix, int ctr_x, int ctr_y, float loc_score,
                                 hf6d_ref_hypothesis* out, int cap) {
    const int W = p->W, H = p->H;
    const int half = p->centers_nms_wsize / 2;
    const float z_bin_size = 0.01f; /* HFTest.cpp:743-745 */
    const int zbins = HF6D_REF_Z_BINS, z_nms = 20, NB = HF6D_REF_POSE_BINS;
    int produced = 0;

    uint64_t zacc[HF6D_REF_Z_BINS];
    memset(zacc, 0, sizeof zacc);
    uint64_t* ypacc = (uint64_t*)calloc((size_t)NB * NB, sizeof(uint64_t));
    roll_entry* rl = NULL;
    int64_t rn = 0, rcap = 0;

    for (int row = ctr_y - half; row < ctr_y + half; ++row)
      for (int col = ctr_x - half; col < ctr_x + half; ++col) { /* HFTest.cpp:757-762 */
        const int64_t gx = (int64_t)col - ix->x0, gy = (int64_t)row - ix->y0;
        if (gx < 0 || gx >= ix->gw || gy < 0 || gy >= ix->gh) continue;
      for (int64_t e = ix->start[gy * ix->gw + gx]; e < ix->start[gy * ix->gw + gx + 1]; ++e) {
        const ref_node* leaf = ix->sorted[e].leaf;
        const uint32_t w = qweight(leaf->class_prob[c]);
        const int inside = row >= 0 && row < H && col >= 0 && col < W; /* reference reads out of bounds: skip */
        const uint16_t dpix = inside ? depth[(size_t)row * W + col] : 0;
        for (int v = 0; v < leaf->nvotes[c]; ++v) {
            const float* vote = leaf->votes[c] + 6 * v;
            if (dpix != 0) { /* HFTest.cpp:770-776: the WINDOW pixel stands in for the patch centre */
                float c3[3];
                centre3d(vote, col, row, dpix, p, c3);
                int zb = f2i_x86(c3[2] / z_bin_size);
                if (zb < zbins && zb >= 0) zacc[zb] += w;
            }
            /* HFTest.cpp:779-780: int = float / M_PI(double) * 180.0f -> double expression, truncated */
            const int yaw = d2i_x86((double)vote[0] / M_PI * 180.0);
            const int pitch = d2i_x86((double)vote[1] / M_PI * 180.0);
            for (int k1 = 0; k1 < 2; ++k1)
                for (int k2 = 0; k2 < 2; ++k2) {
                    int sy = yaw < 0 ? -1 : 1, sp = pitch < 0 ? -1 : 1; /* copysign(1,(float)int): sign(0)=+1 */
                    int cy_ = yaw + (-1) * sy * k1 * 360 + 360;
                    int cp_ = pitch + (-1) * sp * k2 * 360 + 360;
                    if (cy_ >= 0 && cy_ < NB && cp_ >= 0 && cp_ < NB) /* reference would write out of bounds */
                        ypacc[(size_t)cy_ * NB + cp_] += w;
                    if (k1 == 0 && k2 == 0) {
                        if (rn == rcap) {
                            rcap = rcap ? rcap * 2 : 4096;
                            rl = (roll_entry*)realloc(rl, sizeof(roll_entry) * (size_t)rcap);
                        }
                        rl[rn].Y = cy_; rl[rn].Pp = cp_; rl[rn].leaf = leaf;
                        ++rn;
                    }
                }
        }
      }
      }

    /* mode of z: NMS (1 wide, 20 tall) on the 300x1 histogram, HFTest.cpp:803-812 */
    float zf[HF6D_REF_Z_BINS];
    for (int i = 0; i < zbins; ++i) zf[i] = (float)((double)zacc[i] / 65536.0);
    int nz = 0;
    nms_hit* zh = nms_run(zf, zbins, 1, 1, z_nms, &nz);
    if (nz == 0) { free(zh); free(ypacc); free(rl); return 0; }
    const float mode_z = (float)zh[0].y * z_bin_size;
    free(zh);

    /* yaw/pitch: blur 35x35 + NMS 35x35, keep [180,540]^2, HFTest.cpp:817-836 */
    float* ypf = (float*)malloc(sizeof(float) * (size_t)NB * NB);
    hf6d_ref_blur(ypacc, NB, NB, p->pose_blur_size, p->pose_blur_size, ypf);
    int nyp = 0;
    nms_hit* yph = nms_run(ypf, NB, NB, p->pose_nms_wsize, p->pose_nms_wsize, &nyp);
    int kept = 0;
    for (int i = 0; i < nyp; ++i)
        if (!(yph[i].x < 180 || yph[i].x > 360 + 180 || yph[i].y < 180 || yph[i].y > 360 + 180)) yph[kept++] = yph[i];
    nyp = kept;
    free(ypf);
    free(ypacc);

    const int max_yp = nyp < p->max_yaw_pitch_hypotheses ? nyp : p->max_yaw_pitch_hypotheses;
    const int bh = p->pose_blur_size / 2;
    for (int h2 = 0; h2 < max_yp; ++h2) {
        const float yp_score = yph[h2].score / yph[0].score;
        if (yp_score < p->min_yaw_pitch_drop_ratio) break;
        const int Yp = yph[h2].y, Pp = yph[h2].x; /* row = yaw, col = pitch, HFTest.cpp:855-856, 922-923 */
        uint64_t racc[HF6D_REF_POSE_BINS];
        memset(racc, 0, sizeof racc);
        for (int64_t e = 0; e < rn; ++e) {
            if (rl[e].Y < Yp - bh || rl[e].Y >= Yp + bh || rl[e].Pp < Pp - bh || rl[e].Pp >= Pp + bh) continue;
            const ref_node* leaf = rl[e].leaf;
            const uint32_t w = qweight(leaf->class_prob[c]);
            for (int v = 0; v < leaf->nvotes[c]; ++v) {
                /* :863  float * 180.0f (float) / M_PI (double) */
                const int r = d2i_x86((double)(leaf->votes[c][6 * v + 2] * 180.0f) / M_PI);
                int b0 = r + 360, b1 = r < 0 ? r + 720 : r;
                if (b0 >= 0 && b0 < NB) racc[b0] += w;
                if (b1 >= 0 && b1 < NB) racc[b1] += w;
            }
        }
        float rf[HF6D_REF_POSE_BINS];
        hf6d_ref_blur(racc, NB, 1, 1, p->pose_blur_size, rf);
        int nr = 0;
        nms_hit* rh = nms_run(rf, NB, 1, 1, p->pose_nms_wsize, &nr);
        kept = 0;
        for (int i = 0; i < nr; ++i)
            if (!(rh[i].y < 180 || rh[i].y > 360 + 180)) rh[kept++] = rh[i];
        nr = kept;
        int h_roll = 0;
        float prev = 3.402823466e+38f;
        for (int i = 0; i < nr && h_roll < p->max_roll_hypotheses; ++i) {
            const float roll_score = rh[i].score / rh[0].score;
            const int ry = rh[i].y;
            /* HFTest.cpp:918-921: double expression narrowed to a float `dot`, acos() of that float */
            float dot = (float)(cos(prev / 180.0f * M_PI) * cos(ry / 180.0f * M_PI) +
                                sin(prev / 180.0f * M_PI) * sin(ry / 180.0f * M_PI));
            if (h_roll == 0 || acos(dot) / M_PI * 180.0f > 7) {
                if (produced < cap) {
                    hf6d_ref_hypothesis* o = &out[produced];
                    o->cls = c; o->cx = ctr_x; o->cy = ctr_y; o->z = mode_z;
                    o->yaw_deg = Yp - 360; o->pitch_deg = Pp - 360; o->roll_deg = ry - 360;
                    o->loc_score = loc_score; o->yawpitch_score = yp_score; o->roll_score = roll_score;
                    pose_from_tuple(p, ctr_x, ctr_y, mode_z, o->yaw_deg, o->pitch_deg, o->roll_deg, o->pose);
                }
                ++produced;
                prev = (float)ry;
                ++h_roll;
            }
        }
        free(rh);
    }
    free(yph);
    free(rl);
    return produced;
}

int32_t hf6d_ref_hypotheses(const hf6d_ref_forest* f, const int32_t* leaf_ord, const int32_t* locs,
                            const uint16_t* depth, int32_t P, const hf6d_ref_params* p, const uint8_t* should_detect,
                            const int32_t* max_location_hypotheses, const uint64_t* maps_in,
                            hf6d_ref_hypothesis* hyps, int32_t cap) {
    const int W = p->W, H = p->H, K = f->K;
    entry_list* entries = (entry_list*)calloc(K, sizeof(entry_list));
    uint64_t* maps = NULL;
    if (!maps_in) maps = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)K * W * H);
    cast_votes(f, leaf_ord, locs, depth, P, p, should_detect, maps, entries);
    const uint64_t* M = maps_in ? maps_in : maps;
    float* blurred = (float*)malloc(sizeof(float) * (size_t)W * H);
    int total = 0;
    for (int c = 0; c < K; ++c) {
        if (should_detect && !should_detect[c]) continue;
        hf6d_ref_blur(M + (size_t)c * W * H, H, W, p->centers_blur_size, p->centers_blur_size, blurred);
        entry_index ix;
        entry_index_build(&ix, &entries[c], W, H, p->centers_nms_wsize / 2 + 1);
        int nc = 0;
        nms_hit* ch = nms_run(blurred, H, W, p->centers_nms_wsize, p->centers_nms_wsize, &nc);
        int max_loc = max_location_hypotheses ? max_location_hypotheses[c] : 12;
        if (nc < max_loc) max_loc = nc;
        /* per-centre work is independent (omp for in the reference, HFTest.cpp:718-723); keep rank order */
        int* counts = (int*)calloc(max_loc > 0 ? max_loc : 1, sizeof(int));
        hf6d_ref_hypothesis* tmp =
            (hf6d_ref_hypothesis*)malloc(sizeof(hf6d_ref_hypothesis) * (size_t)(max_loc > 0 ? max_loc : 1) * 32);
#pragma omp parallel for schedule(dynamic)
        for (int k = 0; k < max_loc; ++k) {
            if (ch[k].score / ch[0].score < p->min_location_score_ratio) continue; /* HFTest.cpp:726 */
            counts[k] = hypotheses_for_centre(p, c, depth, &ix, ch[k].x, ch[k].y, ch[k].score,
                                              tmp + (size_t)k * 32, 32);
            if (counts[k] > 32) counts[k] = 32;
        }
        for (int k = 0; k < max_loc; ++k)
            for (int i = 0; i < counts[k]; ++i) {
                if (total < cap) hyps[total] = tmp[(size_t)k * 32 + i];
                ++total;
            }
        free(tmp);
        free(counts);
        free(ch);
        entry_index_free(&ix);
    }
    free(blurred);
    free(maps);
    for (int c = 0; c < K; ++c) free(entries[c].e);
    free(entries);
    return total;
}

/* ------------------------------------------------------------------------------------------------ whole frame */
static double now_s(void) {
#ifdef _OPENMP
    return omp_get_wtime();
#else
    return 0.0;
#endif
}

int32_t hf6d_ref_detect(const hf6d_ref_forest* f, const uint8_t* bgr, const uint16_t* depth, const hf6d_ref_params* p,
                        const float* const* weights, const int32_t* dims, const uint8_t* should_detect,
                        const int32_t* max_location_hypotheses, const float* features_override,
                        hf6d_ref_hypothesis* hyps, int32_t cap, int32_t* n_patches_out, double* st) {
    const int ps = p->patch_vox;
    const int maxP = ((p->W + p->stride - 1) / p->stride) * ((p->H + p->stride - 1) / p->stride);
    int32_t* locs = (int32_t*)malloc(sizeof(int32_t) * 2 * (size_t)maxP);
    double t0 = now_s();
    int P = hf6d_ref_scan_centres(depth, p, locs, maxP);
    const int Pp = (P / p->batch_size) * p->batch_size; /* tail dropped, HFTest.cpp:433 */
    if (n_patches_out) { n_patches_out[0] = P; n_patches_out[1] = Pp; }
    float* feat = NULL;
    double t1 = t0, t2 = t0, t3 = t0;
    if (!features_override) {
        const int nch = p->patch_mode ? 6 : 4;
        float* patches = (float*)malloc(sizeof(float) * (size_t)(Pp > 0 ? Pp : 1) * ps * ps * nch);
        uint8_t* q = (uint8_t*)malloc((size_t)(Pp > 0 ? Pp : 1) * ps * ps * nch);
        if (p->patch_mode) {
            float* normals = (float*)malloc(sizeof(float) * (size_t)p->W * p->H * 3);
            hf6d_ref_normals(depth, p->W, p->H, p->normals_focal, normals);
            hf6d_ref_gather_normals(bgr, depth, normals, p, locs, Pp, patches);
            t1 = now_s();
            hf6d_ref_quantise_normals(patches, Pp, ps, q);
            free(normals);
        } else {
            hf6d_ref_gather(bgr, depth, p, locs, Pp, patches);
            t1 = now_s();
            hf6d_ref_normalise(patches, Pp, ps, q);
        }
        t2 = now_s();
        feat = (float*)malloc(sizeof(float) * (size_t)(Pp > 0 ? Pp : 1) * dims[3]);
        hf6d_ref_encode(q, Pp, dims[0], weights[0], weights[1], dims[1], weights[2], weights[3], dims[2], weights[4],
                        weights[5], dims[3], feat);
        t3 = now_s();
        free(patches);
        free(q);
    }
    const float* F = features_override ? features_override : feat;
    int32_t* ord = (int32_t*)malloc(sizeof(int32_t) * (size_t)(Pp > 0 ? Pp : 1) * f->T);
    hf6d_ref_traverse(f, F, Pp, NULL, ord);
    double t4 = now_s();
    int n = hf6d_ref_hypotheses(f, ord, locs, depth, Pp, p, should_detect, max_location_hypotheses, NULL, hyps, cap);
    double t5 = now_s();
    if (st) {
        st[0] = t1 - t0; st[1] = t2 - t1; st[2] = t3 - t2; st[3] = t4 - t3; st[4] = t5 - t4; st[5] = t5 - t0;
    }
    free(ord);
    free(feat);
    free(locs);
    return n;
}
