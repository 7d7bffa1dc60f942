// Stage REFINE: what the reference does with every hypothesis tuple after the Hough stage:
//   MeshUtils::setScene             HoughForest/src/MeshUtils.cpp:340-420  scene cloud, VoxelGrid, normals, smooth clusters
//   MeshUtils::insertObjectFromPLY  HoughForest/include/MeshUtils.h:213-247 object cloud, VoxelGrid, normals
//   MeshUtils::icp                  MeshUtils.cpp:423-464                  pcl::IterativeClosestPoint from the Hough pose
//   MeshUtils::evaluate_hypothesis  MeshUtils.cpp:629-793                  similarity / inliers / clutter / final score
//   MeshUtils::optimize_hypotheses  MeshUtils.cpp:864-1168                 mutual exclusion, groups, best solution per group
// The arithmetic the reference leaves to PCL 1.7 (not vendored) is restated from PCL's published algorithms; the choices are
// listed as R1..R9 in oracle/refine.py, which is the checker of this file.
//
// B200 design
//  * one spatial structure serves every neighbour query (VoxelGrid, normal estimation, clustering, ICP correspondences, radius
//    search of the scoring): a BITMAP of the occupied 5 mm voxels in PCL's own voxel order (x fastest) plus an exclusive prefix
//    of its popcounts.  rank(voxel) = prefix[word] + popc(lower bits) IS the index of the voxel's point in PCL's output order,
//    so downsampling needs no sort and no hash, a radius query is a walk over a few bitmap rows (two words per row), and the
//    neighbours come out in ascending index order -- deterministic, whatever the thread schedule.  3 MB per 640x480 frame.
//  * VoxelGrid centroids are accumulated as 64-bit fixed point (2^-40 m): exact sums, no float atomics, no order dependence.
//  * clusters are connected components (lock-free union-find with atomicCAS hooking, smaller root wins, so a cluster is named
//    by its lowest point exactly as the reference's seed walk names it).
//  * ICP: one CTA per hypothesis, every iteration inside the kernel -- correspondences by the bitmap walk, the 3x3
//    cross-covariance reduced in double, the rotation from Horn's quaternion form (Jacobi on a 4x4 in one thread), the
//    convergence test of pcl::DefaultConvergenceCriteria -- no host round trip between iterations.
//  * scoring: one CTA per hypothesis; the scene points a hypothesis explains are a bitset in global memory (atomicOr), which
//    is also what the joint optimisation intersects.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "../../include/hf6d.h"

namespace hf6d {

struct RfGrid {
    float inv;        // 1 / leaf, as PCL computes it (float)
    int bx, by, bz;   // voxel coordinates of the grid's corner
    int dx, dy, dz;   // extent in voxels; dx is a multiple of 32 (a row of the bitmap is whole words)
    long long words;  // dx / 32 * dy * dz
};

struct RfCam {
    float fx, fy, cx, cy;
    int W, H;
    float dist_thr_mm;  // distance_threshold * 1000 (MeshUtils.cpp:352)
};

constexpr int RF_THREADS = 256;
constexpr double RF_FIX = 1099511627776.0;  // 2^40: fixed-point scale of the centroid sums

__device__ __forceinline__ bool rf_voxel(const RfGrid& g, float x, float y, float z, long long& v) {
    const int i = (int)floorf(x * g.inv) - g.bx, j = (int)floorf(y * g.inv) - g.by, k = (int)floorf(z * g.inv) - g.bz;
    if (i < 0 || i >= g.dx || j < 0 || j >= g.dy || k < 0 || k >= g.dz) return false;
    v = ((long long)k * g.dy + j) * g.dx + i;
    return true;
}

__device__ __forceinline__ unsigned rf_pack_rgb(unsigned r, unsigned g, unsigned b) { return r | (g << 8) | (b << 16); }

// ---------------------------------------------------------------------------------------------- scene cloud from the frame
// MeshUtils.cpp:346-364: an organised cloud with one point per pixel; pixels without a usable depth keep the value-initialised
// point (0, 0, 0, black) -- oracle choice R4.  pts[i] = (x, y, z, rgb bits).
__global__ void rf_scene_points_kernel(const uint8_t* __restrict__ bgr, const uint16_t* __restrict__ depth, RfCam cam,
                                       float4* __restrict__ pts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cam.W * cam.H) return;
    const int row = i / cam.W, col = i - row * cam.W;
    const unsigned d = depth[i];
    float4 p = make_float4(0.f, 0.f, 0.f, __uint_as_float(0u));
    if (d != 0 && (float)d < cam.dist_thr_mm) {
        const float z = __fdiv_rn((float)d, 1000.0f);
        p.x = __fdiv_rn(__fmul_rn(__fsub_rn((float)col, cam.cx), z), cam.fx);
        p.y = __fdiv_rn(__fmul_rn(__fsub_rn((float)row, cam.cy), z), cam.fy);
        p.z = z;
        const uint8_t* c = bgr + (size_t)i * 3;
        p.w = __uint_as_float(rf_pack_rgb(c[2], c[1], c[0]));
    }
    pts[i] = p;
}

// ---------------------------------------------------------------------------------------------- VoxelGrid
__global__ void rf_mark_kernel(const float4* __restrict__ in, const int* __restrict__ n_in, int n_in_host, RfGrid g,
                               unsigned* __restrict__ bits) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = n_in ? *n_in : n_in_host;
    if (i >= n) return;
    const float4 p = in[i];
    long long v;
    if (rf_voxel(g, p.x, p.y, p.z, v)) atomicOr(bits + (v >> 5), 1u << (v & 31));
}

// Exclusive prefix of the words' popcounts in three launches: per-tile sums (a tile = 4096 words, one CTA), an exclusive scan of
// the tile sums by one CTA (a 640x480 frame has ~200 tiles), and the tile-local scan plus its offset.  total[0] = set bits.
constexpr int RF_SCAN_TILE = 4096;

__global__ void __launch_bounds__(1024) rf_tile_sums_kernel(const unsigned* __restrict__ bits, long long words,
                                                            unsigned* __restrict__ tile_sum) {
    __shared__ unsigned s_sum;
    if (threadIdx.x == 0) s_sum = 0;
    __syncthreads();
    const long long w = (long long)blockIdx.x * RF_SCAN_TILE + (long long)threadIdx.x * 4;
    unsigned c = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) c += w + q < words ? (unsigned)__popc(bits[w + q]) : 0u;
#pragma unroll
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_sum, c);
    __syncthreads();
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = s_sum;
}

// one CTA: tile_sum -> exclusive prefix in place, total[0] = grand total
__global__ void __launch_bounds__(1024) rf_tile_scan_kernel(unsigned* __restrict__ tile_sum, int n_tiles, int* __restrict__ total) {
    __shared__ unsigned s_warp[32];
    __shared__ unsigned s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n_tiles; base += 1024) {
        const int i = base + threadIdx.x;
        const unsigned mine = i < n_tiles ? tile_sum[i] : 0u;
        unsigned incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            unsigned v = s_warp[lane], iv = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, iv, o);
                if (lane >= o) iv += t;
            }
            s_warp[lane] = iv - v;
        }
        __syncthreads();
        const unsigned excl = s_carry + s_warp[warp] + incl - mine;
        if (i < n_tiles) tile_sum[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + mine;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = (int)s_carry;
}

__global__ void __launch_bounds__(1024) rf_prefix_kernel(const unsigned* __restrict__ bits, long long words,
                                                         const unsigned* __restrict__ tile_off, unsigned* __restrict__ prefix) {
    __shared__ unsigned s_warp[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long w = (long long)blockIdx.x * RF_SCAN_TILE + (long long)threadIdx.x * 4;
    unsigned c[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) c[q] = w + q < words ? (unsigned)__popc(bits[w + q]) : 0u;
    const unsigned mine = c[0] + c[1] + c[2] + c[3];
    unsigned incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        unsigned v = s_warp[lane], iv = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, iv, o);
            if (lane >= o) iv += t;
        }
        s_warp[lane] = iv - v;  // exclusive over warps
    }
    __syncthreads();
    unsigned run = tile_off[blockIdx.x] + s_warp[warp] + incl - mine;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (w + q < words) prefix[w + q] = run;
        run += c[q];
    }
}

__device__ __forceinline__ int rf_rank(const unsigned* __restrict__ bits, const unsigned* __restrict__ prefix, long long v) {
    const unsigned w = bits[v >> 5];
    return (int)(prefix[v >> 5] + (unsigned)__popc(w & ((1u << (v & 31)) - 1u)));
}

struct RfAccum {
    long long* sum;  // [cap][3] fixed-point coordinate sums
    int* cnt;        // [cap]
    int* rgb;        // [cap][3]
};

__global__ void rf_accumulate_kernel(const float4* __restrict__ in, const int* __restrict__ n_in, int n_in_host, RfGrid g,
                                     const unsigned* __restrict__ bits, const unsigned* __restrict__ prefix, RfAccum a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = n_in ? *n_in : n_in_host;
    if (i >= n) return;
    const float4 p = in[i];
    long long v;
    if (!rf_voxel(g, p.x, p.y, p.z, v)) return;
    const int r = rf_rank(bits, prefix, v);
    atomicAdd(reinterpret_cast<unsigned long long*>(a.sum + 3 * (size_t)r + 0), (unsigned long long)__double2ll_rn((double)p.x * RF_FIX));
    atomicAdd(reinterpret_cast<unsigned long long*>(a.sum + 3 * (size_t)r + 1), (unsigned long long)__double2ll_rn((double)p.y * RF_FIX));
    atomicAdd(reinterpret_cast<unsigned long long*>(a.sum + 3 * (size_t)r + 2), (unsigned long long)__double2ll_rn((double)p.z * RF_FIX));
    atomicAdd(a.cnt + r, 1);
    const unsigned c = __float_as_uint(p.w);
    atomicAdd(a.rgb + 3 * (size_t)r + 0, (int)(c & 255u));
    atomicAdd(a.rgb + 3 * (size_t)r + 1, (int)((c >> 8) & 255u));
    atomicAdd(a.rgb + 3 * (size_t)r + 2, (int)((c >> 16) & 255u));
}

// centroid of every occupied voxel (oracle choice R2), colours = truncated channel means
__global__ void rf_centroid_kernel(RfAccum a, const int* __restrict__ n_pts, float4* __restrict__ pts) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= *n_pts) return;
    const double n = (double)a.cnt[r];
    float4 p;
    p.x = (float)((double)a.sum[3 * (size_t)r + 0] / (n * RF_FIX));
    p.y = (float)((double)a.sum[3 * (size_t)r + 1] / (n * RF_FIX));
    p.z = (float)((double)a.sum[3 * (size_t)r + 2] / (n * RF_FIX));
    const int c = a.cnt[r];
    p.w = __uint_as_float(rf_pack_rgb((unsigned)(a.rgb[3 * (size_t)r + 0] / c), (unsigned)(a.rgb[3 * (size_t)r + 1] / c),
                                      (unsigned)(a.rgb[3 * (size_t)r + 2] / c)));
    pts[r] = p;
}

// ---------------------------------------------------------------------------------------------- neighbour walk
// f(index) for every point whose VOXEL lies in the box of voxels that the ball (q, r) touches, in ascending index order; the
// caller tests the distance.  A point is the centroid of its voxel's samples, so it lies inside its voxel.
template <bool PREFETCH = false, class F>
__device__ __forceinline__ void rf_for_box(const RfGrid& g, const unsigned* __restrict__ bits, const unsigned* __restrict__ prefix,
                                           float qx, float qy, float qz, float r, F&& f) {
    const int i0 = max(0, (int)floorf((qx - r) * g.inv) - g.bx), i1 = min(g.dx - 1, (int)floorf((qx + r) * g.inv) - g.bx);
    const int j0 = max(0, (int)floorf((qy - r) * g.inv) - g.by), j1 = min(g.dy - 1, (int)floorf((qy + r) * g.inv) - g.by);
    const int k0 = max(0, (int)floorf((qz - r) * g.inv) - g.bz), k1 = min(g.dz - 1, (int)floorf((qz + r) * g.inv) - g.bz);
    if (i0 > i1 || j0 > j1 || k0 > k1) return;
    const int w0 = i0 >> 5, w1 = i1 >> 5, wpr = g.dx >> 5;
    const unsigned m0 = 0xffffffffu << (i0 & 31), m1 = 0xffffffffu >> (31 - (i1 & 31));
    auto visit = [&](long long at, unsigned full, unsigned m) {
        if (!m) return;
        const unsigned base = __ldg(prefix + at);
        while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            f((int)(base + (unsigned)__popc(full & ((1u << b) - 1u))));
        }
    };
    if (PREFETCH && w1 - w0 <= 1) {
        // ICP's form, for the usual case (boxes narrower than 32 voxels): a row of the box is one or two words.  The words of four rows are
        // fetched before any is looked at -- the walk is a chain of dependent L2 loads otherwise, and its latency, not its
        // instruction count, is what an ICP iteration waits for (ncu: half of all stall samples at the iteration's barriers;
        // 21 -> 14 ms per frame).  The one-pass kernels (normals, clusters, scoring) have the parallelism to hide it and lose
        // more to the extra registers than they gain.
        const bool two = w1 > w0;
        for (int k = k0; k <= k1; ++k)
            for (int jb = j0; jb <= j1; jb += 4) {
                unsigned a[4], c[4];
                long long row[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const bool in = jb + q <= j1;
                    row[q] = ((long long)k * g.dy + (in ? jb + q : jb)) * wpr;
                    a[q] = in ? __ldg(bits + row[q] + w0) : 0u;
                    c[q] = in && two ? __ldg(bits + row[q] + w1) : 0u;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    visit(row[q] + w0, a[q], two ? a[q] & m0 : a[q] & m0 & m1);
                    visit(row[q] + w1, c[q], c[q] & m1);
                }
            }
        return;
    }
    for (int k = k0; k <= k1; ++k)
        for (int j = j0; j <= j1; ++j) {
            const long long row = ((long long)k * g.dy + j) * wpr;
            for (int w = w0; w <= w1; ++w) {
                const unsigned full = __ldg(bits + row + w);
                unsigned m = full;
                if (w == w0) m &= m0;
                if (w == w1) m &= m1;
                visit(row + w, full, m);
            }
        }
}

__device__ __forceinline__ float rf_dist2(float ax, float ay, float az, float bx, float by, float bz) {
    const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// ---------------------------------------------------------------------------------------------- reach map
// The voxel bitmap dilated by R voxels along every axis (a cube of 2R + 1 voxels, three separable passes of word operations):
// bit v is set when some occupied voxel lies within R voxels of v in every coordinate.  A query point whose voxel has a clear
// bit has no scene point within R voxel sizes (every point lies inside its voxel), so the far side of a solid object and most
// of a misplaced hypothesis are answered "no neighbour" by one bit test instead of a walk over the whole search box.
struct RfReach {
    const unsigned* dil;  // dilated bitmap, the layout of the voxel bitmap; nullptr: no shortcut
    float reach;          // R * leaf: radii up to this may use the shortcut
};

// along x: a row is dx / 32 consecutive words
__global__ void rf_dilate_x_kernel(const unsigned* __restrict__ in, RfGrid g, int R, unsigned* __restrict__ out) {
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= g.words) return;
    const int wpr = g.dx >> 5;
    const int wi = (int)(w % wpr);
    const unsigned c = in[w], l = wi > 0 ? in[w - 1] : 0u, r = wi + 1 < wpr ? in[w + 1] : 0u;
    unsigned v = c;
    for (int s = 1; s <= R; ++s) v |= (c << s) | (l >> (32 - s)) | (c >> s) | (r << (32 - s));
    out[w] = v;
}
// along y (stride = words per row, extent dy) or z (stride = words per slice, extent dz)
__global__ void rf_dilate_axis_kernel(const unsigned* __restrict__ in, long long words, long long stride, int extent, int R,
                                      unsigned* __restrict__ out) {
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= words) return;
    const int pos = (int)((w / stride) % extent);
    unsigned v = 0;
    for (int d = -R; d <= R; ++d)
        if (pos + d >= 0 && pos + d < extent) v |= in[w + d * stride];
    out[w] = v;
}

// false: certainly no scene point within re.reach of (x, y, z).  Points outside the grid say true (the walk decides).
__device__ __forceinline__ bool rf_maybe_near(const RfGrid& g, const RfReach& re, float x, float y, float z) {
    long long v;
    if (!rf_voxel(g, x, y, z, v)) return true;
    return (__ldg(re.dil + (v >> 5)) >> (v & 31)) & 1u;
}

// ---------------------------------------------------------------------------------------------- small dense eigen-solvers
// Cyclic Jacobi for a symmetric N x N matrix in double: a is destroyed (its diagonal becomes the eigenvalues), v receives the
// eigenvectors as columns.  Every loop is unrolled, so both matrices live in registers; a sweep stops the iteration once the
// off-diagonal mass is below 1e-30 of the diagonal's (quadratic convergence: 5-7 sweeps).
template <int N>
__device__ __forceinline__ void rf_jacobi(double (&a)[N][N], double (&v)[N][N]) {
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) v[i][j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 24; ++sweep) {
        double off = 0.0, diag = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            diag += a[i][i] * a[i][i];
#pragma unroll
            for (int j = i + 1; j < N; ++j) off += a[i][j] * a[i][j];
        }
        if (!(off > 1e-30 * diag)) break;
#pragma unroll
        for (int p = 0; p < N; ++p)
#pragma unroll
            for (int q = p + 1; q < N; ++q) {
                const double apq = a[p][q];
                if (apq == 0.0) continue;
                const double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = rsqrt(t * t + 1.0), s = t * c;
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    const double akp = a[k][p], akq = a[k][q];
                    a[k][p] = c * akp - s * akq;
                    a[k][q] = s * akp + c * akq;
                }
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    const double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = c * apk - s * aqk;
                    a[q][k] = s * apk + c * aqk;
                }
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    const double vkp = v[k][p], vkq = v[k][q];
                    v[k][p] = c * vkp - s * vkq;
                    v[k][q] = s * vkp + c * vkq;
                }
            }
    }
}

// ---------------------------------------------------------------------------------------------- normals
// pcl::NormalEstimation with a radius search, viewpoint (0, 0, 0) (MeshUtils.cpp:196-211; oracle choice R3): covariance of the
// neighbours with d^2 < r^2 (the point itself included) about the query point in double, smallest eigenvector, flipped towards
// the viewpoint, curvature = lambda_0 / trace.  Fewer than 3 neighbours -> NaN.  nrm[i] = (nx, ny, nz, curvature).
__global__ void rf_normals_kernel(const float4* __restrict__ pts, const int* __restrict__ n_pts, RfGrid g,
                                  const unsigned* __restrict__ bits, const unsigned* __restrict__ prefix, float radius,
                                  float4* __restrict__ nrm) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *n_pts) return;
    const float4 q = pts[i];
    const float r2 = __fmul_rn(radius, radius);
    int n = 0;
    double s[3] = {0, 0, 0}, ss[6] = {0, 0, 0, 0, 0, 0};
    rf_for_box(g, bits, prefix, q.x, q.y, q.z, radius, [&](int j) {
        const float4 p = __ldg(pts + j);
        if (!(rf_dist2(p.x, p.y, p.z, q.x, q.y, q.z) < r2)) return;
        const double dx = (double)p.x - (double)q.x, dy = (double)p.y - (double)q.y, dz = (double)p.z - (double)q.z;
        ++n;
        s[0] += dx; s[1] += dy; s[2] += dz;
        ss[0] += dx * dx; ss[1] += dx * dy; ss[2] += dx * dz; ss[3] += dy * dy; ss[4] += dy * dz; ss[5] += dz * dz;
    });
    const float nan = __int_as_float(0x7fc00000);
    if (n < 3) { nrm[i] = make_float4(nan, nan, nan, nan); return; }
    const double inv = 1.0 / n, mx = s[0] * inv, my = s[1] * inv, mz = s[2] * inv;
    double a[3][3], v[3][3];
    a[0][0] = ss[0] * inv - mx * mx; a[0][1] = a[1][0] = ss[1] * inv - mx * my; a[0][2] = a[2][0] = ss[2] * inv - mx * mz;
    a[1][1] = ss[3] * inv - my * my; a[1][2] = a[2][1] = ss[4] * inv - my * mz; a[2][2] = ss[5] * inv - mz * mz;
    const double tr = a[0][0] + a[1][1] + a[2][2];
    rf_jacobi<3>(a, v);
    double lam = a[0][0], nx = v[0][0], ny = v[1][0], nz = v[2][0];  // smallest eigenvalue, static indexing
    if (a[1][1] < lam) { lam = a[1][1]; nx = v[0][1]; ny = v[1][1]; nz = v[2][1]; }
    if (a[2][2] < lam) { lam = a[2][2]; nx = v[0][2]; ny = v[1][2]; nz = v[2][2]; }
    if (-(double)q.x * nx - (double)q.y * ny - (double)q.z * nz < 0) { nx = -nx; ny = -ny; nz = -nz; }  // towards the origin
    nrm[i] = make_float4((float)nx, (float)ny, (float)nz, tr != 0.0 ? (float)fabs(lam / tr) : 0.f);
}

// MeshUtils::get_normals_not_nan (MeshUtils.cpp:213-231): rows with a non-finite normal leave the cloud.  Step 1 clears their
// bits (every point knows its voxel: vox[i]); after a new prefix, step 2 moves the surviving rows to their new ranks.
__global__ void rf_voxel_of_rank_kernel(const unsigned* __restrict__ bits, const unsigned* __restrict__ prefix, long long words,
                                        long long* __restrict__ vox) {
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= words) return;
    unsigned m = bits[w];
    unsigned r = prefix[w];
    while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        vox[r++] = w * 32 + b;
    }
}

__global__ void rf_drop_nan_bits_kernel(const float4* __restrict__ nrm, const int* __restrict__ n_pts,
                                        const long long* __restrict__ vox, unsigned* __restrict__ bits) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *n_pts) return;
    const float4 n = nrm[i];
    if (isfinite(n.x) && isfinite(n.y) && isfinite(n.z)) return;
    const long long v = vox[i];
    atomicAnd(bits + (v >> 5), ~(1u << (v & 31)));
}

__global__ void rf_compact_kernel(const float4* __restrict__ pts_in, const float4* __restrict__ nrm_in, const int* __restrict__ n_in,
                                  const long long* __restrict__ vox, const unsigned* __restrict__ bits,
                                  const unsigned* __restrict__ prefix, float4* __restrict__ pts_out, float4* __restrict__ nrm_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *n_in) return;
    const long long v = vox[i];
    if (!((bits[v >> 5] >> (v & 31)) & 1u)) return;
    const int r = rf_rank(bits, prefix, v);
    pts_out[r] = pts_in[i];
    nrm_out[r] = nrm_in[i];
}

// ---------------------------------------------------------------------------------------------- smooth clusters
// MeshUtils::extractEuclideanClustersSmooth (MeshUtils.cpp:245-337) as connected components: an edge joins two points with
// curvature <= the threshold, closer than the tolerance of either end (0.03 m, 0.05 m beyond z = 1.3 m; the reference's walk
// applies the seed's tolerance, which differs only across the 1.3 m line) and normals within eps_angle.
struct RfClusterParams {
    float tol_near, tol_far, curvature;
    double cos_eps;  // cos(eps_angle): `fabs(acos(dot)) < eps_angle` is `dot > cos(eps_angle)` for dot in [-1, 1]
    int min_points;
};

__device__ __forceinline__ int rf_find(int* parent, int i) {
    int p = parent[i];
    while (p != i) {
        const int gp = parent[p];
        parent[i] = gp;  // path halving; a benign race: every value written is an ancestor
        i = p;
        p = gp;
    }
    return i;
}

__global__ void rf_cluster_init_kernel(const int* __restrict__ n_pts, int* __restrict__ parent, int* __restrict__ size) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *n_pts) return;
    parent[i] = i;
    size[i] = 0;
}

// The smooth-neighbour test of the reference's walk (MeshUtils.cpp:292-310) as an undirected edge.
struct RfEdge {
    float4 q, ni;
    float ti2, reach;
    __device__ __forceinline__ RfEdge(const float4& q_, const float4& ni_, const RfClusterParams& cp) : q(q_), ni(ni_) {
        const float tol_i = q.z > 1.3f ? cp.tol_far : cp.tol_near;
        ti2 = __fmul_rn(tol_i, tol_i);
        // the edge exists when the distance is below the tolerance of EITHER end, so the box reaches as far as a neighbour's could
        reach = q.z + cp.tol_far > 1.3f ? fmaxf(cp.tol_far, cp.tol_near) : cp.tol_near;
    }
    __device__ __forceinline__ bool joins(const float4& p, const float4& nj, const RfClusterParams& cp) const {
        const float d2 = rf_dist2(p.x, p.y, p.z, q.x, q.y, q.z);
        const float tol_j = p.z > 1.3f ? cp.tol_far : cp.tol_near;
        if (!(d2 < ti2 || d2 < __fmul_rn(tol_j, tol_j))) return false;
        if (nj.w > cp.curvature) return false;
        const float dot = __fadd_rn(__fadd_rn(__fmul_rn(ni.x, nj.x), __fmul_rn(ni.y, nj.y)), __fmul_rn(ni.z, nj.z));
        // `fabs(acos(dot)) < eps` without a double-precision acos per neighbour: acos falls monotonically, and acos of a dot
        // product above 1 is NaN, which compares false in the reference (R5)
        return (double)dot <= 1.0 && (double)dot > cp.cos_eps;
    }
};

// Components in three steps: (1) every point links to its lowest smooth neighbour -- plain stores, no atomics; the links
// strictly decrease, so they form a forest whose roots are local minima; (2) the forest is flattened; (3) the edges whose ends
// still have different roots are united with the lock-free hook.  Step 3 is the 2.9 ms of the 4 ms scene preparation (ncu:
// 5 % achieved occupancy -- a tail of a few warps walking the chain of basin roots that hook-by-index builds across the table,
// one component of tens of thousands of points with thousands of basins under the 0.05 rad normal test); rounds of
// min-label propagation with a parallel flatten in between are the known cure and are not built.
__global__ void rf_cluster_link_kernel(const float4* __restrict__ pts, const float4* __restrict__ nrm, const int* __restrict__ n_pts,
                                       RfGrid g, const unsigned* __restrict__ bits, const unsigned* __restrict__ prefix,
                                       RfClusterParams cp, int* __restrict__ parent) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *n_pts) return;
    const float4 ni = nrm[i];
    int lowest = i;
    if (!(ni.w > cp.curvature)) {
        const RfEdge e(pts[i], ni, cp);
        rf_for_box(g, bits, prefix, e.q.x, e.q.y, e.q.z, e.reach, [&](int j) {
            if (j >= lowest) return;  // ascending walk: only the first joining neighbour below i matters
            if (e.joins(__ldg(pts + j), __ldg(nrm + j), cp)) lowest = j;
        });
    }
    parent[i] = lowest;
}

__global__ void rf_cluster_flatten_kernel(const int* __restrict__ n_pts, int* __restrict__ parent) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *n_pts) return;
    int r = parent[i];
    while (true) {
        const int p = parent[r];
        if (p == r) break;
        r = p;
    }
    parent[i] = r;  // concurrent writers store roots or ancestors: every value read above is on i's path to its root
}

__global__ void rf_cluster_hook_kernel(const float4* __restrict__ pts, const float4* __restrict__ nrm, const int* __restrict__ n_pts,
                                       RfGrid g, const unsigned* __restrict__ bits, const unsigned* __restrict__ prefix,
                                       RfClusterParams cp, int* __restrict__ parent) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *n_pts) return;
    const float4 ni = nrm[i];
    if (ni.w > cp.curvature) return;
    const RfEdge e(pts[i], ni, cp);
    int root = parent[i];
    rf_for_box(g, bits, prefix, e.q.x, e.q.y, e.q.z, e.reach, [&](int j) {
        if (j >= i) return;               // every edge once, from its larger end
        if (parent[j] == root) return;    // flattened: same basin (or already united)
        if (!e.joins(__ldg(pts + j), __ldg(nrm + j), cp)) return;
        int a = root, b = j;
        for (;;) {
            a = rf_find(parent, a);
            b = rf_find(parent, b);
            if (a == b) break;
            if (a < b) { const int t = a; a = b; b = t; }  // hook the larger root under the smaller
            const int old = atomicCAS(parent + a, a, b);
            if (old == a) { a = b; break; }
            a = old;
        }
        root = a;
    });
}

__global__ void rf_cluster_count_kernel(const int* __restrict__ n_pts, int* __restrict__ parent, int* __restrict__ size) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *n_pts) return;
    const int r = rf_find(parent, i);
    parent[i] = r;
    atomicAdd(size + r, 1);
}

// cluster ids in the order of the clusters' lowest points (the reference's creation order); one CTA.
// label[i] = id or -1; csize[id]; n_clusters[0].
__global__ void __launch_bounds__(1024) rf_cluster_label_kernel(const int* __restrict__ n_pts, const int* __restrict__ parent,
                                                                const int* __restrict__ size, int min_points,
                                                                int* __restrict__ root_id, int* __restrict__ label,
                                                                int* __restrict__ csize, int* __restrict__ n_clusters) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int n = *n_pts, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const int flag = i < n && parent[i] == i && size[i] >= min_points;
        int incl = flag;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int v = s_warp[lane], iv = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, iv, o);
                if (lane >= o) iv += t;
            }
            s_warp[lane] = iv - v;
        }
        __syncthreads();
        const int id = s_carry + s_warp[warp] + incl - flag;
        if (i < n) root_id[i] = flag ? id : -1;
        if (flag) csize[id] = size[i];
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = id + flag;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_clusters = s_carry;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += 1024) label[i] = root_id[parent[i]];
}

// ---------------------------------------------------------------------------------------------- object models
struct RfModel {          // one object class (MeshUtils::insertObjectFromPLY)
    int first, count;     // rows of the concatenated model arrays
    float max_dist;       // ICP correspondence distance = obj_nn_search_radius_[id] (0 when the options give none)
    float radius;         // scoring search radius: the object's nn_search_radius, or the global one
    int iterations;       // ICP iterations
    float max_center_length;
};

// ---------------------------------------------------------------------------------------------- ICP
struct RfIcpOut {
    float pose[16];   // final pose (row-major); the input pose when ICP did not converge
    int converged, iterations;
    float mse;
    int correspondences;
};

__device__ __forceinline__ double rf_block_sum(double v, double* s_red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < RF_THREADS / 32; ++w) t += s_red[w];  // fixed order: deterministic
    return t;
}

// Rotation that maximises sum q_i . (R p_i) from the cross-covariance H = sum p q^T (Horn 1987): the eigenvector of the largest
// eigenvalue of a symmetric 4x4 built from H is the unit quaternion.  Same optimum as the SVD solution with the determinant
// correction (pcl::TransformationEstimationSVD), always a proper rotation.
__device__ __noinline__ void rf_rotation_from_covariance(const double H[3][3], double R[3][3]) {
    double a[4][4], v[4][4];
    const double Sxx = H[0][0], Sxy = H[0][1], Sxz = H[0][2], Syx = H[1][0], Syy = H[1][1], Syz = H[1][2], Szx = H[2][0],
                 Szy = H[2][1], Szz = H[2][2];
    a[0][0] = Sxx + Syy + Szz; a[0][1] = Syz - Szy; a[0][2] = Szx - Sxz; a[0][3] = Sxy - Syx;
    a[1][1] = Sxx - Syy - Szz; a[1][2] = Sxy + Syx; a[1][3] = Szx + Sxz;
    a[2][2] = -Sxx + Syy - Szz; a[2][3] = Syz + Szy;
    a[3][3] = -Sxx - Syy + Szz;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < i; ++j) a[i][j] = a[j][i];
    rf_jacobi<4>(a, v);
    double lam = a[0][0], w = v[0][0], x = v[1][0], y = v[2][0], z = v[3][0];  // largest eigenvalue, static indexing
#pragma unroll
    for (int i = 1; i < 4; ++i)
        if (a[i][i] > lam) { lam = a[i][i]; w = v[0][i]; x = v[1][i]; y = v[2][i]; z = v[3][i]; }
    R[0][0] = w * w + x * x - y * y - z * z; R[0][1] = 2 * (x * y - w * z); R[0][2] = 2 * (x * z + w * y);
    R[1][0] = 2 * (x * y + w * z); R[1][1] = w * w - x * x + y * y - z * z; R[1][2] = 2 * (y * z - w * x);
    R[2][0] = 2 * (x * z - w * y); R[2][1] = 2 * (y * z + w * x); R[2][2] = w * w - x * x - y * y + z * z;
}

// pcl::IterativeClosestPoint::computeTransformation + DefaultConvergenceCriteria (oracle choice R8).  A hypothesis is one
// THREAD-BLOCK CLUSTER of RF_ICP_CLUSTER CTAs: a frame has a few hundred hypotheses, fewer than three CTAs per SM, so the
// model points of one hypothesis are spread over the CTAs of a cluster; every iteration the CTAs leave their partial sums in
// their own shared memory, CTA 0 adds them through distributed shared memory (fixed order), solves for the increment and
// writes the new transform and the loop state into every CTA's shared memory.  Two cluster barriers per iteration, no global
// memory traffic, no host round trip.
constexpr int RF_ICP_CLUSTER = 4;
#ifndef RF_ICP_MIN_CTAS
#define RF_ICP_MIN_CTAS 4  // registers capped at 64: the serial solve of CTA 0 spills, every other thread gains occupancy
#endif

struct RfIcpShared {
    double T[12];      // current total increment (3x4): p_now = T p_0, p_0 = pose0 * model point
    double part[17];   // this CTA's partial sums of the iteration
    int state;         // 0 running, 1 converged, 2 failed
    int iter;
};

__global__ void __cluster_dims__(RF_ICP_CLUSTER, 1, 1) __launch_bounds__(RF_THREADS, RF_ICP_MIN_CTAS)
rf_icp_kernel(const float* __restrict__ pose0 /*[n][16]*/, const int* __restrict__ cls, const RfModel* __restrict__ models,
              const float4* __restrict__ mpts, const float4* __restrict__ spts, const int* __restrict__ n_scene, RfGrid g,
              const unsigned* __restrict__ bits, const unsigned* __restrict__ prefix, RfReach co,
              int* __restrict__ nn_cache /*[n][nn_stride]*/, int nn_stride, RfIcpOut* __restrict__ out) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ RfIcpShared sh;
    __shared__ double s_red[RF_THREADS / 32];
    __shared__ double s_prev_mse;   // CTA 0 only
    __shared__ int s_corr;
    __shared__ float s_mse;
    const int h = blockIdx.x / RF_ICP_CLUSTER;
    const int crank = (int)cluster.block_rank();
    const RfModel m = models[cls[h]];
    const float* P0 = pose0 + (size_t)h * 16;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 12; ++i) sh.T[i] = (i % 5 == 0) ? 1.0 : 0.0;  // identity 3x4: entries 0, 5, 10
        sh.state = (*n_scene == 0 || m.count == 0) ? 2 : 0;
        sh.iter = 0;
        s_prev_mse = 3.4028234663852886e38;
        s_corr = 0;
        s_mse = 0.f;
    }
    cluster.sync();
    const float max_d2 = __fmul_rn(m.max_dist, m.max_dist);
    while (sh.state == 0) {
        double acc[17];
#pragma unroll
        for (int i = 0; i < 17; ++i) acc[i] = 0.0;
        for (int k = crank * RF_THREADS + threadIdx.x; k < m.count; k += RF_ICP_CLUSTER * RF_THREADS) {
            const float4 mp = __ldg(mpts + m.first + k);
            // p_0 in float as pcl::transformPointCloud computes it, then the accumulated increment in double
            const float x0 = P0[0] * mp.x + P0[1] * mp.y + P0[2] * mp.z + P0[3];
            const float y0 = P0[4] * mp.x + P0[5] * mp.y + P0[6] * mp.z + P0[7];
            const float z0 = P0[8] * mp.x + P0[9] * mp.y + P0[10] * mp.z + P0[11];
            const double px = sh.T[0] * x0 + sh.T[1] * y0 + sh.T[2] * z0 + sh.T[3];
            const double py = sh.T[4] * x0 + sh.T[5] * y0 + sh.T[6] * z0 + sh.T[7];
            const double pz = sh.T[8] * x0 + sh.T[9] * y0 + sh.T[10] * z0 + sh.T[11];
            const float fx = (float)px, fy = (float)py, fz = (float)pz;
            // The nearest scene point moves little between iterations: the previous one bounds the search (it is itself a
            // candidate, so the nearest point is no further away) -- after the first iteration the walk covers a box of a few
            // voxels instead of the whole correspondence distance.  The result is the same point either way.
            int* cache = nn_cache + (size_t)h * nn_stride + k;
            if (co.dil && m.max_dist <= co.reach && !rf_maybe_near(g, co, fx, fy, fz)) {  // nothing within reach
                *cache = -1;
                continue;
            }
            float reach = m.max_dist;
            const int pj = sh.iter > 0 ? *cache : -1;
            if (pj >= 0) {
                const float4 sp = __ldg(spts + pj);
                const float dp = sqrtf(rf_dist2(sp.x, sp.y, sp.z, fx, fy, fz));
                reach = fminf(reach, dp * 1.000001f + 1e-7f);
            }
            float best = 3.4e38f;
            int bj = -1;
            rf_for_box<true>(g, bits, prefix, fx, fy, fz, reach, [&](int j) {
                const float4 sp = __ldg(spts + j);
                const float d2 = rf_dist2(sp.x, sp.y, sp.z, fx, fy, fz);
                if (d2 < best) { best = d2; bj = j; }
            });
            *cache = (bj >= 0 && !(best > max_d2)) ? bj : -1;
            if (bj < 0 || best > max_d2) continue;
            const float4 sp = __ldg(spts + bj);
            acc[0] += 1.0;
            acc[1] += px; acc[2] += py; acc[3] += pz;
            acc[4] += sp.x; acc[5] += sp.y; acc[6] += sp.z;
            acc[7] += px * sp.x; acc[8] += px * sp.y; acc[9] += px * sp.z;
            acc[10] += py * sp.x; acc[11] += py * sp.y; acc[12] += py * sp.z;
            acc[13] += pz * sp.x; acc[14] += pz * sp.y; acc[15] += pz * sp.z;
            acc[16] += (double)best;
        }
#pragma unroll
        for (int i = 0; i < 17; ++i) {
            const double t = rf_block_sum(acc[i], s_red);
            if (threadIdx.x == 0) sh.part[i] = t;
        }
        cluster.sync();  // every CTA's partial sums are in its shared memory
        if (crank == 0 && threadIdx.x == 0) {
            double tot[17];
            for (int i = 0; i < 17; ++i) tot[i] = 0.0;
            for (int r = 0; r < RF_ICP_CLUSTER; ++r) {
                const RfIcpShared* peer = cluster.map_shared_rank(&sh, r);
                for (int i = 0; i < 17; ++i) tot[i] += peer->part[i];
            }
            const double n = tot[0];
            int state = 0;
            int it = sh.iter;
            double Tn[12];
            for (int i = 0; i < 12; ++i) Tn[i] = sh.T[i];
            s_corr = (int)n;
            if (n < 3.0) {
                state = 2;  // "Not enough correspondences found": converged_ = false
            } else {
                const double cp[3] = {tot[1] / n, tot[2] / n, tot[3] / n}, cq[3] = {tot[4] / n, tot[5] / n, tot[6] / n};
                double H[3][3], R[3][3];
                for (int a = 0; a < 3; ++a)
                    for (int b = 0; b < 3; ++b) H[a][b] = tot[7 + 3 * a + b] - n * cp[a] * cq[b];
                rf_rotation_from_covariance(H, R);
                double t[3];
                for (int a = 0; a < 3; ++a) t[a] = cq[a] - (R[a][0] * cp[0] + R[a][1] * cp[1] + R[a][2] * cp[2]);
                for (int a = 0; a < 3; ++a)
                    for (int b = 0; b < 4; ++b)
                        Tn[4 * a + b] = R[a][0] * sh.T[b] + R[a][1] * sh.T[4 + b] + R[a][2] * sh.T[8 + b] + (b == 3 ? t[a] : 0.0);
                ++it;
                const double mse = tot[16] / n;
                s_mse = (float)mse;
                const double cos_angle = 0.5 * (R[0][0] + R[1][1] + R[2][2] - 1.0);
                const double tr2 = t[0] * t[0] + t[1] * t[1] + t[2] * t[2];
                if (it >= m.iterations) state = 1;
                else if (cos_angle >= 1.0 && tr2 <= 0.0) state = 1;
                else if (fabs(mse - s_prev_mse) < 1e-12) state = 1;
                else s_prev_mse = mse;
            }
            for (int r = 0; r < RF_ICP_CLUSTER; ++r) {
                RfIcpShared* peer = cluster.map_shared_rank(&sh, r);
                for (int i = 0; i < 12; ++i) peer->T[i] = Tn[i];
                peer->iter = it;
                peer->state = state;
            }
        }
        cluster.sync();  // the new transform and state are in every CTA's shared memory
    }
    if (crank == 0 && threadIdx.x == 0) {
        RfIcpOut o;
        o.converged = sh.state == 1;
        o.iterations = sh.iter;
        o.mse = s_mse;
        o.correspondences = s_corr;
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 4; ++b) {
                double v = P0[4 * a + b];
                if (o.converged)
                    v = sh.T[4 * a + 0] * P0[b] + sh.T[4 * a + 1] * P0[4 + b] + sh.T[4 * a + 2] * P0[8 + b] + (b == 3 ? sh.T[4 * a + 3] : 0.0);
                o.pose[4 * a + b] = (float)v;
            }
        o.pose[12] = 0.f; o.pose[13] = 0.f; o.pose[14] = 0.f; o.pose[15] = 1.f;
        out[h] = o;
    }
}

// ---------------------------------------------------------------------------------------------- scoring
struct RfScoreParams {
    RfCam cam;
    float occlusion, nn_global;  // occlusion_threshold_, nn_search_radius_ (the divisor of the depth score)
    float similarity, inliers, clutter, location, pose;  // *_reg_
    float inliers_thr, clutter_thr, final_thr;
    int use_color, use_normal;
};

struct RfEval {
    float similarity, inliers_ratio, clutter, location_score, pose_score, final_score;
    int accepted, visible, inliers, explained;
};

// MeshUtils::evaluate_hypothesis (MeshUtils.cpp:629-793), one CTA per hypothesis.  mnrm = normals of the model cloud estimated
// in the object frame (a rotation carries them over; they are flipped towards the camera per hypothesis, as the reference's
// re-estimation on the transformed cloud does); rows with NaN are the points get_normals_not_nan would drop again.
__global__ void __launch_bounds__(RF_THREADS)
rf_evaluate_kernel(const RfIcpOut* __restrict__ icp, const int* __restrict__ cls, const float* __restrict__ loc_score,
                   const float* __restrict__ pose_score, const RfModel* __restrict__ models, const float4* __restrict__ mpts,
                   const float4* __restrict__ mnrm, const uint16_t* __restrict__ depth, const float4* __restrict__ spts,
                   const float4* __restrict__ snrm, const int* __restrict__ label, const int* __restrict__ csize,
                   const int* __restrict__ n_clusters, int cl_cap, RfGrid g, const unsigned* __restrict__ bits,
                   const unsigned* __restrict__ prefix, RfReach co, RfScoreParams sp, int words_per_hyp, unsigned* __restrict__ explained,
                   int* __restrict__ scene_cl /*[n][cl_cap]*/, int* __restrict__ model_cl, RfEval* __restrict__ out) {
    __shared__ double s_red[RF_THREADS / 32];
    __shared__ int s_cnt[4];  // visible, inliers, not_in_cluster, explained
    const int h = blockIdx.x;
    const RfModel m = models[cls[h]];
    const float* P = icp[h].pose;
    unsigned* ex = explained + (size_t)h * words_per_hyp;
    int* scl = scene_cl + (size_t)h * cl_cap;
    int* mcl = model_cl + (size_t)h * cl_cap;
    const int ncl = min(*n_clusters, cl_cap);
    if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    RfEval ev;
    ev.location_score = loc_score[h];
    ev.pose_score = pose_score[h];
    ev.similarity = 0.f; ev.inliers_ratio = 0.f; ev.clutter = 0.f; ev.final_score = 0.f;
    ev.accepted = 0; ev.visible = 0; ev.inliers = 0; ev.explained = 0;
    if (P[11] > 1.5f) {  // `if (h.rotmat(2,3) > 1.5f) return false`
        if (threadIdx.x == 0) out[h] = ev;
        return;
    }
    const float r2 = __fmul_rn(m.radius, m.radius);
    double sim = 0.0;
    int visible = 0, inliers = 0, nic = 0, expl = 0;
    for (int k = threadIdx.x; k < m.count; k += RF_THREADS) {
        const float4 mn = __ldg(mnrm + m.first + k);
        if (!(isfinite(mn.x) && isfinite(mn.y) && isfinite(mn.z))) continue;  // dropped by the re-estimation
        const float4 mp = __ldg(mpts + m.first + k);
        const float x = P[0] * mp.x + P[1] * mp.y + P[2] * mp.z + P[3];
        const float y = P[4] * mp.x + P[5] * mp.y + P[6] * mp.z + P[7];
        const float z = P[8] * mp.x + P[9] * mp.y + P[10] * mp.z + P[11];
        float nx = P[0] * mn.x + P[1] * mn.y + P[2] * mn.z, ny = P[4] * mn.x + P[5] * mn.y + P[6] * mn.z,
              nz = P[8] * mn.x + P[9] * mn.y + P[10] * mn.z;
        if (-(x * nx + y * ny + z * nz) < 0.f) { nx = -nx; ny = -ny; nz = -nz; }
        // world_to_image_coords (MeshUtils.cpp:62-65); outside the image: no scene depth (oracle choice R7)
        const float rowf = __fadd_rn(__fdiv_rn(__fmul_rn(y, sp.cam.fy), z), sp.cam.cy);
        const float colf = __fadd_rn(__fdiv_rn(__fmul_rn(x, sp.cam.fx), z), sp.cam.cx);
        float scene_d = 0.f;
        if (fabsf(rowf) < 2e9f && fabsf(colf) < 2e9f) {
            const int row = (int)rowf, col = (int)colf;
            if (row >= 0 && row < sp.cam.H && col >= 0 && col < sp.cam.W)
                scene_d = __fdiv_rn((float)depth[(size_t)row * sp.cam.W + col], 1000.0f);
        }
        if (!(scene_d == 0.f || z < __fadd_rn(scene_d, sp.occlusion))) continue;
        ++visible;
        const unsigned mc = __float_as_uint(mp.w);
        const float mr = (float)(mc & 255u), mg = (float)((mc >> 8) & 255u), mb = (float)((mc >> 16) & 255u);
        float best = 0.f, best_d2 = 3.4e38f;
        int best_id = -1, found = 0;
        if (co.dil && m.radius <= co.reach && !rf_maybe_near(g, co, x, y, z)) continue;  // visible, no scene point in reach
        rf_for_box(g, bits, prefix, x, y, z, m.radius, [&](int j) {
            const float4 q = __ldg(spts + j);
            const float d2 = rf_dist2(q.x, q.y, q.z, x, y, z);
            if (!(d2 < r2)) return;
            ++found;
            const float depth_score = __fsub_rn(1.0f, __fdiv_rn(d2, sp.nn_global));
            float normal_score = 1.f;
            if (sp.use_normal) {
                const float4 qn = __ldg(snrm + j);
                const float dot = __fadd_rn(__fadd_rn(__fmul_rn(qn.x, nx), __fmul_rn(qn.y, ny)), __fmul_rn(qn.z, nz));
                normal_score = __fadd_rn(__fdiv_rn(dot, 2.0f), 0.5f);
            }
            const unsigned sc = __float_as_uint(q.w);
            const float dr = fabsf((float)(sc & 255u) - mr), dg = fabsf((float)((sc >> 8) & 255u) - mg),
                        db = fabsf((float)((sc >> 16) & 255u) - mb);
            const float color_score = (float)(1.0 - (double)fmaxf(fmaxf(dr, dg), db) / 255.0);
            const float score = sp.use_color ? __fdiv_rn(__fadd_rn(__fadd_rn(normal_score, depth_score), color_score), 3.0f)
                                             : __fdiv_rn(__fadd_rn(normal_score, depth_score), 2.0f);
            // the reference walks the neighbours nearest first and keeps the first strict maximum
            if (score > best || (score == best && best_id >= 0 && d2 < best_d2)) { best = score; best_id = j; best_d2 = d2; }
            const unsigned bit = 1u << (j & 31);
            if (!(atomicOr(ex + (j >> 5), bit) & bit)) {
                ++expl;
                const int c = label[j];
                if (c >= 0 && c < cl_cap) atomicAdd(scl + c, 1);
            }
        });
        if (!found) continue;
        sim += (double)best;
        const int c = best_id >= 0 ? label[best_id] : -1;
        if (c >= 0 && c < cl_cap) atomicAdd(mcl + c, 1); else ++nic;
        ++inliers;
    }
    atomicAdd(&s_cnt[0], visible);
    atomicAdd(&s_cnt[1], inliers);
    atomicAdd(&s_cnt[2], nic);
    atomicAdd(&s_cnt[3], expl);
    sim = rf_block_sum(sim, s_red);
    __syncthreads();
    const int n_in = s_cnt[1], n_nic = s_cnt[2];
    double cl = 0.0;
    if (n_in - n_nic > 0)
        for (int c = threadIdx.x; c < ncl; c += RF_THREADS) {
            const int e = scl[c];
            if (e != 0) {
                const float a = __fdiv_rn((float)(csize[c] - e), (float)csize[c]);
                const float b = __fdiv_rn((float)mcl[c], (float)(n_in - n_nic));
                cl += (double)__fmul_rn(a, b);
            }
        }
    cl = rf_block_sum(cl, s_red);
    if (threadIdx.x == 0) {
        ev.visible = s_cnt[0];
        ev.inliers = n_in;
        ev.explained = s_cnt[3];
        ev.similarity = __fdiv_rn((float)sim, (float)n_in);
        ev.clutter = n_in - n_nic <= 0 ? 1.f : (float)cl;
        ev.inliers_ratio = __fdiv_rn((float)n_in, (float)ev.visible);
        float fs = __fmul_rn(ev.similarity, sp.similarity);
        fs = __fadd_rn(fs, __fmul_rn(ev.inliers_ratio, sp.inliers));
        fs = __fsub_rn(fs, __fmul_rn(ev.clutter, sp.clutter));
        fs = __fadd_rn(fs, __fmul_rn(ev.pose_score, sp.pose));
        fs = __fadd_rn(fs, __fmul_rn(ev.location_score, sp.location));
        ev.final_score = fs;
        ev.accepted = (ev.clutter > sp.clutter_thr || ev.inliers_ratio < sp.inliers_thr) ? 0 : (fs > sp.final_thr ? 1 : 0);
        out[h] = ev;
    }
}

// ---------------------------------------------------------------------------------------------- joint optimisation
// common[a][b] = scene points explained by both accepted hypotheses a and b (MeshUtils.cpp:889-897); grid (n, n).
__global__ void rf_common_kernel(const unsigned* __restrict__ explained, int words_per_hyp, const int* __restrict__ list, int n,
                                 int* __restrict__ common) {
    __shared__ int s_sum;
    const int a = blockIdx.x, b = blockIdx.y;
    if (b < a) return;
    if (threadIdx.x == 0) s_sum = 0;
    __syncthreads();
    const unsigned* ea = explained + (size_t)list[a] * words_per_hyp;
    const unsigned* eb = explained + (size_t)list[b] * words_per_hyp;
    int c = 0;
    for (int w = threadIdx.x; w < words_per_hyp; w += blockDim.x) c += __popc(ea[w] & eb[w]);
#pragma unroll
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_sum, c);
    __syncthreads();
    if (threadIdx.x == 0) { common[a * n + b] = s_sum; common[b * n + a] = s_sum; }
}

// Per solution of a group (a subset of its members as a bit vector of `sol_words` 32-bit words): scene points explained by at
// least one chosen member and the surplus of multiply explained ones (MeshUtils.cpp:982-996).  members[i] = hypothesis of group
// member i.  Grid = solutions; counts[2 * s] = total, counts[2 * s + 1] = common cost.
__global__ void rf_solution_kernel(const unsigned* __restrict__ explained, int words_per_hyp, const int* __restrict__ members,
                                   int n_members, const unsigned* __restrict__ solutions, int sol_words, int* __restrict__ counts) {
    __shared__ int s_tot, s_com;
    if (threadIdx.x == 0) { s_tot = 0; s_com = 0; }
    __syncthreads();
    const unsigned* sol = solutions + (size_t)blockIdx.x * sol_words;
    int tot = 0, com = 0;
    for (int w = threadIdx.x; w < words_per_hyp; w += blockDim.x) {
        unsigned any = 0;  // points explained so far
        int surplus = 0;
        for (int i = 0; i < n_members; ++i) {
            if (!((sol[i >> 5] >> (i & 31)) & 1u)) continue;
            const unsigned e = explained[(size_t)members[i] * words_per_hyp + w];
            surplus += __popc(any & e);  // every further explanation of an already explained point costs one
            any |= e;
        }
        tot += __popc(any);
        com += surplus;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) { tot += __shfl_xor_sync(0xffffffffu, tot, o); com += __shfl_xor_sync(0xffffffffu, com, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&s_tot, tot); atomicAdd(&s_com, com); }
    __syncthreads();
    if (threadIdx.x == 0) { counts[2 * blockIdx.x] = s_tot; counts[2 * blockIdx.x + 1] = s_com; }
}

}  // namespace hf6d
