// Shared device/host definitions for libhf6d kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "../../include/hf6d.h"
#include "model.hpp"

namespace hf6d {

// Per-frame constants every kernel needs (passed by value).
struct FrameGeom {
    int W, H, stride;
    int gw, gh;          // stride-grid size: ceil(W/stride), ceil(H/stride)
    float fx, fy, cx, cy;
    int ps;              // patch_size_in_voxels (8)
    float vox, range, dist_thr;
    int fill_random;
    unsigned long long fill_seed;
    int batch;
    int cap;             // patch capacity of the per-slot buffers (multiple of 128)
    float focal;         // focal length of the adaptive patch size: fx (RGB-D patches) or normals_focal (normals variant)
};

// Device view of the flattened forest (model.hpp::HostForest).
struct DevForest {
    int T, K, F;
    const PackedRecord* recs;  // two tree levels per 48-byte record
    const int32_t* root;       // [T] record-numbered root entries
    const int32_t* leaf_base;  // [T+1]
    const int32_t* group_off;  // [L+1]
    const VoteGroup* groups;
    const float* oz;           // per vote: z of R(yaw,pitch,roll)*(-x,-y,-z) (the z-histogram re-walk reads only this)
    const float4* vote4;       // per vote: (ox, oy, oz, bits: cls | w << 5) -- one 16-byte load casts a vote
    const int32_t* vgroup;     // per vote: its group (read only for votes that land in a centre window)
    const int2* leaf_votes;    // per leaf: (first vote, number of gated votes) over all its groups (contiguous)
    const short4* bins;        // per vote: integer-degree yaw, pitch, roll bins (HFTest.cpp:779-780, :863)
    const float2* oz_range;    // per group: (min, max) of its votes' oz; (inf, -inf) when unordered (a NaN among them)
    const float* oz_sorted;    // per vote: oz, ascending within every ordered group
};

// x86 cvttss2si semantics: NaN / out of range -> INT_MIN (CUDA's cast saturates and maps NaN to 0).
__device__ __forceinline__ int f2i_x86(float y) {
    return fabsf(y) < 2147483648.0f ? (int)y : INT_MIN;  // the comparison is false for NaN; -2^31 itself maps to INT_MIN either way
}

// x / Y for a compile-time constant Y without the IEEE division sequence: q = x * RN(1/Y), one FMA for the exact
// remainder, one FMA to correct.  Bit-identical to x / Y for every x with |x| in [1e-30, 1e30) and for x = +0 (checked
// exhaustively over all 2^32 floats for Y = 3, 64, 192, 255, 1000, 0.01 by tools/check_const_division.c); callers
// guarantee the range, or only use the truncated integer part (tiny / huge / infinite x then give the same integer).
constexpr int HF6D_MAX_PEERS = 8;  // ranks of a tree-sharded group on one NVLink / NVSwitch box

template <int NUM, int DEN>
__device__ __forceinline__ float div_const(float x) {
    constexpr float Y = (float)NUM / (float)DEN;
    constexpr float R = 1.0f / Y;
    const float q = __fmul_rn(x, R);
    return __fmaf_rn(__fmaf_rn(-Y, q, x), R, q);
}

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

// Patch sharding (one stream of frames over several GPUs): rank r of `world` gathers, encodes, traverses and votes the patches
// [lo, hi) of the frame's P' processed patches.  The cuts fall on multiples of 128 patches, the encoder's row-block size.
struct PatchShard {
    int rank, world;  // world == 1: everything
};
__host__ __device__ __forceinline__ void patch_shard_range(int Pp, PatchShard ps, int& lo, int& hi) {
    const int mb = (Pp + 127) / 128;
    lo = min(Pp, (int)((long long)mb * ps.rank / ps.world) * 128);
    hi = min(Pp, (int)((long long)mb * (ps.rank + 1) / ps.world) * 128);
}
// the rank whose range holds patch p
__device__ __forceinline__ int patch_shard_owner(int Pp, int world, int p) {
    const int mb = (Pp + 127) / 128, b = p >> 7;
    int r = 0;
    while (r + 1 < world && (int)((long long)mb * (r + 1) / world) <= b) ++r;
    return r;
}

// Blocking host -> device copy whose data is IN device memory on return.  A plain cudaMemcpy from pageable memory returns once
// the bytes are staged; the DMA may still be in flight, and the kernels that read the destination run on non-blocking streams
// that do not order themselves behind the legacy stream -- so wait for the legacy stream as well.
inline cudaError_t memcpy_h2d_done(void* dst, const void* src, size_t bytes) {
    const cudaError_t e = cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
    return e != cudaSuccess ? e : cudaStreamSynchronize(cudaStreamLegacy);
}

// patch_extractor.cu:257 / :378 -- ((ps*vox)/d)*f in fp32, truncated
__device__ __forceinline__ int adaptive_size(const FrameGeom& g, float depth_m) {
    return (int)__fmul_rn(__fdiv_rn(__fmul_rn((float)g.ps, g.vox), depth_m), g.focal);
}

}  // namespace hf6d
