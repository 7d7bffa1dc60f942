// Cross-check of the software texture filter in gather.cuh against the hardware texture unit.
//
// The reference's gather IS the texture unit: a 3-D float texture of extent (4 channels, W, H), unnormalised
// coordinates, linear filter, border addressing, fetched at (ch + 0.5, u + 0.5, v + 0.5)
// (PatchGen/src/cuda/patch_extractor.cu:339-343, 230-309).  The legacy texture<> reference API it uses no longer exists
// in CUDA 12, so this file rebuilds the same fetch with a cudaTextureObject_t and reproduces extract_rgbd's arithmetic
// around it (fill = 0, i.e. are_objects_segmented).  It is a diagnostic entry point (hf6d_debug_texture_gather): the
// GPU parity suite compares its fp32 patches with the oracle's software filter.  Not on the hot path.
#pragma once
#include "common.cuh"

namespace hf6d {

__global__ void texture_build_kernel(const uint8_t* __restrict__ bgr, const uint16_t* __restrict__ depth, int W, int H,
                                     float* __restrict__ vol /*[H][W][4]*/) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W * H) return;
    // HFTest.cpp:370-379: B,G,R / 255.0f, depth in millimetres as float
    vol[4 * i + 0] = __fdiv_rn((float)bgr[3 * i + 0], 255.0f);
    vol[4 * i + 1] = __fdiv_rn((float)bgr[3 * i + 1], 255.0f);
    vol[4 * i + 2] = __fdiv_rn((float)bgr[3 * i + 2], 255.0f);
    vol[4 * i + 3] = (float)depth[i];
}

// One 64-thread block per patch, as the reference launches it.
__global__ void texture_gather_kernel(cudaTextureObject_t tex, FrameGeom g, const int* __restrict__ locs,
                                      const int* __restrict__ counts, float* __restrict__ out /*[P'][ps][ps][4]*/) {
    const int p = blockIdx.x;
    if (p >= counts[1]) return;
    const int cx = locs[2 * p], cy = locs[2 * p + 1];
    const float dc = __fdiv_rn(tex3D<float>(tex, 3 + 0.5f, cx + 0.5f, cy + 0.5f), 1000.0f);
    const int a = adaptive_size(g, dc);
    const int x0 = cx - a / 2, y0 = cy - a / 2;
    const float step = __fdiv_rn((float)a, (float)g.ps);
    for (int v = threadIdx.x; v < g.ps * g.ps; v += blockDim.x) {
        const int tx = v % g.ps, ty = v / g.ps;
        const float u = __fadd_rn((float)x0, __fmul_rn((float)tx, step));
        const float w = __fadd_rn((float)y0, __fmul_rn((float)ty, step));
        const float d = __fdiv_rn(tex3D<float>(tex, 3 + 0.5f, u + 0.5f, w + 0.5f), 1000.0f);
        float* o = out + ((size_t)p * g.ps * g.ps + v) * 4;
        if (d > 0.f) {
            o[0] = tex3D<float>(tex, 0 + 0.5f, u + 0.5f, w + 0.5f);
            o[1] = tex3D<float>(tex, 1 + 0.5f, u + 0.5f, w + 0.5f);
            o[2] = tex3D<float>(tex, 2 + 0.5f, u + 0.5f, w + 0.5f);
            float td = __fadd_rn(__fdiv_rn(__fsub_rn(d, dc), g.range), 0.5f);
            if (td > 1.0f) td = 1.0f;
            if (td < 0.0f) td = 0.0f;
            o[3] = td;
        } else {
            o[0] = o[1] = o[2] = o[3] = 0.f;
        }
    }
}

}  // namespace hf6d
