// View renderer: what PatchGen --render produces per view (PatchGen/src/render_views_tesselated_sphere_mod.cpp),
// without VTK / OpenGL: a z-buffer rasteriser for triangle meshes with per-vertex colour.
//
//  * pass 1, one thread per triangle: project, walk the pixels of its bounding box whose centre lies inside, and race for the
//    pixel with a 64-bit atomicMin on (float depth bits << 32 | triangle index) -- the nearest surface wins, the lower triangle on
//    exact ties, whatever the schedule;
//  * pass 2, one thread per pixel: the winning triangle's barycentric weights again, perspective-correct depth and colour, the
//    headlight shading, uint16 millimetres and 8-bit BGR on a white background.
// Every expression is evaluated in double in the order oracle/render.py writes it (the library is built with --fmad=false), so
// the two agree to the pixel.  The meshes of this domain have 10^4..10^6 small triangles; a triangle that covers a large part of
// the image is walked by one thread and is merely slow.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace hf6d {

struct RdView {
    double R[9], t[3];  // world -> camera (camera looks down -z, y up): the reference's pose<N>.txt
    double f, cx, cy;
    double ambient;
    int W, H;
};

struct RdVertex {
    double px, py, pz;  // camera coordinates
    double z, sx, sy;   // depth along the view direction, continuous pixel coordinates
};

__global__ void rd_project_kernel(const float* __restrict__ xyz, int n, RdView v, RdVertex* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double X = xyz[3 * i], Y = xyz[3 * i + 1], Z = xyz[3 * i + 2];
    RdVertex o;
    o.px = ((X * v.R[0] + Y * v.R[1]) + Z * v.R[2]) + v.t[0];
    o.py = ((X * v.R[3] + Y * v.R[4]) + Z * v.R[5]) + v.t[1];
    o.pz = ((X * v.R[6] + Y * v.R[7]) + Z * v.R[8]) + v.t[2];
    o.z = -o.pz;
    o.sx = (v.f * o.px) / o.z + v.cx;
    o.sy = v.cy - (v.f * o.py) / o.z;
    out[i] = o;
}

struct RdBary {
    double w0, w1, w2;
    bool inside;
};
__device__ __forceinline__ RdBary rd_bary(const RdVertex& a, const RdVertex& b, const RdVertex& c, double area, double px, double py) {
    RdBary r;
    r.w0 = ((b.sx - px) * (c.sy - py) - (c.sx - px) * (b.sy - py)) / area;
    r.w1 = ((c.sx - px) * (a.sy - py) - (a.sx - px) * (c.sy - py)) / area;
    r.w2 = (1.0 - r.w0) - r.w1;
    r.inside = r.w0 >= 0 && r.w1 >= 0 && r.w2 >= 0;
    return r;
}
__device__ __forceinline__ double rd_area(const RdVertex& a, const RdVertex& b, const RdVertex& c) {
    return (b.sx - a.sx) * (c.sy - a.sy) - (c.sx - a.sx) * (b.sy - a.sy);
}

__global__ void rd_raster_kernel(const RdVertex* __restrict__ vtx, const int* __restrict__ faces, int n_faces, int W, int H,
                                 unsigned long long* __restrict__ zbuf) {
    const int ti = blockIdx.x * blockDim.x + threadIdx.x;
    if (ti >= n_faces) return;
    const RdVertex a = vtx[faces[3 * ti]], b = vtx[faces[3 * ti + 1]], c = vtx[faces[3 * ti + 2]];
    if (!(a.z > 0 && b.z > 0 && c.z > 0)) return;
    const double xmin = fmin(fmin(a.sx, b.sx), c.sx), xmax = fmax(fmax(a.sx, b.sx), c.sx);
    const double ymin = fmin(fmin(a.sy, b.sy), c.sy), ymax = fmax(fmax(a.sy, b.sy), c.sy);
    if (!(xmax >= -1.0 && xmin <= W + 1.0 && ymax >= -1.0 && ymin <= H + 1.0)) return;  // also rejects NaN
    const int x0 = max((int)floor(xmin - 0.5), 0), x1 = min((int)ceil(xmax - 0.5), W - 1);
    const int y0 = max((int)floor(ymin - 0.5), 0), y1 = min((int)ceil(ymax - 0.5), H - 1);
    if (x0 > x1 || y0 > y1) return;
    const double area = rd_area(a, b, c);
    if (area == 0) return;
    for (int y = y0; y <= y1; ++y)
        for (int x = x0; x <= x1; ++x) {
            const RdBary w = rd_bary(a, b, c, area, x + 0.5, y + 0.5);
            if (!w.inside) continue;
            const double d = 1.0 / ((w.w0 / a.z + w.w1 / b.z) + w.w2 / c.z);
            const float df = (float)d;
            if (!(df > 0.f)) continue;
            atomicMin(zbuf + (size_t)y * W + x, ((unsigned long long)__float_as_uint(df) << 32) | (unsigned)ti);
        }
}

__global__ void rd_resolve_kernel(const RdVertex* __restrict__ vtx, const int* __restrict__ faces, const uint8_t* __restrict__ rgb,
                                  const unsigned long long* __restrict__ zbuf, RdView v, uint8_t* __restrict__ bgr,
                                  uint16_t* __restrict__ depth_mm) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v.W * v.H) return;
    const unsigned long long key = zbuf[i];
    if (key == ~0ull) {
        bgr[3 * i] = 255; bgr[3 * i + 1] = 255; bgr[3 * i + 2] = 255;  // white background (:260)
        depth_mm[i] = 0;
        return;
    }
    const int ti = (int)(unsigned)key, y = i / v.W, x = i - y * v.W;
    const int ia = faces[3 * ti], ib = faces[3 * ti + 1], ic = faces[3 * ti + 2];
    const RdVertex a = vtx[ia], b = vtx[ib], c = vtx[ic];
    const RdBary w = rd_bary(a, b, c, rd_area(a, b, c), x + 0.5, y + 0.5);
    const double q0 = w.w0 / a.z, q1 = w.w1 / b.z, q2 = w.w2 / c.z;
    const double d = 1.0 / ((q0 + q1) + q2);
    const double e1x = b.px - a.px, e1y = b.py - a.py, e1z = b.pz - a.pz, e2x = c.px - a.px, e2y = c.py - a.py, e2z = c.pz - a.pz;
    const double nx = e1y * e2z - e1z * e2y, ny = e1z * e2x - e1x * e2z, nz = e1x * e2y - e1y * e2x;
    const double nn = sqrt((nx * nx + ny * ny) + nz * nz);
    const double mx = ((a.px + b.px) + c.px) / 3.0, my = ((a.py + b.py) + c.py) / 3.0, mz = ((a.pz + b.pz) + c.pz) / 3.0;
    const double cl = sqrt((mx * mx + my * my) + mz * mz);
    double shade = v.ambient;
    if (nn > 0 && cl > 0)
        shade = fmin(1.0, v.ambient + fabs(((nx / nn) * (-mx / cl) + (ny / nn) * (-my / cl)) + (nz / nn) * (-mz / cl)));
    const double dm = trunc(d * 1000.0);
    depth_mm[i] = (uint16_t)(dm < 0 ? 0 : dm > 65535.0 ? 65535 : (int)dm);
    for (int ch = 0; ch < 3; ++ch) {
        const double val = (((q0 * (double)rgb[3 * ia + ch] + q1 * (double)rgb[3 * ib + ch]) + q2 * (double)rgb[3 * ic + ch]) * d) * shade;
        const double r = rint(val);  // round half to even, as numpy
        bgr[3 * i + (2 - ch)] = (uint8_t)(r < 0 ? 0 : r > 255 ? 255 : (int)r);
    }
}

}  // namespace hf6d
