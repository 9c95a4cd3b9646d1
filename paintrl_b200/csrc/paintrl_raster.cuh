// paintrl_raster.cuh -- load-time texel rasterisation on the GPU.
//
// Part.preprocess (bullet_paint_wrapper.py:604-618) walks the triangles of a side in bary_list order;
// BarycentricInterpolator.get_uv_pixels (:191-212) gives each triangle its three corner pixels and
// every pixel of the corners' bounding box whose normalised coordinate (u / W, v / H) lies inside the
// UV triangle, and `profile_dicts[side].update(pixel_dict)` lets a later triangle overwrite an
// earlier one (inside one triangle the interior value overwrites a corner's, and of equal corner
// pixels the later corner wins: dict literal at :198).  The owner of a pixel is therefore the maximum
// of  tri * 4 + kind  (kind 0/1/2 corner a/b/c, 3 interior) over everything that touches it, which is
// what `raster_owner_kernel` computes with one atomicMax per touch -- order-free, so the result is
// the sequential one.  `raster_position_kernel` then evaluates the owner's value for every texel
// (get_coordinate_in_barycentric, :220-224).
//
// Arithmetic (bit-exact against oracle_rasterize / the reference-derived 240x240 packs): 2-vector
// np.dot = fma(x1, y1, x0 * y0); everything else is rounded separately (-fmad=false); u / W is an
// IEEE division; Python round() = rint().
#pragma once
#include "paintrl_device.cuh"

namespace paintrl {

struct UvBary {       // BarycentricInterpolator(uva, uvb, uvc), bullet_paint_wrapper.py:123-134
    double ax, ay, v0x, v0y, v1x, v1y, d00, d01, d11, inv;
};

__device__ __forceinline__ double npdot2(double x0, double x1, double y0, double y1) { return fma(x1, y1, x0 * y0); }

__device__ __forceinline__ UvBary make_uv_bary(const double *uv) {
    UvBary b;
    b.ax = uv[0]; b.ay = uv[1];
    b.v0x = uv[2] - uv[0]; b.v0y = uv[3] - uv[1];
    b.v1x = uv[4] - uv[0]; b.v1y = uv[5] - uv[1];
    b.d00 = npdot2(b.v0x, b.v0y, b.v0x, b.v0y);
    b.d01 = npdot2(b.v0x, b.v0y, b.v1x, b.v1y);
    b.d11 = npdot2(b.v1x, b.v1y, b.v1x, b.v1y);
    const double denom = b.d00 * b.d11 - b.d01 * b.d01;
    b.inv = denom != 0.0 ? 1.0 / denom : 0.0;
    return b;
}

// _get_bary_coordinate + is_inside_triangle (:154-163, :183-185) of the pixel (u, v)
__device__ __forceinline__ bool uv_inside(const UvBary &b, int u, int v, int width, int height, double &bu, double &bv, double &bw) {
    const double px = (double)u / (double)width, py = (double)v / (double)height;
    const double v2x = px - b.ax, v2y = py - b.ay;
    const double d20 = npdot2(v2x, v2y, b.v0x, b.v0y), d21 = npdot2(v2x, v2y, b.v1x, b.v1y);
    bv = (b.d11 * d20 - b.d01 * d21) * b.inv;
    bw = (b.d00 * d21 - b.d01 * d20) * b.inv;
    bu = 1.0 - bv - bw;
    if (b.inv == 0.0) return false;
    return 0.0 <= bu && bu <= 1.0 && 0.0 <= bv && bv <= 1.0 && 0.0 <= bw && bw <= 1.0;
}

// _get_pixel_coordinate (:165-171)
__device__ __forceinline__ void pixel_coordinate(double u, double v, int width, int height, int &i, int &j) {
    const double ri = rint((double)width * u), rj = rint((double)height * v);
    i = ri < (double)(width - 1) ? (int)ri : width - 1;
    j = rj < (double)(height - 1) ? (int)rj : height - 1;
}

// One warp per triangle.  owner: [width * height] (index u * height + v), preset to -1.
// *bad is raised when a corner maps to a negative pixel coordinate (the reference would index the
// texture from its end there; not supported).
__global__ void __launch_bounds__(128)
raster_owner_kernel(const double *tri_uv, int n_tris, int width, int height, int *owner, int *bad) {
    const int t = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (t >= n_tris) return;
    const double *uv = tri_uv + (size_t)t * 6;
    int ci[3], cj[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) pixel_coordinate(__ldg(uv + 2 * k), __ldg(uv + 2 * k + 1), width, height, ci[k], cj[k]);
    if (min(min(ci[0], ci[1]), ci[2]) < 0 || min(min(cj[0], cj[1]), cj[2]) < 0) {
        if (lane == 0) atomicExch(bad, 1);
        return;
    }
    if (lane < 3) atomicMax(owner + (size_t)ci[lane] * height + cj[lane], t * 4 + lane);
    double uvr[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) uvr[k] = __ldg(uv + k);
    const UvBary b = make_uv_bary(uvr);
    const int x_min = min(min(ci[0], ci[1]), ci[2]), x_max = max(max(ci[0], ci[1]), ci[2]);
    const int y_min = min(min(cj[0], cj[1]), cj[2]), y_max = max(max(cj[0], cj[1]), cj[2]);
    const int ny = y_max - y_min + 1, total = (x_max - x_min + 1) * ny;
    for (int k = lane; k < total; k += 32) {
        const int u = x_min + k / ny, v = y_min + k % ny;
        double bu, bv, bw;
        if (uv_inside(b, u, v, width, height, bu, bv, bw)) atomicMax(owner + (size_t)u * height + v, t * 4 + 3);
    }
}

// One thread per texel of the compacted list `pix` (= u * height + v, ascending).
__global__ void __launch_bounds__(256)
raster_position_kernel(const double *tri_a, const double *tri_b, const double *tri_c, const double *tri_uv, int width, int height,
                       const int *owner, const int *pix, int n, int *ij_out, double *pos_out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int p = __ldg(pix + k), o = __ldg(owner + p);
    const int t = o >> 2, kind = o & 3;
    const int u = p / height, v = p % height;
    ij_out[2 * k] = u;
    ij_out[2 * k + 1] = v;
    const double *a = tri_a + (size_t)t * 3, *b = tri_b + (size_t)t * 3, *c = tri_c + (size_t)t * 3;
    double *out = pos_out + (size_t)k * 3;
    if (kind < 3) {
        const double *src = kind == 0 ? a : (kind == 1 ? b : c);
        out[0] = __ldg(src); out[1] = __ldg(src + 1); out[2] = __ldg(src + 2);
        return;
    }
    double uvr[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) uvr[q] = __ldg(tri_uv + (size_t)t * 6 + q);
    const UvBary ub = make_uv_bary(uvr);
    double bu, bv, bw;
    uv_inside(ub, u, v, width, height, bu, bv, bw);
#pragma unroll
    for (int q = 0; q < 3; ++q) out[q] = (bu * __ldg(a + q) + bv * __ldg(b + q)) + bw * __ldg(c + q);
}

// ---- silhouette scans (Part._get_exact_boundary, bullet_paint_wrapper.py:906-920) ------------------------------------
// A scan marches from `point` along the first principal axis (`proof`) in 1 mm steps, outwards (is_min: towards smaller
// values), and at every step tests a ray along the non-principal axis `npa` against the collision hull; it reports the
// coordinate of the first step whose ray misses.  The reference moves the ray's end points 1 further out at every step
// (start -= 1, end += 1, cumulatively rounded) -- kept.  One warp per scan, 32 consecutive steps per round (lane = step),
// every lane scanning all hull planes for its own ray with shim S1's slab arithmetic (products and sums rounded one by
// one: the translation unit is compiled with -fmad=false), the first missing lane found by ballot.
__global__ void __launch_bounds__(128)
silhouette_march_kernel(const double4 *planes, int n_planes, const double *points, const signed char *is_min, int n_scans, int proof,
                        int npa, int steps_range, double *boundary_out, signed char *found_out) {
    const int scan = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (scan >= n_scans) return;
    const double p[3] = {points[3 * scan], points[3 * scan + 1], points[3 * scan + 2]};
    const double step = is_min[scan] ? -1e-3 : 1e-3;
    double s_base = p[npa], e_base = p[npa];             // the ray's end points after the steps of the previous rounds
    for (int i0 = 0; i0 < steps_range; i0 += 32) {
        const int i = i0 + lane;
        double s = s_base, e = e_base;
        for (int k = 0; k <= lane; ++k) { s -= 1.0; e += 1.0; }
        for (int k = 0; k < 32; ++k) { s_base -= 1.0; e_base += 1.0; }
        const double bound = p[proof] + (double)i * step;
        double frm[3] = {p[0], p[1], p[2]}, to[3] = {p[0], p[1], p[2]};
        frm[proof] = bound; to[proof] = bound;
        frm[npa] = s; to[npa] = e;
        const double d0 = to[0] - frm[0], d1 = to[1] - frm[1], d2 = to[2] - frm[2];
        double t_in = -INFINITY, t_out = INFINITY;
        bool parallel_out = false;
        for (int q = 0; q < n_planes; ++q) {
            const double2 pa = __ldg(reinterpret_cast<const double2 *>(planes + q)), pb = __ldg(reinterpret_cast<const double2 *>(planes + q) + 1);
            const double4 pl = make_double4(pa.x, pa.y, pb.x, pb.y);
            const double den = (pl.x * d0 + pl.y * d1) + pl.z * d2;
            const double num = pl.w - ((pl.x * frm[0] + pl.y * frm[1]) + pl.z * frm[2]);
            if (den == 0.0) { parallel_out |= num < 0.0; continue; }
            const double t = num / den;
            if (den < 0.0) t_in = fmax(t_in, t); else t_out = fmin(t_out, t);
        }
        const bool hit = !parallel_out && t_in <= t_out && 0.0 <= t_in && t_in <= 1.0;
        const unsigned miss = __ballot_sync(0xffffffffu, i < steps_range && !hit);
        if (miss) {
            const int src = __ffs(miss) - 1;
            const double b = __shfl_sync(0xffffffffu, bound, src);
            if (lane == 0) { boundary_out[scan] = b; found_out[scan] = 1; }
            return;
        }
    }
    if (lane == 0) { boundary_out[scan] = 0.0; found_out[scan] = 0; }
}

}  // namespace paintrl
