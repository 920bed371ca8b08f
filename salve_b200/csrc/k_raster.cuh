// k_raster.cuh -- one CTA per BEV image: barycentric rasterisation of the Delaunay triangles,
// fused with the hallucination mask and the vertical flip.
//
// Replaces scipy's LinearNDInterpolator evaluation + the uint8 store of
// salve/utils/interpolation_utils.py:46-53, the mask multiply of :114-121 and np.flipud
// (bev_rendering_utils.py:319).  Barycentric weights are exact integers (twice the sub-triangle
// areas), so value = floor(sum(w_i*c_i) / sum(w_i)) is the truncation of the exact interpolant;
// pixels on a shared edge get the same value from both triangles, so duplicate writes are benign
// and the image is deterministic.  Site pixels were already written by the site kernel.
#pragma once
#include "bev_common.cuh"

namespace bev {

constexpr int RASTER_NT = 512;
constexpr int RASTER_SMALL_AREA = 48;
constexpr int RASTER_QUEUE = 4096;

struct RasterArgs {
    GridParams G;
    const Tri* tris; size_t tris_stride;
    const uint32_t* color; size_t color_stride;
    const uint32_t* keep; size_t bits_stride;
    const ImgHeader* headers;
    int32_t* counts;
    int32_t* status;  // optional: set to SALVE_BEV_IMG_COLLINEAR when the mesh has no real triangle
    uint8_t* out; size_t out_stride;
    uint8_t* hull; size_t hull_stride;  // optional (tap): 1 inside the closed hull
    int32_t raw_mode;
};

struct TriSetup {
    int ax, ay, bx, by, cx, cy;
    int x0, y0, bw, bh;
    uint32_t ca, cb, cc;
    uint32_t A2;
};

__device__ __forceinline__ void raster_pixels(const RasterArgs& A, const TriSetup& s, const uint32_t* keep, uint8_t* out, uint8_t* hull,
                                              int start, int stride) {
    const int w = A.G.grid_w, h = A.G.grid_h, wpr = A.G.wpr;
    const int bw1 = s.bw + 1, area = bw1 * (s.bh + 1);
    for (int p = start; p < area; p += stride) {
        const int dy = p / bw1, dx = p - dy * bw1;
        const int x = s.x0 + dx, y = s.y0 + dy;
        const int wa = (s.cx - s.bx) * (y - s.by) - (s.cy - s.by) * (x - s.bx);
        const int wb = (s.ax - s.cx) * (y - s.cy) - (s.ay - s.cy) * (x - s.cx);
        const int wc = (s.bx - s.ax) * (y - s.ay) - (s.by - s.ay) * (x - s.ax);
        if ((wa | wb | wc) < 0) continue;
        if (hull) hull[y * w + x] = 1;
        if ((uint32_t)wa == s.A2 || (uint32_t)wb == s.A2 || (uint32_t)wc == s.A2) continue;  // a site: already written
        size_t o;
        if (A.raw_mode) o = ((size_t)y * w + x) * 3;
        else {
            if (!((keep[y * wpr + (x >> 5)] >> (x & 31)) & 1u)) continue;
            o = ((size_t)(h - 1 - y) * w + x) * 3;
        }
        const uint32_t ua = (uint32_t)wa, ub = (uint32_t)wb, uc = (uint32_t)wc;
        out[o + 0] = (uint8_t)((ua * (s.ca & 0xFF) + ub * (s.cb & 0xFF) + uc * (s.cc & 0xFF)) / s.A2);
        out[o + 1] = (uint8_t)((ua * ((s.ca >> 8) & 0xFF) + ub * ((s.cb >> 8) & 0xFF) + uc * ((s.cc >> 8) & 0xFF)) / s.A2);
        out[o + 2] = (uint8_t)((ua * ((s.ca >> 16) & 0xFF) + ub * ((s.cb >> 16) & 0xFF) + uc * ((s.cc >> 16) & 0xFF)) / s.A2);
    }
}

__device__ __forceinline__ bool tri_setup(const RasterArgs& A, const Tri& T, const uint32_t* color, TriSetup& s, bool want_all) {
    const uint32_t va = tri_v(T, 0), vb = tri_v(T, 1), vc = tri_v(T, 2);
    if (va == GHOST || vb == GHOST || vc == GHOST) return false;
    s.ax = vcol(va); s.ay = vrow(va); s.bx = vcol(vb); s.by = vrow(vb); s.cx = vcol(vc); s.cy = vrow(vc);
    s.x0 = min(s.ax, min(s.bx, s.cx)); s.y0 = min(s.ay, min(s.by, s.cy));
    s.bw = max(s.ax, max(s.bx, s.cx)) - s.x0; s.bh = max(s.ay, max(s.by, s.cy)) - s.y0;
    if (!want_all && s.bw <= 1 && s.bh <= 1) return false;  // unit cell: lattice points are its own vertices
    s.A2 = (uint32_t)((s.bx - s.ax) * (s.cy - s.ay) - (s.by - s.ay) * (s.cx - s.ax));
    const int w = A.G.grid_w;
    s.ca = color[s.ay * w + s.ax]; s.cb = color[s.by * w + s.bx]; s.cc = color[s.cy * w + s.cx];
    return true;
}

__global__ void __launch_bounds__(RASTER_NT) raster_kernel(RasterArgs A) {
    const int img = blockIdx.x;
    const ImgHeader hd = A.headers[img];
    const int nt = hd.n_tris;
    if (nt <= 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const Tri* tris = A.tris + (size_t)img * A.tris_stride;
    const uint32_t* color = A.color + (size_t)img * A.color_stride;
    const uint32_t* keep = A.keep + (size_t)img * A.bits_stride;
    uint8_t* out = A.out + (size_t)img * A.out_stride;
    uint8_t* hull = A.hull ? A.hull + (size_t)img * A.hull_stride : nullptr;

    __shared__ uint32_t s_queue[RASTER_QUEUE];
    __shared__ int s_nq, s_real;
    if (tid == 0) { s_nq = 0; s_real = 0; }
    __syncthreads();
    int n_real = 0;
    for (int t = tid; t < nt; t += RASTER_NT) {
        const Tri T = ld_tri(tris + t);
        n_real += (tri_v(T, 0) != GHOST && tri_v(T, 1) != GHOST && tri_v(T, 2) != GHOST);
        TriSetup s;
        if (!tri_setup(A, T, color, s, hull != nullptr)) continue;
        if ((s.bw + 1) * (s.bh + 1) > RASTER_SMALL_AREA) {
            const int q = atomicAdd(&s_nq, 1);
            if (q < RASTER_QUEUE) { s_queue[q] = (uint32_t)t; continue; }
        }
        raster_pixels(A, s, keep, out, hull, 0, 1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n_real += __shfl_xor_sync(0xffffffffu, n_real, o);
    if (lane == 0 && n_real) atomicAdd(&s_real, n_real);
    __syncthreads();
    const int nq = min(s_nq, RASTER_QUEUE);
    for (int q = warp; q < nq; q += RASTER_NT / 32) {  // large triangles: one warp each
        const Tri T = ld_tri(tris + s_queue[q]);
        TriSetup s;
        if (!tri_setup(A, T, color, s, true)) continue;
        raster_pixels(A, s, keep, out, hull, lane, 32);
    }
    if (tid == 0) {
        A.counts[img * 8 + 5] = s_real;
        if (s_real == 0 && A.status) A.status[img] = 3;  // all sites on one oblique line
    }
}

// ---- stand-alone hallucination mask on explicit images (remove_hallucinated_content) ------------
// nonempty(y,x) = uint8(r*g*b) != 0 of `sparse`; out = interp where any nonempty within the K/2 Chebyshev ball.
__global__ void halluc_nonempty_kernel(const uint8_t* __restrict__ sparse, int h, int w, uint8_t* __restrict__ ne) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= h * w) return;
    const uint32_t r = sparse[p * 3], g = sparse[p * 3 + 1], b = sparse[p * 3 + 2];
    ne[p] = ((r * g * b) & 0xFFu) != 0u;
}
__global__ void halluc_rowdilate_kernel(const uint8_t* __restrict__ ne, int h, int w, int rad, uint8_t* __restrict__ tmp) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= h * w) return;
    const int y = p / w, x = p - y * w;
    uint8_t o = 0;
    for (int xx = max(x - rad, 0); xx <= min(x + rad, w - 1); xx++) o |= ne[y * w + xx];
    tmp[p] = o;
}
__global__ void halluc_apply_kernel(const uint8_t* __restrict__ tmp, const uint8_t* __restrict__ interp, int h, int w, int rad,
                                    uint8_t* __restrict__ out) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= h * w) return;
    const int y = p / w, x = p - y * w;
    uint8_t o = 0;
    for (int yy = max(y - rad, 0); yy <= min(y + rad, h - 1); yy++) o |= tmp[yy * w + x];
    out[p * 3 + 0] = o ? interp[p * 3 + 0] : 0;
    out[p * 3 + 1] = o ? interp[p * 3 + 1] : 0;
    out[p * 3 + 2] = o ? interp[p * 3 + 2] : 0;
}

// ---- taps -----------------------------------------------------------------------------------------
__global__ void tap_tris_kernel(const Tri* __restrict__ tris, int nt, int grid_w, int32_t* __restrict__ outv) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nt) return;
    const Tri T = ld_tri(tris + t);
    for (int i = 0; i < 3; i++) {
        const uint32_t v = tri_v(T, i);
        outv[t * 3 + i] = (v == GHOST) ? -1 : vrow(v) * grid_w + vcol(v);
    }
}

// points -> key grid for the generic interp_dense path: key = index + 1 (slice 0)
__global__ void points_to_keys_kernel(const long long* __restrict__ xy, const double* __restrict__ values, long long n, int grid_h,
                                      int grid_w, uint32_t* __restrict__ keygrid, uint8_t* __restrict__ rgb_u8, int* __restrict__ err) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long x = xy[i * 2], y = xy[i * 2 + 1];
    for (int ch = 0; ch < 3; ch++) rgb_u8[i * 3 + ch] = (uint8_t)(long long)values[i * 3 + ch];
    if (x < 0 || y < 0 || x >= grid_w || y >= grid_h) { atomicExch(err, 1); return; }
    atomicMax(keygrid + y * grid_w + x, (uint32_t)i + 1u);
}

}  // namespace bev
