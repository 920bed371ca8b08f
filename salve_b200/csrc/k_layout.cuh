// k_layout.cuh -- the "layout" modality: room polygon + window / door / opening strokes rasterised into a BEV image.
//
// Replaces the cv2 calls behind rasterize_single_layout (reference salve/utils/bev_rendering_utils.py:101-156):
//   cv2.fillPoly(image, [points], color)                                   :159-181   (draw_polygon_cv2)
//   cv2.line(image, p1, p2, color, thickness, lineType=cv2.LINE_AA)        :220-251   (draw_polyline_cv2)
//   np.flipud                                                              :155
//
// Polygon: cv2.fillPoly (8-connected) = the polygon's edges drawn as Bresenham lines (cv::LineIterator, left to right, the step
// rule `err < 0`) + an even-odd scan-line fill whose span between two consecutive edge crossings xa <= xb of a row is
// floor(xa + 1/2) .. floor(xb), crossings taken on rows y0 <= y < y1 of each non-horizontal edge.  Reproduced here with exact
// rational crossings: bit-identical to cv2 (4.13) for polygons whose vertices lie inside the image (tests/test_layout_*); a
// vertex outside the image makes cv2 re-derive the edge from its integer-clipped end points, which this kernel does not imitate
// (it fills the exact polygon clipped to the image; the differing pixels hug the clipped edges and are reported by the test).
// Strokes: cv2 builds a thick anti-aliased line from a convex quad + two 12-gon end caps, each outlined with its 1-px
// anti-aliased line filter (table driven).  Here a stroke is the capsule around the segment with radius thickness/2 + 3/4 and a
// one-pixel linear edge ramp: identical to cv2 in the stroke's interior and outside it, within the stated tolerance on the
// anti-aliased rim (tests report the differing fraction).
#pragma once
#include "bev_common.cuh"

namespace bev {

constexpr int LAYOUT_NT = 256;
constexpr int LAYOUT_MAX_POLY = 128;  // vertices of the room polygon
// descriptor of one image (int32 words): header, polygon vertices (x, y), strokes (x0, y0, x1, y1, rgb, thickness)
enum { LD_NPOLY = 0, LD_NSEG, LD_POLY_RGB, LD_FLIP, LD_HDR };

struct LayoutArgs {
    const int32_t* desc;       // all descriptors
    const long long* offset;   // [n_img + 1] word offsets into desc
    uint8_t* out; size_t out_stride;
    const uint8_t* init;       // optional initial image(s) (same stride), else zero
    int32_t h, w;
};

__device__ __forceinline__ long long floor_div(long long a, long long b) {  // b > 0
    long long q = a / b;
    return (a % b != 0 && a < 0) ? q - 1 : q;
}

__global__ void __launch_bounds__(LAYOUT_NT) layout_raster_kernel(LayoutArgs A) {
    const int img = blockIdx.x;
    const int h = A.h, w = A.w;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int32_t* D = A.desc + A.offset[img];
    const int n_poly = D[LD_NPOLY], n_seg = D[LD_NSEG];
    const uint32_t poly_rgb = (uint32_t)D[LD_POLY_RGB];
    const bool flip = D[LD_FLIP] != 0;
    const int32_t* pv = D + LD_HDR;
    const int32_t* sg = pv + 2 * n_poly;
    uint8_t* out = A.out + (size_t)img * A.out_stride;
    auto px = [&](int x, int y) -> uint8_t* { return out + ((size_t)(flip ? h - 1 - y : y) * w + x) * 3; };

    // ---- 0. background
    const size_t nbytes = (size_t)h * w * 3;
    if (A.init) {
        const uint8_t* src = A.init + (size_t)img * A.out_stride;
        for (size_t i = tid; i < nbytes; i += LAYOUT_NT) {
            const size_t p = i / 3, c = i - p * 3, y = p / w, x = p - y * w;
            px((int)x, (int)y)[c] = src[i];
        }
    } else {
        for (size_t i = tid; i < nbytes; i += LAYOUT_NT) out[i] = 0;
    }
    __syncthreads();

    __shared__ long long w_num[LAYOUT_NT / 32][LAYOUT_MAX_POLY], w_den[LAYOUT_NT / 32][LAYOUT_MAX_POLY];
    __shared__ int w_x1[LAYOUT_NT / 32][LAYOUT_MAX_POLY], w_x2[LAYOUT_NT / 32][LAYOUT_MAX_POLY];
    const uint8_t pr = (uint8_t)(poly_rgb & 0xFF), pg = (uint8_t)((poly_rgb >> 8) & 0xFF), pb = (uint8_t)(poly_rgb >> 16);

    if (n_poly >= 1) {
        // ---- 1. scan-line fill, a warp per row
        for (int y = warp; y < h; y += LAYOUT_NT / 32) {
            // crossings of this row: lanes take edges (v[i-1] -> v[i]); compacted into the warp's arrays
            int n = 0;
            for (int e0 = 0; e0 < n_poly; e0 += 32) {
                const int e = e0 + lane;
                bool hit = false;
                long long num = 0, den = 1;
                if (e < n_poly) {
                    const int i0 = e == 0 ? n_poly - 1 : e - 1;
                    int ax = pv[2 * i0], ay = pv[2 * i0 + 1], bx = pv[2 * e], by = pv[2 * e + 1];
                    if (ay != by) {
                        if (ay > by) { const int tx = ax, ty = ay; ax = bx; ay = by; bx = tx; by = ty; }
                        if (ay <= y && y < by) { hit = true; den = by - ay; num = (long long)ax * den + (long long)(y - ay) * (bx - ax); }
                    }
                }
                const unsigned m = __ballot_sync(0xffffffffu, hit);
                if (hit) { const int k = n + __popc(m & ((1u << lane) - 1u)); w_num[warp][k] = num; w_den[warp][k] = den; }
                n += __popc(m);
            }
            __syncwarp();
            // rank every crossing (exact comparison of the rationals, ties by index), then the spans of the even-odd pairs
            for (int k0 = 0; k0 < n; k0 += 32) {
                const int k = k0 + lane;
                if (k < n) {
                    const long long nk = w_num[warp][k], dk = w_den[warp][k];
                    int rank = 0;
                    for (int j = 0; j < n; j++) {
                        const long long l = w_num[warp][j] * dk, r = nk * w_den[warp][j];  // x_j < x_k  <=>  num_j * den_k < num_k * den_j
                        rank += (l < r || (l == r && j < k)) ? 1 : 0;
                    }
                    if (rank & 1) w_x2[warp][rank >> 1] = (int)floor_div(nk, dk);               // right end: floor(x)
                    else w_x1[warp][rank >> 1] = (int)floor_div(2 * nk + dk, 2 * dk);            // left end: floor(x + 1/2)
                }
            }
            __syncwarp();
            for (int p = 0; p < n / 2; p++) {
                const int x1 = max(w_x1[warp][p], 0), x2 = min(w_x2[warp][p], w - 1);
                for (int x = x1 + lane; x <= x2; x += 32) { uint8_t* o = px(x, y); o[0] = pr; o[1] = pg; o[2] = pb; }
            }
            __syncwarp();
        }
        __syncthreads();
        // ---- 2. the polygon's edges as 8-connected Bresenham lines (cv::LineIterator, left to right), a thread per edge
        for (int e = tid; e < n_poly; e += LAYOUT_NT) {
            const int i0 = e == 0 ? n_poly - 1 : e - 1;
            int x0 = pv[2 * i0], y0 = pv[2 * i0 + 1], x1 = pv[2 * e], y1 = pv[2 * e + 1];
            if (x0 > x1) { const int tx = x0, ty = y0; x0 = x1; y0 = y1; x1 = tx; y1 = ty; }
            const int dx = x1 - x0, ady = abs(y1 - y0), sy = y1 < y0 ? -1 : 1;
            int x = x0, y = y0;
            if (dx >= ady) {
                int err = dx - 2 * ady;
                for (int i = 0; i <= dx; i++) {
                    if (x >= 0 && x < w && y >= 0 && y < h) { uint8_t* o = px(x, y); o[0] = pr; o[1] = pg; o[2] = pb; }
                    if (err < 0) { err += 2 * dx - 2 * ady; y += sy; } else err -= 2 * ady;
                    x++;
                }
            } else {
                int err = ady - 2 * dx;
                for (int i = 0; i <= ady; i++) {
                    if (x >= 0 && x < w && y >= 0 && y < h) { uint8_t* o = px(x, y); o[0] = pr; o[1] = pg; o[2] = pb; }
                    if (err < 0) { err += 2 * ady - 2 * dx; x++; } else err -= 2 * dx;
                    y += sy;
                }
            }
        }
        __syncthreads();
    }

    // ---- 3. strokes, in order (a later stroke is blended over an earlier one)
    for (int s = 0; s < n_seg; s++) {
        const int32_t* q = sg + 6 * s;
        const float ax = (float)q[0], ay = (float)q[1], bx = (float)q[2], by = (float)q[3];
        const uint32_t rgb = (uint32_t)q[4];
        const float R = 0.5f * (float)q[5] + 0.75f;
        const int xlo = max((int)floorf(fminf(ax, bx) - R - 1.f), 0), xhi = min((int)ceilf(fmaxf(ax, bx) + R + 1.f), w - 1);
        const int ylo = max((int)floorf(fminf(ay, by) - R - 1.f), 0), yhi = min((int)ceilf(fmaxf(ay, by) + R + 1.f), h - 1);
        const int bw = xhi - xlo + 1, bh = yhi - ylo + 1;
        const float dx = bx - ax, dy = by - ay, L2 = dx * dx + dy * dy;
        if (bw > 0 && bh > 0) {
            for (int i = tid; i < bw * bh; i += LAYOUT_NT) {
                const int y = ylo + i / bw, x = xlo + i % bw;
                float t = L2 > 0.f ? ((x - ax) * dx + (y - ay) * dy) / L2 : 0.f;
                t = fminf(fmaxf(t, 0.f), 1.f);
                const float ex = x - (ax + t * dx), ey = y - (ay + t * dy);
                const float a = fminf(fmaxf(R - sqrtf(ex * ex + ey * ey) + 0.5f, 0.f), 1.f);
                const int ai = __float2int_rn(a * 255.f);
                if (ai == 0) continue;
                uint8_t* o = px(x, y);
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const int col = (int)((rgb >> (8 * c)) & 0xFF), bg = (int)o[c];
                    o[c] = (uint8_t)((bg * (255 - ai) + col * ai + 127) / 255);
                }
            }
        }
        __syncthreads();
    }
}

}  // namespace bev
