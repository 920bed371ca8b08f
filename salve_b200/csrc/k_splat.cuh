// k_splat.cuh -- back-projection + pose + height crop + BEV splat (deterministic atomicMax),
// and the order-preserving crop/compaction kernels behind get_xyzrgb_from_depth.
//
// Reference arithmetic reproduced bit-for-bit (float64 unless noted):
//   depth  d  = float32(u16) * float32(0.001)                     bev_rendering_utils.py:367
//   xyz       = d * (cos_phi*cos_theta, cos_phi*sin_theta, -sin_phi)   hohonet_pano_utils.py:27-43, :392
//   rows      v in [crop, H-crop)                                 bev_rendering_utils.py:397-401
//   band      lo < z <= hi                                        bev_rendering_utils.py:408-413
//   xy <- xy @ rotmat2d(-90).T                                    bev_rendering_utils.py:443-446
//   xy <- xy @ R32.T + t32*1.5 (pano 1 only; t*1.5 in float32)    bev_rendering_utils.py:448-451
//   bbox      xmin <= x <= xmax, ymin <= y <= ymax                bev_rendering_utils.py:38-45
//   pixel     rint((x - xmin) * px_per_m)                         sim2.py:157-160, bev_rendering_utils.py:287
//   z-order   winner = max (z-slice, point index)                 zorder_utils.py:49-65
// numpy's (N,2)@(2,2) evaluates out_j = fma(p1, M[j][1], round(p0*M[j][0])) (measured),
// which is what rot_pose() spells out with explicit _rn intrinsics so nvcc cannot re-associate.
#pragma once
#include "bev_common.cuh"

namespace bev {

struct SplatJob {
    int32_t pano_slot;
    int32_t posed;
    float R[4];
    float t[2];
    int32_t img_floor;  // image index within the chunk, or -1
    int32_t img_ceil;   // image index within the chunk, or -1
};

struct SplatParams {
    int32_t H, W, crop_rows;
    int32_t rows_per_thread;  // pano rows one thread of splat_pano_kernel walks (a multiple of SPLAT_BATCH)
    float depth_scale;
    double xmin, ymin, xmax, ymax, px_per_m;
    double a_lo, a_hi;  // band A ("floor")   keeps a_lo < z <= a_hi; reference (-inf, -1.0]
    double b_lo, b_hi;  // band B ("ceiling") keeps b_lo < z <= b_hi; reference (0.5, +inf)
    int32_t grid_w, g;
    const double* cos_phi;
    const double* neg_sin_phi;
    const double* cos_theta;
    const double* sin_theta;
    const uint16_t* const* depth;  // per-slot device pointers
};

constexpr double C90 = 6.123233995736766e-17;  // np.cos(np.deg2rad(-90)), rotation_utils.py:14-29

// z-slice of zorder_utils.py for zmin=-2, zmax=2, num_slices=4: planes are exactly -2,-1,0,1,2.
__device__ __forceinline__ int z_slice4(double z) {
    if (!(z >= -2.0) || !(z < 2.0)) return -1;
    return (z >= 1.0) ? 3 : (z >= 0.0) ? 2 : (z >= -1.0) ? 1 : 0;
}

// xy in the HoHoNet frame -> ZInD frame -> (optionally) pano 2's frame.
__device__ __forceinline__ void rot_pose(double x, double y, bool posed, const float* R, double tx, double ty, double& ox, double& oy) {
    double x1 = __dadd_rn(y, __dmul_rn(x, C90));   // fma(y, 1.0, round(x*c90))
    double y1 = __fma_rn(y, C90, -x);              // fma(y, c90, round(x*-1.0))
    if (posed) {
        double x2 = __dadd_rn(__fma_rn(y1, (double)R[1], __dmul_rn(x1, (double)R[0])), tx);
        double y2 = __dadd_rn(__fma_rn(y1, (double)R[3], __dmul_rn(x1, (double)R[2])), ty);
        x1 = x2; y1 = y2;
    }
    ox = x1; oy = y1;
}

__device__ __forceinline__ bool bbox_pixel(const SplatParams& P, double x, double y, int& row, int& col) {
    if (!(P.xmin <= x && x <= P.xmax && P.ymin <= y && y <= P.ymax)) return false;
    col = (int)rint(__dmul_rn(__dadd_rn(x, -P.xmin), P.px_per_m));
    row = (int)rint(__dmul_rn(__dadd_rn(y, -P.ymin), P.px_per_m));
    return true;
}

// One thread = 4 consecutive pano columns (one 8-byte depth load per row) x P.rows_per_thread consecutive rows: the column factors
// cos/sin theta stay in registers across the rows, and the depth loads of a batch of rows are issued together.
// grid = (ceil(rows / P.rows_per_thread) * ceil(W / 1024), n_jobs), block = 256.
// Reads only the depth map (2 B/px); colours are gathered later for winners only.
// Rows per thread are chosen per launch (splat_rows_for): 32 for the batches of a building (measured 1.26 -> 1.10 ms per 679 pano
// passes against 8; the column factors and the job's pose are loaded once per 32 rows), fewer when there are too few jobs to fill
// the GPU.  Depth loads are issued two rows at a time (2 beats 4 and 8: fewer live registers, no spills at 80).
#ifndef SPLAT_BATCH_DEF
#define SPLAT_BATCH_DEF 2
#endif
constexpr int SPLAT_BATCH = SPLAT_BATCH_DEF;
__host__ inline int splat_rows_for(size_t n_jobs) { return n_jobs >= 128 ? 32 : (n_jobs >= 32 ? 16 : 8); }

#ifndef SPLAT_CTAS
#define SPLAT_CTAS 3
#endif
__global__ void __launch_bounds__(256, SPLAT_CTAS) splat_pano_kernel(SplatParams P, const SplatJob* __restrict__ jobs,
                                                         uint32_t* __restrict__ keygrid_base, size_t keygrid_stride,
                                                         int32_t* __restrict__ counts /* [n_img][8] */) {
    const SplatJob job = jobs[blockIdx.y];
    const int rows = P.H - 2 * P.crop_rows;
    const int ncg = (P.W + 1023) >> 10;  // column groups of 1024 columns
    const int cg = blockIdx.x % ncg, rg = blockIdx.x / ncg;
    const int u0 = ((cg << 8) + threadIdx.x) << 2;
    int n_crop_f = 0, n_crop_c = 0, n_box_f = 0, n_box_c = 0;
    if (u0 < P.W) {
        double ct[4], st[4];
#pragma unroll
        for (int k = 0; k < 4; k++) { ct[k] = __ldg(P.cos_theta + u0 + k); st[k] = __ldg(P.sin_theta + u0 + k); }
        const double tx = (double)__fmul_rn(job.t[0], 1.5f), ty = (double)__fmul_rn(job.t[1], 1.5f);
        uint32_t* kg_f = job.img_floor >= 0 ? keygrid_base + (size_t)job.img_floor * keygrid_stride : nullptr;
        uint32_t* kg_c = job.img_ceil >= 0 ? keygrid_base + (size_t)job.img_ceil * keygrid_stride : nullptr;
        const uint16_t* dbase = P.depth[job.pano_slot];
        const int r0 = rg * P.rows_per_thread;
#pragma unroll 1
        for (int rb = 0; rb < P.rows_per_thread; rb += SPLAT_BATCH) {
            uint2 raw[SPLAT_BATCH];
#pragma unroll
            for (int j = 0; j < SPLAT_BATCH; j++) {
                const int rr = r0 + rb + j;
                raw[j] = (rr < rows) ? __ldg(reinterpret_cast<const uint2*>(dbase + (size_t)(P.crop_rows + rr) * P.W + u0)) : make_uint2(0u, 0u);
            }
#pragma unroll
            for (int j = 0; j < SPLAT_BATCH; j++) {
                const int rr = r0 + rb + j;
                if (rr >= rows) break;
                const int v = P.crop_rows + rr;
                const uint32_t d16[4] = {raw[j].x & 0xFFFFu, raw[j].x >> 16, raw[j].y & 0xFFFFu, raw[j].y >> 16};
                const double cphi = __ldg(P.cos_phi + v), sz = __ldg(P.neg_sin_phi + v);
                // targets of the 4 pixels; consecutive pano pixels often land on the same BEV pixel (near the camera up to 4 of
                // them): merged here (max of the keys) they cost one atomic instead of several on the same address
                int pix[4]; uint32_t key[4]; bool ff[4], cf[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    ff[k] = false; cf[k] = false; pix[k] = -1; key[k] = 0u;
                    const double d = (double)__fmul_rn((float)d16[k], P.depth_scale);
                    const double z = __dmul_rn(d, sz);
                    const bool is_f = (z > P.a_lo && z <= P.a_hi);
                    const bool is_c = (z > P.b_lo && z <= P.b_hi);
                    if (is_f) n_crop_f++;
                    if (is_c) n_crop_c++;
                    const bool do_f = is_f && kg_f != nullptr, do_c = is_c && kg_c != nullptr;
                    if (!do_f && !do_c) continue;
                    const double x = __dmul_rn(d, __dmul_rn(cphi, ct[k]));
                    const double y = __dmul_rn(d, __dmul_rn(cphi, st[k]));
                    double wx, wy;
                    rot_pose(x, y, job.posed != 0, job.R, tx, ty, wx, wy);
                    int row, col;
                    if (!bbox_pixel(P, wx, wy, row, col)) continue;
                    if (do_f) n_box_f++;
                    if (do_c) n_box_c++;
                    const int sl = z_slice4(z);
                    if (sl < 0) continue;
                    key[k] = (((uint32_t)sl << KEY_IDX_BITS) | (uint32_t)(v * P.W + u0 + k)) + 1u;
                    pix[k] = row * P.grid_w + col;
                    ff[k] = do_f; cf[k] = do_c;
                }
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    if (pix[k] >= 0 && pix[k] == pix[k + 1] && ff[k] == ff[k + 1] && cf[k] == cf[k + 1]) {
                        key[k + 1] = max(key[k + 1], key[k]);
                        pix[k] = -1;
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (pix[k] < 0) continue;
                    if (ff[k]) atomicMax(kg_f + pix[k], key[k]);
                    if (cf[k]) atomicMax(kg_c + pix[k], key[k]);
                }
            }
        }
    }
    if (counts != nullptr) {
        // warp-aggregated counters
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            n_crop_f += __shfl_xor_sync(0xffffffffu, n_crop_f, o);
            n_crop_c += __shfl_xor_sync(0xffffffffu, n_crop_c, o);
            n_box_f += __shfl_xor_sync(0xffffffffu, n_box_f, o);
            n_box_c += __shfl_xor_sync(0xffffffffu, n_box_c, o);
        }
        if ((threadIdx.x & 31) == 0) {
            if (job.img_floor >= 0) {
                if (n_crop_f) atomicAdd(counts + job.img_floor * 8 + 0, n_crop_f);
                if (n_box_f) atomicAdd(counts + job.img_floor * 8 + 1, n_box_f);
            }
            if (job.img_ceil >= 0) {
                if (n_crop_c) atomicAdd(counts + job.img_ceil * 8 + 0, n_crop_c);
                if (n_box_c) atomicAdd(counts + job.img_ceil * 8 + 1, n_box_c);
            }
        }
    }
}

// ---- arbitrary cloud (render_bev_image on an (N,6) float64 array) -----------------------------
// Also converts the cloud's colours to the u8 triple the reference stores in the sparse image
// (rgb*255 truncated to uint8, bev_rendering_utils.py:266,307-308).
__global__ void __launch_bounds__(256) splat_cloud_kernel(SplatParams P, const double* __restrict__ xyzrgb, long long n,
                                                          uint32_t* __restrict__ keygrid, uint8_t* __restrict__ rgb_u8,
                                                          int32_t* __restrict__ counts) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int inside = 0;
    if (i < n) {
        const double* p = xyzrgb + i * 6;
        const double x = p[0], y = p[1], z = p[2];
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            double c = __dmul_rn(p[3 + ch], 255.0);
            rgb_u8[i * 3 + ch] = (uint8_t)(long long)c;  // truncation, wraps like numpy's f64->u8 cast
        }
        int row, col;
        if (bbox_pixel(P, x, y, row, col)) {
            inside = 1;
            const int sl = z_slice4(z);
            if (sl >= 0) atomicMax(keygrid + row * P.grid_w + col, (((uint32_t)sl << KEY_IDX_BITS) | (uint32_t)i) + 1u);
        }
    }
    const unsigned b = __ballot_sync(0xffffffffu, inside);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(counts + 1, __popc(b));
    if (i == 0) atomicAdd(counts + 0, (int)n);
}

// ---- crop + stream compaction (get_xyzrgb_from_depth) ------------------------------------------
// Pass 1: per-block survivor counts (warp ballot + popc).  Pass 2 (after an exclusive scan of the
// block counts): each warp re-evaluates, takes its offset from ballot prefix, writes rows in pano
// raster order -- the order numpy's boolean indexing produces.
constexpr int COMPACT_BLOCK = 256;

__device__ __forceinline__ bool crop_point(const SplatParams& P, const uint16_t* depth, int idx, double lo, double hi, double& x,
                                           double& y, double& z, int& src) {
    const int rows = P.H - 2 * P.crop_rows;
    if (idx >= rows * P.W) return false;
    const int v = P.crop_rows + idx / P.W, u = idx % P.W;
    src = v * P.W + u;
    const double d = (double)__fmul_rn((float)depth[src], P.depth_scale);
    const double cphi = P.cos_phi[v];
    z = __dmul_rn(d, P.neg_sin_phi[v]);
    if (!(z > lo && z <= hi)) return false;
    x = __dmul_rn(d, __dmul_rn(cphi, P.cos_theta[u]));
    y = __dmul_rn(d, __dmul_rn(cphi, P.sin_theta[u]));
    return true;
}

__global__ void __launch_bounds__(COMPACT_BLOCK) crop_count_kernel(SplatParams P, const uint16_t* __restrict__ depth, double lo,
                                                                   double hi, int32_t* __restrict__ block_counts) {
    __shared__ int warp_cnt[COMPACT_BLOCK / 32];
    const int idx = blockIdx.x * COMPACT_BLOCK + threadIdx.x;
    double x, y, z; int src;
    const bool keep = crop_point(P, depth, idx, lo, hi, x, y, z, src);
    const unsigned b = __ballot_sync(0xffffffffu, keep);
    if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5] = __popc(b);
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int w = 0; w < COMPACT_BLOCK / 32; w++) s += warp_cnt[w];
        block_counts[blockIdx.x] = s;
    }
}

// single-block exclusive scan of n int32 (n up to a few thousand blocks), total -> out[n]
__global__ void __launch_bounds__(1024) exclusive_scan_kernel(const int32_t* __restrict__ in, long long* __restrict__ out, int n) {
    __shared__ long long warp_tot[32];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        long long v = (i < n) ? in[i] : 0, incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += t;
        }
        if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            long long w = warp_tot[threadIdx.x], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                long long t = __shfl_up_sync(0xffffffffu, wi, o);
                if (threadIdx.x >= o) wi += t;
            }
            warp_tot[threadIdx.x] = wi - w;  // exclusive
        }
        __syncthreads();
        const long long excl = carry + warp_tot[threadIdx.x >> 5] + incl - v;
        if (i < n) out[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = carry;
}

// frame: 0 = HoHoNet frame (get_xyzrgb_from_depth), 1 = ZInD frame (after rotmat2d(-90)),
// 2 = posed into pano 2's frame (get_bev_pair_xyzrgb, bev_rendering_utils.py:483-522)
struct PoseArg { float R[4]; float t[2]; };
__global__ void __launch_bounds__(COMPACT_BLOCK) crop_write_kernel(SplatParams P, const uint16_t* __restrict__ depth,
                                                                   const uint8_t* __restrict__ rgb, double lo, double hi,
                                                                   const long long* __restrict__ block_offsets, int frame, PoseArg pose,
                                                                   double* __restrict__ out_xyzrgb) {
    __shared__ int warp_cnt[COMPACT_BLOCK / 32];
    const int idx = blockIdx.x * COMPACT_BLOCK + threadIdx.x;
    double x, y, z; int src = 0;
    const bool keep = crop_point(P, depth, idx, lo, hi, x, y, z, src);
    const unsigned b = __ballot_sync(0xffffffffu, keep);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) warp_cnt[warp] = __popc(b);
    __syncthreads();
    if (!keep) return;
    long long off = block_offsets[blockIdx.x];
    for (int w = 0; w < warp; w++) off += warp_cnt[w];
    off += __popc(b & ((1u << lane) - 1u));
    double* o = out_xyzrgb + off * 6;
    if (frame > 0) {
        const double tx = (double)__fmul_rn(pose.t[0], 1.5f), ty = (double)__fmul_rn(pose.t[1], 1.5f);
        double wx, wy;
        rot_pose(x, y, frame == 2, pose.R, tx, ty, wx, wy);
        x = wx; y = wy;
    }
    o[0] = x; o[1] = y; o[2] = z;
    const uint32_t c3 = gather_rgb(rgb, (uint32_t)src, P.W);
    o[3] = __ddiv_rn((double)(c3 & 0xFF), 255.0);  // rgb / 255.0, bev_rendering_utils.py:394
    o[4] = __ddiv_rn((double)((c3 >> 8) & 0xFF), 255.0);
    o[5] = __ddiv_rn((double)(c3 >> 16), 255.0);
}

// ---- z-order rule on explicit arrays (choose_elevated_repeated_vals) ----------------------------
// pass A: atomicMax of (slice, index) per pixel into a u64 grid; pass B: valid[i] = (grid[pix] == my key).
__global__ void zorder_mark_kernel(const long long* __restrict__ x, const long long* __restrict__ y, const double* __restrict__ z,
                                   long long n, const double* __restrict__ planes, int num_slices, long long w,
                                   unsigned long long* __restrict__ grid) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double zi = z[i];
    int sl = -1;
    for (int k = 0; k < num_slices; k++)
        if (zi >= planes[k] && zi < planes[k + 1]) sl = k;  // z >= z_start && z < z_end, zorder_utils.py:55
    if (sl < 0) return;
    atomicMax(grid + y[i] * w + x[i], (((unsigned long long)sl << 40) | (unsigned long long)i) + 1ull);
}
__global__ void zorder_resolve_kernel(const long long* __restrict__ x, const long long* __restrict__ y, const double* __restrict__ z,
                                      long long n, const double* __restrict__ planes, int num_slices, long long w,
                                      const unsigned long long* __restrict__ grid, uint8_t* __restrict__ valid) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double zi = z[i];
    int sl = -1;
    for (int k = 0; k < num_slices; k++)
        if (zi >= planes[k] && zi < planes[k + 1]) sl = k;
    uint8_t ok = 0;
    if (sl >= 0) ok = grid[y[i] * w + x[i]] == ((((unsigned long long)sl << 40) | (unsigned long long)i) + 1ull);
    valid[i] = ok;
}

}  // namespace bev
