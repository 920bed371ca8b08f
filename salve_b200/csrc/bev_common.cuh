// bev_common.cuh -- shared device helpers for the BEV render kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bev {

// ---- vertex / triangle packing --------------------------------------------------------------
// A vertex is a BEV pixel, labelled row<<11 | col (row <= 1022, col <= 2047); GHOST is the single
// vertex at infinity that closes the hull.  A triangle is 128 bits: three 21-bit vertex labels in
// `lo`, three 21-bit neighbour triangle ids in `hi` (n[i] = triangle across the edge opposite v[i]).
// One 16-byte load fetches everything a flip test needs; coordinates need no gather.
constexpr uint32_t M21 = 0x1FFFFFu;
constexpr uint32_t GHOST = M21;
constexpr int COL_BITS = 11;
constexpr uint32_t COL_MASK = (1u << COL_BITS) - 1;

struct __align__(16) Tri {
    unsigned long long lo, hi;
};

__host__ __device__ __forceinline__ uint32_t tri_v(const Tri& t, int i) { return (uint32_t)(t.lo >> (21 * i)) & M21; }
__host__ __device__ __forceinline__ uint32_t tri_n(const Tri& t, int i) { return (uint32_t)(t.hi >> (21 * i)) & M21; }
__host__ __device__ __forceinline__ Tri make_tri(uint32_t v0, uint32_t v1, uint32_t v2, uint32_t n0, uint32_t n1, uint32_t n2) {
    Tri t;
    t.lo = (unsigned long long)v0 | ((unsigned long long)v1 << 21) | ((unsigned long long)v2 << 42);
    t.hi = (unsigned long long)n0 | ((unsigned long long)n1 << 21) | ((unsigned long long)n2 << 42);
    return t;
}
__host__ __device__ __forceinline__ uint32_t vlabel(int row, int col) { return ((uint32_t)row << COL_BITS) | (uint32_t)col; }
__host__ __device__ __forceinline__ int vrow(uint32_t v) { return (int)(v >> COL_BITS); }
__host__ __device__ __forceinline__ int vcol(uint32_t v) { return (int)(v & COL_MASK); }

__device__ __forceinline__ Tri ld_tri(const Tri* p) {
    ulonglong2 r = *reinterpret_cast<const ulonglong2*>(p);
    Tri t; t.lo = r.x; t.hi = r.y; return t;
}
__device__ __forceinline__ void st_tri(Tri* p, const Tri& t) {
    *reinterpret_cast<ulonglong2*>(p) = make_ulonglong2(t.lo, t.hi);
}

// ---- exact predicates ------------------------------------------------------------------------
__host__ __device__ __forceinline__ long long orient2d(int ax, int ay, int bx, int by, int cx, int cy) {
    return (long long)(bx - ax) * (cy - ay) - (long long)(by - ay) * (cx - ax);
}
__host__ __device__ __forceinline__ long long orient_v(uint32_t a, uint32_t b, uint32_t c) {
    return orient2d(vcol(a), vrow(a), vcol(b), vrow(b), vcol(c), vrow(c));
}
// > 0: d strictly inside the circumcircle of CCW (a,b,c).  |result| <= 6*2047^4 ~ 1e14: exact in int64.
__host__ __device__ __forceinline__ long long incircle_v(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    long long adx = vcol(a) - vcol(d), ady = vrow(a) - vrow(d);
    long long bdx = vcol(b) - vcol(d), bdy = vrow(b) - vrow(d);
    long long cdx = vcol(c) - vcol(d), cdy = vrow(c) - vrow(d);
    long long ad = adx * adx + ady * ady, bd = bdx * bdx + bdy * bdy, cd = cdx * cdx + cdy * cdy;
    return adx * (bdy * cd - bd * cdy) - ady * (bdx * cd - bd * cdx) + ad * (bdx * cdy - bdy * cdx);
}
// 20-bit symbolic-perturbation weight of a pixel (same definition as oracle/canonical_dt.c).
__host__ __device__ __forceinline__ long long pert_weight(uint32_t v, int grid_w) {
    uint32_t h = (uint32_t)(vrow(v) * grid_w + vcol(v));
    h ^= h >> 16; h *= 0x7feb352dU; h ^= h >> 15; h *= 0x846ca68bU; h ^= h >> 16;
    return (long long)(h & 0xFFFFF);
}
// the same weight from the row-major pixel index row * grid_w + col (no label to pack and unpack)
__host__ __device__ __forceinline__ int pert_weight_idx(uint32_t h) {
    h ^= h >> 16; h *= 0x7feb352dU; h ^= h >> 15; h *= 0x846ca68bU; h ^= h >> 16;
    return (int)(h & 0xFFFFF);
}
__host__ __device__ __forceinline__ long long incircle_pert(uint32_t a, uint32_t b, uint32_t c, uint32_t d, int grid_w) {
    return pert_weight(a, grid_w) * orient_v(b, c, d) - pert_weight(b, grid_w) * orient_v(a, c, d) +
           pert_weight(c, grid_w) * orient_v(a, b, d) - pert_weight(d, grid_w) * orient_v(a, b, c);
}
// Flip rule for the edge (b,c) shared by CCW t=(a,b,c) and u=(d,c,b).
__host__ __device__ __forceinline__ bool flip_rule(uint32_t a, uint32_t b, uint32_t c, uint32_t d, int grid_w) {
    if (a == GHOST || d == GHOST) return false;          // hull edge: never removed
    if (c == GHOST) return orient_v(a, b, d) > 0;        // ghost-ghost edge: fill a strictly reflex pocket
    if (b == GHOST) return orient_v(a, d, c) > 0;
    long long ic = incircle_v(a, b, c, d);
    if (ic != 0) return ic > 0;
    return incircle_pert(a, b, c, d, grid_w) > 0;        // co-circular: canonical tie-break
}

__host__ __device__ __forceinline__ uint32_t hash32(uint32_t h) {
    h ^= h >> 16; h *= 0x7feb352dU; h ^= h >> 15; h *= 0x846ca68bU; h ^= h >> 16;
    return h;
}

#ifndef GATHER_WORDS
#define GATHER_WORDS 1
#endif
// ---- colour gather ---------------------------------------------------------------------------
// A colour source is a u8 rgb array indexed by the key's source index.  Bit 0 of the pointer (sources are 2-byte aligned)
// tags a FULL-RESOLUTION pano of (2H, 2W): the colour of pano pixel (v, u) is then the rounded mean of its 2x2 block,
// (sum + 2) >> 2 -- exactly what cv2.resize(rgb, (W, H), INTER_LINEAR) gives at a scale of 2, which is how the reference
// brings a 2048x1024 ZInD pano to 1024x512 (bev_rendering_utils.py:373-375).  No down-sampled copy is ever materialised.
__device__ __forceinline__ uint32_t load_rgb(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16); }
__device__ __forceinline__ uint32_t gather_rgb(const uint8_t* tagged, uint32_t idx, int pano_w) {
    const uintptr_t a = (uintptr_t)tagged;
    const uint8_t* base = reinterpret_cast<const uint8_t*>(a & ~(uintptr_t)1);
    if (!(a & 1)) {
#if GATHER_WORDS
        // the 3 bytes lie in one aligned word or straddle two: aligned 32-bit loads + funnel shift (one or two L1 wavefronts instead
        // of three).  An aligned word that overlaps the array never leaves its allocation (cudaMalloc granularity).
        const uintptr_t p = (uintptr_t)base + (size_t)idx * 3;
        const uint32_t* wp = reinterpret_cast<const uint32_t*>(p & ~(uintptr_t)3);
        const uint32_t sh = (uint32_t)(p & 3) * 8;
        const uint32_t w0 = __ldg(wp), w1 = sh > 8 ? __ldg(wp + 1) : 0u;
        return __funnelshift_r(w0, w1, sh) & 0xFFFFFFu;
#else
        return load_rgb(base + (size_t)idx * 3);
#endif
    }
    const uint32_t v = idx / (uint32_t)pano_w, u = idx - v * (uint32_t)pano_w;
    const size_t pitch = (size_t)pano_w * 6;
    const uint8_t* p = base + (size_t)(2 * v) * pitch + (size_t)u * 6;
    const uint8_t* q = p + pitch;
    const uint32_t r = ((uint32_t)p[0] + p[3] + q[0] + q[3] + 2u) >> 2;
    const uint32_t g = ((uint32_t)p[1] + p[4] + q[1] + q[4] + 2u) >> 2;
    const uint32_t b = ((uint32_t)p[2] + p[5] + q[2] + q[5] + 2u) >> 2;
    return r | (g << 8) | (b << 16);
}


// ---- 1-D bulk copies (TMA engine) with mbarrier completion: sm_90+ PTX, used for the contiguous streams of the path ----------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; streaming data: evict-first in L2
__device__ __forceinline__ void bulk_g2s_stream(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MBAR_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MBAR_DONE_%=;\n"
        "bra MBAR_WAIT_%=;\n"
        "MBAR_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// ---- splat key -----------------------------------------------------------------------------
// key = (z-slice << 29 | source index) + 1; 0 = empty.  max over keys == the reference's rule
// "later z-slice wins, then later point index wins" (salve/utils/zorder_utils.py:49-65).
constexpr int KEY_IDX_BITS = 29;
constexpr uint32_t KEY_IDX_MASK = (1u << KEY_IDX_BITS) - 1;

// ---- per-image scratch layout ---------------------------------------------------------------
struct ImgScratch {
    uint32_t* keygrid;   // [g]
    uint32_t* color;     // [g]
    uint32_t* occ;       // [grid_h * wpr]
    uint32_t* nonempty;  // [grid_h * wpr]
    uint32_t* keep;      // [grid_h * wpr]
    uint32_t* tmpbits;   // [grid_h * wpr]
    uint16_t* wprefix;   // [grid_h * wpr] exclusive popcount prefix of occ within the row
    Tri* tris;           // [2*g]
    unsigned long long* owner;  // [2*g]
    uint32_t* list0;     // [2*g]
    uint32_t* list1;     // [2*g]
    uint32_t* cand;      // [3*g]
};

// per-image header written by the site kernel, read by flip / raster
struct ImgHeader {
    int32_t n_sites;
    int32_t n_tris;      // 2*n_sites - 2, or 0 when there is nothing to triangulate
    int32_t status;      // SALVE_BEV_IMG_*
    int32_t pad;
};

struct GridParams {
    int32_t grid_h, grid_w, wpr, g;  // g = grid_h*grid_w
    int32_t K;
};

}  // namespace bev
