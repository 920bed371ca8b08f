// k_image.cuh -- everything after the splat, as a pipeline of six kernels over a chunk of BEV images:
//   winners -> site bit rows, sparse colours, emptiness / keep masks, convex hull,
//   and the densification itself, with NO triangle mesh in memory.
//
// Reference stages covered:
//   sparse_bev_img[y, x] = rgb of the z-order winner            bev_rendering_utils.py:298-308
//   degenerate-input guards (<4 points, one row, one column)     interpolation_utils.py:37-42, 57-71
//   griddata(points, values, grid, "linear") + u8 truncation     interpolation_utils.py:46-53
//   nonempty = uint8(r*g*b) > 0  (wraps mod 256)                 interpolation_utils.py:95-98
//   keep = KxK zero-padded box count > 0, out = keep * interp    interpolation_utils.py:101-121
//   np.flipud of the result                                      bev_rendering_utils.py:319
//
// Densification = "query-driven Lawson flips".  The value of linear interpolation at a non-site pixel q
// is fixed by the Delaunay triangle that contains q.  For each such pixel a thread holds ONE triangle
// (a,b,c) of sites with q inside it and repeats the Lawson step restricted to q: find a site d that
// violates the empty-circumcircle property of (a,b,c) (exact in-circle test; co-circular ties
// by the same symbolic perturbation as the mesh path and the CPU checker), flip inside the 4-point
// configuration {a,b,c,d} and keep the one new triangle that still contains q.  The lifted plane over
// q rises strictly with every flip, so the descent ends at the unique triangle of the canonical
// (perturbed) Delaunay triangulation over q -- bit-identical to rasterising the full mesh, but the
// only state is the 32 KB occupancy bitmap of the image: no mesh, no atomics, no rounds.
//
// The six stages (one launch each per chunk; state between them lives in per-image global scratch that the next stage
// reads through L2):
//   1 sites_stage_kernel   grid (rows / 32, images), a warp per BEV row: key grid (bulk copies, TMA engine) -> winner colours
//                          (gather from the pano), occupancy / non-empty bit rows, row summaries, the sparse image written with
//                          aligned 16-byte stores from a row staged in shared memory.  Streaming, HBM bound.
//   2 prep_stage_kernel    a CTA per image: guards, exact convex hull (pre-filtered monotone chains), keep mask (separable
//                          dilation), EDGE RULE (queries between two opposite 4-neighbour sites: mean of two colours) and the
//                          list of the remaining query pixels.
//   3 local_stage_kernel   LOCAL RULE: a table look-up on 12 neighbour bits + the empty-circle / perturbation certificate on the
//                          5 x 5 neighbourhood resolves (and shades) two thirds of those queries; the rest goes on.
//   4 window_stage_kernel  persistent warps over ONE query list for the whole chunk (lanes are filled across image
//                          boundaries): small triangles, 7 x 32 window in registers, exact float interval classification;
//                          what it cannot certify is handed on with the triangle it has reached.
//   5 shade_stage_kernel   one thread per list entry: edge-rule means, exact barycentrics of what the window pass resolved.
//   6 finish_stage_kernel  a CTA per image (most handed-on queries first): the COOPERATIVE PASS (one warp per query, 32 rows
//                          per wave, cached violators, previous final triangle as a start; a final triangle is shared by all
//                          deferred pixels inside it), masked-out sites, counters, status.
// The closed convex hull decides which pixels are queries at all (outside it griddata gives NaN -> 0).
#pragma once
#include <type_traits>

#include "bev_common.cuh"

namespace bev {

// ---- tuning knobs.  `salve_b200.build.build(defines=[...], out=...)` compiles variants and scripts/variant_bench.py times
// them against each other with an identity check of the output.
#ifndef SITES_BATCH_DEF
#define SITES_BATCH_DEF 4     // sites stage: words per lane and round trip (4 beats 8 and 16: smaller code, same latency hiding)
#endif
#ifndef IMAGE_STREAM_HINTS
#define IMAGE_STREAM_HINTS 3  // bit 0: key grid loads, bit 1: first-pass output stores are streaming (evict first)
#endif
#ifndef IMAGE_WIN_NR
#define IMAGE_WIN_NR 3        // window pass: rows above and below the query held in registers (2 is slower, 4 does not pack)
#endif
#ifndef IMAGE_WIN_REFILL
#define IMAGE_WIN_REFILL 1    // window pass: idle lanes that trigger a refill
#endif
#ifndef IMAGE_WIN_VINIT
#define IMAGE_WIN_VINIT 1     // window pass: initial triangle on the column pair D-U when it is shorter than the row pair L-R
#endif
#ifndef IMAGE_WIN_MAXGAP
#define IMAGE_WIN_MAXGAP 30   // window pass: widest row pair L-R a descent starts from (14: +0.03 ms in the finish stage, which then starts those queries from scratch)
#endif
#ifndef IMAGE_WIN_APEX_MASK
#define IMAGE_WIN_APEX_MASK 0x01FFFF00u  // window pass: columns (bit 16 + dx) searched for the apex of the initial triangle: |dx| <= 8
#endif
#ifndef IMAGE_WIN_BLOCK
#define IMAGE_WIN_BLOCK 256   // window pass: queries a warp takes from the chunk-wide list per atomic
#endif
#ifndef IMAGE_WIN_THREADS
#define IMAGE_WIN_THREADS 256 // window pass: threads per CTA
#endif
#ifndef IMAGE_WIN_CTAS
#define IMAGE_WIN_CTAS 4      // window pass: persistent CTAs per SM
#endif
#ifndef IMAGE_COOP_CHAIN
#define IMAGE_COOP_CHAIN 1    // cooperative pass: start from the previous final triangle's edge
#endif
#ifndef IMAGE_COOP_CACHE
#define IMAGE_COOP_CACHE 1    // cooperative pass: retry the last scan's violators before scanning again
#endif
#ifndef IMAGE_COOP_MIN_BAND
#define IMAGE_COOP_MIN_BAND 4 // cooperative pass: smallest band of the guided self-scheduling
#endif
#ifndef IMAGE_COOP_BAND_DIV
#define IMAGE_COOP_BAND_DIV 2 // cooperative pass: band = what is left / (IMAGE_COOP_BAND_DIV * warps)
#endif
#ifndef IMAGE_FINISH_THREADS
#define IMAGE_FINISH_THREADS 384  // finish stage: threads per CTA (3 x 384 with the deferred plane in global memory beats 2 x 512 with both planes in shared memory by 2 %; 4 x 256 ties, 5 x 128 loses 20 %)
#endif
#ifndef IMAGE_FINISH_CTAS
#define IMAGE_FINISH_CTAS 3       // finish stage: CTAs per SM the launch bounds ask for
#endif
#ifndef IMAGE_FINISH_DEFER_GLOBAL
#define IMAGE_FINISH_DEFER_GLOBAL 1  // finish stage: 1 = the deferred-query bit plane lives in global memory
#endif
#ifndef IMAGE_PREP_THREADS
#define IMAGE_PREP_THREADS 256    // prep stage: threads per CTA
#endif
#if IMAGE_STREAM_HINTS & 1
#define IMAGE_KEY_LD(p) __ldcs(p)
#else
#define IMAGE_KEY_LD(p) __ldg(p)
#endif
#if IMAGE_STREAM_HINTS & 2
#define IMAGE_OUT_ST(p, v) __stcs(p, v)
#else
#define IMAGE_OUT_ST(p, v) (*(p) = (v))
#endif
constexpr int SITES_BATCH = SITES_BATCH_DEF;
constexpr int SITES_WARPS = 8;            // sites stage: rows (= warps) per group
#ifndef SITES_DBG
#define SITES_DBG 0
#endif
#ifndef SITES_CLEAR_WARP
#define SITES_CLEAR_WARP 0
#endif
#ifndef SITES_GROUPS_DEF
#define SITES_GROUPS_DEF 4
#endif
constexpr int SITES_GROUPS = SITES_GROUPS_DEF;  // sites stage: groups of rows a CTA works through (double-buffered key fetch)
constexpr int PREP_NT = IMAGE_PREP_THREADS;
constexpr int WIN_NT = IMAGE_WIN_THREADS;
constexpr int FINISH_NT = IMAGE_FINISH_THREADS;
constexpr int WIN_MAX_IMAGES = 2048;      // images one window-stage launch can index (prefix table in shared memory)
constexpr int IMAGE_MAX_FLIPS = 100000;   // safety cap on one descent (never reached: the lift is strictly monotone)

// per-image row arrays kept in global memory between the stages: RA_N16 int16 arrays of `hp` entries, then RA_NF float arrays
enum { RA_CNT = 0, RA_FIRST, RA_LAST, RA_NE, RA_UP, RA_DN, RA_HL0, RA_HL1, RA_HR0, RA_HR1, RA_N16 };
enum { RA_HLF = 0, RA_HRF, RA_NF };
// per-image header (16 int32)
enum { HD_STATUS = 0, HD_NQ, HD_EDGE, HD_MASKED, HD_PEND, HD_XTRA, HD_LOCAL, HD_NRAW, HD_N };
constexpr int HD_STRIDE = 16;

struct ImageArgs {
    GridParams G;
    int32_t n_img;                    // images of this launch
    const int32_t* order;             // finish stage: CTA k works on image order[k] (longest expected first), or null: k
    uint32_t* keygrid; size_t keygrid_stride;
    const uint8_t* const* color_src;  // per image: u8 rgb triples indexed by the key's source index (tagged, see gather_rgb)
    int32_t pano_w;                   // width of the key's index space (for tagged full-resolution sources)
    int32_t* counts;                  // [n_img][8] working counters, indexed like the key grids ([0], [1] come from the splat)
    int32_t* status;                  // by destination, or null
    uint8_t* out; size_t out_stride;  // final images by destination (bytes per image)
    // Destination of image i (null: i).  d >= 0: out / counts_out / status slot d.  d < 0: slot -1-d of the context's cache of
    // hypothesis-independent renders (cache_out / cache_counts / cache_status).
    const int32_t* dest;
    int32_t* counts_out;              // final counters by destination, or null (then `counts` is the final array)
    uint8_t* cache_out; int32_t* cache_counts; int32_t* cache_status;
    uint8_t* hull; size_t hull_stride;        // optional tap: 1 inside the closed convex hull
    int32_t* qtri; size_t qtri_stride;        // optional tap: per pixel the 3 vertex pixel ids of its triangle (pre-filled with -1)
    // state between the stages, per image of the chunk
    uint32_t* planes; size_t plane_stride;    // [n_img][3][plane_stride] bit rows: occupancy, non-empty, keep
    uint32_t* defer;                          // [n_img][plane_stride] finish stage: queries handed to the cooperative pass (bit rows)
    unsigned char* rows; size_t rows_stride;  // [n_img] row arrays (RA_*), rows_stride bytes per image
    int32_t hp;                               // entries per row array (grid_h rounded up to 16)
    int32_t* hdr;                             // [n_img][HD_STRIDE]
    uint32_t* qlist; size_t qlist_stride;     // per image: query pixels (row << 11 | col) for the window pass, capacity g
    unsigned long long* qres;                 // per image, per list entry: triangle (3 x 21-bit vertex labels), final if | QRES_DONE
    uint32_t* clist;                          // per image: edge-rule pixels (prep), then the entries handed to the cooperative pass
    int32_t* work_counter;                    // window stage: next block of the chunk-wide list (zeroed by the host)
    const uint32_t* local_lut;                // local rule tables (LocalRule, built by the host), or null: rule off
    int32_t raw_mode;                 // 1: no keep mask, no flip (interp_dense_grid_from_sparse semantics)
    int32_t skip_empty_check;         // 1: generic interp path (no EMPTY status)
    int32_t clear_keys;               // 1: the sites stage zeroes the keys it consumes (the next chunk needs no memset)
};

__device__ __forceinline__ uint8_t* image_out(const ImageArgs& A, int img, int& dst) {
    dst = A.dest ? A.dest[img] : img;
    return dst >= 0 ? A.out + (size_t)dst * A.out_stride : A.cache_out + (size_t)(-1 - dst) * A.out_stride;
}
__device__ __forceinline__ int16_t* row_arr(const ImageArgs& A, int img, int k) {
    return reinterpret_cast<int16_t*>(A.rows + (size_t)img * A.rows_stride) + (size_t)k * A.hp;
}
__device__ __forceinline__ float* row_arr_f(const ImageArgs& A, int img, int k) {
    return reinterpret_cast<float*>(A.rows + (size_t)img * A.rows_stride + (size_t)RA_N16 * A.hp * 2) + (size_t)k * A.hp;
}
__host__ __device__ inline size_t image_rows_stride(int hp) { return (size_t)RA_N16 * hp * 2 + (size_t)RA_NF * hp * 4; }

// shared memory (bytes) of the stages for a grid of (h, wpr)
__host__ __device__ inline size_t image_row_bytes(int h) { return ((size_t)h * 2 + 15) & ~(size_t)15; }
__host__ __device__ inline size_t sites_row_words(int wpr) { return (size_t)((wpr + SITES_BATCH - 1) / SITES_BATCH) * SITES_BATCH * 24 + 8; }
__host__ __device__ inline size_t sites_key_words(int w) { return ((size_t)SITES_WARPS * w + 3 + 3) & ~(size_t)3; }  // rounded-up bulk copy, 16-byte multiple
__host__ __device__ inline size_t sites_smem_bytes(int w, int wpr) { return (2 * sites_key_words(w) + SITES_WARPS * sites_row_words(wpr)) * 4; }
__host__ __device__ inline size_t prep_smem_bytes(int h, int wpr) {
    return (size_t)h * wpr * 4 + 13 * image_row_bytes(h) + 2 * (((size_t)h * 4 + 15) & ~(size_t)15) + 64;
}
__host__ __device__ inline size_t finish_smem_bytes(int h, int wpr) {
    return (IMAGE_FINISH_DEFER_GLOBAL ? 1 : 2) * (size_t)h * wpr * 4 + 9 * image_row_bytes(h) + 2 * (((size_t)h * 4 + 15) & ~(size_t)15) + 64;
}
// the largest of them decides whether a grid can take this path at all (else: explicit-mesh kernels)
__host__ __device__ inline size_t image_smem_bytes(int h, int wpr) {
    const size_t a = prep_smem_bytes(h, wpr), b = finish_smem_bytes(h, wpr), c = sites_smem_bytes(wpr * 32, wpr);
    return a > b ? (a > c ? a : c) : (b > c ? b : c);
}

#if IMAGE_FINISH_DEFER_GLOBAL
#define DEFER_LD(p) __ldcg(p)   // other warps clear bits with atomics (performed in L2): never read them through L1
#else
#define DEFER_LD(p) (*(p))
#endif

// ---- local rule: queries whose Delaunay triangle has its three vertices among 12 near neighbours --------------------------
// After the edge rule, two thirds of the remaining queries lie in a triangle whose vertices are all among the 8 neighbours and the
// 4 pixels at distance 2 on q's row and column, with a circumcircle that stays inside the 5 x 5 neighbourhood.  For them the
// whole search is a table look-up on the 12 neighbour bits: the table gives (up to 4 of) the triangles of those neighbours that
// contain q and have none of the 12 strictly inside; a candidate is THE triangle of the canonical triangulation over q iff the
// other lattice points strictly inside its circle (a 25-bit mask of the neighbourhood) are no sites and every site ON the circle
// loses the symbolic-perturbation test -- the same certificate the window pass and the cooperative pass end with, so the result
// is identical bit for bit.  The prep stage runs it with one thread per query; what it resolves never reaches the window pass.
// Neighbourhood bit (dy + 2) * 5 + (dx + 2) <-> pixel (x + dx, r + dy).
struct LocalRule {
    static constexpr int NPOS = 12, NPAT = 1 << NPOS, MAXCAND = 128;
    // words of the table: [NPAT] up to four candidate ids per pattern (a byte each, 0xFF: none), then per candidate
    // inside mask, on-circle mask, the three vertices (5 bits each: neighbourhood bit indices, counter-clockwise)
    static constexpr int OFF_INSIDE = NPAT, OFF_ON = NPAT + MAXCAND, OFF_VERTS = NPAT + 2 * MAXCAND, WORDS = NPAT + 3 * MAXCAND;
    // pattern bit k <-> neighbourhood bit: (0,-2); (-1,-1) (0,-1) (1,-1); (-2,0) (-1,0); (1,0) (2,0); (-1,1) (0,1) (1,1); (0,2)
    __host__ __device__ static constexpr int pos_bit(int k) {
        return k == 0 ? 2 : k <= 3 ? 5 + k : k <= 5 ? 6 + k : k <= 7 ? 7 + k : k <= 10 ? 8 + k : 22;
    }
    __host__ __device__ static uint32_t pattern_of(uint32_t nb) {
        return ((nb >> 2) & 1u) | (((nb >> 6) & 7u) << 1) | (((nb >> 10) & 3u) << 4) | (((nb >> 13) & 3u) << 6) | (((nb >> 16) & 7u) << 8) |
               (((nb >> 22) & 1u) << 11);
    }
};

struct Tri2 { int ax, ay, bx, by, cx, cy; };

__device__ __forceinline__ int orient_i(int ax, int ay, int bx, int by, int cx, int cy) {
    return (bx - ax) * (cy - ay) - (by - ay) * (cx - ax);  // |.| <= 2*2047^2: exact in int32
}
__device__ __forceinline__ bool tri_contains(const Tri2& t, int qx, int qy) {
    return orient_i(t.ax, t.ay, t.bx, t.by, qx, qy) >= 0 && orient_i(t.bx, t.by, t.cx, t.cy, qx, qy) >= 0 &&
           orient_i(t.cx, t.cy, t.ax, t.ay, qx, qy) >= 0;
}
__device__ __forceinline__ bool ccw_contains(int ax, int ay, int bx, int by, int cx, int cy, int qx, int qy) {
    return orient_i(ax, ay, bx, by, cx, cy) > 0 && orient_i(ax, ay, bx, by, qx, qy) >= 0 && orient_i(bx, by, cx, cy, qx, qy) >= 0 &&
           orient_i(cx, cy, ax, ay, qx, qy) >= 0;
}

// bits lo..hi (absolute columns, inclusive) that fall into word wi
__device__ __forceinline__ uint32_t range_mask(int wi, int lo, int hi) {
    int a = lo - wi * 32, b = hi - wi * 32;
    if (lo > hi || b < 0 || a > 31) return 0u;
    a = max(a, 0); b = min(b, 31);
    return (0xFFFFFFFFu >> (31 - b)) & (0xFFFFFFFFu << a);
}
// highest set bit p with xmin <= p <= x, or -1
__device__ __forceinline__ int prev_bit(const uint32_t* row, int x, int xmin) {
    if (x < xmin) return -1;
    int wi = x >> 5;
    const int wmin = xmin >> 5;
    uint32_t m = row[wi] & (0xFFFFFFFFu >> (31 - (x & 31)));
    while (true) {
        if (m) { const int p = wi * 32 + 31 - __clz(m); return p >= xmin ? p : -1; }
        if (--wi < wmin) return -1;
        m = row[wi];
    }
}
// lowest set bit p with x <= p <= xmax, or -1
__device__ __forceinline__ int next_bit(const uint32_t* row, int x, int xmax) {
    if (x > xmax) return -1;
    int wi = x >> 5;
    const int wmax = xmax >> 5;
    uint32_t m = row[wi] & (0xFFFFFFFFu << (x & 31));
    while (true) {
        if (m) { const int p = wi * 32 + __ffs(m) - 1; return p <= xmax ? p : -1; }
        if (++wi > wmax) return -1;
        m = row[wi];
    }
}


struct ImageShared {
    uint32_t* occ; uint32_t* keep; uint32_t* tmp;
    int16_t *cnt, *first, *last, *up, *dn, *hlo, *hhi, *hl0, *hl1, *hr0, *hr1, *stk_l, *stk_r;
    float *hlf, *hrf;  // real-valued hull cross-section per row (left, right); empty rows: +inf, -inf
};

// Initial triangle for pixel (x, r) lying strictly between two sites of its own row: nearest site left, nearest right,
// and the nearest site of the closest non-empty row above (else below).
__device__ __forceinline__ bool init_tri_row(const ImageShared& S, int wpr, int w, int x, int r, Tri2& t) {
    const uint32_t* row = S.occ + r * wpr;
    const int xl = prev_bit(row, x - 1, 0), xr = next_bit(row, x + 1, w - 1);
    if (xl < 0 || xr < 0) return false;
    int yy = S.up[r];
    const bool above = yy >= 0;
    if (!above) yy = S.dn[r];
    if (yy < 0) return false;
    const uint32_t* r2 = S.occ + yy * wpr;
    const int pl = prev_bit(r2, x, 0), pr = next_bit(r2, x + 1, w - 1);
    const int xc = (pr < 0 || (pl >= 0 && x - pl <= pr - x)) ? pl : pr;
    if (above) { t.ax = xl; t.ay = r; t.bx = xr; t.by = r; }
    else { t.ax = xr; t.ay = r; t.bx = xl; t.by = r; }
    t.cx = xc; t.cy = yy;
    return true;
}
// Initial triangle for a pixel inside the hull but outside its row's site extent: the hull edge on its side plus the
// nearest site of its own row; if the row is empty, split the quad spanned by the two hull edges that cross row r.
__device__ __forceinline__ bool init_tri_hull(const ImageShared& S, int x, int r, Tri2& t) {
    const int l0 = S.hl0[r], l1 = S.hl1[r], r0 = S.hr0[r], r1 = S.hr1[r];
    const int L0x = S.first[l0], L1x = S.first[l1], R0x = S.last[r0], R1x = S.last[r1];
    if (S.cnt[r] > 0) {
        if (x < S.first[r]) {
            const int px = S.first[r];
            if (ccw_contains(L0x, l0, px, r, L1x, l1, x, r)) { t = {L0x, l0, px, r, L1x, l1}; return true; }
        } else {
            const int px = S.last[r];
            if (ccw_contains(R0x, r0, R1x, r1, px, r, x, r)) { t = {R0x, r0, R1x, r1, px, r}; return true; }
        }
    }
    if (ccw_contains(L0x, l0, R0x, r0, R1x, r1, x, r)) { t = {L0x, l0, R0x, r0, R1x, r1}; return true; }
    if (ccw_contains(L0x, l0, R1x, r1, L1x, l1, x, r)) { t = {L0x, l0, R1x, r1, L1x, l1}; return true; }
    if (ccw_contains(L0x, l0, R0x, r0, L1x, l1, x, r)) { t = {L0x, l0, R0x, r0, L1x, l1}; return true; }
    if (ccw_contains(R0x, r0, R1x, r1, L1x, l1, x, r)) { t = {R0x, r0, R1x, r1, L1x, l1}; return true; }
    return false;
}

constexpr unsigned long long QRES_DONE = 1ull << 63;     // the entry holds the query's final triangle
// an entry without QRES_DONE is either 0 (no triangle) or a valid triangle around the query that is not final yet: three distinct
// 21-bit labels fill bits 0..62, so a triangle never encodes to 0

// bits lo..hi of a 32-bit window (any lo, hi; empty when lo > hi): clamped funnel shifts make every out-of-range case come out right
__device__ __forceinline__ uint32_t bit_span(int lo, int hi) {
    return __funnelshift_lc(0u, 0xFFFFFFFFu, (uint32_t)max(lo, 0)) & __funnelshift_rc(0xFFFFFFFFu, 0u, (uint32_t)(31 - min(hi, 31)));
}
__device__ __forceinline__ float sqrt_approx(float v) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }

// ---- stage 3: small triangles in a register window --------------------------------------------------------------------------
// 92 % of the queries that are left after the edge rule end in a triangle within 2 px of the pixel with a circumradius of
// at most 2 px.  For them the whole Lawson descent runs on a (2*NR+1)-row x 32-column window of the occupancy bitmap held in
// registers, in coordinates relative to the query q = (0, 0):
//   * circle of (a, b, c): A2 = twice the area, (U, V) = 2*A2 * (centre - a); a lattice point p has
//     |p - centre|^2 - R^2 = -inc(p) / A2 with inc an integer, so points off the circle miss it by at least 1/A2 in squared
//     distance.  With |coordinates| <= 16 every float below is exact to ~1e-5, hence "strictly inside" (margin > 1/(2 A2)) and
//     "on the circle" (|.| <= 1/(2 A2)) are decided EXACTLY by float interval arithmetic per row: no per-point test at all.
//   * one trip = the strict-interior masks of all rows (unrolled; every lane executes the same code), then either a flip towards
//     the violator nearest to q or, if the circle is empty, the on-circle sites and their symbolic-perturbation tests.
// A query whose triangle or circle leaves the window is handed to the cooperative pass.
//
// Work distribution: the query lists of all images of the chunk form one index space (prefix sums of the per-image counts,
// rebuilt in shared memory by every CTA).  A warp takes IMAGE_WIN_BLOCK consecutive indices per atomic and refills its idle
// lanes from that block, across image boundaries: there is no per-image barrier and no per-image tail.
template <int NR>
__global__ void __launch_bounds__(WIN_NT, IMAGE_WIN_CTAS) window_stage_kernel(ImageArgs A) {
    constexpr int NROW = 2 * NR + 1;
    static_assert(NROW * 9 <= 64, "on-circle candidates are packed 9 bits per row into one 64-bit word");
    static_assert(WIN_MAX_IMAGES % WIN_NT == 0, "prefix table: whole entries per thread");
    constexpr int MAXGAP = IMAGE_WIN_MAXGAP, MAXFLIPS = 16;
    constexpr int PER = WIN_MAX_IMAGES / WIN_NT;
    const unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31;
    const int W = A.G.grid_w, H = A.G.grid_h, wpr = A.G.wpr;
    const int n_img = A.n_img;

    // ---- exclusive prefix of the per-image list lengths: s_off[i] = first global index of image i, s_off[n_img] = total
    __shared__ int s_off[WIN_MAX_IMAGES + 1];
    __shared__ int s_warp[WIN_NT / 32];
    {
        int v[PER], sum = 0;
#pragma unroll
        for (int k = 0; k < PER; k++) {
            const int i = tid * PER + k;
            v[k] = i < n_img ? __ldg(A.hdr + (size_t)i * HD_STRIDE + HD_NQ) : 0;
            sum += v[k];
        }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) s_warp[tid >> 5] = incl;
        __syncthreads();
        int base = 0;
        for (int k = 0; k < (tid >> 5); k++) base += s_warp[k];
        int run = base + incl - sum;
#pragma unroll
        for (int k = 0; k < PER; k++) {
            const int i = tid * PER + k;
            if (i <= n_img) s_off[i] = run;  // entry n_img: every later v is 0, so run is already the total there
            run += v[k];
        }
        __syncthreads();
    }
    const int total = s_off[n_img];

    bool active = false, exhausted = false;
    int x = 0, r = 0, flips = 0, img = 0;
    uint32_t gidx = 0;  // img * qlist_stride + index within the image's list
    int acc_img = -1, acc_flips = 0, acc_max = 0;  // flip statistics of the image this lane last worked on (diagnostic counters)
    int cur = 0, end = 0, blk_img = 0;  // warp-uniform: this warp's block of the chunk-wide list
    uint32_t wr[NROW];  // wr[dy + NR]: bit 16 + dx <-> pixel (x + dx, r + dy)
#pragma unroll
    for (int k = 0; k < NROW; k++) wr[k] = 0u;
    int ax = 0, ay = 0, bx = 0, by = 0, cx = 0, cy = 0;

    auto flush_stats = [&]() {
        if (acc_img >= 0 && acc_flips) {
            atomicAdd(A.counts + (size_t)acc_img * 8 + 7, acc_flips);
            atomicMax(A.counts + (size_t)acc_img * 8 + 6, acc_max);
        }
        acc_flips = 0; acc_max = 0;
    };
    // hand the query to the cooperative pass, with the triangle the descent has reached so far (if there is one)
    auto give_up = [&](bool with_tri) {
        unsigned long long v = 0ull;
        if (with_tri)
            v = (unsigned long long)vlabel(r + ay, x + ax) | ((unsigned long long)vlabel(r + by, x + bx) << 21) |
                ((unsigned long long)vlabel(r + cy, x + cx) << 42);
        A.qres[gidx] = v;
        atomicAdd(A.hdr + (size_t)img * HD_STRIDE + HD_PEND, 1);
        active = false;
    };
    // nearest set bit to dx = 0 in a window row (m != 0): returns dx
    auto nearest = [](uint32_t m) -> int {
        const uint32_t lo = m & 0x1FFFFu, hi = m >> 17;  // dx <= 0, dx >= 1
        const int dl = lo ? 16 - (31 - __clz(lo)) : 64, dr = hi ? __ffs(hi) : 64;  // distances
        return dl <= dr ? -dl : dr;
    };
    auto cross = [](int px, int py, int qx, int qy) { return px * qy - py * qx; };

    while (true) {
        // ---- refill idle lanes
        const unsigned idle = __ballot_sync(FULL, !active);
        if (idle && !exhausted && (__popc(idle) >= IMAGE_WIN_REFILL || idle == FULL)) {
            if (cur == end) {  // this warp's block is used up: take the next one
                int b = 0;
                if (lane == 0) b = atomicAdd(A.work_counter, IMAGE_WIN_BLOCK);
                b = __shfl_sync(FULL, b, 0);
                if (b >= total) exhausted = true;
                else {
                    cur = b; end = min(b + IMAGE_WIN_BLOCK, total);
                    int lo = 0, hi = n_img - 1;  // image that holds index b: the last i with s_off[i] <= b
                    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (s_off[mid] <= b) lo = mid; else hi = mid - 1; }
                    blk_img = lo;
                }
            }
            if (!exhausted) {
                const int take = min(__popc(idle), end - cur);
                const int rank = __popc(idle & ((1u << lane) - 1u));
                if (!active && rank < take) {
                    const int gi = cur + rank;
                    int im = blk_img;
                    while (gi >= s_off[im + 1]) im++;  // gi < total = s_off[n_img]: stops at the image that holds gi
                    if (im != acc_img) { flush_stats(); acc_img = im; }
                    img = im;
                    gidx = (uint32_t)((size_t)im * A.qlist_stride) + (uint32_t)(gi - s_off[im]);
                    const uint32_t code = A.qlist[gidx];
                    x = (int)(code & COL_MASK); r = (int)(code >> COL_BITS);
                    active = true; flips = 0;
                    const uint32_t* occ = A.planes + (size_t)im * 3 * A.plane_stride;
                    // window: branch-free with clamped row / word indices and masks for what lies outside the grid
                    const int c0 = x - 16;
                    const int w0 = c0 >> 5, sh = c0 & 31;  // c0 may be negative: arithmetic shift = floor
                    const int wlo = max(w0, 0), whi = min(w0 + 1, wpr - 1);
                    const uint32_t mlo = w0 >= 0 ? 0xFFFFFFFFu : 0u, mhi = w0 + 1 < wpr ? 0xFFFFFFFFu : 0u;
                    if (r >= NR && r + NR < H && w0 >= 0 && w0 + 1 < wpr) {  // the window lies inside the grid (almost always)
                        const uint32_t* row = occ + (r - NR) * wpr + w0;
#pragma unroll
                        for (int k = 0; k < NROW; k++) wr[k] = __funnelshift_r(__ldg(row + k * wpr), __ldg(row + k * wpr + 1), sh);
                    } else {
#pragma unroll
                        for (int k = 0; k < NROW; k++) {
                            const int y = r + k - NR;
                            const int yc = min(max(y, 0), H - 1);
                            const uint32_t* row = occ + yc * wpr;
                            const uint32_t vm = y == yc ? 0xFFFFFFFFu : 0u;
                            wr[k] = __funnelshift_r(__ldg(row + wlo) & mlo & vm, __ldg(row + whi) & mhi & vm, sh);
                        }
                    }
                    // initial triangle.  Row pair: nearest sites left and right of q (64: none in the window).
                    const uint32_t ml = wr[NR] & 0xFFFFu, mr = wr[NR] >> 17;
                    const int xl = ml ? (31 - __clz(ml)) - 16 : -64, xr = mr ? __ffs(mr) : 64;
                    bool ok = xr - xl <= MAXGAP;
                    // apex candidates: nearest site (|dx| <= 8) of the closest non-empty row above and of the closest below
                    uint32_t mu = 0u, md = 0u;
                    int ku = 0, kd = 0, yu = 0, yd = 0;  // yu, yd: nearest sites in q's own column
#pragma unroll
                    for (int k = NR; k >= 1; k--) {
                        const uint32_t u = wr[NR + k], d = wr[NR - k];
                        if (u & IMAGE_WIN_APEX_MASK) { mu = u & IMAGE_WIN_APEX_MASK; ku = k; }
                        if (d & IMAGE_WIN_APEX_MASK) { md = d & IMAGE_WIN_APEX_MASK; kd = k; }
                        if (u & 0x10000u) yu = k;
                        if (d & 0x10000u) yd = k;
                    }
                    int px = 0, py = 0;
                    if (ku) { px = nearest(mu); py = ku; }
                    if (kd) {
                        const int pxd = nearest(md);
                        if (!ku || pxd * pxd + kd * kd < px * px + py * py) { px = pxd; py = -kd; }  // the nearer of the two
                    }
                    const bool found = (ku | kd) != 0;
                    ok = ok && found;
                    // Column pair: q lies on the segment D-U as it lies on L-R.  The shorter of the two is the likelier Delaunay
                    // edge (fewer flips to come), and D-U serves when L or R is missing.
                    bool vert = false;
                    if (IMAGE_WIN_VINIT && yu != 0 && yd != 0 && (!ok || yu + yd < xr - xl)) {
                        int qx = 0, qy = 0;  // apex: the nearer of L and R, else the site found in a neighbouring row
                        if (ml != 0u && -xl <= xr) qx = xl;
                        else if (mr != 0u) qx = xr;
                        else if (px != 0) { qx = px; qy = py; }
                        if (qx != 0) {
                            vert = true; ok = true;
                            ax = 0; bx = 0; cx = qx; cy = qy;
                            if (qx < 0) { ay = -yd; by = yu; } else { ay = yu; by = -yd; }  // (D, U, P) with P on the left, (U, D, P) on the right
                        }
                    }
                    if (ok && !vert) {
                        if (py > 0) { ax = xl; bx = xr; } else { ax = xr; bx = xl; }
                        ay = 0; by = 0; cx = px; cy = py;
                    }
                    if (!ok) give_up(false);
                }
                cur += take;
                while (blk_img + 1 < n_img && cur >= s_off[blk_img + 1]) blk_img++;
            }
        }
        if (!__any_sync(FULL, active)) { if (exhausted) break; continue; }
        if (!active) continue;

        // ---- circle of (a, b, c)
        const int ux = bx - ax, uy = by - ay, vx = cx - ax, vy = cy - ay;
        const int b2 = ux * ux + uy * uy, c2 = vx * vx + vy * vy;
        const int A2 = ux * vy - uy * vx;  // > 0
        const int U = b2 * vy - uy * c2, V = ux * c2 - b2 * vx;
        const float inv = 0.5f / (float)A2;
        const float fU = (float)U, fV = (float)V;
        const float ccx = (float)ax + fU * inv, ccy = (float)ay + fV * inv;
        const float R2 = (fU * fU + fV * fV) * inv * inv;
        const float thr = inv;  // half the smallest possible |distance^2 - R^2| of a lattice point off the circle
        const float ccx16 = ccx + 16.0f;  // window column of the centre (bit 16 is q's column)
        // does the closed disc stay inside rows +-NR and columns +-15 (conservative, float)?  Only then can the window certify
        // an empty circle; a larger circle can still be searched for violators inside the window (any violator is a valid flip)
        const float Rr = sqrt_approx(R2) + 0.01f;
        const bool fits = fabsf(ccy) + Rr < (float)NR + 0.99f && fabsf(ccx) + Rr < 15.0f;
        // ---- one pass over the window rows (q's row first, then +-1, +-2, ...): strict interior (the violator nearest to q wins)
        // and the sites ON the circle.  Every lane runs the same code whether its circle turns out empty or not.
        bool have = false;
        int dx = 0, dy = 0;
        uint32_t on[NROW], vm = 0u, anyon = 0u;
#pragma unroll
        for (int k = 0; k < NROW; k++) {
            const int yy = (k == 0) ? 0 : ((k & 1) ? (k + 1) / 2 : -(k / 2));
            const float e = (float)yy - ccy;
            const float base = R2 - e * e;
            const float ti = base - thr, to = base + thr;
            uint32_t im = 0u, om = 0u;
            if (to >= 0.0f) {
                const float hwo = sqrt_approx(to);
                om = bit_span(__float2int_ru(ccx16 - hwo), __float2int_rd(ccx16 + hwo));
                if (ti > 0.0f) {
                    const float hwi = sqrt_approx(ti);
                    im = bit_span(__float2int_ru(ccx16 - hwi), __float2int_rd(ccx16 + hwi));
                }
            }
            const uint32_t sites = wr[yy + NR];
            const uint32_t m = sites & im;
            if (m && !have) { have = true; vm = m; dy = yy; }
            // sites on the circle (a, b, c among them: they are taken out when the candidates are packed)
            const uint32_t o = sites & om & ~im;
            on[k] = o; anyon += __popc(o);
        }
        if (have) dx = nearest(vm);
        if (have && !fits) {
            // large circle: float may misjudge points near it -- confirm the violator exactly (int32: |coordinates| <= 16)
            const int ex = dx - ax, ey = dy - ay;
            if (U * ex + V * ey - A2 * (ex * ex + ey * ey) <= 0) { give_up(true); continue; }
        }
        if (!have && !fits) { give_up(true); continue; }
        if (!have) {
            // ---- empty circle: sites ON it decide by the symbolic perturbation (same rule as incircle_pert(), in window
            // coordinates: weights < 2^20, |orient| <= 2 * 31 * 6, so int32 holds every term and the sum)
            // The circle fits (R < 4), so an on-circle site has |dx - rint(ccx)| <= 4: 9 bits per row, NROW rows in one 64-bit
            // word, rows in scan order.
            const int icx = __float2int_rn(ccx);
            unsigned long long cand = 0ull;
            if (anyon > 3u) {
#pragma unroll
                for (int k = 0; k < NROW; k++) cand |= (unsigned long long)((uint32_t)(((unsigned long long)on[k] << 4) >> (icx + 16)) & 0x1FFu) << (9 * k);
                // row index in scan order of a vertex row vy: 0, +1, -1, +2, ... -> 0, 1, 2, 3, ...
                cand &= ~(1ull << (9 * (ay > 0 ? 2 * ay - 1 : -2 * ay) + ax - icx + 4));
                cand &= ~(1ull << (9 * (by > 0 ? 2 * by - 1 : -2 * by) + bx - icx + 4));
                cand &= ~(1ull << (9 * (cy > 0 ? 2 * cy - 1 : -2 * cy) + cx - icx + 4));
            }
            if (cand) {
                const int qi = r * W + x;  // row-major index of q: a window point (dx, dy) is pixel qi + dy * W + dx
                const int wa = pert_weight_idx((uint32_t)(qi + ay * W + ax)), wb = pert_weight_idx((uint32_t)(qi + by * W + bx)),
                          wc = pert_weight_idx((uint32_t)(qi + cy * W + cx));
                while (cand && !have) {
                    const int b = __ffsll((long long)cand) - 1;
                    cand &= cand - 1ull;
                    const int k = (b * 57) >> 9;  // b / 9 for b < 63
                    const int ddx = b - 9 * k - 4 + icx;
                    const int yy = (k & 1) ? (k + 1) >> 1 : -(k >> 1);
                    const int wd = pert_weight_idx((uint32_t)(qi + yy * W + ddx));
                    const int obcd = (cx - bx) * (yy - by) - (cy - by) * (ddx - bx);
                    const int oacd = (cx - ax) * (yy - ay) - (cy - ay) * (ddx - ax);
                    const int oabd = (bx - ax) * (yy - ay) - (by - ay) * (ddx - ax);
                    const int pert = wa * obcd - wb * oacd + wc * oabd - wd * A2;
                    if (pert > 0) { have = true; dx = ddx; dy = yy; }
                }
            }
            if (!have) {  // t is the triangle of the canonical triangulation over q
                const uint32_t va = vlabel(r + ay, x + ax), vb = vlabel(r + by, x + bx), vc = vlabel(r + cy, x + cx);
                A.qres[gidx] = QRES_DONE | (unsigned long long)va | ((unsigned long long)vb << 21) | ((unsigned long long)vc << 42);
                acc_flips += flips; acc_max = max(acc_max, flips);
                active = false;
                continue;
            }
        }
        // ---- Lawson flip inside {a, b, c, d}: keep the new triangle that contains q (the origin).  The three candidates
        // (d,b,c), (a,d,c), (a,b,d) share six cross products; orient(p,q,s) = p x q + q x s + s x p.  No branches.
        {
            const int Xab = cross(ax, ay, bx, by), Xbc = cross(bx, by, cx, cy), Xca = cross(cx, cy, ax, ay);
            const int Xad = cross(ax, ay, dx, dy), Xbd = cross(bx, by, dx, dy), Xcd = cross(cx, cy, dx, dy);
            const bool fa = (-Xbd >= 0) & (Xbc >= 0) & (Xcd >= 0) & (Xbc + Xcd - Xbd > 0);
            const bool fb = (Xad >= 0) & (-Xcd >= 0) & (Xca >= 0) & (Xad - Xcd + Xca > 0);
            const bool fc = (Xab >= 0) & (Xbd >= 0) & (-Xad >= 0) & (Xab + Xbd - Xad > 0);
            if (fa) { ax = dx; ay = dy; }
            else if (fb) { bx = dx; by = dy; }
            else if (fc) { cx = dx; cy = dy; }
            else { give_up(false); continue; }  // cannot happen
        }
        if (++flips > MAXFLIPS) give_up(true);
    }
    flush_stats();
}
// ---- warp-cooperative versions for queries whose triangles are large (wide gaps, hull pockets) -------------------------
// One lane per row of each 32-row wave (row offsets 0, -1, +1, -2, ... from qy).  Returns, in every lane, the violator
// nearest to q found in the first wave that has one (x | y << 16), or -1.
// SG (grid_h, grid_w <= 512): |U|, |V| <= 2 * (2 * 511^2) * 511 < 2^31, so the circle's integers are int32 and only the in-circle
// determinant itself is widened (32 x 32 -> 64 bit multiplies); larger grids keep everything in int64.
template <bool SG>
__device__ __forceinline__ int coop_find_violator(const uint32_t* __restrict__ occ, const float* __restrict__ hlf, const float* __restrict__ hrf, int wpr,
                                                  int W, int H, const Tri2& t, int qx, int qy, int grid_w, int lane, int& waves, unsigned long long& cache) {
    typedef typename std::conditional<SG, int, long long>::type I;
    const I bx = t.bx - t.ax, by = t.by - t.ay, cx = t.cx - t.ax, cy = t.cy - t.ay;
    const I b2 = bx * bx + by * by, c2 = cx * cx + cy * cy;
    const I A2 = bx * cy - by * cx;
    const I U = b2 * cy - by * c2, V = bx * c2 - b2 * cx;
    const double inv = 1.0 / (double)A2;
    const double ux = 0.5 * (double)U * inv, uy2 = (double)V * inv, ux2 = ux * ux, cxa = (double)t.ax + ux;
    const uint32_t va = vlabel(t.ay, t.ax), vb = vlabel(t.by, t.bx), vc = vlabel(t.cy, t.cx);
    bool up_dead = false, dn_dead = false;
    for (int wave = 0;; wave++) {
        waves++;
        const int o = wave * 32 + lane;
        const int k = (o + 1) >> 1;
        const bool down = (o & 1) != 0;
        const int y = down ? qy - k : qy + k;
        bool row_dead = false;
        unsigned long long best = ~0ull;
        if (!(down ? dn_dead : up_dead)) {
            if (y < 0 || y >= H) row_dead = true;
            else {
                const double dyr = (double)(y - t.ay);
                const double tt = ux2 + dyr * (uy2 - dyr);
                if (tt < -1e-6) row_dead = true;
                else {
                    const double hw = sqrt(fmax(tt, 0.0)) + 1e-3;
                    const double lo = fmax(cxa - hw, (double)hlf[y] - 1e-2), hi = fmin(cxa + hw, (double)hrf[y] + 1e-2);
                    if (!(lo <= hi)) row_dead = true;
                    else {
                        const int x0 = (int)ceil(lo), x1 = (int)floor(hi);
                        if (x0 <= x1) {
                            const uint32_t* row = occ + y * wpr;
                            int pl = prev_bit(row, min(qx, x1), x0);
                            int pr = next_bit(row, max(qx + 1, x0), x1);
                            while (pl >= 0 || pr >= 0) {
                                const bool take_l = pr < 0 || (pl >= 0 && qx - pl <= pr - qx);
                                const int x = take_l ? pl : pr;
                                const uint32_t vd = vlabel(y, x);
                                if (vd != va && vd != vb && vd != vc) {
                                    const int dx = x - t.ax, dy = y - t.ay;
                                    const long long inc = (long long)U * dx + (long long)V * dy - (long long)A2 * (long long)(dx * dx + dy * dy);
                                    if (inc > 0 || (inc == 0 && incircle_pert(va, vb, vc, vd, grid_w) > 0)) {
                                        const int ddx = x - qx, ddy = y - qy;
                                        best = ((unsigned long long)(uint32_t)(ddx * ddx + ddy * ddy) << 32) | (uint32_t)(x | (y << 16));
                                        break;
                                    }
                                }
                                if (take_l) pl = prev_bit(row, x - 1, x0); else pr = next_bit(row, x + 1, x1);
                            }
                        }
                    }
                }
            }
        }
        // nearest violator of the wave: one 32-bit min reduction on the squared distance, first lane that holds it
        const uint32_t d2 = (uint32_t)(best >> 32);  // 0xFFFFFFFF: none in this lane's row
        const uint32_t dmin = __reduce_min_sync(0xffffffffu, d2);
        if (dmin != 0xFFFFFFFFu) {
            const int src = __ffs(__ballot_sync(0xffffffffu, d2 == dmin)) - 1;
            const int v = __shfl_sync(0xffffffffu, (int)(uint32_t)best, src);
            cache = lane == src ? ~0ull : best;  // the other rows' violators are tried first against the next circle
            return v;
        }
        // convexity of circle /\ hull: a dead row kills every row beyond it in its direction
        if (__ballot_sync(0xffffffffu, row_dead && !down)) up_dead = true;
        if (__ballot_sync(0xffffffffu, row_dead && down)) dn_dead = true;
        if (up_dead && dn_dead) return -1;
    }
}

// ---- stage 1: winners -> occupancy / non-empty bit rows, sparse image (site colours, rest zero) ---------------------------
// A CTA per SITES_WARPS consecutive BEV rows, a warp per row.  The rows' keys are one contiguous, 16-byte aligned piece of the
// key grid: one thread brings it to shared memory with a single bulk copy (TMA engine, mbarrier completion, evict-first in L2),
// so a row pays one memory round trip for its keys instead of one per batch, and the colour gathers of a batch (SITES_BATCH
// words per lane) are all in flight together.  Lane k keeps the bit words of word k of the current 32-word chunk, so that the row
// summary (count, first, last) and the stores of the bit rows cost a few instructions per row instead of per word.  The 3-byte
// pixels of the row are packed into words with two shuffles per 32 pixels, staged in shared memory and written out as aligned
// 16-byte vectors (an image row is 1 503 bytes at an arbitrary alignment).
__global__ void __launch_bounds__(SITES_WARPS * 32) sites_stage_kernel(ImageArgs A) {
    extern __shared__ __align__(16) uint32_t sites_smem[];
    __shared__ __align__(8) uint64_t s_bar[2];
    const int img = blockIdx.y;
    const int h = A.G.grid_h, w = A.G.grid_w, wpr = A.G.wpr;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned FULL = 0xffffffffu;
    uint32_t* keygrid = A.keygrid + (size_t)img * A.keygrid_stride;
    const size_t kwords = sites_key_words(w);
    uint32_t* sb = sites_smem + 2 * kwords + (size_t)warp * sites_row_words(wpr);  // this warp's packed output row
    // groups of SITES_WARPS rows this CTA works through, the keys of the next group in flight while the current one is processed
    const int g_first = blockIdx.x * SITES_GROUPS;
    const int n_groups = min(SITES_GROUPS, (h + SITES_WARPS - 1) / SITES_WARPS - g_first);
    auto fetch = [&](int g) {  // thread 0: keys of group g -> buffer g & 1
        const int r0 = (g_first + g) * SITES_WARPS;
        const int rows_here = min(SITES_WARPS, h - r0);
        const uint32_t bytes = ((uint32_t)(rows_here * w) * 4u + 15u) & ~15u;  // may reach into the padding behind the image's grid
        mbar_expect_tx(&s_bar[g & 1], bytes);
        bulk_g2s_stream(sites_smem + (size_t)(g & 1) * kwords, keygrid + (size_t)r0 * w, bytes, &s_bar[g & 1]);
    };
    if (threadIdx.x == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); }
    __syncthreads();
    if (threadIdx.x == 0) { fetch(0); if (n_groups > 1) fetch(1); }
    const bool raw = A.raw_mode != 0;
    uint32_t* occ = A.planes + (size_t)img * 3 * A.plane_stride;
    uint32_t* nonempty = occ + A.plane_stride;
    // word of the packed row this lane builds from 32 pixels: bytes 4*lane .. 4*lane+3 of the 96 = pixels p0 and p0 + 1
    const int p0 = (lane * 4) / 3, o8 = ((lane * 4) % 3) * 8;
    const uint8_t* csrc = A.color_src[img];
    int dst;
    uint8_t* out = image_out(A, img, dst);

  for (int g = 0; g < n_groups; g++) {
    const int r = (g_first + g) * SITES_WARPS + warp;
    mbar_wait(&s_bar[g & 1], (uint32_t)(g >> 1) & 1u);
    if (r < h) {
    int running = 0, first = -1, last = -1, ne_cnt = 0;
    const uint32_t* kr = sites_smem + (size_t)(g & 1) * kwords + warp * w + lane;  // this lane's key of word 0 (shared memory)
    uint32_t* kp = keygrid + (size_t)r * w + lane;                                 // ... and its home in the key grid
    for (int wc = 0; wc < wpr; wc += 32) {  // chunks of 32 words (one chunk for grids up to 1 024 pixels wide)
        uint32_t my_ob = 0u, my_nb = 0u;
        const int wend = min(wpr, wc + 32);
        for (int wi0 = wc; wi0 < wend; wi0 += SITES_BATCH, kr += SITES_BATCH * 32, kp += SITES_BATCH * 32) {
            uint32_t key[SITES_BATCH], col[SITES_BATCH];  // col: rgb in bits 0..23, bit 31 = the pixel is a site
            const int c0 = wi0 * 32 + lane;
            uint32_t any = 0u;
#pragma unroll
            for (int j = 0; j < SITES_BATCH; j++) { key[j] = (c0 + j * 32 < w) ? kr[j * 32] : 0u; any |= key[j]; }
            if (!__any_sync(FULL, any != 0u)) {  // nothing here (70 % of an image lies outside the footprint): zero bytes, zero bits
                if (lane < 24) {
#pragma unroll
                    for (int j = 0; j < SITES_BATCH; j++) sb[(wi0 + j) * 24 + lane] = 0u;
                }
                continue;
            }
#if !(SITES_DBG & 4)
            if (A.clear_keys) {
#if SITES_CLEAR_WARP
#pragma unroll
                for (int j = 0; j < SITES_BATCH; j++) if (__any_sync(FULL, key[j] != 0u) && c0 + j * 32 < w) kp[j * 32] = 0u;  // whole sectors: no read-modify-write in L2
#else
#pragma unroll
                for (int j = 0; j < SITES_BATCH; j++) if (key[j]) kp[j * 32] = 0u;
#endif
            }
#endif
#pragma unroll
            for (int j = 0; j < SITES_BATCH; j++) {
                col[j] = 0u;
#if SITES_DBG & 1
                if (key[j]) col[j] = (key[j] & 0xFFFFFFu) | 0x80000000u;
#else
                if (key[j]) col[j] = gather_rgb(csrc, (key[j] - 1u) & KEY_IDX_MASK, A.pano_w) | 0x80000000u;
#endif
            }
#pragma unroll
            for (int j = 0; j < SITES_BATCH; j++) {
                const uint32_t cr = col[j] & 0xFF, cg = (col[j] >> 8) & 0xFF, cb = (col[j] >> 16) & 0xFF;
                const bool ne = ((cr * cg * cb) & 0xFFu) != 0u;  // uint8 product wraps (interpolation_utils.py:95)
                const uint32_t ob = __ballot_sync(FULL, (col[j] >> 31) != 0u);
                const uint32_t nb = __ballot_sync(FULL, ne);
                if (lane == ((wi0 + j) & 31)) { my_ob = ob; my_nb = nb; }
                const uint32_t c24 = col[j] & 0xFFFFFFu;
                const uint32_t a = __shfl_sync(FULL, c24, p0 & 31), b = __shfl_sync(FULL, c24, (p0 + 1) & 31);
                if (lane < 24) sb[(wi0 + j) * 24 + lane] = (a >> o8) | (b << (24 - o8));
            }
        }
        // bit rows of the chunk and the row summary
        if (wc + lane < wpr) { occ[r * wpr + wc + lane] = my_ob; nonempty[r * wpr + wc + lane] = my_nb; }
        const uint32_t nz = __ballot_sync(FULL, my_ob != 0u);
        if (nz) {
            const int fw = __ffs(nz) - 1, lw = 31 - __clz(nz);
            const uint32_t fo = __shfl_sync(FULL, my_ob, fw), lo = __shfl_sync(FULL, my_ob, lw);
            if (first < 0) first = (wc + fw) * 32 + __ffs(fo) - 1;
            last = (wc + lw) * 32 + 31 - __clz(lo);
            running += __reduce_add_sync(FULL, __popc(my_ob));
            ne_cnt += __reduce_add_sync(FULL, __popc(my_nb));
        }
    }
    if (lane == 0) {
        row_arr(A, img, RA_CNT)[r] = (int16_t)running; row_arr(A, img, RA_FIRST)[r] = (int16_t)first;
        row_arr(A, img, RA_LAST)[r] = (int16_t)last; row_arr(A, img, RA_NE)[r] = (int16_t)ne_cnt;
    }
    __syncwarp();
    // ---- the staged row -> global memory: head bytes, aligned 16-byte vectors, tail bytes
    uint8_t* orow = out + (size_t)(raw ? r : (h - 1 - r)) * w * 3;
    const uint8_t* sb8 = reinterpret_cast<const uint8_t*>(sb);
    const int nb = 3 * w;
    int head = (int)((16u - (uint32_t)((uintptr_t)orow & 15u)) & 15u);
    if (head > nb) head = nb;
    const int nv = (nb - head) >> 4, tail0 = head + (nv << 4);
    const int sw0 = head >> 2, sh = (head & 3) * 8;
#if SITES_DBG & 2
    if (nv < 0)
#endif
    for (int k = lane; k < nv; k += 32) {
        const uint32_t* s = sb + sw0 + 4 * k;
        const uint32_t a0 = s[0], a1 = s[1], a2 = s[2], a3 = s[3], a4 = s[4];  // s[4]: padding words follow the row
        uint4 v;
        v.x = __funnelshift_r(a0, a1, sh); v.y = __funnelshift_r(a1, a2, sh); v.z = __funnelshift_r(a2, a3, sh); v.w = __funnelshift_r(a3, a4, sh);
        IMAGE_OUT_ST(reinterpret_cast<uint4*>(orow + head + 16 * k), v);
    }
    if (lane < head) orow[lane] = sb8[lane];
    if (tail0 + lane < nb) orow[tail0 + lane] = sb8[tail0 + lane];
    }  // r < h
    if (g + 2 < n_groups) {  // every warp is done with buffer g & 1: refill it with the keys of group g + 2
        __syncthreads();
        if (threadIdx.x == 0) fetch(g + 2);
    }
    __syncwarp();  // the packed row is read back by other lanes before the next group overwrites it
  }
}

// ---- stage 2: guards, hull, keep mask, edge rule, query list: a CTA per image -------------------------------------------
__global__ void __launch_bounds__(PREP_NT) prep_stage_kernel(ImageArgs A) {
    const int img = blockIdx.x;
    const int h = A.G.grid_h, w = A.G.grid_w, wpr = A.G.wpr;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwords = h * wpr;
    const unsigned FULL = 0xffffffffu;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    ImageShared S;
    {
        unsigned char* p = smem_raw;
        S.tmp = (uint32_t*)p; p += (size_t)nwords * 4;
        const size_t rows = image_row_bytes(h);
        int16_t** arr[13] = {&S.cnt, &S.first, &S.last, &S.up, &S.dn, &S.hlo, &S.hhi, &S.hl0, &S.hl1, &S.hr0, &S.hr1, &S.stk_l, &S.stk_r};
        for (int i = 0; i < 13; i++) { *arr[i] = (int16_t*)p; p += rows; }
        const size_t frows = ((size_t)h * 4 + 15) & ~(size_t)15;
        S.hlf = (float*)p; p += frows; S.hrf = (float*)p; p += frows;
    }
    uint32_t* planes = A.planes + (size_t)img * 3 * A.plane_stride;
    S.occ = planes;                            // global, read-only here (L1)
    const uint32_t* nonempty = planes + A.plane_stride;
    S.keep = planes + 2 * A.plane_stride;      // written here
    __shared__ uint32_t s_rownz[32];  // rows with at least one site, one bit per row (grid_h <= 1023)
    __shared__ int s_S, s_M, s_mincol, s_maxcol, s_ne_cnt, s_keep_cnt, s_nitems, s_hull_ok, s_nedge, s_masked;

    int32_t* counts = A.counts + (size_t)img * 8;
    const bool raw = A.raw_mode != 0;
    if (tid == 0) {
        s_S = 0; s_M = 0; s_mincol = 1 << 30; s_maxcol = -1; s_ne_cnt = 0; s_keep_cnt = 0; s_nitems = 0;
        s_hull_ok = 0; s_nedge = 0; s_masked = 0;
    }
    __syncthreads();
    // row summaries of the sites stage
    {
        const int16_t *g_cnt = row_arr(A, img, RA_CNT), *g_first = row_arr(A, img, RA_FIRST), *g_last = row_arr(A, img, RA_LAST),
                      *g_ne = row_arr(A, img, RA_NE);
        int nS = 0, nM = 0, mn = 1 << 30, mx = -1, ne = 0;
        for (int r = tid; r < h; r += PREP_NT) {
            const int c = g_cnt[r], f = g_first[r], l = g_last[r];
            S.cnt[r] = (int16_t)c; S.first[r] = (int16_t)f; S.last[r] = (int16_t)l;
            if (c > 0) { nS += c; nM++; mn = min(mn, f); mx = max(mx, l); }
            ne += g_ne[r];
        }
        nS = __reduce_add_sync(FULL, nS); nM = __reduce_add_sync(FULL, nM); ne = __reduce_add_sync(FULL, ne);
        mn = __reduce_min_sync(FULL, mn); mx = __reduce_max_sync(FULL, mx);
        if (lane == 0) {
            if (nS) { atomicAdd(&s_S, nS); atomicAdd(&s_M, nM); atomicMin(&s_mincol, mn); atomicMax(&s_maxcol, mx); }
            if (ne) atomicAdd(&s_ne_cnt, ne);
        }
    }
    __syncthreads();

    // ---- guards (interpolation_utils.py:37-42, 57-71) ---------------------------------------------
    const int nS = s_S, M = s_M;
    int status = 0;  // SALVE_BEV_IMG_OK
    if (!A.skip_empty_check && counts[1] == 0) status = 1;                         // EMPTY -> None
    else if (nS < 4 || M < 2 || s_mincol == s_maxcol) status = 2;                  // DEGENERATE -> zeros

    // nearest non-empty row above / below every row: one bit per row (a warp ballot per 32 rows), then two bit scans per row
    for (int r0 = warp * 32; r0 < h; r0 += PREP_NT) {
        const uint32_t nz = __ballot_sync(FULL, r0 + lane < h && S.cnt[r0 + lane] > 0);
        if (lane == 0) s_rownz[r0 >> 5] = nz;
    }
    __syncthreads();
    for (int r = tid; r < h; r += PREP_NT) {
        const int u = next_bit(s_rownz, r + 1, h - 1), d = prev_bit(s_rownz, r - 1, 0);
        S.up[r] = (int16_t)u; S.dn[r] = (int16_t)d;
        S.hlo[r] = 1; S.hhi[r] = 0;  // rows outside the hull: empty range
        S.hl0[r] = 0; S.hl1[r] = 0; S.hr0[r] = 0; S.hr1[r] = 0;
        S.hlf[r] = __int_as_float(0x7f800000); S.hrf[r] = __int_as_float(0xff800000);
    }

    // ---- exact convex hull from the per-row first / last sites: two monotone chains (warps 0 and 1) ---------------
    // The warp pre-filters the rows, lane 0 builds the chain from the survivors; the whole warp then fills the per-row bounds edge by edge.
    __syncthreads();
    if (status == 0 && warp < 2) {
        const bool left = warp == 0;
        const int16_t* xs = left ? S.first : S.last;
        int16_t* e0 = left ? S.hl0 : S.hr0;   // per row: rows of the two end points of the hull edge crossing it
        int16_t* e1 = left ? S.hl1 : S.hr1;
        int16_t* stk = left ? S.stk_l : S.stk_r;
        int16_t* bound = left ? S.hlo : S.hhi;
        float* bf = left ? S.hlf : S.hrf;
        // Parallel pre-filter: a row's extreme site that lies on or inside the segment between the extreme sites of two other
        // rows (here: the rows at distance 1, 2, 4, ... on either side) is not a strict vertex of this chain.  What survives
        // (the corners and a few dozen rows of a ragged wall) goes to the serial monotone-chain scan in row order.
        int16_t* cand = e0;  // free until the fill below
        int ncand = 0;
        for (int r0 = 0; r0 < h; r0 += 32) {
            const int r = r0 + lane;
            bool alive = r < h && S.cnt[r] > 0;
            if (alive) {
                const int x = xs[r];
                for (int d = 1; d < h && alive; d <<= 1) {
                    const int ra = r - d, rb = r + d;
                    if (ra < 0 || rb >= h) break;
                    if (S.cnt[ra] > 0 && S.cnt[rb] > 0) {
                        const int o = orient_i(xs[ra], ra, x, r, xs[rb], rb);
                        if (left ? (o >= 0) : (o <= 0)) alive = false;
                    }
                }
            }
            const uint32_t m = __ballot_sync(FULL, alive);
            if (alive) cand[ncand + __popc(m & ((1u << lane) - 1u))] = (int16_t)r;
            ncand += __popc(m);
        }
        __syncwarp();
        int top = 0;
        if (lane == 0) {
            for (int i = 0; i < ncand; i++) {
                const int r = cand[i];
                const int x = xs[r];
                while (top >= 2) {
                    const int r1 = stk[top - 1], r0 = stk[top - 2];
                    const int o = orient_i(xs[r0], r0, xs[r1], r1, x, r);
                    // going up, the left chain turns clockwise at every vertex and the right chain counter-clockwise
                    if (left ? (o >= 0) : (o <= 0)) top--; else break;
                }
                stk[top++] = (int16_t)r;
            }
            if (top <= 2) atomicAdd(&s_hull_ok, left ? 1 : 2);  // this chain has no interior vertex
        }
        top = __shfl_sync(FULL, top, 0);
        __syncwarp();
        for (int r = lane; r < h; r += 32) { e0[r] = 0; e1[r] = 0; }  // the pre-filter's scratch: rows outside the hull read 0
        __syncwarp();
        for (int k = 1; k < top; k++) {
            const int r0 = stk[k - 1], r1 = stk[k];
            const int x0 = xs[r0], x1 = xs[r1], dy = r1 - r0;
            for (int rr = r0 + lane; rr <= r1; rr += 32) {
                const int num = x0 * dy + (x1 - x0) * (rr - r0);  // >= 0: a convex combination scaled by dy
                bound[rr] = (int16_t)(left ? (num + dy - 1) / dy : num / dy);  // ceil on the left, floor on the right
                e0[rr] = (int16_t)r0; e1[rr] = (int16_t)r1;
                bf[rr] = (float)num / (float)dy;
            }
            __syncwarp();  // the shared end row of consecutive edges is written by both: keep their order
        }
    }

    // ---- keep mask = Chebyshev dilation of `nonempty` by K/2, zero padded: rows first (into shared memory) -------------
    const int rad = A.G.K / 2;
    if (!raw) {
        for (int item = tid; item < nwords; item += PREP_NT) {
            const int r = item / wpr, wi = item - r * wpr;
            const uint32_t cur = __ldg(nonempty + item);
            const uint32_t prv = (wi > 0) ? __ldg(nonempty + item - 1) : 0u;
            const uint32_t nxt = (wi + 1 < wpr) ? __ldg(nonempty + item + 1) : 0u;
            const unsigned long long L = ((unsigned long long)cur << 32) | prv, R = ((unsigned long long)nxt << 32) | cur;
            uint32_t o = cur;
            for (int d = 1; d <= rad; d++) o |= (uint32_t)(L >> (32 - d)) | (uint32_t)(R >> d);
            const int valid = w - wi * 32;
            if (valid < 32) o &= (1u << valid) - 1u;
            S.tmp[item] = o;
        }
    }
    __syncthreads();
    // all sites on one oblique line: every row has one site and neither chain has an interior vertex
    if (status == 0 && s_hull_ok == 3 && nS == M) status = 3;  // COLLINEAR: the reference's Qhull call raises

    // ---- columns of the dilation, fused with the EDGE RULE and the work list of query pixels (row << 11 | col) ----------
    // query = kept, not a site, inside the closed hull.
    // Edge rule: a query whose W and E (or N and S) neighbours are both sites lies at the midpoint of a Delaunay edge --
    // the circle of radius 1 around it has no lattice point strictly inside but the query itself -- so its interpolated value is
    // the exact mean of the two site colours whichever triangle it is attributed to (third barycentric weight 0).  With all
    // four neighbours present they are co-circular and the canonical diagonal is decided by the symbolic perturbation
    // (incircle_pert on (E, N, W, S) = 2 * (wE + wW - wN - wS): positive keeps N-S).  About half of all queries end here.
    uint32_t* qlist = A.qlist + (size_t)img * A.qlist_stride;
    uint32_t* elist = A.clist + (size_t)img * A.qlist_stride;  // edge-rule pixels for the shade stage (the cooperative pass's list is built later)
    const bool edge_rule = A.qtri == nullptr;  // the triangle tap wants every query resolved to a triangle
    // With the local rule the queries first go to a scratch list (the free top end of the edge list's array, downwards: queries
    // and edge-rule pixels are disjoint, so both fit), from which the local stage feeds the window pass's list.
    const bool local_rule = edge_rule && A.local_lut != nullptr && status == 0;
    const int cap = (int)A.qlist_stride;
    {
        int kc = 0;
        const int nw_pad = (nwords + 31) & ~31;
        for (int item = tid; item < nw_pad; item += PREP_NT) {  // a warp takes 32 consecutive words
            uint32_t q = 0, he = 0, ve = 0;
            int r = 0, wi = 0;
            if (item < nwords) {
                r = item / wpr; wi = item - r * wpr;
                uint32_t kp;
                if (!raw) {
                    kp = 0u;
                    const int r0 = max(r - rad, 0), r1 = min(r + rad, h - 1);
                    for (int rr = r0; rr <= r1; rr++) kp |= S.tmp[rr * wpr + wi];
                    kc += __popc(kp);
                } else {
                    kp = range_mask(wi, 0, w - 1);  // raw mode: every pixel is kept
                }
                S.keep[item] = kp;
                if (status == 0) {
                    const uint32_t hm = range_mask(wi, S.hlo[r], S.hhi[r]);
                    const uint32_t oc = __ldg(S.occ + item);
                    q = kp & ~oc & hm;
                    if (oc & ~kp) s_masked = 1;  // a site the hallucination mask removes: the finish stage has work (rare)
                    if (A.hull && hm) {
                        uint8_t* hp = A.hull + (size_t)img * A.hull_stride + (size_t)r * w + wi * 32;
                        uint32_t m = hm;
                        while (m) { const int b = __ffs(m) - 1; m &= m - 1; hp[b] = 1; }
                    }
                    if (q && edge_rule) {
                        const uint32_t ol = wi > 0 ? __ldg(S.occ + item - 1) : 0u, orr = wi + 1 < wpr ? __ldg(S.occ + item + 1) : 0u;
                        he = q & ((oc << 1) | (ol >> 31)) & ((oc >> 1) | (orr << 31));
                        ve = q & (r + 1 < h ? __ldg(S.occ + item + wpr) : 0u) & (r > 0 ? __ldg(S.occ + item - wpr) : 0u);
                        q &= ~(he | ve);
                    }
                }
            }
            // one warp scan for both lists: queries in the low half, edge-rule pixels in the high half (<= 1024 each per warp)
            const uint32_t e = he | ve;
            const int n = __popc(q) | (__popc(e) << 16);
            int incl = n;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += t; }
            const int total = __shfl_sync(FULL, incl, 31);
            if (total == 0) continue;
            int base_q = 0, base_e = 0;
            if (lane == 0) {
                if (total & 0xFFFF) base_q = atomicAdd(&s_nitems, total & 0xFFFF);
                if (total >> 16) base_e = atomicAdd(&s_nedge, total >> 16);
            }
            const int excl = incl - n;
            base_q = __shfl_sync(FULL, base_q, 0) + (excl & 0xFFFF);
            base_e = __shfl_sync(FULL, base_e, 0) + (excl >> 16);
            const uint32_t code0 = ((uint32_t)r << COL_BITS) | (uint32_t)(wi * 32);
            if (local_rule) { while (q) { const int b = __ffs(q) - 1; q &= q - 1; elist[cap - 1 - base_q++] = code0 + b; } }
            else { while (q) { const int b = __ffs(q) - 1; q &= q - 1; qlist[base_q++] = code0 + b; } }
            uint32_t m = e;
            while (m) {  // bit 31: W-E pair present, bit 30: N-S pair present
                const int b = __ffs(m) - 1; m &= m - 1;
                elist[base_e++] = (code0 + b) | (((he >> b) & 1u) << 31) | (((ve >> b) & 1u) << 30);
            }
        }
        if (!raw) {
            kc = __reduce_add_sync(FULL, kc);
            if (lane == 0 && kc) atomicAdd(&s_keep_cnt, kc);
        }
    }
    // row arrays the finish stage needs
    {
        int16_t* dsts[6] = {row_arr(A, img, RA_UP), row_arr(A, img, RA_DN), row_arr(A, img, RA_HL0), row_arr(A, img, RA_HL1),
                            row_arr(A, img, RA_HR0), row_arr(A, img, RA_HR1)};
        const int16_t* srcs[6] = {S.up, S.dn, S.hl0, S.hl1, S.hr0, S.hr1};
        float *g_hlf = row_arr_f(A, img, RA_HLF), *g_hrf = row_arr_f(A, img, RA_HRF);
        for (int r = tid; r < h; r += PREP_NT) {
#pragma unroll
            for (int k = 0; k < 6; k++) dsts[k][r] = srcs[k][r];
            g_hlf[r] = S.hlf[r]; g_hrf[r] = S.hrf[r];
        }
    }
    __syncthreads();
    if (tid == 0) {
        counts[2] = nS; counts[3] = s_ne_cnt; counts[4] = raw ? 0 : s_keep_cnt; counts[5] = 0; counts[6] = 0; counts[7] = 0;
        int32_t* hd = A.hdr + (size_t)img * HD_STRIDE;
        // with the local rule the queries are still in the scratch list: the local stage fills the window list and counts it
        hd[HD_STATUS] = status; hd[HD_NQ] = status == 0 && !local_rule ? s_nitems : 0; hd[HD_EDGE] = status == 0 ? s_nedge : 0;
        hd[HD_MASKED] = s_masked; hd[HD_PEND] = 0; hd[HD_XTRA] = 0; hd[HD_LOCAL] = 0; hd[HD_NRAW] = local_rule ? s_nitems : 0;
    }
}

// ---- exact barycentric value of a query pixel from its final triangle (local stage, shade stage, cooperative pass) ----------------------
__device__ __forceinline__ uint32_t site_rgb_at(const uint8_t* out, bool raw, int h, int w, int sx, int sy) {
    return load_rgb(out + ((size_t)(raw ? sy : h - 1 - sy) * w + sx) * 3);
}
__device__ __forceinline__ void write_px_at(uint8_t* out, int32_t* qtri, bool raw, int h, int w, const Tri2& t, uint32_t ca, uint32_t cb,
                                            uint32_t cc, int x, int r) {
    const uint32_t ua = (uint32_t)orient_i(t.bx, t.by, t.cx, t.cy, x, r), ub = (uint32_t)orient_i(t.cx, t.cy, t.ax, t.ay, x, r),
                   uc = (uint32_t)orient_i(t.ax, t.ay, t.bx, t.by, x, r);
    const uint32_t A2 = ua + ub + uc;
    uint8_t* o = out + ((size_t)(raw ? r : h - 1 - r) * w + x) * 3;
    o[0] = (uint8_t)((ua * (ca & 0xFF) + ub * (cb & 0xFF) + uc * (cc & 0xFF)) / A2);
    o[1] = (uint8_t)((ua * ((ca >> 8) & 0xFF) + ub * ((cb >> 8) & 0xFF) + uc * ((cc >> 8) & 0xFF)) / A2);
    o[2] = (uint8_t)((ua * ((ca >> 16) & 0xFF) + ub * ((cb >> 16) & 0xFF) + uc * ((cc >> 16) & 0xFF)) / A2);
    if (qtri) {
        // ascending vertex ids: two warps of the cooperative pass may reach the same triangle with different rotations and
        // write the same pixel concurrently; in canonical order their stores are identical word for word
        int32_t* q = qtri + ((size_t)r * w + x) * 3;
        const int i0 = t.ay * w + t.ax, i1 = t.by * w + t.bx, i2 = t.cy * w + t.cx;
        const int lo = min(i0, min(i1, i2)), hi = max(i0, max(i1, i2));
        q[0] = lo; q[1] = i0 + i1 + i2 - lo - hi; q[2] = hi;
    }
}

// ---- local stage (see LocalRule): the queries of the prep stage's scratch list ------------------------------------------------
// Resolved queries are shaded on the spot (three colour gathers, exact integer barycentrics), the rest is appended to the window
// pass's list.  grid = (LOCAL_SPLIT, images).
#ifndef LOCAL_SPLIT_DEF
#define LOCAL_SPLIT_DEF 16
#endif
constexpr int LOCAL_SPLIT = LOCAL_SPLIT_DEF;
constexpr int LOCAL_NT = 256;
// Four in ten queries have a candidate with sites ON its circle and go through the perturbation tests, a loop of a few to a dozen
// trips; with one thread per query the warp walks through the longest of them at 5 to 10 lanes (measured: 534 us per C2 step
// against 454 us this way).  So a warp takes LOCAL_Q entries per lane and round: the straight-line part (neighbourhood, pattern, the four
// inside masks, the first candidate's on-circle mask) classifies them into "no candidate" (window list), "first candidate has an
// empty circle and nothing on it" (done) and "jobs"; the jobs are packed into a warp-private array and every lane then takes
// one job per trip.  The results of a round are staged per warp: one atomic for the window list's slots, and the resolved
// queries are shaded one per lane, every gather of a trip in flight together.
#ifndef LOCAL_Q_DEF
#define LOCAL_Q_DEF 4
#endif
constexpr int LOCAL_Q = LOCAL_Q_DEF;
#ifndef LOCAL_MIN_CTAS
#define LOCAL_MIN_CTAS 5
#endif
__global__ void __launch_bounds__(LOCAL_NT, LOCAL_MIN_CTAS) local_stage_kernel(ImageArgs A) {
    const int img = blockIdx.y;
    int32_t* hd = A.hdr + (size_t)img * HD_STRIDE;
    const int n = hd[HD_NRAW];
    if (n == 0) return;
    const unsigned FULL = 0xffffffffu;
    const int h = A.G.grid_h, w = A.G.grid_w, wpr = A.G.wpr;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cap = (int)A.qlist_stride;
    const uint32_t* lut = A.local_lut;
    const uint32_t* occ = A.planes + (size_t)img * 3 * A.plane_stride;
    const uint32_t* raw_list = A.clist + (size_t)img * A.qlist_stride;
    uint32_t* qlist = A.qlist + (size_t)img * A.qlist_stride;
    const bool raw = A.raw_mode != 0;
    int dst;
    uint8_t* out = image_out(A, img, dst);
    constexpr int RQ = 32 * LOCAL_Q;  // entries of a warp's round
    __shared__ uint32_t s_job_code[LOCAL_NT / 32][RQ], s_job_nb[LOCAL_NT / 32][RQ], s_job_e[LOCAL_NT / 32][RQ];
    __shared__ uint32_t s_win[LOCAL_NT / 32][RQ], s_loc_code[LOCAL_NT / 32][RQ];
    __shared__ unsigned long long s_loc_tri[LOCAL_NT / 32][RQ];
    uint32_t *job_code = s_job_code[warp], *job_nb = s_job_nb[warp], *job_e = s_job_e[warp], *win = s_win[warp], *loc_code = s_loc_code[warp];
    unsigned long long* loc_tri = s_loc_tri[warp];
    const uint32_t below = (1u << lane) - 1u;
    auto tri_of = [&](uint32_t verts, int x, int r) {
        const int ia = (int)(verts & 31u), ib = (int)((verts >> 5) & 31u), ic = (int)((verts >> 10) & 31u);
        const int ay = (ia * 13 >> 6) - 2, ax = ia - 5 * (ay + 2) - 2, by = (ib * 13 >> 6) - 2, bx = ib - 5 * (by + 2) - 2,
                  cy = (ic * 13 >> 6) - 2, cx = ic - 5 * (cy + 2) - 2;  // i / 5 = i * 13 >> 6 for i < 25
        return QRES_DONE | (unsigned long long)vlabel(r + ay, x + ax) | ((unsigned long long)vlabel(r + by, x + bx) << 21) |
               ((unsigned long long)vlabel(r + cy, x + cx) << 42);
    };
    const int n_warps = LOCAL_SPLIT * (LOCAL_NT / 32);
    for (int base = (blockIdx.x * (LOCAL_NT / 32) + warp) * RQ; base < n; base += n_warps * RQ) {
        int n_job = 0, n_win = 0, n_loc = 0;  // warp-uniform
        // ---- classify LOCAL_Q entries per lane
#pragma unroll
        for (int u = 0; u < LOCAL_Q; u++) {
            const int i = base + u * 32 + lane;
            const bool valid = i < n;
            uint32_t code = 0u, nb = 0u, e = 0xFFFFFFFFu, surv = 0u, ties = 0u, verts = 0u;
            if (valid) {
                code = __ldg(raw_list + cap - 1 - i);
                const int x = (int)(code & COL_MASK), r = (int)(code >> COL_BITS);
                // 5 x 5 neighbourhood bits, branch-free with clamped row / word indices and masks for what lies outside the grid
                const int c0 = x - 2;
                const int w0 = c0 >> 5, sh = c0 & 31;  // c0 may be negative: arithmetic shift = floor
                const int wlo = max(w0, 0), whi = min(w0 + 1, wpr - 1);
                const uint32_t mlo = w0 >= 0 ? 0xFFFFFFFFu : 0u, mhi = w0 + 1 < wpr ? 0xFFFFFFFFu : 0u;
#pragma unroll
                for (int k = 0; k < 5; k++) {
                    const int y = r + k - 2;
                    const int yc = min(max(y, 0), h - 1);
                    const uint32_t vm = y == yc ? 0xFFFFFFFFu : 0u;
                    const uint32_t* row = occ + yc * wpr;
                    nb |= (__funnelshift_r(__ldg(row + wlo) & mlo & vm, __ldg(row + whi) & mhi & vm, sh) & 31u) << (5 * k);
                }
                e = __ldg(lut + LocalRule::pattern_of(nb));
                // candidates with no site strictly inside their circle: all four slots at once (unused slots hold 0xFF: the
                // mask index wraps to a table entry, the id test discards it)
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint32_t cid = (e >> (8 * k)) & 0xFFu;
                    const uint32_t inside = __ldg(lut + LocalRule::OFF_INSIDE + (cid & (LocalRule::MAXCAND - 1)));
                    if (cid != 0xFFu && !(nb & inside)) surv |= 1u << k;
                }
                if (surv) {
                    const uint32_t cid = (e >> (8 * (__ffs(surv) - 1))) & 0xFFu;
                    ties = nb & __ldg(lut + LocalRule::OFF_ON + cid);
                    verts = __ldg(lut + LocalRule::OFF_VERTS + cid);
                }
            }
            const bool is_win = valid && !surv, is_loc = valid && surv && !ties, is_job = valid && surv && ties;
            const uint32_t mw = __ballot_sync(FULL, is_win), ml = __ballot_sync(FULL, is_loc), mj = __ballot_sync(FULL, is_job);
            if (is_win) win[n_win + __popc(mw & below)] = code;
            if (is_loc) {
                const int k = n_loc + __popc(ml & below);
                loc_code[k] = code; loc_tri[k] = tri_of(verts, (int)(code & COL_MASK), (int)(code >> COL_BITS));
            }
            if (is_job) {
                const int k = n_job + __popc(mj & below);
                job_code[k] = code | (surv << 24); job_nb[k] = nb; job_e[k] = e;
            }
            n_win += __popc(mw); n_loc += __popc(ml); n_job += __popc(mj);
        }
        __syncwarp();
        // ---- the jobs, one per lane and trip
        for (int j0 = 0; j0 < n_job; j0 += 32) {
            const bool valid = j0 + lane < n_job;
            bool solved = false;
            uint32_t code = 0u;
            unsigned long long tri = 0ull;
            if (valid) {
                const uint32_t cs = job_code[j0 + lane], nb = job_nb[j0 + lane], e = job_e[j0 + lane];
                code = cs & 0xFFFFFFu;
                uint32_t surv = cs >> 24;
                const int x = (int)(code & COL_MASK), r = (int)(code >> COL_BITS);
                const int qi = r * w + x;
                while (surv && !solved) {
                    const int k = __ffs(surv) - 1; surv &= surv - 1;
                    const uint32_t cid = (e >> (8 * k)) & 0xFFu;
                    const uint32_t verts = __ldg(lut + LocalRule::OFF_VERTS + cid);
                    uint32_t ties = nb & __ldg(lut + LocalRule::OFF_ON + cid);
                    bool ok = true;
                    if (ties) {
                        // Sites on the circle: symbolic perturbation, as incircle_pert().  Its first three terms are linear in
                        // the tested point d: wa * orient(b,c,d) - wb * orient(a,c,d) + wc * orient(a,b,d) = PA * dx + PB * dy + PC
                        // (weights < 2^20, |coordinates| <= 2: every term and the sum fit int32).
                        const int ia = (int)(verts & 31u), ib = (int)((verts >> 5) & 31u), ic = (int)((verts >> 10) & 31u);
                        const int ay = (ia * 13 >> 6) - 2, ax = ia - 5 * (ay + 2) - 2, by = (ib * 13 >> 6) - 2, bx = ib - 5 * (by + 2) - 2,
                                  cy = (ic * 13 >> 6) - 2, cx = ic - 5 * (cy + 2) - 2;
                        const int wa = pert_weight_idx((uint32_t)(qi + ay * w + ax)), wb = pert_weight_idx((uint32_t)(qi + by * w + bx)),
                                  wc = pert_weight_idx((uint32_t)(qi + cy * w + cx));
                        const int PA = -wa * (cy - by) + wb * (cy - ay) - wc * (by - ay);
                        const int PB = wa * (cx - bx) - wb * (cx - ax) + wc * (bx - ax);
                        const int PC = wa * (bx * cy - by * cx) - wb * (ax * cy - ay * cx) + wc * (ax * by - ay * bx);
                        const int oabc = (bx - ax) * (cy - ay) - (by - ay) * (cx - ax);
                        while (ties && ok) {
                            const int b = __ffs(ties) - 1; ties &= ties - 1;
                            const int dy = (b * 13 >> 6) - 2, dx = b - 5 * (dy + 2) - 2;
                            const int wd = pert_weight_idx((uint32_t)(qi + dy * w + dx));
                            // > 0: the site is inside; = 0: a residual tie, left to the passes that know how the oracle breaks it
                            if (PA * dx + PB * dy + PC - wd * oabc >= 0) ok = false;
                        }
                    }
                    if (ok) { solved = true; tri = tri_of(verts, x, r); }
                }
            }
            const uint32_t ml = __ballot_sync(FULL, valid && solved), mw = __ballot_sync(FULL, valid && !solved);
            if (valid && solved) { const int k = n_loc + __popc(ml & below); loc_code[k] = code; loc_tri[k] = tri; }
            else if (valid) win[n_win + __popc(mw & below)] = code;
            n_loc += __popc(ml); n_win += __popc(mw);
        }
        // ---- write the round out
        int base_w = 0;
        if (lane == 0) {
            if (n_loc) atomicAdd(hd + HD_LOCAL, n_loc);  // (a count only: the finish stage reports the filled pixels)
            if (n_win) base_w = atomicAdd(hd + HD_NQ, n_win);
        }
        __syncwarp();
        for (int k = lane; k < n_loc; k += 32) {
            const uint32_t code = loc_code[k];
            const unsigned long long t3 = loc_tri[k];
            const uint32_t a = (uint32_t)t3 & M21, b = (uint32_t)(t3 >> 21) & M21, c = (uint32_t)(t3 >> 42) & M21;
            const Tri2 t = {vcol(a), vrow(a), vcol(b), vrow(b), vcol(c), vrow(c)};
            const uint32_t ca = site_rgb_at(out, raw, h, w, t.ax, t.ay), cb = site_rgb_at(out, raw, h, w, t.bx, t.by),
                           cc = site_rgb_at(out, raw, h, w, t.cx, t.cy);
            write_px_at(out, nullptr, raw, h, w, t, ca, cb, cc, (int)(code & COL_MASK), (int)(code >> COL_BITS));
        }
        base_w = __shfl_sync(FULL, base_w, 0);
        for (int k = lane; k < n_win; k += 32) qlist[base_w + k] = win[k];
        __syncwarp();  // the staging arrays are reused by the next round
    }
}

// ---- shade stage: one thread per list entry of the chunk, every lane busy -------------------------------------------------
// Entries of an image: its edge-rule pixels (the exact mean of the two site colours), then what the window pass resolved
// (three gathers, exact integer barycentrics).  grid = (SHADE_SPLIT, images).
#ifndef SHADE_UNROLL_DEF
#define SHADE_UNROLL_DEF 4
#endif
constexpr int SHADE_UNROLL = SHADE_UNROLL_DEF;
constexpr int SHADE_SPLIT = 8;
constexpr int SHADE_NT = 256;
__global__ void __launch_bounds__(SHADE_NT) shade_stage_kernel(ImageArgs A) {
    const int img = blockIdx.y;
    int32_t* hd = A.hdr + (size_t)img * HD_STRIDE;
    if (hd[HD_STATUS] != 0) return;
    const int n_edge = hd[HD_EDGE], n_win = hd[HD_NQ];
    const int h = A.G.grid_h, w = A.G.grid_w;
    const bool raw = A.raw_mode != 0;
    int dst;
    uint8_t* out = image_out(A, img, dst);
    int32_t* qtri = A.qtri ? A.qtri + (size_t)img * A.qtri_stride : nullptr;
    const uint32_t* elist = A.clist + (size_t)img * A.qlist_stride;
    uint32_t* qlist = A.qlist + (size_t)img * A.qlist_stride;
    unsigned long long* qres = A.qres + (size_t)img * A.qlist_stride;
    const ptrdiff_t dn = raw ? (ptrdiff_t)w * 3 : -(ptrdiff_t)w * 3;  // address step to row r + 1 in the (flipped) output
    constexpr int STRIDE = SHADE_SPLIT * SHADE_NT;
    // Every entry is a chain of dependent memory round trips (list entry -> colours -> store): SHADE_UNROLL entries per thread are in
    // flight together.
    for (int i0 = blockIdx.x * SHADE_NT + threadIdx.x; i0 < n_edge; i0 += SHADE_UNROLL * STRIDE) {
        uint32_t code[SHADE_UNROLL];
        uint8_t* p[SHADE_UNROLL];
        uint32_t ca[SHADE_UNROLL], cb[SHADE_UNROLL];
#pragma unroll
        for (int u = 0; u < SHADE_UNROLL; u++) code[u] = i0 + u * STRIDE < n_edge ? __ldg(elist + i0 + u * STRIDE) : 0u;  // 0: no pair bits = nothing
#pragma unroll
        for (int u = 0; u < SHADE_UNROLL; u++) {
            p[u] = nullptr;
            if (!(code[u] >> 30)) continue;
            const int x = (int)(code[u] & COL_MASK), r = (int)((code[u] >> COL_BITS) & 0x3FFu);
            bool horiz = (code[u] >> 31) != 0u;
            if (horiz && ((code[u] >> 30) & 1u)) {
                const uint32_t qi = (uint32_t)(r * w + x);
                const int wh = pert_weight_idx(qi - 1u) + pert_weight_idx(qi + 1u);
                const int wv = pert_weight_idx(qi - (uint32_t)w) + pert_weight_idx(qi + (uint32_t)w);
                if (wh == wv) {  // residual tie of the perturbation: the cooperative pass decides (an entry without a triangle)
                    const int slot = n_win + atomicAdd(hd + HD_XTRA, 1);
                    qlist[slot] = code[u] & ((1u << 21) - 1u);
                    qres[slot] = 0ull;
                    continue;
                }
                horiz = wh < wv;
            }
            p[u] = out + ((size_t)(raw ? r : h - 1 - r) * w + x) * 3;
            ca[u] = load_rgb(horiz ? p[u] - 3 : p[u] + dn);
            cb[u] = load_rgb(horiz ? p[u] + 3 : p[u] - dn);
        }
#pragma unroll
        for (int u = 0; u < SHADE_UNROLL; u++) {
            if (!p[u]) continue;
            p[u][0] = (uint8_t)(((ca[u] & 0xFF) + (cb[u] & 0xFF)) >> 1);
            p[u][1] = (uint8_t)((((ca[u] >> 8) & 0xFF) + ((cb[u] >> 8) & 0xFF)) >> 1);
            p[u][2] = (uint8_t)(((ca[u] >> 16) + (cb[u] >> 16)) >> 1);
        }
    }
    for (int j0 = blockIdx.x * SHADE_NT + threadIdx.x; j0 < n_win; j0 += SHADE_UNROLL * STRIDE) {
        unsigned long long rs[SHADE_UNROLL];
        uint32_t code[SHADE_UNROLL], ca[SHADE_UNROLL], cb[SHADE_UNROLL], cc[SHADE_UNROLL];
#pragma unroll
        for (int u = 0; u < SHADE_UNROLL; u++) {
            const int j = j0 + u * STRIDE;
            rs[u] = j < n_win ? qres[j] : 0ull;  // without QRES_DONE: handed on to the cooperative pass (or beyond the list)
            code[u] = j < n_win ? __ldg(qlist + j) : 0u;
        }
#pragma unroll
        for (int u = 0; u < SHADE_UNROLL; u++) {
            if (!(rs[u] & QRES_DONE)) continue;
            const uint32_t a = (uint32_t)rs[u] & M21, b = (uint32_t)(rs[u] >> 21) & M21, c = (uint32_t)(rs[u] >> 42) & M21;
            ca[u] = site_rgb_at(out, raw, h, w, vcol(a), vrow(a));
            cb[u] = site_rgb_at(out, raw, h, w, vcol(b), vrow(b));
            cc[u] = site_rgb_at(out, raw, h, w, vcol(c), vrow(c));
        }
#pragma unroll
        for (int u = 0; u < SHADE_UNROLL; u++) {
            if (!(rs[u] & QRES_DONE)) continue;
            const uint32_t a = (uint32_t)rs[u] & M21, b = (uint32_t)(rs[u] >> 21) & M21, c = (uint32_t)(rs[u] >> 42) & M21;
            const Tri2 t = {vcol(a), vrow(a), vcol(b), vrow(b), vcol(c), vrow(c)};
            write_px_at(out, qtri, raw, h, w, t, ca[u], cb[u], cc[u], (int)(code[u] & COL_MASK), (int)(code[u] >> COL_BITS));
        }
    }
}

// Hand-out order of the finish stage: longest expected image first, so that the launch does not end with a few CTAs working on
// long images while the rest of the GPU idles.  The predictor is what the stage's cooperative pass has to do: the queries the
// window pass handed on (ties by the number of window queries, which the stage shades).  rank = number of images that come first.
__global__ void __launch_bounds__(256) image_order_kernel(const int32_t* __restrict__ hdr, int n, int32_t* __restrict__ order) {
    __shared__ uint32_t s_key[WIN_MAX_IMAGES];  // n <= WIN_MAX_IMAGES (run_image_stage cuts larger chunks into groups)
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const uint32_t pend = (uint32_t)min(hdr[(size_t)j * HD_STRIDE + HD_PEND] + hdr[(size_t)j * HD_STRIDE + HD_XTRA], (1 << 20) - 1);
        s_key[j] = (pend << 12) | (uint32_t)min(hdr[(size_t)j * HD_STRIDE + HD_NQ] >> 6, 4095);
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t key = s_key[i];
    int rank = 0;
    for (int j = 0; j < n; j++) {
        const uint32_t kj = s_key[j];
        rank += (kj > key || (kj == key && j < i)) ? 1 : 0;
    }
    order[rank] = i;
}

// ---- stage 4: shade the window pass's results, cooperative pass for what it handed on, masked-out sites, counters ---------
template <bool SG>
__global__ void __launch_bounds__(FINISH_NT, IMAGE_FINISH_CTAS) finish_stage_kernel(ImageArgs A) {
    const int img = A.order ? A.order[blockIdx.x] : blockIdx.x;
    const int h = A.G.grid_h, w = A.G.grid_w, wpr = A.G.wpr;
    const int tid = threadIdx.x, lane = tid & 31;
    constexpr int NW = FINISH_NT / 32;
    const int nwords = h * wpr;
    const unsigned FULL = 0xffffffffu;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    ImageShared S;
    {
        unsigned char* p = smem_raw;
        S.occ = (uint32_t*)p; p += (size_t)nwords * 4;
#if IMAGE_FINISH_DEFER_GLOBAL
        S.tmp = A.defer + (size_t)img * A.plane_stride;
#else
        S.tmp = (uint32_t*)p; p += (size_t)nwords * 4;
#endif
        const size_t rows = image_row_bytes(h);
        int16_t** arr[9] = {&S.cnt, &S.first, &S.last, &S.up, &S.dn, &S.hl0, &S.hl1, &S.hr0, &S.hr1};
        for (int i = 0; i < 9; i++) { *arr[i] = (int16_t*)p; p += rows; }
        const size_t frows = ((size_t)h * 4 + 15) & ~(size_t)15;
        S.hlf = (float*)p; p += frows; S.hrf = (float*)p; p += frows;
        S.hlo = nullptr; S.hhi = nullptr; S.stk_l = nullptr; S.stk_r = nullptr;
    }
    const uint32_t* planes = A.planes + (size_t)img * 3 * A.plane_stride;
    S.keep = const_cast<uint32_t*>(planes) + 2 * A.plane_stride;  // global (L2): only the masking of sites reads it
    __shared__ int s_nitems, s_next, s_filled, s_flips, s_maxflips;

    int dst;
    uint8_t* out = image_out(A, img, dst);
    int32_t* counts = A.counts + (size_t)img * 8;
    int32_t* counts_final = dst >= 0 ? (A.counts_out ? A.counts_out + (size_t)dst * 8 : nullptr) : A.cache_counts + (size_t)(-1 - dst) * 8;
    int32_t* status_final = dst >= 0 ? (A.status ? A.status + dst : nullptr) : A.cache_status + (-1 - dst);
    const bool raw = A.raw_mode != 0;
    const int32_t* hd = A.hdr + (size_t)img * HD_STRIDE;
    const int status = hd[HD_STATUS], masked = hd[HD_MASKED];
    const int n_win = hd[HD_NQ] + hd[HD_XTRA];  // window-pass entries + residual ties of the edge rule (shade stage)
    const bool coop = status == 0 && hd[HD_PEND] + hd[HD_XTRA] > 0;  // something was handed on: the bit planes and row arrays are needed

    if (tid == 0) { s_nitems = 0; s_next = 0; s_filled = 0; s_flips = 0; s_maxflips = 0; }
    uint32_t* defer = S.tmp;  // bit plane: queries handed to the cooperative pass
    if (coop) {
        for (int i = tid; i < nwords; i += FINISH_NT) { S.occ[i] = __ldg(planes + i); defer[i] = 0u; }
        int16_t* d16[9] = {S.cnt, S.first, S.last, S.up, S.dn, S.hl0, S.hl1, S.hr0, S.hr1};
        const int ks[9] = {RA_CNT, RA_FIRST, RA_LAST, RA_UP, RA_DN, RA_HL0, RA_HL1, RA_HR0, RA_HR1};
#pragma unroll
        for (int k = 0; k < 9; k++) {
            const int16_t* g = row_arr(A, img, ks[k]);
            for (int r = tid; r < h; r += FINISH_NT) d16[k][r] = g[r];
        }
        const float *g_hlf = row_arr_f(A, img, RA_HLF), *g_hrf = row_arr_f(A, img, RA_HRF);
        for (int r = tid; r < h; r += FINISH_NT) { S.hlf[r] = g_hlf[r]; S.hrf[r] = g_hrf[r]; }
    }
    __syncthreads();

    int my_filled = 0, my_flips = 0, my_maxflips = 0;
    int32_t* qtri = A.qtri ? A.qtri + (size_t)img * A.qtri_stride : nullptr;
    const uint32_t* qlist = A.qlist + (size_t)img * A.qlist_stride;
    unsigned long long* qres = A.qres + (size_t)img * A.qlist_stride;
    uint32_t* clist = A.clist + (size_t)img * A.qlist_stride;

    auto site_rgb = [&](int sx, int sy) { return site_rgb_at(out, raw, h, w, sx, sy); };
    auto write_px = [&](const Tri2& t, uint32_t ca, uint32_t cb, uint32_t cc, int x, int r) { write_px_at(out, qtri, raw, h, w, t, ca, cb, cc, x, r); };
    auto flip_to = [&](Tri2& t, int v, int x, int r) {
        const int dx = v & 0xFFFF, dy = v >> 16;
        // flip inside {a,b,c,d}: keep the new triangle that contains q.  In coordinates relative to q the candidates (d,b,c),
        // (a,d,c), (a,b,d) share six cross products (orient(p,q,s) = p x q + q x s + s x p); |products| <= 2 * 2047^2 < 2^31.
        const int ax = t.ax - x, ay = t.ay - r, bx = t.bx - x, by = t.by - r, cx = t.cx - x, cy = t.cy - r, ex = dx - x, ey = dy - r;
        const int Xab = ax * by - ay * bx, Xbc = bx * cy - by * cx, Xca = cx * ay - cy * ax;
        const int Xad = ax * ey - ay * ex, Xbd = bx * ey - by * ex, Xcd = cx * ey - cy * ex;
        const bool fa = (Xbd <= 0) & (Xbc >= 0) & (Xcd >= 0) & ((long long)Xbc + Xcd - Xbd > 0);
        const bool fb = (Xad >= 0) & (Xcd <= 0) & (Xca >= 0) & ((long long)Xad - Xcd + Xca > 0);
        const bool fc = (Xab >= 0) & (Xbd >= 0) & (Xad <= 0) & ((long long)Xab + Xbd - Xad > 0);
        if (fa) { t.ax = dx; t.ay = dy; return true; }
        if (fb) { t.bx = dx; t.by = dy; return true; }
        if (fc) { t.cx = dx; t.cy = dy; return true; }
        return false;  // cannot happen (d is inside the triangle or across exactly one edge)
    };

    // ---- list what the window pass handed on (list order is kept within 32 entries) and mark those pixels in the deferred plane
    if (coop) {
        const int n_pad = (n_win + 31) & ~31;
        for (int i = tid; i < n_pad; i += FINISH_NT) {
            const bool pend = i < n_win && !(qres[i] & QRES_DONE);
            if (pend) {
                const uint32_t code = qlist[i];
                const int x = (int)(code & COL_MASK), r = (int)(code >> COL_BITS);
                atomicOr(&defer[r * wpr + (x >> 5)], 1u << (x & 31));
            }
            const uint32_t m = __ballot_sync(FULL, pend);
            if (!m) continue;
            int base = 0;
            if (lane == 0) base = atomicAdd(&s_nitems, __popc(m));
            base = __shfl_sync(FULL, base, 0);
            if (pend) clist[base + __popc(m & ((1u << lane) - 1u))] = (uint32_t)i;
        }
    }
    __syncthreads();

    // ---- cooperative pass: what the window pass handed on (hull pockets, wide gaps, the hole under the camera), one warp
    // per query ----------------------------------------------------------------------------------------------------------
    // The final triangle of a descent is rasterised over ALL deferred pixels it contains (they share it), which are then
    // taken off the list: a big triangle across a hole is found about once instead of once per pixel.  Warps take
    // row-major bands of the list, so that the pixels of one triangle mostly meet the same warp.
    if (coop) {
        const int n = s_nitems;
        // guided self-scheduling: bands shrink with what is left (long row-major bands first, so that the pixels of one
        // triangle mostly meet the same warp; short ones at the end, so that no warp is left alone with a long band)
        Tri2 tp = {0, 0, 0, 0, 0, 0};  // final triangle of this warp's previous descent
        bool have_prev = false;
        while (true) {
            int i0 = 0, band = 0;
            if (lane == 0) {
                const int seen = *(volatile int*)&s_next;
                band = min(max(IMAGE_COOP_MIN_BAND, (n - seen) / (IMAGE_COOP_BAND_DIV * NW)), max(n - seen, 1));
                i0 = atomicAdd(&s_next, band);
            }
            i0 = __shfl_sync(FULL, i0, 0); band = __shfl_sync(FULL, band, 0);
            if (i0 >= n) break;
            const int i_end = min(n, i0 + band);
          // 32 entries of the band at a time: every lane fetches one entry and watches its deferred bit; the warp takes the
          // first entry that is still pending (most are filled by the triangle of an earlier descent: skipping them costs one
          // memory round trip per descent instead of three per entry)
          for (int g0 = i0; g0 < i_end; g0 += 32) {
           bool mine = g0 + lane < i_end;
           uint32_t my_idx = 0u, my_code = 0u;
           if (mine) { my_idx = clist[g0 + lane]; my_code = qlist[my_idx]; }
           const int my_x = (int)(my_code & COL_MASK), my_r = (int)(my_code >> COL_BITS);
           while (true) {
            const bool still = mine && ((DEFER_LD(&defer[my_r * wpr + (my_x >> 5)]) >> (my_x & 31)) & 1u);
            const uint32_t pm = __ballot_sync(FULL, still);
            if (!pm) break;
            const int src = __ffs(pm) - 1;
            if (lane == src) mine = false;  // taken once, whatever becomes of it
            const uint32_t idx = __shfl_sync(FULL, my_idx, src);
            const uint32_t code = __shfl_sync(FULL, my_code, src);
            const int x = (int)(code & COL_MASK), r = (int)(code >> COL_BITS);
            Tri2 t;
            const unsigned long long part = qres[idx];
            if (part != 0ull) {  // continue the window pass's descent (the entry is not QRES_DONE: it is on this list)
                const uint32_t a = (uint32_t)part & M21, b = (uint32_t)(part >> 21) & M21, c = (uint32_t)(part >> 42) & M21;
                t = {vcol(a), vrow(a), vcol(b), vrow(b), vcol(c), vrow(c)};
            } else {
                const bool in_row = S.cnt[r] > 1 && x > S.first[r] && x < S.last[r];
                if (!(in_row ? init_tri_row(S, wpr, w, x, r, t) : init_tri_hull(S, x, r, t))) continue;
            }
#if IMAGE_COOP_CHAIN
            if (have_prev) {
                // The previous query of this warp is usually a neighbour, and its final triangle tp a neighbour of the triangle
                // wanted now.  If q lies beyond exactly one edge (u, v) of tp -- a Delaunay edge -- start from (v, u, s) with s a
                // vertex of the triangle found above: two of the three vertices are then final more often than not.
                const int o0 = orient_i(tp.ax, tp.ay, tp.bx, tp.by, x, r), o1 = orient_i(tp.bx, tp.by, tp.cx, tp.cy, x, r),
                          o2 = orient_i(tp.cx, tp.cy, tp.ax, tp.ay, x, r);
                if ((o0 < 0) + (o1 < 0) + (o2 < 0) == 1) {
                    int ux, uy, vx, vy;
                    if (o0 < 0) { ux = tp.ax; uy = tp.ay; vx = tp.bx; vy = tp.by; }
                    else if (o1 < 0) { ux = tp.bx; uy = tp.by; vx = tp.cx; vy = tp.cy; }
                    else { ux = tp.cx; uy = tp.cy; vx = tp.ax; vy = tp.ay; }
                    const int sx[3] = {t.ax, t.bx, t.cx}, sy[3] = {t.ay, t.by, t.cy};
#pragma unroll
                    for (int k = 2; k >= 0; k--)
                        if (ccw_contains(vx, vy, ux, uy, sx[k], sy[k], x, r)) t = {vx, vy, ux, uy, sx[k], sy[k]};
                }
            }
#endif
            int flips = 0, waves = 0;
#ifdef IMAGE_FINISH_DBG
            const int dbg_before = my_filled;
#endif
            // Each lane keeps the violator its row produced in the last scan.  After a flip these candidates are tested (exactly)
            // against the new circle before any row is scanned again: consecutive circles overlap, so about half of the flips
            // are found this way at a tenth of the cost of a scan.  The descent still ends with a full scan that finds nothing.
            unsigned long long cache = ~0ull;
            while (flips < IMAGE_MAX_FLIPS) {
                int v = -1;
#if IMAGE_COOP_CACHE
                if (__any_sync(FULL, cache != ~0ull)) {
                    bool viol = false;
                    if (cache != ~0ull) {
                        const uint32_t va = vlabel(t.ay, t.ax), vb = vlabel(t.by, t.bx), vc = vlabel(t.cy, t.cx);
                        const uint32_t vd = vlabel((int)((uint32_t)cache >> 16), (int)((uint32_t)cache & 0xFFFFu));
                        if (vd != va && vd != vb && vd != vc) {
                            const long long inc = incircle_v(va, vb, vc, vd);
                            viol = inc > 0 || (inc == 0 && incircle_pert(va, vb, vc, vd, w) > 0);
                        }
                        if (!viol) cache = ~0ull;
                    }
                    const uint32_t d2 = viol ? (uint32_t)(cache >> 32) : 0xFFFFFFFFu;
                    const uint32_t dmin = __reduce_min_sync(FULL, d2);
                    if (dmin != 0xFFFFFFFFu) {
                        const int src = __ffs(__ballot_sync(FULL, d2 == dmin)) - 1;
                        v = __shfl_sync(FULL, (int)(uint32_t)cache, src);
                        if (lane == src) cache = ~0ull;
                    }
                }
#endif
                if (v < 0) v = coop_find_violator<SG>(S.occ, S.hlf, S.hrf, wpr, w, h, t, x, r, w, lane, waves, cache);
                if (v < 0 || !flip_to(t, v, x, r)) break;
                flips++;
            }
#ifdef IMAGE_FINISH_DBG
            if (lane == 0) { my_flips += 1 | (waves << 12); }  // debug build: descents and waves instead of flips
#else
            if (lane == 0) { my_flips += flips; my_maxflips = max(my_maxflips, flips); }
#endif
            tp = t; have_prev = true;
            // rasterise t over the deferred pixels it contains: one lane per row of its bounding box
            const uint32_t ca = site_rgb(t.ax, t.ay), cb = site_rgb(t.bx, t.by), cc = site_rgb(t.cx, t.cy);
            const int y0 = min(t.ay, min(t.by, t.cy)), y1 = max(t.ay, max(t.by, t.cy));
            const int bx0 = min(t.ax, min(t.bx, t.cx)), bx1 = max(t.ax, max(t.bx, t.cx));
            for (int y = y0 + lane; y <= y1; y += 32) {
                for (int wi = bx0 >> 5; wi <= (bx1 >> 5); wi++) {
                    uint32_t m = DEFER_LD(&defer[y * wpr + wi]) & range_mask(wi, bx0, bx1);
                    uint32_t done = 0;
                    while (m) {
                        const int b = __ffs(m) - 1; m &= m - 1;
                        const int px = wi * 32 + b;
                        if (tri_contains(t, px, y)) { write_px(t, ca, cb, cc, px, y); done |= 1u << b; }
                    }
                    // count a pixel once even when two concurrent descents reach the same triangle (same value either way)
                    if (done) my_filled += __popc(atomicAnd(&defer[y * wpr + wi], ~done) & done);
                }
            }
#ifdef IMAGE_FINISH_DBG
            { const int got = __reduce_add_sync(FULL, my_filled - dbg_before); if (lane == 0 && got <= 2) my_maxflips += 1 | (waves << 10) | (flips << 20); }
#endif
            __syncwarp();
           }
          }
        }
    }
    __syncthreads();

    // ---- masked-out sites, degenerate images, counters ------------------------------------------------------------------
    const uint32_t* g_occ = planes;
    if (status == 1 || status == 2) {
        // the reference returns None (empty) or an all-zero interpolation (degenerate): clear the site colours
        for (int item = tid; item < nwords; item += FINISH_NT) {
            uint32_t m = __ldg(g_occ + item);
            const int r = item / wpr, wi = item - r * wpr;
            while (m) {
                const int b = __ffs(m) - 1; m &= m - 1;
                uint8_t* o = out + ((size_t)(raw ? r : h - 1 - r) * w + wi * 32 + b) * 3;
                o[0] = 0; o[1] = 0; o[2] = 0;
            }
        }
    } else if (!raw && (masked || status != 0)) {
        for (int item = tid; item < nwords; item += FINISH_NT) {
            uint32_t m = __ldg(g_occ + item) & ~__ldg(S.keep + item);  // sites the hallucination mask removes
            const int r = item / wpr, wi = item - r * wpr;
            while (m) {
                const int b = __ffs(m) - 1; m &= m - 1;
                uint8_t* o = out + ((size_t)(h - 1 - r) * w + wi * 32 + b) * 3;
                o[0] = 0; o[1] = 0; o[2] = 0;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        my_filled += __shfl_xor_sync(FULL, my_filled, o);
        my_flips += __shfl_xor_sync(FULL, my_flips, o);
#ifdef IMAGE_FINISH_DBG
        my_maxflips += __shfl_xor_sync(FULL, my_maxflips, o);
#else
        my_maxflips = max(my_maxflips, __shfl_xor_sync(FULL, my_maxflips, o));
#endif
    }
    #ifdef IMAGE_FINISH_DBG
    if (lane == 0) { atomicAdd(&s_filled, my_filled); atomicAdd(&s_flips, my_flips); atomicAdd(&s_maxflips, my_maxflips); }
#else
    if (lane == 0) { atomicAdd(&s_filled, my_filled); atomicAdd(&s_flips, my_flips); atomicMax(&s_maxflips, my_maxflips); }
#endif
    __syncthreads();
    if (tid == 0) {
        // filled = edge-rule pixels + window-pass pixels (what was handed on, residual ties included, is counted by the cooperative pass)
        counts[5] = status == 0 ? hd[HD_EDGE] + hd[HD_LOCAL] + hd[HD_NQ] - s_nitems + s_filled : 0; counts[6] = max(counts[6], s_maxflips); counts[7] += s_flips;
#ifdef IMAGE_FINISH_DBG
        counts[6] = s_maxflips; counts[7] = s_flips;  // debug build: entries handed on | small descents << 16, descents | waves << 12
#endif
        if (status_final) *status_final = status;
        if (counts_final) {
#pragma unroll
            for (int k = 0; k < 8; k++) counts_final[k] = counts[k];
        }
    }
}

// colour word tap: r | g<<8 | b<<16 | 0xFF<<24 at sites, 0 elsewhere
__global__ void tap_color_kernel(const uint32_t* __restrict__ keygrid, const uint8_t* __restrict__ csrc, int g, int pano_w, uint32_t* __restrict__ outc) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= g) return;
    const uint32_t key = keygrid[p];
    uint32_t cw = 0;
    if (key) {
        cw = gather_rgb(csrc, (key - 1u) & KEY_IDX_MASK, pano_w) | 0xFF000000u;
    }
    outc[p] = cw;
}

}  // namespace bev
