// k_sites.cuh -- one CTA per BEV image: winners -> site bit rows, colours, emptiness / keep masks,
// base output image, and the initial ("zipper") triangulation that the flip kernel legalises.
//
// Reference stages covered:
//   sparse_bev_img[y, x] = rgb of the z-order winner            bev_rendering_utils.py:298-308
//   degenerate-input guards (<4 points, one row, one column)     interpolation_utils.py:37-42, 57-71
//   nonempty = uint8(r*g*b) > 0  (wraps mod 256)                 interpolation_utils.py:95-98
//   keep = KxK zero-padded box count > 0                         interpolation_utils.py:101-115
//   np.flipud of the result                                      bev_rendering_utils.py:319
//
// Initial triangulation.  Sites live on integer pixels, so every non-empty image row is a sorted
// list of collinear points.  Between consecutive non-empty rows A (below) and B (above) the strip
// conv(A u B) is triangulated by a merge ("zipper") ordered by column; the ring of boundary edges
// gets ghost triangles to the vertex at infinity.  All indices -- triangle ids, apexes, neighbours --
// follow from popcounts over the bit rows, so the whole mesh (2S-2 triangles, closed sphere
// topology) is written in parallel with no sort and no hash.
#pragma once
#include "bev_common.cuh"

namespace bev {

constexpr int SITES_NT = 512;
constexpr int MAX_GRID_H = 1023;

struct SitesArgs {
    GridParams G;
    // per-image scratch: base pointers + per-image strides (elements)
    uint32_t* keygrid; size_t keygrid_stride;
    uint32_t* color; size_t color_stride;
    uint32_t* occ; uint32_t* nonempty; uint32_t* keep; uint32_t* tmpbits; size_t bits_stride;
    uint16_t* wprefix;
    Tri* tris; size_t tris_stride;
    ImgHeader* headers;
    int32_t* counts;                 // [n_img][8]
    int32_t* status;                 // [n_img] or null
    const uint8_t* const* color_src; // per image: u8 rgb triples indexed by the key's source index (tagged, see gather_rgb)
    int32_t pano_w;
    uint8_t* out; size_t out_stride; // final images (bytes per image)
    int32_t raw_mode;                // 1: no keep mask, no flip (interp_dense_grid_from_sparse semantics)
    int32_t skip_empty_check;        // 1: generic interp path (no EMPTY status)
};

__device__ __forceinline__ int bits_rank_lt(const uint32_t* bits, const uint16_t* pre, int wpr, int x) {
    if (x <= 0) return 0;
    int wi = x >> 5;
    if (wi >= wpr) return pre[wpr - 1] + __popc(bits[wpr - 1]);
    return pre[wi] + __popc(bits[wi] & ((1u << (x & 31)) - 1u));
}
// highest set bit with column < x, or the row's first set bit when there is none
__device__ __forceinline__ int bits_pred_or_first(const uint32_t* bits, int wpr, int x) {
    int wi = x >> 5;
    uint32_t m = 0;
    if (wi >= wpr) wi = wpr; else m = bits[wi] & ((1u << (x & 31)) - 1u);
    while (m == 0 && wi > 0) { wi--; m = bits[wi]; }
    if (m) return wi * 32 + 31 - __clz(m);
    for (wi = 0; wi < wpr; wi++) { m = bits[wi]; if (m) return wi * 32 + __ffs(m) - 1; }
    return -1;
}
__device__ __forceinline__ int bits_next_after(const uint32_t* bits, int wpr, int col) {
    int wi = col >> 5;
    uint32_t m = bits[wi] & ~((2u << (col & 31)) - 1u);
    while (m == 0) { wi++; if (wi >= wpr) return -1; m = bits[wi]; }
    return wi * 32 + __ffs(m) - 1;
}

// block-wide exclusive scan of one int per thread (blockDim = SITES_NT)
__device__ __forceinline__ int block_excl_scan(int v, int* s_warp, int& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    __syncthreads();
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = (lane < SITES_NT / 32) ? s_warp[lane] : 0, wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
        if (lane < SITES_NT / 32) s_warp[lane] = wi - w;
        if (lane == 31) s_warp[32] = wi;
    }
    __syncthreads();
    total = s_warp[32];
    return s_warp[warp] + incl - v;
}

__global__ void __launch_bounds__(SITES_NT) sites_kernel(SitesArgs A) {
    const int img = blockIdx.x;
    const int h = A.G.grid_h, w = A.G.grid_w, wpr = A.G.wpr;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = SITES_NT / 32;

    uint32_t* keygrid = A.keygrid + (size_t)img * A.keygrid_stride;
    uint32_t* color = A.color + (size_t)img * A.color_stride;
    uint32_t* occ = A.occ + (size_t)img * A.bits_stride;
    uint32_t* nonempty = A.nonempty + (size_t)img * A.bits_stride;
    uint32_t* keep = A.keep + (size_t)img * A.bits_stride;
    uint32_t* tmpbits = A.tmpbits + (size_t)img * A.bits_stride;
    uint16_t* wprefix = A.wprefix + (size_t)img * A.bits_stride;
    Tri* tris = A.tris + (size_t)img * A.tris_stride;
    const uint8_t* csrc = A.color_src[img];
    uint8_t* out = A.out ? A.out + (size_t)img * A.out_stride : nullptr;
    int32_t* counts = A.counts + img * 8;

    __shared__ int s_cnt[MAX_GRID_H + 1], s_rank[MAX_GRID_H + 1], s_rows[MAX_GRID_H + 1], s_base[MAX_GRID_H + 1];
    __shared__ int s_first[MAX_GRID_H + 1], s_last[MAX_GRID_H + 1];
    __shared__ int s_warp[33];
    __shared__ int s_mincol, s_maxcol, s_nonempty_cnt, s_keep_cnt;
    if (tid == 0) { s_mincol = 1 << 30; s_maxcol = -1; s_nonempty_cnt = 0; s_keep_cnt = 0; }
    __syncthreads();

    // ---- A. winners -> colour grid, occupancy / non-empty bit rows, per-row counts ------------
    for (int r = warp; r < h; r += NW) {
        int running = 0, first = -1, last = -1, ne_cnt = 0;
        for (int wi = 0; wi < wpr; wi++) {
            const int c = wi * 32 + lane;
            uint32_t key = (c < w) ? keygrid[r * w + c] : 0u;
            uint32_t cw = 0; bool ne = false;
            if (key) {
                const uint32_t c3 = gather_rgb(csrc, (key - 1u) & KEY_IDX_MASK, A.pano_w);
                const uint32_t cr = c3 & 0xFF, cg = (c3 >> 8) & 0xFF, cb = c3 >> 16;
                cw = cr | (cg << 8) | (cb << 16) | 0xFF000000u;
                ne = ((cr * cg * cb) & 0xFFu) != 0u;  // uint8 product wraps (interpolation_utils.py:95)
            }
            if (c < w) color[r * w + c] = cw;
            const uint32_t ob = __ballot_sync(0xffffffffu, key != 0u);
            const uint32_t nb = __ballot_sync(0xffffffffu, ne);
            if (lane == 0) {
                occ[r * wpr + wi] = ob; nonempty[r * wpr + wi] = nb; wprefix[r * wpr + wi] = (uint16_t)running;
            }
            if (ob) { if (first < 0) first = wi * 32 + __ffs(ob) - 1; last = wi * 32 + 31 - __clz(ob); }
            running += __popc(ob); ne_cnt += __popc(nb);
        }
        if (lane == 0) {
            s_cnt[r] = running; s_first[r] = first; s_last[r] = last;
            if (running) { atomicMin(&s_mincol, first); atomicMax(&s_maxcol, last); }
            if (ne_cnt) atomicAdd(&s_nonempty_cnt, ne_cnt);
        }
    }
    __syncthreads();

    // ---- B. row bookkeeping: site total, non-empty row list, strip bases -----------------------
    int S = 0, M = 0;
    {
        int carry_s = 0, carry_m = 0;
        for (int base = 0; base < h; base += SITES_NT) {
            const int r = base + tid;
            const int c = (r < h) ? s_cnt[r] : 0;
            int tot_s, tot_m;
            (void)block_excl_scan(c, s_warp, tot_s);
            const int rk = block_excl_scan(c > 0 ? 1 : 0, s_warp, tot_m);
            if (r < h) { s_rank[r] = carry_m + rk; if (c > 0) s_rows[carry_m + rk] = r; }
            carry_s += tot_s; carry_m += tot_m;
        }
        S = carry_s; M = carry_m;
    }
    __syncthreads();
    int NT0 = 0;
    {
        int carry = 0;
        for (int base = 0; base < M - 1; base += SITES_NT) {
            const int k = base + tid;
            const int t = (k < M - 1) ? (s_cnt[s_rows[k]] + s_cnt[s_rows[k + 1]] - 2) : 0;
            int tot;
            const int ex = block_excl_scan(t, s_warp, tot);
            if (k < M - 1) s_base[k] = carry + ex;
            carry += tot;
        }
        NT0 = carry;
    }
    __syncthreads();

    int status = 0;  // SALVE_BEV_IMG_OK
    if (!A.skip_empty_check && counts[1] == 0) status = 1;                       // EMPTY -> None
    else if (S < 4 || M < 2 || s_mincol == s_maxcol) status = 2;                 // DEGENERATE -> zeros
    const bool triangulate = (status == 0);
    if (tid == 0) {
        ImgHeader hd; hd.n_sites = S; hd.n_tris = triangulate ? 2 * S - 2 : 0; hd.status = status; hd.pad = 0;
        A.headers[img] = hd;
        counts[2] = S; counts[3] = s_nonempty_cnt;
        if (A.status) A.status[img] = status;
    }

    // ---- C. zipper triangulation -------------------------------------------------------------
    if (triangulate) {
        const int p0 = s_cnt[s_rows[0]], pl = s_cnt[s_rows[M - 1]];
        const int Gn = (p0 - 1) + (M - 1) + (pl - 1) + (M - 1);
        const int g_bot = NT0, g_right = g_bot + (p0 - 1), g_top = g_right + (M - 1), g_left = g_top + (pl - 1);
        auto prev_g = [&](int g) { return NT0 + ((g - NT0 + Gn - 1) % Gn); };
        auto next_g = [&](int g) { return NT0 + ((g - NT0 + 1) % Gn); };
        // segments: one work item per (row, word); loop over the word's set bits
        for (int item = tid; item < h * wpr; item += SITES_NT) {
            const int r = item / wpr, wi = item - r * wpr;
            uint32_t word = occ[r * wpr + wi];
            if (!word) continue;
            const int k = s_rank[r];
            const uint32_t* rowbits = occ + r * wpr;
            const int p = s_cnt[r];
            int j = wprefix[r * wpr + wi];
            // neighbours rows
            const int rn = (k + 1 < M) ? s_rows[k + 1] : -1;
            const int rp = (k > 0) ? s_rows[k - 1] : -1;
            const int rnn = (k + 2 < M) ? s_rows[k + 2] : -1;
            const int rpp = (k > 1) ? s_rows[k - 2] : -1;
            while (word) {
                const int b = __ffs(word) - 1; word &= word - 1;
                const int c0 = wi * 32 + b;
                if (j + 1 < p) {
                    const int c1 = word ? (wi * 32 + __ffs(word) - 1) : bits_next_after(rowbits, wpr, c0);
                    const uint32_t va = vlabel(r, c0), vb = vlabel(r, c1);
                    int up_id = -1, down_id = -1;
                    // up-triangle of strip k (this row is the lower row)
                    if (rn >= 0) {
                        const uint32_t* B = occ + rn * wpr; const uint16_t* Bp = wprefix + rn * wpr;
                        const int q = s_cnt[rn];
                        const int cb = max(bits_rank_lt(B, Bp, wpr, c1) - 1, 0);
                        const int pos = j + cb, T = p + q - 2;
                        up_id = s_base[k] + pos;
                        const int lg = g_left + (M - 2 - k), rg = g_right + k;
                        int n2;
                        if (k == 0) n2 = g_bot + j;
                        else {
                            const uint32_t* A2 = occ + rp * wpr; const uint16_t* A2p = wprefix + rp * wpr;
                            n2 = s_base[k - 1] + j + max(bits_rank_lt(A2, A2p, wpr, c1 + 1) - 1, 0);
                        }
                        const uint32_t apex = vlabel(rn, bits_pred_or_first(B, wpr, c1));
                        st_tri(tris + up_id, make_tri(va, vb, apex, (pos + 1 < T) ? up_id + 1 : rg, (pos > 0) ? up_id - 1 : lg, n2));
                    }
                    // down-triangle of strip k-1 (this row is the upper row)
                    if (rp >= 0) {
                        const uint32_t* A2 = occ + rp * wpr; const uint16_t* A2p = wprefix + rp * wpr;
                        const int pp = s_cnt[rp];
                        const int ca = max(bits_rank_lt(A2, A2p, wpr, c1 + 1) - 1, 0);
                        const int pos = j + ca, T = pp + p - 2;
                        down_id = s_base[k - 1] + pos;
                        const int lg = g_left + (M - 2 - (k - 1)), rg = g_right + (k - 1);
                        int n2;
                        if (k == M - 1) n2 = g_top + (p - 2 - j);
                        else {
                            const uint32_t* B = occ + rn * wpr; const uint16_t* Bp = wprefix + rn * wpr;
                            n2 = s_base[k] + j + max(bits_rank_lt(B, Bp, wpr, c1) - 1, 0);
                        }
                        const uint32_t apex = vlabel(rp, bits_pred_or_first(A2, wpr, c1 + 1));
                        st_tri(tris + down_id, make_tri(vb, va, apex, (pos > 0) ? down_id - 1 : lg, (pos + 1 < T) ? down_id + 1 : rg, n2));
                    }
                    if (k == 0) {  // bottom ghost: boundary edge a_j -> a_{j+1}
                        const int g = g_bot + j;
                        st_tri(tris + g, make_tri(vb, va, GHOST, prev_g(g), next_g(g), up_id));
                    }
                    if (k == M - 1) {  // top ghost: boundary edge b_{i+1} -> b_i
                        const int g = g_top + (p - 2 - j);
                        st_tri(tris + g, make_tri(va, vb, GHOST, prev_g(g), next_g(g), down_id));
                    }
                    (void)rnn; (void)rpp;
                }
                j++;
            }
        }
        // side ghosts: one per strip
        for (int k = tid; k < M - 1; k += SITES_NT) {
            const int ra = s_rows[k], rb = s_rows[k + 1];
            const int T = s_cnt[ra] + s_cnt[rb] - 2;
            const int rg = g_right + k, lg = g_left + (M - 2 - k);
            st_tri(tris + rg, make_tri(vlabel(rb, s_last[rb]), vlabel(ra, s_last[ra]), GHOST, prev_g(rg), next_g(rg),
                                       (T > 0) ? s_base[k] + T - 1 : lg));
            st_tri(tris + lg, make_tri(vlabel(ra, s_first[ra]), vlabel(rb, s_first[rb]), GHOST, prev_g(lg), next_g(lg),
                                       (T > 0) ? s_base[k] : rg));
        }
    }

    // ---- D. keep mask = Chebyshev dilation of `nonempty` by K/2, zero padded -------------------
    const int rad = A.G.K / 2;
    for (int item = tid; item < h * wpr; item += SITES_NT) {
        const int r = item / wpr, wi = item - r * wpr;
        const uint32_t cur = nonempty[item];
        const uint32_t prv = (wi > 0) ? nonempty[item - 1] : 0u;
        const uint32_t nxt = (wi + 1 < wpr) ? nonempty[item + 1] : 0u;
        const unsigned long long L = ((unsigned long long)cur << 32) | prv, R = ((unsigned long long)nxt << 32) | cur;
        uint32_t o = cur;
        for (int d = 1; d <= rad; d++) o |= (uint32_t)(L >> (32 - d)) | (uint32_t)(R >> d);
        const int valid = w - wi * 32;
        if (valid < 32) o &= (1u << valid) - 1u;
        tmpbits[item] = o;
    }
    __syncthreads();
    {
        int kc = 0;
        for (int item = tid; item < h * wpr; item += SITES_NT) {
            const int r = item / wpr, wi = item - r * wpr;
            uint32_t o = 0;
            const int r0 = max(r - rad, 0), r1 = min(r + rad, h - 1);
            for (int rr = r0; rr <= r1; rr++) o |= tmpbits[rr * wpr + wi];
            keep[item] = o;
            kc += __popc(o);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) kc += __shfl_xor_sync(0xffffffffu, kc, o);
        if (lane == 0 && kc) atomicAdd(&s_keep_cnt, kc);
    }
    __syncthreads();
    if (tid == 0) counts[4] = s_keep_cnt;

    // ---- E. base image: zeros + site colours (masked and flipped unless raw_mode) --------------
    if (out != nullptr) {
        const bool zero_all = (status != 0);
        for (int r = warp; r < h; r += NW) {
            uint8_t* orow = out + (size_t)(A.raw_mode ? r : (h - 1 - r)) * w * 3;
            for (int wi = 0; wi < wpr; wi++) {
                const int c = wi * 32 + lane;
                if (c >= w) continue;
                uint32_t cw = zero_all ? 0u : color[r * w + c];
                if (!A.raw_mode && !((keep[r * wpr + wi] >> lane) & 1u)) cw = 0u;
                orow[c * 3 + 0] = (uint8_t)(cw & 0xFF);
                orow[c * 3 + 1] = (uint8_t)((cw >> 8) & 0xFF);
                orow[c * 3 + 2] = (uint8_t)((cw >> 16) & 0xFF);
            }
        }
    }
}

}  // namespace bev
