// k_flip.cuh -- one CTA per BEV image: legalise the zipper mesh into the Delaunay triangulation
// by synchronous rounds of parallel edge flips (Lawson), replacing the Qhull call inside
// scipy.interpolate.griddata (salve/utils/interpolation_utils.py:46-48).
//
// Round structure (all inside one kernel; __syncthreads between phases):
//   1. every active triangle tests its edges with the exact int64 in-circle predicate (+ symbolic
//      perturbation for co-circular ties, + orientation tests on ghost edges = hull convexification);
//      an illegal edge becomes a candidate and bids for the 4 triangles its flip would write
//      (the pair and the two outer neighbours whose back-pointers change) with atomicMax of a
//      unique (round, hash, edge) key -- deterministic: max is order independent.
//   2. a candidate that holds all 4 bids flips; winners and losers re-queue their triangles.
// The mesh after each round is a function of the mesh before it, and the fixed point is the unique
// regular triangulation of the perturbed lift, so the result does not depend on thread timing.
#pragma once
#include <cooperative_groups.h>
#include "bev_common.cuh"

namespace bev {
namespace cg = cooperative_groups;

constexpr int FLIP_NT = 512;
constexpr int FLIP_MAX_ROUNDS = 60000;

struct FlipArgs {
    int32_t grid_w;
    Tri* tris; size_t tris_stride;
    unsigned long long* owner; size_t owner_stride;
    uint32_t* list0; uint32_t* list1; size_t list_stride;
    uint32_t* cand; size_t cand_stride;
    const ImgHeader* headers;
    int32_t* counts;
};

__device__ __forceinline__ unsigned long long flip_key(uint32_t round, uint32_t e) {
    return ((unsigned long long)round << 48) | ((unsigned long long)(hash32(e * 0x9E3779B1u + round) & 0xFFFFFFu) << 24) | e;
}

// locate in U the edge (c,b) shared with t; returns local index j or -1
__device__ __forceinline__ int find_shared(const Tri& U, uint32_t t, uint32_t b, uint32_t c) {
    const uint32_t u0 = tri_v(U, 0), u1 = tri_v(U, 1), u2 = tri_v(U, 2);
    if (u1 == c && u2 == b && tri_n(U, 0) == t) return 0;
    if (u2 == c && u0 == b && tri_n(U, 1) == t) return 1;
    if (u0 == c && u1 == b && tri_n(U, 2) == t) return 2;
    return -1;
}

__device__ __forceinline__ void relink(Tri* tris, uint32_t x, uint32_t e0, uint32_t e1, uint32_t to) {
    Tri X = ld_tri(tris + x);
    const uint32_t x0 = tri_v(X, 0), x1 = tri_v(X, 1), x2 = tri_v(X, 2);
    int k = -1;
    if (x1 == e0 && x2 == e1) k = 0;
    else if (x2 == e0 && x0 == e1) k = 1;
    else if (x0 == e0 && x1 == e1) k = 2;
    if (k < 0) return;
    X.hi = (X.hi & ~((unsigned long long)M21 << (21 * k))) | ((unsigned long long)to << (21 * k));
    tris[x].hi = X.hi;
}

__global__ void __launch_bounds__(FLIP_NT) flip_kernel(FlipArgs A) {
    const int img = blockIdx.x;
    const ImgHeader hd = A.headers[img];
    const int nt = hd.n_tris;
    if (nt <= 0) return;
    const int tid = threadIdx.x;
    Tri* tris = A.tris + (size_t)img * A.tris_stride;
    unsigned long long* owner = A.owner + (size_t)img * A.owner_stride;
    uint32_t* cur = A.list0 + (size_t)img * A.list_stride;
    uint32_t* nxt = A.list1 + (size_t)img * A.list_stride;
    uint32_t* cand = A.cand + (size_t)img * A.cand_stride;
    const int gw = A.grid_w;

    extern __shared__ uint32_t s_bits[];  // membership of the active list, nt bits
    __shared__ int s_ncand, s_nnext, s_flips;
    for (int i = tid; i < (nt + 31) / 32; i += FLIP_NT) s_bits[i] = 0u;
    for (int i = tid; i < nt; i += FLIP_NT) owner[i] = 0ull;
    if (tid == 0) { s_ncand = 0; s_nnext = 0; s_flips = 0; }
    __syncthreads();

    int n_cur = nt, my_flips = 0;
    uint32_t round = 0;
    while (true) {
        round++;
        // ---- phase 1: detect illegal edges, bid ------------------------------------------------
        for (int idx = tid; idx < n_cur; idx += FLIP_NT) {
            const uint32_t t = (round == 1) ? (uint32_t)idx : cur[idx];
            const Tri T = ld_tri(tris + t);
#pragma unroll
            for (int i = 0; i < 3; i++) {
                const uint32_t u = tri_n(T, i);
                if (u < t && (round == 1 || ((s_bits[u >> 5] >> (u & 31)) & 1u))) continue;  // u tests this edge itself
                const uint32_t a = tri_v(T, i), b = tri_v(T, (i + 1) % 3), c = tri_v(T, (i + 2) % 3);
                if (a == GHOST) continue;
                const Tri U = ld_tri(tris + u);
                const int j = find_shared(U, t, b, c);
                if (j < 0) continue;
                const uint32_t d = tri_v(U, j);
                if (!flip_rule(a, b, c, d, gw)) continue;
                const uint32_t e = t * 3u + (uint32_t)i;
                const unsigned long long key = flip_key(round, e);
                atomicMax(owner + t, key);
                atomicMax(owner + u, key);
                atomicMax(owner + tri_n(U, (j + 1) % 3), key);  // x_bd
                atomicMax(owner + tri_n(T, (i + 1) % 3), key);  // x_ca
                auto g = cg::coalesced_threads();
                int pos = 0;
                if (g.thread_rank() == 0) pos = atomicAdd(&s_ncand, (int)g.size());
                pos = g.shfl(pos, 0) + (int)g.thread_rank();
                cand[pos] = e;
            }
        }
        __syncthreads();
        const int ncand = s_ncand;
        if (ncand == 0 || round >= FLIP_MAX_ROUNDS) break;
        // ---- phase 1.5: retire the current active list ---------------------------------------------
        if (round > 1)
            for (int idx = tid; idx < n_cur; idx += FLIP_NT) { const uint32_t t = cur[idx]; atomicAnd(&s_bits[t >> 5], ~(1u << (t & 31))); }
        __syncthreads();
        // ---- phase 2: winners flip -------------------------------------------------------------------
        for (int idx = tid; idx < ncand; idx += FLIP_NT) {
            const uint32_t e = cand[idx];
            const uint32_t t = e / 3u; const int i = (int)(e - t * 3u);
            const unsigned long long key = flip_key(round, e);
            bool won = false; uint32_t u = 0;
            if (__ldcg(owner + t) == key) {
                const Tri T = ld_tri(tris + t);
                u = tri_n(T, i);
                if (__ldcg(owner + u) == key) {
                    const Tri U = ld_tri(tris + u);
                    const uint32_t a = tri_v(T, i), b = tri_v(T, (i + 1) % 3), c = tri_v(T, (i + 2) % 3);
                    const int j = find_shared(U, t, b, c);
                    const uint32_t d = tri_v(U, j);
                    const uint32_t x_ca = tri_n(T, (i + 1) % 3), x_ab = tri_n(T, (i + 2) % 3);
                    const uint32_t x_bd = tri_n(U, (j + 1) % 3), x_dc = tri_n(U, (j + 2) % 3);
                    if (__ldcg(owner + x_bd) == key && __ldcg(owner + x_ca) == key) {
                        st_tri(tris + t, make_tri(a, b, d, x_bd, u, x_ab));
                        st_tri(tris + u, make_tri(a, d, c, x_dc, x_ca, t));
                        relink(tris, x_bd, d, b, t);
                        relink(tris, x_ca, a, c, u);
                        won = true; my_flips++;
                    }
                }
            }
            {
                const uint32_t m = 1u << (t & 31);
                if (!(atomicOr(&s_bits[t >> 5], m) & m)) nxt[atomicAdd(&s_nnext, 1)] = t;
            }
            if (won) {
                const uint32_t m = 1u << (u & 31);
                if (!(atomicOr(&s_bits[u >> 5], m) & m)) nxt[atomicAdd(&s_nnext, 1)] = u;
            }
        }
        __syncthreads();
        n_cur = s_nnext;
        uint32_t* tmp = cur; cur = nxt; nxt = tmp;
        __syncthreads();
        if (tid == 0) { s_ncand = 0; s_nnext = 0; }
        __syncthreads();
    }
    if (my_flips) atomicAdd(&s_flips, my_flips);
    __syncthreads();
    if (tid == 0) {
        int32_t* counts = A.counts + img * 8;
        counts[6] = (int)round;
        counts[7] = s_flips;
    }
}

}  // namespace bev
