/*
 * oracle/canonical_dt.c -- CPU checker for the densification stage.  TEST INFRASTRUCTURE ONLY.
 *
 * The reference densifies with scipy.interpolate.griddata(method="linear")
 * (salve/utils/interpolation_utils.py:46-48): Qhull Delaunay + barycentric evaluation.  That
 * code is a third-party dependency, not under /root/reference, and on integer pixel sites its
 * triangulation is massively degenerate (co-circular cells), so *which* Delaunay triangulation
 * comes out depends on Qhull's insertion order.
 *
 * This file restates the published algorithm in a form whose result is a pure function of the
 * site SET:
 *   - Delaunay triangulation by Lawson edge flipping with an exact int64 in-circle predicate;
 *   - co-circular ties broken by an infinitesimal perturbation of the lifted height,
 *         h(p) = |p|^2 + eps * w(p),   w(p) = 20-bit hash of the pixel index,
 *     so the outcome is the (generically unique) regular triangulation of the perturbed lift
 *     -- independent of the starting triangulation and of the flip order;
 *   - the convex hull handled by one ghost vertex at infinity (ghost-ghost flips = Graham scan);
 *   - interpolation with exact integer barycentrics: value = floor(sum(w_i*c_i) / sum(w_i)),
 *     the truncation (interpolation_utils.py:53) of the exact value.
 *
 * It deliberately uses a different schedule from the CUDA path (sequential work stack here,
 * synchronous parallel rounds with atomic arbitration there); both must reach the same
 * triangulation.  tests/ additionally pin it against SciPy itself: validity (no site strictly
 * inside any circumcircle), identical hull mask, identical set of tie-independent triangles,
 * and +-1 on tie-independent pixels.
 *
 * Build: gcc -O2 -shared -fPIC -o oracle/libcanonical_dt.so oracle/canonical_dt.c
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define GHOST (-1)

typedef struct {
    int32_t v[3]; /* CCW (x = col, y = row); exactly one may be GHOST */
    int32_t n[3]; /* n[i] = triangle across the edge opposite v[i] */
} Tri;

static inline int64_t orient2d(int64_t ax, int64_t ay, int64_t bx, int64_t by, int64_t cx, int64_t cy) {
    return (bx - ax) * (cy - ay) - (by - ay) * (cx - ax);
}

/* 20-bit perturbation weight of a pixel (shared definition with the CUDA path). */
static inline int64_t pert_weight(int32_t row, int32_t col, int32_t img_w) {
    uint32_t h = (uint32_t)(row * img_w + col);
    h ^= h >> 16; h *= 0x7feb352dU; h ^= h >> 15; h *= 0x846ca68bU; h ^= h >> 16;
    return (int64_t)(h & 0xFFFFF);
}

typedef struct {
    int32_t S, img_w;
    const int32_t *row, *col;
    Tri *tri;
    int32_t ntri;
} Mesh;

/* >0: d strictly inside circumcircle of CCW (a,b,c);  0: co-circular. */
static int64_t incircle(const Mesh *m, int a, int b, int c, int d) {
    int64_t adx = m->col[a] - m->col[d], ady = m->row[a] - m->row[d];
    int64_t bdx = m->col[b] - m->col[d], bdy = m->row[b] - m->row[d];
    int64_t cdx = m->col[c] - m->col[d], cdy = m->row[c] - m->row[d];
    int64_t ad = adx * adx + ady * ady, bd = bdx * bdx + bdy * bdy, cd = cdx * cdx + cdy * cdy;
    return adx * (bdy * cd - bd * cdy) - ady * (bdx * cd - bd * cdx) + ad * (bdx * cdy - bdy * cdx);
}

static int64_t orient_s(const Mesh *m, int a, int b, int c) {
    return orient2d(m->col[a], m->row[a], m->col[b], m->row[b], m->col[c], m->row[c]);
}

/* first-order term of the perturbed in-circle determinant */
static int64_t incircle_pert(const Mesh *m, int a, int b, int c, int d) {
    int64_t wa = pert_weight(m->row[a], m->col[a], m->img_w), wb = pert_weight(m->row[b], m->col[b], m->img_w);
    int64_t wc = pert_weight(m->row[c], m->col[c], m->img_w), wd = pert_weight(m->row[d], m->col[d], m->img_w);
    return wa * orient_s(m, b, c, d) - wb * orient_s(m, a, c, d) + wc * orient_s(m, a, b, d) - wd * orient_s(m, a, b, c);
}

/* Should the edge opposite t.v[i] be flipped?  Finds (u, j).  Returns 1/0. */
static int needs_flip(const Mesh *m, int t, int i, int *pu, int *pj) {
    const Tri *T = &m->tri[t];
    int u = T->n[i];
    if (u < 0) return 0;
    int a = T->v[i], b = T->v[(i + 1) % 3], c = T->v[(i + 2) % 3];
    const Tri *U = &m->tri[u];
    int j = -1;
    for (int k = 0; k < 3; k++)
        if (U->v[(k + 1) % 3] == c && U->v[(k + 2) % 3] == b && U->n[k] == t) j = k;
    if (j < 0) return 0; /* cannot happen in a consistent mesh */
    int d = U->v[j];
    *pu = u; *pj = j;
    if (a == GHOST || d == GHOST) return 0;
    if (c == GHOST) return orient_s(m, a, b, d) > 0;
    if (b == GHOST) return orient_s(m, a, d, c) > 0;
    int64_t ic = incircle(m, a, b, c, d);
    if (ic != 0) return ic > 0;
    return incircle_pert(m, a, b, c, d) > 0;
}

static void relink(Mesh *m, int x, int from, int to, int e0, int e1) {
    /* in triangle x, the neighbour slot for edge (e0,e1) [as x sees it: e0->e1] that points to `from` -> `to` */
    if (x < 0) return;
    Tri *X = &m->tri[x];
    for (int k = 0; k < 3; k++)
        if (X->n[k] == from && X->v[(k + 1) % 3] == e0 && X->v[(k + 2) % 3] == e1) { X->n[k] = to; return; }
}

/* t=(a,b,c), u=(d,c,b) sharing (b,c)  ->  t=(a,b,d), u=(a,d,c) */
static void do_flip(Mesh *m, int t, int i, int u, int j) {
    Tri *T = &m->tri[t], *U = &m->tri[u];
    int a = T->v[i], b = T->v[(i + 1) % 3], c = T->v[(i + 2) % 3], d = U->v[j];
    int x_ca = T->n[(i + 1) % 3], x_ab = T->n[(i + 2) % 3];
    int x_bd = U->n[(j + 1) % 3], x_dc = U->n[(j + 2) % 3];
    T->v[0] = a; T->v[1] = b; T->v[2] = d;
    T->n[0] = x_bd; T->n[1] = u; T->n[2] = x_ab;
    U->v[0] = a; U->v[1] = d; U->v[2] = c;
    U->n[0] = x_dc; U->n[1] = x_ca; U->n[2] = t;
    relink(m, x_bd, u, t, d, b);
    relink(m, x_ca, t, u, a, c);
}

/* number of elements x in sorted arr[0..n) with x < key (strict) or x <= key */
static int count_lt(const int32_t *arr, int n, int key) { int lo = 0, hi = n; while (lo < hi) { int mid = (lo + hi) / 2; if (arr[mid] < key) lo = mid + 1; else hi = mid; } return lo; }
static int count_le(const int32_t *arr, int n, int key) { int lo = 0, hi = n; while (lo < hi) { int mid = (lo + hi) / 2; if (arr[mid] <= key) lo = mid + 1; else hi = mid; } return lo; }

/*
 * Initial triangulation: zipper strips between consecutive non-empty rows + ghost ring.
 * Sites must be sorted row-major (row, then col), all distinct.  Requires >= 2 non-empty rows.
 * Returns number of triangles (2S-2) or <0 on error.
 */
static int build_initial(Mesh *m) {
    int S = m->S;
    const int32_t *row = m->row, *col = m->col;
    int nrows = 0;
    for (int s = 0; s < S; s++) if (s == 0 || row[s] != row[s - 1]) nrows++;
    if (nrows < 2) return -1;
    int *off = (int *)malloc(sizeof(int) * (nrows + 1));
    int *base = (int *)malloc(sizeof(int) * (nrows + 1));
    int k = 0;
    for (int s = 0; s < S; s++) if (s == 0 || row[s] != row[s - 1]) off[k++] = s;
    off[nrows] = S;
    int M = nrows;
    base[0] = 0;
    for (k = 0; k + 1 < M; k++) base[k + 1] = base[k] + (off[k + 1] - off[k] - 1) + (off[k + 2] - off[k + 1] - 1);
    int NT0 = base[M - 1];
    int p0 = off[1] - off[0], pl = off[M] - off[M - 1];
    int G = (p0 - 1) + (M - 1) + (pl - 1) + (M - 1);
    int g_bot = NT0, g_right = g_bot + (p0 - 1), g_top = g_right + (M - 1), g_left = g_top + (pl - 1);
    if (NT0 + G != 2 * S - 2) { free(off); free(base); return -2; }
    Tri *tri = m->tri;
    m->ntri = NT0 + G;
#define NEXT_G(g) (NT0 + (((g) - NT0 + 1) % G))
#define PREV_G(g) (NT0 + (((g) - NT0 + G - 1) % G))
    for (k = 0; k + 1 < M; k++) {
        const int32_t *A = col + off[k], *B = col + off[k + 1];
        int p = off[k + 1] - off[k], q = off[k + 2] - off[k + 1];
        int T = p + q - 2;
        int lg = g_left + (M - 2 - k), rg = g_right + k;
        for (int j = 0; j + 1 < p; j++) { /* up-triangles: base on the lower row */
            int cb = count_lt(B + 1, q - 1, A[j + 1]);
            int pos = j + cb, id = base[k] + pos;
            Tri *t = &tri[id];
            t->v[0] = off[k] + j; t->v[1] = off[k] + j + 1; t->v[2] = off[k + 1] + cb;
            t->n[0] = (pos + 1 < T) ? id + 1 : rg;
            t->n[1] = (pos > 0) ? id - 1 : lg;
            if (k == 0) t->n[2] = g_bot + j;
            else { /* down-triangle of strip k-1 with this base */
                const int32_t *A2 = col + off[k - 1]; int p2 = off[k] - off[k - 1];
                t->n[2] = base[k - 1] + j + count_le(A2 + 1, p2 - 1, A[j + 1]);
            }
        }
        for (int i = 0; i + 1 < q; i++) { /* down-triangles: base on the upper row */
            int ca = count_le(A + 1, p - 1, B[i + 1]);
            int pos = i + ca, id = base[k] + pos;
            Tri *t = &tri[id];
            t->v[0] = off[k + 1] + i + 1; t->v[1] = off[k + 1] + i; t->v[2] = off[k] + ca;
            t->n[0] = (pos > 0) ? id - 1 : lg;
            t->n[1] = (pos + 1 < T) ? id + 1 : rg;
            if (k + 1 == M - 1) t->n[2] = g_top + (q - 2 - i);
            else {
                const int32_t *B2 = col + off[k + 2]; int q2 = off[k + 3] - off[k + 2];
                t->n[2] = base[k + 1] + i + count_lt(B2 + 1, q2 - 1, B[i + 1]);
            }
        }
        /* right ghost k: boundary edge R_k -> R_{k+1};  left ghost k: L_{k+1} -> L_k */
        Tri *gr = &tri[rg], *gl = &tri[lg];
        gr->v[0] = off[k + 2] - 1; gr->v[1] = off[k + 1] - 1; gr->v[2] = GHOST;
        gl->v[0] = off[k]; gl->v[1] = off[k + 1]; gl->v[2] = GHOST;
        gr->n[0] = PREV_G(rg); gr->n[1] = NEXT_G(rg); gr->n[2] = (T > 0) ? base[k] + T - 1 : lg;
        gl->n[0] = PREV_G(lg); gl->n[1] = NEXT_G(lg); gl->n[2] = (T > 0) ? base[k] : rg;
    }
    for (int j = 0; j + 1 < p0; j++) { /* bottom ghosts: edge A_j -> A_{j+1} */
        int g = g_bot + j; Tri *t = &tri[g];
        t->v[0] = off[0] + j + 1; t->v[1] = off[0] + j; t->v[2] = GHOST;
        t->n[0] = PREV_G(g); t->n[1] = NEXT_G(g);
        t->n[2] = base[0] + j + count_lt(col + off[1] + 1, off[2] - off[1] - 1, col[off[0] + j + 1]);
    }
    for (int i = 0; i + 1 < pl; i++) { /* top ghosts: edge B_{i+1} -> B_i */
        int g = g_top + (pl - 2 - i); Tri *t = &tri[g];
        t->v[0] = off[M - 1] + i; t->v[1] = off[M - 1] + i + 1; t->v[2] = GHOST;
        t->n[0] = PREV_G(g); t->n[1] = NEXT_G(g);
        t->n[2] = base[M - 2] + i + count_le(col + off[M - 2] + 1, off[M - 1] - off[M - 2] - 1, col[off[M - 1] + i + 1]);
    }
    free(off); free(base);
    return m->ntri;
}

/* Consistency check of the mesh: neighbour symmetry, orientation.  0 = ok. */
static int check_mesh(const Mesh *m) {
    for (int t = 0; t < m->ntri; t++) {
        const Tri *T = &m->tri[t];
        int ng = 0;
        for (int i = 0; i < 3; i++) ng += (T->v[i] == GHOST);
        if (ng > 1) return 1;
        if (ng == 0 && orient_s(m, T->v[0], T->v[1], T->v[2]) <= 0) return 2;
        for (int i = 0; i < 3; i++) {
            int u = T->n[i];
            if (u < 0 || u >= m->ntri) return 3;
            int b = T->v[(i + 1) % 3], c = T->v[(i + 2) % 3], ok = 0;
            const Tri *U = &m->tri[u];
            for (int k = 0; k < 3; k++)
                if (U->n[k] == t && U->v[(k + 1) % 3] == c && U->v[(k + 2) % 3] == b) ok = 1;
            if (!ok) return 4;
        }
    }
    return 0;
}

/*
 * Triangulate.  row/col: S distinct sites sorted row-major.  tri_v: out, (2S-2)*3 int32
 * (GHOST=-1 marks hull ghosts).  stats[0]=flips, stats[1]=mesh check of the initial mesh,
 * stats[2]=mesh check of the final mesh, stats[3]=residual ties (edges with incircle==0 and pert==0).
 * Returns the triangle count (2S-2), or <0.
 */
int cdt_triangulate(int S, int img_w, const int32_t *row, const int32_t *col, int32_t *tri_v, int32_t *tri_n, int64_t *stats) {
    Mesh m; m.S = S; m.img_w = img_w; m.row = row; m.col = col;
    m.tri = (Tri *)malloc(sizeof(Tri) * (size_t)(2 * S));
    int nt = build_initial(&m);
    if (nt < 0) { free(m.tri); return nt; }
    stats[1] = check_mesh(&m);
    /* Lawson flips, LIFO work stack of (t*3+i) */
    size_t cap = (size_t)nt * 3 + 16, top = 0;
    int32_t *stack = (int32_t *)malloc(sizeof(int32_t) * cap);
    for (int t = nt - 1; t >= 0; t--) for (int i = 0; i < 3; i++) stack[top++] = t * 3 + i;
    int64_t flips = 0;
    while (top > 0) {
        int e = stack[--top]; int t = e / 3, i = e % 3, u, j;
        if (!needs_flip(&m, t, i, &u, &j)) continue;
        do_flip(&m, t, i, u, j);
        flips++;
        if (top + 6 > cap) { cap *= 2; stack = (int32_t *)realloc(stack, sizeof(int32_t) * cap); }
        for (int k = 0; k < 3; k++) { stack[top++] = t * 3 + k; stack[top++] = u * 3 + k; }
    }
    stats[0] = flips;
    stats[2] = check_mesh(&m);
    int64_t ties = 0;
    for (int t = 0; t < nt; t++) for (int i = 0; i < 3; i++) {
        const Tri *T = &m.tri[t]; int u = T->n[i];
        if (u < t) continue;
        int a = T->v[i], b = T->v[(i + 1) % 3], c = T->v[(i + 2) % 3];
        if (a == GHOST || b == GHOST || c == GHOST) continue;
        const Tri *U = &m.tri[u];
        for (int k = 0; k < 3; k++) if (U->n[k] == t && U->v[(k + 1) % 3] == c && U->v[(k + 2) % 3] == b) {
            int d = U->v[k];
            if (d != GHOST && incircle(&m, a, b, c, d) == 0 && incircle_pert(&m, a, b, c, d) == 0) ties++;
        }
    }
    stats[3] = ties;
    for (int t = 0; t < nt; t++) for (int i = 0; i < 3; i++) { tri_v[t * 3 + i] = m.tri[t].v[i]; if (tri_n) tri_n[t * 3 + i] = m.tri[t].n[i]; }
    free(stack); free(m.tri);
    return nt;
}

/*
 * Validity check of ANY triangulation given as vertex triples (ghost rows allowed, skipped):
 * returns the number of violations = (triangle, site) pairs where an adjacent triangle's opposite
 * vertex lies strictly inside the circumcircle, plus non-CCW triangles.  Also returns in
 * *n_strict the number of triangles none of whose neighbours' opposite vertices is co-circular.
 * strict_flag (may be NULL): per triangle 1 if tie-independent.
 * Adjacency is rebuilt here from the triples with a hash so that the check does not trust tri_n.
 */
typedef struct { uint64_t key; int32_t val; } HEnt;
static inline uint64_t mix64(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }

int64_t cdt_check_delaunay(int S, const int32_t *row, const int32_t *col, int nt, const int32_t *tri_v,
                           int64_t *n_strict, uint8_t *strict_flag, int64_t *n_unmatched_edges) {
    size_t hcap = 1; while (hcap < (size_t)nt * 6 + 16) hcap <<= 1;
    HEnt *h = (HEnt *)malloc(sizeof(HEnt) * hcap);
    for (size_t i = 0; i < hcap; i++) h[i].val = -1;
    Mesh m; m.S = S; m.img_w = 0; m.row = row; m.col = col; m.tri = NULL; m.ntri = nt;
    for (int t = 0; t < nt; t++) {
        const int32_t *v = tri_v + t * 3;
        if (v[0] < 0 || v[1] < 0 || v[2] < 0) continue;
        for (int i = 0; i < 3; i++) {
            uint64_t key = ((uint64_t)(uint32_t)v[(i + 1) % 3] << 32) | (uint32_t)v[(i + 2) % 3];
            size_t p = mix64(key) & (hcap - 1);
            while (h[p].val >= 0) p = (p + 1) & (hcap - 1);
            h[p].key = key; h[p].val = t * 3 + i;
        }
    }
    int64_t viol = 0, strict = 0, unmatched = 0;
    for (int t = 0; t < nt; t++) {
        const int32_t *v = tri_v + t * 3;
        if (v[0] < 0 || v[1] < 0 || v[2] < 0) { if (strict_flag) strict_flag[t] = 0; continue; }
        if (orient_s(&m, v[0], v[1], v[2]) <= 0) viol++;
        int is_strict = 1;
        for (int i = 0; i < 3; i++) {
            uint64_t key = ((uint64_t)(uint32_t)v[(i + 2) % 3] << 32) | (uint32_t)v[(i + 1) % 3];
            size_t p = mix64(key) & (hcap - 1);
            int found = -1;
            while (h[p].val >= 0) { if (h[p].key == key) { found = h[p].val; break; } p = (p + 1) & (hcap - 1); }
            if (found < 0) { unmatched++; continue; }
            int d = tri_v[found];
            int64_t ic = incircle(&m, v[0], v[1], v[2], d);
            if (ic > 0) viol++;
            if (ic == 0) is_strict = 0;
        }
        strict += is_strict;
        if (strict_flag) strict_flag[t] = (uint8_t)is_strict;
    }
    if (n_strict) *n_strict = strict;
    if (n_unmatched_edges) *n_unmatched_edges = unmatched; /* = number of hull edges */
    free(h);
    return viol;
}

/*
 * Exact integer barycentric rasterisation of real triangles into interp (img_h*img_w*3 u8, pre-zeroed
 * by the caller) and hull (img_h*img_w u8).  rgb: S*3 u8.  tri_id (may be NULL): img_h*img_w int32,
 * last triangle written per pixel (-1 if none).
 */
void cdt_rasterize(int S, int img_h, int img_w, const int32_t *row, const int32_t *col, const uint8_t *rgb,
                   int nt, const int32_t *tri_v, uint8_t *interp, uint8_t *hull, int32_t *tri_id) {
    (void)S;
    if (tri_id) for (int i = 0; i < img_h * img_w; i++) tri_id[i] = -1;
    for (int t = 0; t < nt; t++) {
        const int32_t *v = tri_v + t * 3;
        if (v[0] < 0 || v[1] < 0 || v[2] < 0) continue;
        int64_t ax = col[v[0]], ay = row[v[0]], bx = col[v[1]], by = row[v[1]], cx = col[v[2]], cy = row[v[2]];
        int64_t A2 = orient2d(ax, ay, bx, by, cx, cy);
        if (A2 <= 0) continue;
        int x0 = (int)(ax < bx ? (ax < cx ? ax : cx) : (bx < cx ? bx : cx));
        int x1 = (int)(ax > bx ? (ax > cx ? ax : cx) : (bx > cx ? bx : cx));
        int y0 = (int)(ay < by ? (ay < cy ? ay : cy) : (by < cy ? by : cy));
        int y1 = (int)(ay > by ? (ay > cy ? ay : cy) : (by > cy ? by : cy));
        for (int y = y0; y <= y1; y++) for (int x = x0; x <= x1; x++) {
            int64_t wa = orient2d(bx, by, cx, cy, x, y), wb = orient2d(cx, cy, ax, ay, x, y), wc = orient2d(ax, ay, bx, by, x, y);
            if (wa < 0 || wb < 0 || wc < 0) continue;
            size_t p = (size_t)y * img_w + x;
            for (int ch = 0; ch < 3; ch++)
                interp[p * 3 + ch] = (uint8_t)((wa * rgb[v[0] * 3 + ch] + wb * rgb[v[1] * 3 + ch] + wc * rgb[v[2] * 3 + ch]) / A2);
            hull[p] = 1;
            if (tri_id) tri_id[p] = t;
        }
    }
}
