// api.cu -- context, scratch management and the extern "C" entry points of libsalve_bev.so.
// See include/salve_bev.h for the contract of each function and the reference code it replaces.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/salve_bev.h"
#include "bev_common.cuh"
#include "k_flip.cuh"
#include "k_image.cuh"
#include "k_layout.cuh"
#include "k_preprocess.cuh"
#include "k_raster.cuh"
#include "k_sites.cuh"
#include "k_splat.cuh"

using namespace bev;

static thread_local char g_err[512] = "";
static void set_err(const char* fmt, const char* a, const char* b, int line) { snprintf(g_err, sizeof(g_err), fmt, a, b, line); }
#define CU(call)                                                                        \
    do {                                                                                \
        cudaError_t e_ = (call);                                                        \
        if (e_ != cudaSuccess) {                                                        \
            set_err("CUDA error: %s at %s (api.cu:%d)", cudaGetErrorString(e_), #call, __LINE__); \
            return SALVE_BEV_E_CUDA;                                                    \
        }                                                                               \
    } while (0)
#define FAIL(code, msg)                                 \
    do {                                                \
        set_err("%s%s (api.cu:%d)", msg, "", __LINE__); \
        return (code);                                  \
    } while (0)

constexpr int N_TMP = 8;
#ifndef IMAGE_LOCAL_RULE
#define IMAGE_LOCAL_RULE 1
#endif
constexpr int N_STAGE_EVENTS = 8;  // chunk start, after: splat, sites, prep, local, window, shade, finish

struct salve_bev_ctx {
    salve_bev_config cfg;
    GridParams G;
    size_t img_bytes;
    // pano slots
    uint8_t* pano_rgb_store = nullptr;
    uint16_t* pano_depth_store = nullptr;
    uint8_t* pano_rgb2x_store = nullptr;  // lazily allocated: full-resolution (2H, 2W) rgb of the slots uploaded that way
    std::vector<const uint8_t*> h_rgb_ptr;
    std::vector<const uint16_t*> h_depth_ptr;
    const uint16_t** d_depth_ptr = nullptr;
    const uint16_t** h_depth_pin = nullptr;  // pinned staging of the pointer table: its upload never blocks the host
    cudaEvent_t ev_ptr = nullptr;
    bool ptr_dirty = true;
    int32_t* h_rep = nullptr; size_t h_rep_cap = 0; cudaEvent_t ev_rep = nullptr;  // pinned staging of replicate_images_kernel's index tables
    // sphere tables (cos_phi[H], neg_sin_phi[H], cos_theta[W], sin_theta[W])
    double* d_tables = nullptr;
    // per-image scratch (strides in elements)
    uint32_t *keygrid = nullptr, *color = nullptr, *occ = nullptr, *nonempty = nullptr, *keep = nullptr, *tmpbits = nullptr;
    uint16_t* wprefix = nullptr;
    Tri* tris = nullptr;
    unsigned long long* owner = nullptr;
    uint32_t *list0 = nullptr, *list1 = nullptr, *cand = nullptr;
    // state between the stages of the image pipeline (k_image.cuh), per image of a chunk
    uint32_t* planes = nullptr;          // 3 bit planes (occupancy, non-empty, keep)
    uint32_t* defer_planes = nullptr;    // 1 bit plane: queries handed to the cooperative pass
    unsigned char* rowarr = nullptr;     // row arrays (RA_*), rows_stride bytes per image
    int32_t* hdr = nullptr;              // HD_STRIDE int32 per image
    uint32_t* qlist = nullptr;           // query pixels for the window pass (g entries per image)
    uint32_t* clist = nullptr;           // edge-rule pixels, then the entries handed from the window pass to the cooperative pass
    unsigned long long* qres = nullptr;  // per list entry: the triangle the window pass reached
    int32_t* work_counter = nullptr;     // window stage: next block of the chunk-wide query list
    int32_t* d_order = nullptr;          // finish stage: hand-out order (image_order_kernel)
    uint32_t* local_lut = nullptr;       // prep stage: tables of the local rule (LocalRule::WORDS words)
    bool local_rule = true;
    size_t rows_stride = 0;
    int n_sm = 0;
    // The key grid is all zero between calls: the sites stage zeroes the keys it consumes, so no chunk pays for a memset of
    // 1 MB per image.  Stage taps re-splat the image they look at from the job table of the last chunk.
    std::vector<SplatJob> last_jobs;
    // hypothesis-independent (un-posed pano 2) renders of the current call: max_panos x 2 surfaces
    uint8_t* cache_out = nullptr; int32_t* cache_counts = nullptr; int32_t* cache_status = nullptr;
    int32_t* d_dest = nullptr;           // per image of the chunk: destination (see ImageArgs::dest)
    int32_t* h_dest[2] = {nullptr, nullptr};
    bool dedup_unposed = true;
    // verifier pre-processing: resize taps of the last (src, resize, crop) combination, pinned staging of the pointer table
    ResizeTap* d_taps = nullptr; int taps_key[4] = {0, 0, 0, 0};
    const uint8_t** h_pp_src = nullptr; const uint8_t** d_pp_src = nullptr; size_t pp_cap = 0; cudaEvent_t ev_pp = nullptr;
    size_t g_stride = 0, bits_stride = 0, tris_stride = 0, cand_stride = 0;
    ImgHeader* headers = nullptr;
    int32_t* counts = nullptr;
    int32_t* status = nullptr;
    SplatJob* d_jobs = nullptr;
    const uint8_t** d_color_src = nullptr;
    uint8_t* out_store = nullptr;  // 2 x max_images images (double buffered), for the *_host variants
    // host-output pipeline: chunk k+1 renders while chunk k is copied device->host on copy_stream
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_done[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr}, ev_staged[2] = {nullptr, nullptr};
    cudaEvent_t ev_cache = nullptr;  // the un-posed cache of the current call is complete
    SplatJob* h_jobs[2] = {nullptr, nullptr};          // pinned staging of the per-chunk job tables
    const uint8_t** h_src[2] = {nullptr, nullptr};
    int stage_parity = 0;
    int32_t* h_meta = nullptr;  // pinned staging of counts/status for the *_host variants (user arrays may be pageable,
    size_t h_meta_cap = 0;      //  and a device->pageable async copy would block the host and serialise the pipeline)
    // growable temporaries
    void* tmp[N_TMP] = {nullptr};
    size_t tmp_bytes[N_TMP] = {0};
    // timing
    bool timing = false;
    std::vector<cudaEvent_t> events;  // N_STAGE_EVENTS per chunk of the last call
    size_t events_used = 0;
    int64_t launches = 0;
    int last_chunk_images = 0;
    int32_t* last_counts = nullptr;  // device counters of the last chunk (taps re-run the image kernel on a copy)
    size_t flip_smem = 0;
    int max_smem_optin = 0;
    double band[4] = {-INFINITY, -1.0, 0.5, INFINITY};  // bev_rendering_utils.py:560-566
};

static int tmp_get(salve_bev_ctx* c, int slot, size_t bytes, void** out) {
    if (c->tmp_bytes[slot] < bytes) {
        if (c->tmp[slot]) CU(cudaFree(c->tmp[slot]));
        c->tmp[slot] = nullptr; c->tmp_bytes[slot] = 0;
        size_t want = std::max(bytes, (size_t)1 << 20);
        CU(cudaMalloc(&c->tmp[slot], want));
        c->tmp_bytes[slot] = want;
    }
    *out = c->tmp[slot];
    return SALVE_BEV_OK;
}

extern "C" const char* salve_bev_last_error(void) { return g_err; }

extern "C" void salve_bev_default_config(salve_bev_config* cfg, int32_t pano_h, int32_t pano_w) {
    memset(cfg, 0, sizeof(*cfg));
    cfg->device = 0;
    cfg->pano_h = pano_h; cfg->pano_w = pano_w;
    cfg->max_panos = 64; cfg->max_images = 592;
    cfg->grid_h = 501; cfg->grid_w = 501; cfg->kernel_sz = 11;
    cfg->xmin = -5.0; cfg->ymin = -5.0; cfg->xmax = 5.0; cfg->ymax = 5.0; cfg->px_per_m = 1.0 / 0.02;
    cfg->depth_scale = 0.001f;
    cfg->crop_rows = (int32_t)(pano_h * (80.0 / 512.0));  // int(H * crop_ratio), bev_rendering_utils.py:399,613
}

static void compute_tables_libm(int H, int W, std::vector<double>& t) {
    t.resize(2 * (size_t)H + 2 * (size_t)W);
    const double pi = 3.141592653589793;
    for (int v = 0; v < H; v++) {
        double phi = (v + 0.5) / H; phi -= 0.5; phi *= pi;
        t[v] = cos(phi); t[H + v] = -sin(phi);
    }
    for (int u = 0; u < W; u++) {
        double th = -(u + 0.5) / W; th *= 2 * pi;
        t[2 * H + u] = cos(th); t[2 * H + W + u] = sin(th);
    }
}

// Tables of the local rule (k_image.cuh, LocalRule): every counter-clockwise triangle of three of the 12 near neighbours that
// contains the query (the origin) and whose circle -- the lattice points strictly inside it and on it -- stays within the 5 x 5
// neighbourhood; per 12-bit neighbour pattern the candidates whose vertices are present and which have none of the pattern's
// sites strictly inside (the first four: 224 patterns with co-circular sites at distance 2 have 7 or 14, and a query whose triangle
// is one of the others simply goes on to the window pass).
static const std::vector<uint32_t>& local_rule_tables() {
    static const std::vector<uint32_t> tab = [] {
        std::vector<uint32_t> t(LocalRule::WORDS, 0u);
        struct P { int x, y; };
        auto pos = [](int k) { const int b = LocalRule::pos_bit(k); return P{b % 5 - 2, b / 5 - 2}; };
        auto orient = [](P a, P b, P c) { return (b.x - a.x) * (c.y - a.y) - (b.y - a.y) * (c.x - a.x); };
        auto incircle = [](P a, P b, P c, P d) {
            const long long ax = a.x - d.x, ay = a.y - d.y, bx = b.x - d.x, by = b.y - d.y, cx = c.x - d.x, cy = c.y - d.y;
            return (ax * ax + ay * ay) * (bx * cy - by * cx) - (bx * bx + by * by) * (ax * cy - ay * cx) + (cx * cx + cy * cy) * (ax * by - ay * bx);
        };
        struct Cand { uint32_t vmask12, inside12; };
        std::vector<Cand> cands;
        const P o{0, 0};
        for (int i = 0; i < LocalRule::NPOS; i++)
            for (int j = i + 1; j < LocalRule::NPOS; j++)
                for (int k = j + 1; k < LocalRule::NPOS; k++) {
                    P a = pos(i), b = pos(j), c = pos(k);
                    int ib = j, ic = k;
                    const int orr = orient(a, b, c);
                    if (orr == 0) continue;
                    if (orr < 0) { std::swap(b, c); std::swap(ib, ic); }
                    if (orient(a, b, o) < 0 || orient(b, c, o) < 0 || orient(c, a, o) < 0) continue;
                    uint32_t inside = 0u, on = 0u;
                    bool fits = true;
                    for (int y = -8; y <= 8 && fits; y++)
                        for (int x = -8; x <= 8; x++) {
                            const P d{x, y};
                            if ((x == a.x && y == a.y) || (x == b.x && y == b.y) || (x == c.x && y == c.y)) continue;
                            const long long inc = incircle(a, b, c, d);
                            if (inc < 0) continue;
                            if (x < -2 || x > 2 || y < -2 || y > 2) { fits = false; break; }
                            if (x == 0 && y == 0) continue;  // the query itself is no site
                            (inc > 0 ? inside : on) |= 1u << ((y + 2) * 5 + x + 2);
                        }
                    if (!fits || (int)cands.size() >= LocalRule::MAXCAND) continue;
                    const int id = (int)cands.size();
                    t[LocalRule::OFF_INSIDE + id] = inside;
                    t[LocalRule::OFF_ON + id] = on;
                    t[LocalRule::OFF_VERTS + id] = (uint32_t)LocalRule::pos_bit(i) | ((uint32_t)LocalRule::pos_bit(ib) << 5) | ((uint32_t)LocalRule::pos_bit(ic) << 10);
                    uint32_t in12 = 0u;
                    for (int m = 0; m < LocalRule::NPOS; m++) if (inside & (1u << LocalRule::pos_bit(m))) in12 |= 1u << m;
                    cands.push_back({(1u << i) | (1u << j) | (1u << k), in12});
                }
        for (uint32_t pat = 0; pat < (uint32_t)LocalRule::NPAT; pat++) {
            uint32_t e = 0xFFFFFFFFu;
            int n = 0;
            for (size_t id = 0; id < cands.size() && n < 4; id++)
                if ((pat & cands[id].vmask12) == cands[id].vmask12 && !(pat & cands[id].inside12)) {
                    e = (e & ~(0xFFu << (8 * n))) | ((uint32_t)id << (8 * n));
                    n++;
                }
            t[pat] = e;
        }
        return t;
    }();
    return tab;
}

static int ctx_init(salve_bev_ctx* c, const salve_bev_config* cfg);

extern "C" int salve_bev_ctx_create(const salve_bev_config* cfg, salve_bev_ctx** out) {
    if (!cfg || !out) FAIL(SALVE_BEV_E_INVALID, "null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) FAIL(SALVE_BEV_E_NODEVICE, "no CUDA device: this library has no CPU fallback");
    if (cfg->device < 0 || cfg->device >= ndev) FAIL(SALVE_BEV_E_INVALID, "bad device ordinal");
    if (cfg->pano_h < 1 || cfg->pano_w < 4 || (cfg->pano_w & 3)) FAIL(SALVE_BEV_E_INVALID, "pano_w must be a positive multiple of 4");
    if ((int64_t)cfg->pano_h * cfg->pano_w > ((int64_t)1 << KEY_IDX_BITS)) FAIL(SALVE_BEV_E_INVALID, "pano too large");
    if (cfg->grid_h < 2 || cfg->grid_h > MAX_GRID_H || cfg->grid_w < 2 || cfg->grid_w > 2047 ||
        (int64_t)cfg->grid_h * cfg->grid_w > 800000)
        FAIL(SALVE_BEV_E_INVALID, "grid must satisfy 2 <= grid_h <= 1023, 2 <= grid_w <= 2047, grid_h*grid_w <= 800000");
    if (cfg->kernel_sz < 1 || !(cfg->kernel_sz & 1) || cfg->kernel_sz > 63) FAIL(SALVE_BEV_E_INVALID, "kernel_sz must be odd, 1..63");
    if (cfg->max_images < 1 || cfg->max_panos < 1) FAIL(SALVE_BEV_E_INVALID, "max_images/max_panos must be positive");
    if (cfg->crop_rows < 0 || 2 * cfg->crop_rows >= cfg->pano_h) FAIL(SALVE_BEV_E_INVALID, "bad crop_rows");
    CU(cudaSetDevice(cfg->device));
    salve_bev_ctx* c = new salve_bev_ctx();
    const int rc = ctx_init(c, cfg);
    if (rc != SALVE_BEV_OK) { salve_bev_ctx_destroy(c); return rc; }  // frees whatever was allocated before the failure
    *out = c;
    return SALVE_BEV_OK;
}

static int ctx_init(salve_bev_ctx* c, const salve_bev_config* cfg) {
    c->cfg = *cfg;
    c->G.grid_h = cfg->grid_h; c->G.grid_w = cfg->grid_w; c->G.wpr = (cfg->grid_w + 31) / 32;
    c->G.g = cfg->grid_h * cfg->grid_w; c->G.K = cfg->kernel_sz;
    c->img_bytes = (size_t)c->G.g * 3;
    const size_t H = cfg->pano_h, W = cfg->pano_w, P = cfg->max_panos, N = cfg->max_images, g = c->G.g;
    c->g_stride = (g + 31) & ~(size_t)31;
    c->bits_stride = ((size_t)c->G.grid_h * c->G.wpr + 31) & ~(size_t)31;
    c->tris_stride = 2 * c->g_stride;
    c->cand_stride = 3 * c->g_stride;
#define ALLOC(ptr, count) CU(cudaMalloc((void**)&(ptr), sizeof(*(ptr)) * (size_t)(count)))
    ALLOC(c->pano_rgb_store, P * H * W * 3);
    ALLOC(c->pano_depth_store, P * H * W);
    ALLOC(c->d_depth_ptr, P);
    ALLOC(c->d_tables, 2 * H + 2 * W);
    ALLOC(c->keygrid, N * c->g_stride);
    CU(cudaMemset(c->keygrid, 0, sizeof(uint32_t) * N * c->g_stride));
    CU(cudaDeviceGetAttribute(&c->n_sm, cudaDevAttrMultiProcessorCount, cfg->device));
    // lists are sized for the worst case (every pixel of every image a query); only what a chunk really uses is ever touched
    ALLOC(c->qlist, N * c->g_stride);
    ALLOC(c->clist, N * c->g_stride);
    ALLOC(c->qres, N * c->g_stride);
    ALLOC(c->planes, N * 3 * c->bits_stride);
    ALLOC(c->defer_planes, N * c->bits_stride);
    c->rows_stride = image_rows_stride(MAX_GRID_H + 1);  // any grid height the generic interp entry point accepts
    ALLOC(c->rowarr, N * c->rows_stride);
    ALLOC(c->hdr, N * HD_STRIDE);
    ALLOC(c->work_counter, 1);
    ALLOC(c->d_order, N);
    ALLOC(c->local_lut, LocalRule::WORDS);
    {
        const std::vector<uint32_t>& lut = local_rule_tables();
        CU(cudaMemcpy(c->local_lut, lut.data(), sizeof(uint32_t) * LocalRule::WORDS, cudaMemcpyHostToDevice));
    }
    ALLOC(c->cache_out, 2 * P * c->img_bytes + 64);  // +64: replicate_images_kernel reads whole words
    ALLOC(c->cache_counts, 2 * P * 8);
    ALLOC(c->cache_status, 2 * P);
    ALLOC(c->d_dest, N);
    // mesh scratch (explicit triangulation: grids too large for the image stages' shared memory, and the triangle tap): ONE image
    ALLOC(c->color, c->g_stride);
    ALLOC(c->occ, c->bits_stride);
    ALLOC(c->nonempty, c->bits_stride);
    ALLOC(c->keep, c->bits_stride);
    ALLOC(c->tmpbits, c->bits_stride);
    ALLOC(c->wprefix, c->bits_stride);
    ALLOC(c->tris, c->tris_stride);
    ALLOC(c->owner, c->tris_stride);
    ALLOC(c->list0, c->tris_stride);
    ALLOC(c->list1, c->tris_stride);
    ALLOC(c->cand, c->cand_stride);
    ALLOC(c->headers, N);
    ALLOC(c->counts, 2 * N * 8);
    ALLOC(c->status, 2 * N);
    ALLOC(c->d_jobs, N);
    ALLOC(c->d_color_src, N);
    ALLOC(c->out_store, 2 * N * c->img_bytes);
#undef ALLOC
    CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&c->ev_cache, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_ptr, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_rep, cudaEventDisableTiming));
    CU(cudaMallocHost((void**)&c->h_depth_pin, sizeof(void*) * P));
    for (int k = 0; k < 2; k++) {
        CU(cudaEventCreateWithFlags(&c->ev_done[k], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->ev_copied[k], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->ev_staged[k], cudaEventDisableTiming));
        CU(cudaMallocHost((void**)&c->h_jobs[k], sizeof(SplatJob) * N));
        CU(cudaMallocHost((void**)&c->h_src[k], sizeof(void*) * N));
        CU(cudaMallocHost((void**)&c->h_dest[k], sizeof(int32_t) * N));
    }
    c->h_rgb_ptr.assign(P, nullptr);
    c->h_depth_ptr.assign(P, nullptr);
    for (size_t s = 0; s < P; s++) {
        c->h_rgb_ptr[s] = c->pano_rgb_store + s * H * W * 3;
        c->h_depth_ptr[s] = c->pano_depth_store + s * H * W;
    }
    std::vector<double> t;
    compute_tables_libm((int)H, (int)W, t);
    CU(cudaMemcpy(c->d_tables, t.data(), t.size() * sizeof(double), cudaMemcpyHostToDevice));
    c->flip_smem = ((2 * g + 31) / 32) * 4 + 16;
    CU(cudaFuncSetAttribute(flip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->flip_smem));
    CU(cudaDeviceGetAttribute(&c->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, cfg->device));
    {
        cudaFuncAttributes fa, fb, fc;
        CU(cudaFuncGetAttributes(&fa, finish_stage_kernel<true>));
        CU(cudaFuncGetAttributes(&fb, finish_stage_kernel<false>));
        CU(cudaFuncGetAttributes(&fc, prep_stage_kernel));
        c->max_smem_optin -= (int)std::max(std::max(fa.sharedSizeBytes, fb.sharedSizeBytes), fc.sharedSizeBytes);  // what is left for dynamic shared memory
        CU(cudaFuncSetAttribute(finish_stage_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem_optin));
        CU(cudaFuncSetAttribute(finish_stage_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem_optin));
        CU(cudaFuncSetAttribute(prep_stage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem_optin));
        CU(cudaFuncSetAttribute(sites_stage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem_optin));
    }
    return SALVE_BEV_OK;
}

extern "C" void salve_bev_ctx_destroy(salve_bev_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->cfg.device);
    cudaDeviceSynchronize();
    void* ptrs[] = {c->pano_rgb_store, c->pano_rgb2x_store, c->pano_depth_store, c->d_depth_ptr, c->d_tables, c->keygrid, c->color, c->occ, c->nonempty,
                    c->keep, c->tmpbits, c->wprefix, c->tris, c->owner, c->list0, c->list1, c->cand, c->qlist, c->clist, c->qres, c->planes, c->defer_planes, c->rowarr, c->hdr, c->work_counter, c->d_order, c->local_lut, c->cache_out, c->cache_counts, c->cache_status, c->d_dest, c->headers, c->counts,
                    c->status, c->d_jobs, c->d_color_src, c->out_store};
    for (void* p : ptrs) if (p) cudaFree(p);
    for (int i = 0; i < N_TMP; i++) if (c->tmp[i]) cudaFree(c->tmp[i]);
    for (cudaEvent_t e : c->events) cudaEventDestroy(e);
    for (int k = 0; k < 2; k++) {
        if (c->ev_done[k]) cudaEventDestroy(c->ev_done[k]);
        if (c->ev_copied[k]) cudaEventDestroy(c->ev_copied[k]);
        if (c->ev_staged[k]) cudaEventDestroy(c->ev_staged[k]);
        if (c->h_jobs[k]) cudaFreeHost(c->h_jobs[k]);
        if (c->h_src[k]) cudaFreeHost(c->h_src[k]);
        if (c->h_dest[k]) cudaFreeHost(c->h_dest[k]);
    }
    if (c->ev_cache) cudaEventDestroy(c->ev_cache);
    if (c->ev_ptr) cudaEventDestroy(c->ev_ptr);
    if (c->ev_rep) cudaEventDestroy(c->ev_rep);
    if (c->h_depth_pin) cudaFreeHost(c->h_depth_pin);
    if (c->h_rep) cudaFreeHost(c->h_rep);
    if (c->ev_pp) cudaEventDestroy(c->ev_pp);
    if (c->h_pp_src) cudaFreeHost(c->h_pp_src);
    if (c->d_pp_src) cudaFree(c->d_pp_src);
    if (c->d_taps) cudaFree(c->d_taps);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->h_meta) cudaFreeHost(c->h_meta);
    delete c;
}

extern "C" int salve_bev_set_sphere_tables(salve_bev_ctx* c, const double* cp, const double* nsp, const double* ct, const double* st) {
    if (!c || !cp || !nsp || !ct || !st) FAIL(SALVE_BEV_E_INVALID, "null argument");
    const size_t H = c->cfg.pano_h, W = c->cfg.pano_w;
    CU(cudaSetDevice(c->cfg.device));
    CU(cudaMemcpy(c->d_tables, cp, H * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(c->d_tables + H, nsp, H * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(c->d_tables + 2 * H, ct, W * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(c->d_tables + 2 * H + W, st, W * sizeof(double), cudaMemcpyHostToDevice));
    return SALVE_BEV_OK;
}

__global__ void sphere_xyz_kernel(const double* t, int H, int W, double* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= H * W) return;
    const int v = i / W, u = i % W;
    const double cphi = t[v];
    out[i * 3 + 0] = __dmul_rn(cphi, t[2 * H + u]);
    out[i * 3 + 1] = __dmul_rn(cphi, t[2 * H + W + u]);
    out[i * 3 + 2] = t[H + v];
}

extern "C" int salve_bev_get_uni_sphere_xyz(salve_bev_ctx* c, double* host_out) {
    if (!c || !host_out) FAIL(SALVE_BEV_E_INVALID, "null argument");
    CU(cudaSetDevice(c->cfg.device));
    const int H = c->cfg.pano_h, W = c->cfg.pano_w;
    void* buf;
    int rc = tmp_get(c, 0, (size_t)H * W * 3 * sizeof(double), &buf);
    if (rc) return rc;
    sphere_xyz_kernel<<<(H * W + 255) / 256, 256>>>(c->d_tables, H, W, (double*)buf);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpy(host_out, buf, (size_t)H * W * 3 * sizeof(double), cudaMemcpyDeviceToHost));
    return SALVE_BEV_OK;
}

extern "C" int salve_bev_upload_pano(salve_bev_ctx* c, int32_t slot, const uint8_t* host_rgb, const uint16_t* host_depth, void* stream) {
    if (!c || !host_rgb || !host_depth) FAIL(SALVE_BEV_E_INVALID, "null argument");
    if (slot < 0 || slot >= c->cfg.max_panos) FAIL(SALVE_BEV_E_CAPACITY, "pano slot out of range");
    CU(cudaSetDevice(c->cfg.device));
    const size_t H = c->cfg.pano_h, W = c->cfg.pano_w;
    uint8_t* drgb = c->pano_rgb_store + (size_t)slot * H * W * 3;
    uint16_t* dd = c->pano_depth_store + (size_t)slot * H * W;
    CU(cudaMemcpyAsync(drgb, host_rgb, H * W * 3, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    CU(cudaMemcpyAsync(dd, host_depth, H * W * 2, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    if (c->h_rgb_ptr[slot] != drgb || c->h_depth_ptr[slot] != dd) { c->h_rgb_ptr[slot] = drgb; c->h_depth_ptr[slot] = dd; c->ptr_dirty = true; }  // untagged: 1x
    return SALVE_BEV_OK;
}

static int bind_pano_impl(salve_bev_ctx* c, int32_t slot, const uint8_t* dev_rgb, const uint16_t* dev_depth, bool fullres) {
    if (!c || !dev_rgb || !dev_depth) FAIL(SALVE_BEV_E_INVALID, "null argument");
    if (slot < 0 || slot >= c->cfg.max_panos) FAIL(SALVE_BEV_E_CAPACITY, "pano slot out of range");
    if (((uintptr_t)dev_depth & 7) != 0) FAIL(SALVE_BEV_E_INVALID, "depth pointer must be 8-byte aligned");
    if (((uintptr_t)dev_rgb & 1) != 0) FAIL(SALVE_BEV_E_INVALID, "rgb pointer must be 2-byte aligned");
    // bit 0 of a colour source tags a full-resolution pano (gather_rgb)
    c->h_rgb_ptr[slot] = reinterpret_cast<const uint8_t*>((uintptr_t)dev_rgb | (fullres ? 1u : 0u));
    c->h_depth_ptr[slot] = dev_depth; c->ptr_dirty = true;
    return SALVE_BEV_OK;
}
extern "C" int salve_bev_bind_pano(salve_bev_ctx* c, int32_t slot, const uint8_t* dev_rgb, const uint16_t* dev_depth) {
    return bind_pano_impl(c, slot, dev_rgb, dev_depth, false);
}
extern "C" int salve_bev_bind_pano_fullres(salve_bev_ctx* c, int32_t slot, const uint8_t* dev_rgb_2x, const uint16_t* dev_depth) {
    return bind_pano_impl(c, slot, dev_rgb_2x, dev_depth, true);
}
extern "C" int salve_bev_upload_pano_fullres(salve_bev_ctx* c, int32_t slot, const uint8_t* host_rgb_2x, const uint16_t* host_depth, void* stream) {
    if (!c || !host_rgb_2x || !host_depth) FAIL(SALVE_BEV_E_INVALID, "null argument");
    if (slot < 0 || slot >= c->cfg.max_panos) FAIL(SALVE_BEV_E_CAPACITY, "pano slot out of range");
    CU(cudaSetDevice(c->cfg.device));
    const size_t H = c->cfg.pano_h, W = c->cfg.pano_w, P = c->cfg.max_panos;
    if (!c->pano_rgb2x_store) CU(cudaMalloc((void**)&c->pano_rgb2x_store, P * H * W * 12));
    uint8_t* drgb = c->pano_rgb2x_store + (size_t)slot * H * W * 12;
    uint16_t* dd = c->pano_depth_store + (size_t)slot * H * W;
    CU(cudaMemcpyAsync(drgb, host_rgb_2x, H * W * 12, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    CU(cudaMemcpyAsync(dd, host_depth, H * W * 2, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return bind_pano_impl(c, slot, drgb, dd, true);
}

static SplatParams make_splat_params(salve_bev_ctx* c) {
    SplatParams P;
    const int H = c->cfg.pano_h, W = c->cfg.pano_w;
    P.H = H; P.W = W; P.crop_rows = c->cfg.crop_rows; P.depth_scale = c->cfg.depth_scale;
    P.rows_per_thread = 8;  // splat_pano_kernel launches override it (splat_rows_for)
    P.xmin = c->cfg.xmin; P.ymin = c->cfg.ymin; P.xmax = c->cfg.xmax; P.ymax = c->cfg.ymax; P.px_per_m = c->cfg.px_per_m;
    P.a_lo = c->band[0]; P.a_hi = c->band[1]; P.b_lo = c->band[2]; P.b_hi = c->band[3];
    P.grid_w = c->G.grid_w; P.g = c->G.g;
    P.cos_phi = c->d_tables; P.neg_sin_phi = c->d_tables + H; P.cos_theta = c->d_tables + 2 * H; P.sin_theta = c->d_tables + 2 * H + W;
    P.depth = c->d_depth_ptr;
    return P;
}

static int sync_ptr_tables(salve_bev_ctx* c, cudaStream_t st) {
    if (!c->ptr_dirty) return SALVE_BEV_OK;
    CU(cudaEventSynchronize(c->ev_ptr));  // the previous upload out of the pinned table is done (it was issued a bind ago)
    memcpy(c->h_depth_pin, c->h_depth_ptr.data(), sizeof(void*) * c->cfg.max_panos);
    CU(cudaMemcpyAsync(c->d_depth_ptr, c->h_depth_pin, sizeof(void*) * c->cfg.max_panos, cudaMemcpyHostToDevice, st));
    CU(cudaEventRecord(c->ev_ptr, st));
    c->ptr_dirty = false;
    return SALVE_BEV_OK;
}

static int stage_event(salve_bev_ctx* c, cudaStream_t st) {
    if (!c->timing) return SALVE_BEV_OK;
    if (c->events_used == c->events.size()) { cudaEvent_t e; CU(cudaEventCreate(&e)); c->events.push_back(e); }
    CU(cudaEventRecord(c->events[c->events_used++], st));
    return SALVE_BEV_OK;
}

// Everything after the splat for images [0, n_img): the six stages of k_image.cuh (sites, prep, local, window, shade, finish).
static int run_image_stage(salve_bev_ctx* c, int n_img, const GridParams& G, uint32_t* keygrid, const uint8_t* const* color_src,
                           uint8_t* dev_out, int32_t* dev_counts, int32_t* dev_status, int raw_mode, int skip_empty, uint8_t* hull,
                           int32_t* qtri, cudaStream_t st, const int32_t* dest = nullptr, int32_t* counts_out = nullptr, bool clear_keys = false, bool timed = false) {
    const size_t smem = image_smem_bytes(G.grid_h, G.wpr);
    if (smem > (size_t)c->max_smem_optin) FAIL(SALVE_BEV_E_CAPACITY, "grid too large for the image stages' shared memory");
    if (n_img > c->cfg.max_images) FAIL(SALVE_BEV_E_CAPACITY, "more images than the context's scratch holds");
    // a launch of the window stage indexes at most WIN_MAX_IMAGES images: larger chunks go through in groups
    for (int g0 = 0; g0 < n_img; g0 += WIN_MAX_IMAGES) {
        const int n = std::min(WIN_MAX_IMAGES, n_img - g0);
        ImageArgs IA;
        IA.G = G;
        IA.n_img = n;
        IA.order = nullptr;
        IA.keygrid = keygrid + (size_t)g0 * c->g_stride; IA.keygrid_stride = c->g_stride;
        IA.color_src = color_src + g0; IA.pano_w = c->cfg.pano_w;
        IA.counts = dev_counts + (size_t)g0 * 8; IA.status = dev_status;
        IA.out = dev_out; IA.out_stride = (size_t)G.g * 3;
        IA.dest = dest ? dest + g0 : nullptr; IA.counts_out = counts_out;
        if (!dest) {  // destination = image index: shift the destination arrays with the group
            IA.out = dev_out + (size_t)g0 * IA.out_stride;
            IA.status = dev_status ? dev_status + g0 : nullptr;
            IA.counts_out = counts_out ? counts_out + (size_t)g0 * 8 : nullptr;
        }
        IA.cache_out = c->cache_out; IA.cache_counts = c->cache_counts; IA.cache_status = c->cache_status;
        IA.hull = hull ? hull + (size_t)g0 * G.g : nullptr; IA.hull_stride = (size_t)G.g;
        IA.qtri = qtri ? qtri + (size_t)g0 * G.g * 3 : nullptr; IA.qtri_stride = (size_t)G.g * 3;
        IA.planes = c->planes; IA.plane_stride = c->bits_stride; IA.defer = c->defer_planes;
        IA.rows = c->rowarr; IA.rows_stride = c->rows_stride; IA.hp = (G.grid_h + 15) & ~15;
        IA.hdr = c->hdr;
        IA.qlist = c->qlist; IA.qlist_stride = c->g_stride; IA.qres = c->qres; IA.clist = c->clist;
        IA.work_counter = c->work_counter;
        IA.local_lut = (IMAGE_LOCAL_RULE && c->local_rule) ? c->local_lut : nullptr;
        IA.raw_mode = raw_mode; IA.skip_empty_check = skip_empty; IA.clear_keys = clear_keys ? 1 : 0;
        CU(cudaMemsetAsync(c->work_counter, 0, sizeof(int32_t), st));
        const bool ev = timed && n_img <= WIN_MAX_IMAGES;  // per-stage events: one group per chunk (else only the chunk total is kept)
        int rc;
        sites_stage_kernel<<<dim3((unsigned)((G.grid_h + SITES_WARPS * SITES_GROUPS - 1) / (SITES_WARPS * SITES_GROUPS)), (unsigned)n), SITES_WARPS * 32, sites_smem_bytes(G.grid_w, G.wpr), st>>>(IA);
        if (ev && (rc = stage_event(c, st))) return rc;
        prep_stage_kernel<<<n, PREP_NT, prep_smem_bytes(G.grid_h, G.wpr), st>>>(IA);
        if (ev && (rc = stage_event(c, st))) return rc;
        if (IA.local_lut && !qtri) {  // (with the triangle tap the prep stage leaves the queries in the window list)
            local_stage_kernel<<<dim3(LOCAL_SPLIT, (unsigned)n), LOCAL_NT, 0, st>>>(IA);
            c->launches++;
        }
        if (ev && (rc = stage_event(c, st))) return rc;
        const int win_ctas = std::max(1, std::min(c->n_sm * IMAGE_WIN_CTAS, n * 16));
        window_stage_kernel<IMAGE_WIN_NR><<<win_ctas, WIN_NT, 0, st>>>(IA);
        if (ev && (rc = stage_event(c, st))) return rc;
        shade_stage_kernel<<<dim3(SHADE_SPLIT, (unsigned)n), SHADE_NT, 0, st>>>(IA);
        if (ev && (rc = stage_event(c, st))) return rc;
        c->launches += 4;
        if (n > 2 * c->n_sm) {  // more images than CTA slots: longest expected first
            image_order_kernel<<<(n + 255) / 256, 256, 0, st>>>(c->hdr, n, c->d_order);
            c->launches++;
            IA.order = c->d_order;
        }
        // grids of up to 512 x 512 pixels (the reference's 501 x 501 included) take the instantiation with int32 circle parameters
        if (G.grid_h <= 512 && G.grid_w <= 512) finish_stage_kernel<true><<<n, FINISH_NT, finish_smem_bytes(G.grid_h, G.wpr), st>>>(IA);
        else finish_stage_kernel<false><<<n, FINISH_NT, finish_smem_bytes(G.grid_h, G.wpr), st>>>(IA);
        c->launches++;
        CU(cudaGetLastError());
    }
    if (timed && n_img > WIN_MAX_IMAGES)
        for (int k = 0; k < 5; k++) { int rc = stage_event(c, st); if (rc) return rc; }  // keep the event layout: stages not separated
    return timed ? stage_event(c, st) : SALVE_BEV_OK;
}

// Explicit-mesh path on ONE image (mesh scratch): sites + zipper, parallel Lawson flips, and (optionally) the rasteriser.
// Used for grids that exceed the image stages' shared memory and for the triangle tap.
static int run_mesh_stages(salve_bev_ctx* c, const GridParams& G, const uint32_t* keygrid, const uint8_t* const* color_src, uint8_t* dev_out,
                           int32_t* dev_counts, int32_t* dev_status, int raw_mode, int skip_empty, uint8_t* hull, bool raster, cudaStream_t st) {
    SitesArgs SA;
    SA.G = G;
    SA.keygrid = const_cast<uint32_t*>(keygrid); SA.keygrid_stride = c->g_stride;
    SA.color = c->color; SA.color_stride = c->g_stride;
    SA.occ = c->occ; SA.nonempty = c->nonempty; SA.keep = c->keep; SA.tmpbits = c->tmpbits; SA.bits_stride = c->bits_stride;
    SA.wprefix = c->wprefix;
    SA.tris = c->tris; SA.tris_stride = c->tris_stride;
    SA.headers = c->headers; SA.counts = dev_counts; SA.status = dev_status;
    SA.color_src = color_src; SA.pano_w = c->cfg.pano_w;
    SA.out = dev_out; SA.out_stride = (size_t)G.g * 3;
    SA.raw_mode = raw_mode; SA.skip_empty_check = skip_empty;
    sites_kernel<<<1, SITES_NT, 0, st>>>(SA);
    c->launches++;
    CU(cudaGetLastError());

    FlipArgs FA;
    FA.grid_w = G.grid_w;
    FA.tris = c->tris; FA.tris_stride = c->tris_stride;
    FA.owner = c->owner; FA.owner_stride = c->tris_stride;
    FA.list0 = c->list0; FA.list1 = c->list1; FA.list_stride = c->tris_stride;
    FA.cand = c->cand; FA.cand_stride = c->cand_stride;
    FA.headers = c->headers; FA.counts = dev_counts;
    const size_t flip_smem = ((2 * (size_t)G.g + 31) / 32) * 4 + 16;
    // the attribute is per function, not per context: (re)assert it for this launch
    CU(cudaFuncSetAttribute(flip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(flip_smem, (size_t)49152)));
    flip_kernel<<<1, FLIP_NT, flip_smem, st>>>(FA);
    c->launches++;
    CU(cudaGetLastError());
    if (!raster) return SALVE_BEV_OK;

    RasterArgs RA;
    RA.G = G;
    RA.tris = c->tris; RA.tris_stride = c->tris_stride;
    RA.color = c->color; RA.color_stride = c->g_stride;
    RA.keep = c->keep; RA.bits_stride = c->bits_stride;
    RA.headers = c->headers; RA.counts = dev_counts; RA.status = dev_status;
    RA.out = dev_out; RA.out_stride = (size_t)G.g * 3;
    RA.hull = hull; RA.hull_stride = (size_t)G.g;
    RA.raw_mode = raw_mode;
    raster_kernel<<<1, RASTER_NT, 0, st>>>(RA);
    c->launches++;
    CU(cudaGetLastError());
    return SALVE_BEV_OK;
}

// One chunk of pano-sourced images.  jobs / color slots are host arrays.
static int render_chunk(salve_bev_ctx* c, int n_img, const std::vector<SplatJob>& jobs, const std::vector<int>& img_slot, uint8_t* dev_out,
                        int32_t* dev_counts, int32_t* dev_status, cudaStream_t st, const std::vector<int32_t>* dest = nullptr,
                        int32_t* counts_out = nullptr) {
    if (n_img > c->cfg.max_images || (int)jobs.size() > c->cfg.max_images) FAIL(SALVE_BEV_E_CAPACITY, "chunk exceeds max_images");
    int rc = sync_ptr_tables(c, st); if (rc) return rc;
    // job tables go through pinned, double-buffered staging so that the chunk loop never blocks the host
    const int sp = c->stage_parity; c->stage_parity ^= 1;
    CU(cudaEventSynchronize(c->ev_staged[sp]));  // the copy issued two chunks ago from this staging buffer has been consumed
    for (int i = 0; i < n_img; i++) {
        if (img_slot[i] < 0 || img_slot[i] >= c->cfg.max_panos) FAIL(SALVE_BEV_E_CAPACITY, "pano slot out of range");
        c->h_src[sp][i] = c->h_rgb_ptr[img_slot[i]];
    }
    memcpy(c->h_jobs[sp], jobs.data(), sizeof(SplatJob) * jobs.size());
    CU(cudaMemcpyAsync(c->d_jobs, c->h_jobs[sp], sizeof(SplatJob) * jobs.size(), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(c->d_color_src, c->h_src[sp], sizeof(void*) * n_img, cudaMemcpyHostToDevice, st));
    if (dest) {
        memcpy(c->h_dest[sp], dest->data(), sizeof(int32_t) * n_img);
        CU(cudaMemcpyAsync(c->d_dest, c->h_dest[sp], sizeof(int32_t) * n_img, cudaMemcpyHostToDevice, st));
    }
    CU(cudaEventRecord(c->ev_staged[sp], st));
    if (!dev_counts) dev_counts = c->counts;
    rc = stage_event(c, st); if (rc) return rc;
    CU(cudaMemsetAsync(dev_counts, 0, sizeof(int32_t) * 8 * n_img, st));
    SplatParams P = make_splat_params(c);
    const int rows = P.H - 2 * P.crop_rows;
    P.rows_per_thread = splat_rows_for(jobs.size());
    dim3 grid((unsigned)(((rows + P.rows_per_thread - 1) / P.rows_per_thread) * ((P.W + 1023) >> 10)), (unsigned)jobs.size());
    splat_pano_kernel<<<grid, 256, 0, st>>>(P, c->d_jobs, c->keygrid, c->g_stride, dev_counts);
    c->launches++;
    CU(cudaGetLastError());
    rc = stage_event(c, st); if (rc) return rc;
    c->last_chunk_images = n_img;
    c->last_counts = dev_counts;
    c->last_jobs = jobs;
    return run_image_stage(c, n_img, c->G, c->keygrid, c->d_color_src, dev_out, dev_counts, dev_status, 0, 0, nullptr, nullptr, st,
                           dest ? c->d_dest : nullptr, counts_out, true, true);
}

// Copy images between two device buffers at any alignment (an image is 753 003 bytes: consecutive images share no
// alignment).  Copy k: cache image src_idx[k] -> out image dst_idx[k], plus its counters and status.
#ifndef REPLICATE_SPLIT
#define REPLICATE_SPLIT 16  // CTAs per copied image
#endif
__global__ void __launch_bounds__(256) replicate_images_kernel(const uint8_t* __restrict__ cache, uint8_t* __restrict__ out,
                                                              const int32_t* __restrict__ src_idx, const int32_t* __restrict__ dst_idx,
                                                              size_t bytes, const int32_t* __restrict__ cache_counts,
                                                              int32_t* __restrict__ counts, const int32_t* __restrict__ cache_status,
                                                              int32_t* __restrict__ status) {
    const int k = blockIdx.y;
    const int si = src_idx[k], di = dst_idx[k];
    if (blockIdx.x == 0 && threadIdx.x < 9) {
        if (threadIdx.x < 8) { if (counts) counts[(size_t)di * 8 + threadIdx.x] = cache_counts[(size_t)si * 8 + threadIdx.x]; }
        else if (status) status[di] = cache_status[si];
    }
    if (!out) return;
    const uint8_t* s = cache + (size_t)si * bytes;
    uint8_t* d = out + (size_t)di * bytes;
    size_t head = (16 - ((uintptr_t)d & 15)) & 15;
    if (head > bytes) head = bytes;
    const size_t nchunks = (bytes - head) / 16;
    const size_t tail0 = head + nchunks * 16;
    for (size_t ch = (size_t)blockIdx.x * blockDim.x + threadIdx.x; ch < nchunks; ch += (size_t)gridDim.x * blockDim.x) {
        const uintptr_t sa = (uintptr_t)(s + head + ch * 16);
        const uint32_t* w = reinterpret_cast<const uint32_t*>(sa & ~(uintptr_t)3);
        const int sh = (int)(sa & 3) * 8;
        const uint32_t a0 = w[0], a1 = w[1], a2 = w[2], a3 = w[3], a4 = sh ? w[4] : 0u;
        uint4 v;
        v.x = __funnelshift_r(a0, a1, sh); v.y = __funnelshift_r(a1, a2, sh); v.z = __funnelshift_r(a2, a3, sh); v.w = __funnelshift_r(a3, a4, sh);
        *reinterpret_cast<uint4*>(d + head + ch * 16) = v;
    }
    if (blockIdx.x == 0) {
        if (threadIdx.x < head) d[threadIdx.x] = s[threadIdx.x];
        if (tail0 + threadIdx.x < bytes) d[tail0 + threadIdx.x] = s[tail0 + threadIdx.x];
    }
}

// ---- de-duplicated rendering ----------------------------------------------------------------------------------------------
// img2 of render_bev_pair does not depend on the hypothesis: only xyzrgb1 is posed (bev_rendering_utils.py:451), pano 2 is
// rendered in its own frame (:455).  A call that names the same pano 2 in several hypotheses renders it once per surface
// into the context's cache; every hypothesis then gets a copy (full layout) or an index (compact layout).
// Image stream of a call: first the unique un-posed images (unique pano 2's in order of first appearance x surfaces), then the
// posed images (hypothesis-major, surface-minor), cut into chunks of whole jobs.
struct HypPlan {
    std::vector<int32_t> uniq;        // unique pano-2 slots
    std::vector<int32_t> uniq_of_hyp; // per hypothesis: index into uniq
};
static void make_plan(int32_t n_hyp, const int32_t* p2, int max_panos, HypPlan& P) {
    std::vector<int32_t> where(max_panos, -1);
    P.uniq.clear(); P.uniq_of_hyp.resize(n_hyp);
    for (int h = 0; h < n_hyp; h++) {
        int32_t& w = where[p2[h]];
        if (w < 0) { w = (int32_t)P.uniq.size(); P.uniq.push_back(p2[h]); }
        P.uniq_of_hyp[h] = w;
    }
}

// full = true : out holds n_hyp * nsurf * 2 images in the order of salve_bev_render_hypotheses (posed, un-posed per surface)
// full = false: out holds the n_hyp * nsurf posed images, out_unposed the uniq * nsurf un-posed ones
static int render_hyp_dedup(salve_bev_ctx* c, int32_t n_hyp, const int32_t* p1, const float* R, const float* t, uint32_t surfaces,
                            bool full, bool host_out, uint8_t* out, uint8_t* out_unposed, int32_t* counts, int32_t* counts_unposed,
                            int32_t* status, int32_t* status_unposed, const HypPlan& plan, cudaStream_t st) {
    const bool do_f = surfaces & SALVE_BEV_SURF_FLOOR, do_c = surfaces & SALVE_BEV_SURF_CEILING;
    const int nsurf = (int)do_f + (int)do_c;
    const int nU = (int)plan.uniq.size();
    const int jobs_per_chunk = c->cfg.max_images / nsurf;
    if (jobs_per_chunk < 1) FAIL(SALVE_BEV_E_CAPACITY, "max_images too small for one hypothesis");
    const int n_jobs = nU + n_hyp;
    const size_t n_posed = (size_t)n_hyp * nsurf, n_unposed = (size_t)nU * nsurf;
    const size_t n_user = full ? 2 * n_posed : n_posed;  // images in `out`
    const size_t ib = c->img_bytes;
    const size_t N = c->cfg.max_images;
    int rc;
    int32_t *hm_counts = nullptr, *hm_status = nullptr, *hm_ucounts = nullptr, *hm_ustatus = nullptr;
    if (host_out) {
        const size_t need = (n_user + n_unposed) * 9;
        if (c->h_meta_cap < need) {
            if (c->h_meta) CU(cudaFreeHost(c->h_meta));
            c->h_meta = nullptr; c->h_meta_cap = 0;
            CU(cudaMallocHost((void**)&c->h_meta, sizeof(int32_t) * need));
            c->h_meta_cap = need;
        }
        hm_counts = c->h_meta; hm_status = hm_counts + n_user * 8;
        hm_ucounts = hm_status + n_user; hm_ustatus = hm_ucounts + n_unposed * 8;
    }
    std::vector<SplatJob> jobs;
    std::vector<int> slots;
    std::vector<int32_t> dest;
    int chunk_no = 0;
    for (int j0 = 0; j0 < n_jobs; j0 += jobs_per_chunk, chunk_no++) {
        const int nj = std::min(jobs_per_chunk, n_jobs - j0);
        const int n_img = nj * nsurf;
        const int par = chunk_no & 1;
        jobs.clear(); slots.assign(n_img, 0); dest.assign(n_img, 0);
        int first_posed = -1;  // chunk-local index of the first posed image (posed images are a suffix of the chunk)
        for (int k = 0; k < nj; k++) {
            const int j = j0 + k;
            SplatJob a;
            a.img_floor = do_f ? k * nsurf : -1;
            a.img_ceil = do_c ? k * nsurf + (do_f ? 1 : 0) : -1;
            if (j < nU) {
                a.pano_slot = plan.uniq[j]; a.posed = 0;
                a.R[0] = 1.f; a.R[1] = 0.f; a.R[2] = 0.f; a.R[3] = 1.f; a.t[0] = 0.f; a.t[1] = 0.f;
                for (int s = 0; s < nsurf; s++) dest[k * nsurf + s] = -1 - (j * nsurf + s);
            } else {
                const int h = j - nU;
                a.pano_slot = p1[h]; a.posed = 1;
                memcpy(a.R, R + 4 * (size_t)h, sizeof(float) * 4); memcpy(a.t, t + 2 * (size_t)h, sizeof(float) * 2);
                if (first_posed < 0) first_posed = k * nsurf;
                for (int s = 0; s < nsurf; s++) {
                    const int li = k * nsurf + s;
                    // host output: posed images go through the double-buffered staging store in chunk-local order
                    dest[li] = host_out ? li : (full ? (h * nsurf + s) * 2 : h * nsurf + s);
                }
            }
            if (a.pano_slot < 0 || a.pano_slot >= c->cfg.max_panos) FAIL(SALVE_BEV_E_CAPACITY, "pano slot out of range");
            for (int s = 0; s < nsurf; s++) slots[k * nsurf + s] = a.pano_slot;
            jobs.push_back(a);
        }
        if (host_out) {
            CU(cudaStreamWaitEvent(st, c->ev_copied[par], 0));  // staging buffer `par` was drained (chunk k-2)
            uint8_t* stage = c->out_store + par * N * ib;
            int32_t* cnt_stage = c->counts + par * N * 8;
            int32_t* st_stage = c->status + par * N;
            rc = render_chunk(c, n_img, jobs, slots, stage, cnt_stage, st_stage, st, &dest, nullptr);
            if (rc) return rc;
            if (first_posed >= 0) {
                const int np = n_img - first_posed;
                const size_t fin0 = (size_t)(std::max(j0, nU) - nU) * nsurf;  // index among the posed images
                CU(cudaEventRecord(c->ev_done[par], st));
                CU(cudaStreamWaitEvent(c->copy_stream, c->ev_done[par], 0));
                const uint8_t* src = stage + (size_t)first_posed * ib;
                const size_t k = full ? 2 : 1;  // posed image i of the call is image k*i of `out`
                if (full) CU(cudaMemcpy2DAsync(out + fin0 * 2 * ib, 2 * ib, src, ib, ib, np, cudaMemcpyDeviceToHost, c->copy_stream));
                else CU(cudaMemcpyAsync(out + fin0 * ib, src, (size_t)np * ib, cudaMemcpyDeviceToHost, c->copy_stream));
                CU(cudaMemcpy2DAsync(hm_counts + fin0 * k * 8, k * 32, cnt_stage + (size_t)first_posed * 8, 32, 32, np, cudaMemcpyDeviceToHost, c->copy_stream));
                CU(cudaMemcpy2DAsync(hm_status + fin0 * k, k * 4, st_stage + first_posed, 4, 4, np, cudaMemcpyDeviceToHost, c->copy_stream));
                CU(cudaEventRecord(c->ev_copied[par], c->copy_stream));
            }
            if (j0 < nU && j0 + nj >= nU) {
                // the cache is complete once this chunk is done: its device->host copies overlap the posed chunks that follow
                CU(cudaEventRecord(c->ev_cache, st));
                CU(cudaStreamWaitEvent(c->copy_stream, c->ev_cache, 0));
                if (full) {
                    for (int h = 0; h < n_hyp; h++)
                        for (int s = 0; s < nsurf; s++)
                            CU(cudaMemcpyAsync(out + ((size_t)(h * nsurf + s) * 2 + 1) * ib, c->cache_out + (size_t)(plan.uniq_of_hyp[h] * nsurf + s) * ib, ib,
                                               cudaMemcpyDeviceToHost, c->copy_stream));
                } else {
                    CU(cudaMemcpyAsync(out_unposed, c->cache_out, n_unposed * ib, cudaMemcpyDeviceToHost, c->copy_stream));
                }
                CU(cudaMemcpyAsync(hm_ucounts, c->cache_counts, sizeof(int32_t) * 8 * n_unposed, cudaMemcpyDeviceToHost, c->copy_stream));
                CU(cudaMemcpyAsync(hm_ustatus, c->cache_status, sizeof(int32_t) * n_unposed, cudaMemcpyDeviceToHost, c->copy_stream));
            }
        } else {
            rc = render_chunk(c, n_img, jobs, slots, out, c->counts, status, st, &dest, counts);
            if (rc) return rc;
        }
    }
    // ---- device output: un-posed images from the cache to their destinations
    if (!host_out && full && n_posed) {
        void *dsrc, *ddst;
        if ((rc = tmp_get(c, 5, sizeof(int32_t) * n_posed, &dsrc))) return rc;
        if ((rc = tmp_get(c, 6, sizeof(int32_t) * n_posed, &ddst))) return rc;
        CU(cudaEventSynchronize(c->ev_rep));  // the previous call's copies out of the pinned tables are done
        if (c->h_rep_cap < 2 * n_posed) {
            if (c->h_rep) CU(cudaFreeHost(c->h_rep));
            c->h_rep = nullptr; c->h_rep_cap = 0;
            CU(cudaMallocHost((void**)&c->h_rep, sizeof(int32_t) * 2 * n_posed));
            c->h_rep_cap = 2 * n_posed;
        }
        int32_t *hs = c->h_rep, *hd = c->h_rep + n_posed;
        for (int h = 0; h < n_hyp; h++)
            for (int s = 0; s < nsurf; s++) {
                hs[(size_t)h * nsurf + s] = plan.uniq_of_hyp[h] * nsurf + s;
                hd[(size_t)h * nsurf + s] = (h * nsurf + s) * 2 + 1;
            }
        CU(cudaMemcpyAsync(dsrc, hs, sizeof(int32_t) * n_posed, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(ddst, hd, sizeof(int32_t) * n_posed, cudaMemcpyHostToDevice, st));
        CU(cudaEventRecord(c->ev_rep, st));
        replicate_images_kernel<<<dim3(REPLICATE_SPLIT, (unsigned)n_posed), 256, 0, st>>>(c->cache_out, out, (const int32_t*)dsrc, (const int32_t*)ddst, ib,
                                                                            c->cache_counts, counts, c->cache_status, status);
        c->launches++;
        CU(cudaGetLastError());
    } else if (!host_out && n_unposed) {
        CU(cudaMemcpyAsync(out_unposed, c->cache_out, n_unposed * ib, cudaMemcpyDeviceToDevice, st));
        if (counts_unposed) CU(cudaMemcpyAsync(counts_unposed, c->cache_counts, sizeof(int32_t) * 8 * n_unposed, cudaMemcpyDeviceToDevice, st));
        if (status_unposed) CU(cudaMemcpyAsync(status_unposed, c->cache_status, sizeof(int32_t) * n_unposed, cudaMemcpyDeviceToDevice, st));
    }
    if (host_out) {
        CU(cudaStreamSynchronize(c->copy_stream));
        CU(cudaStreamSynchronize(st));
        if (full) {
            for (int h = 0; h < n_hyp; h++)
                for (int s = 0; s < nsurf; s++) {
                    const size_t d = (size_t)(h * nsurf + s) * 2 + 1, u = (size_t)plan.uniq_of_hyp[h] * nsurf + s;
                    memcpy(hm_counts + d * 8, hm_ucounts + u * 8, 32);
                    hm_status[d] = hm_ustatus[u];
                }
        }
        if (counts) memcpy(counts, hm_counts, sizeof(int32_t) * 8 * n_user);
        if (status) memcpy(status, hm_status, sizeof(int32_t) * n_user);
        if (counts_unposed) memcpy(counts_unposed, hm_ucounts, sizeof(int32_t) * 8 * n_unposed);
        if (status_unposed) memcpy(status_unposed, hm_ustatus, sizeof(int32_t) * n_unposed);
    }
    return SALVE_BEV_OK;
}

static int render_hyp_impl(salve_bev_ctx* c, int32_t n_hyp, const int32_t* p1, const int32_t* p2, const float* R, const float* t,
                           uint32_t surfaces, uint8_t* out, int32_t* counts, int32_t* status, bool host_out, cudaStream_t st) {
    if (!c || !p1 || !p2 || !R || !t || !out) FAIL(SALVE_BEV_E_INVALID, "null argument");
    if (n_hyp < 0) FAIL(SALVE_BEV_E_INVALID, "negative n_hyp");
    const bool do_f = surfaces & SALVE_BEV_SURF_FLOOR, do_c = surfaces & SALVE_BEV_SURF_CEILING;
    const int nsurf = (int)do_f + (int)do_c;
    if (nsurf == 0 || (surfaces & ~3u)) FAIL(SALVE_BEV_E_INVALID, "bad surface mask");
    const int per_hyp = nsurf * 2;
    const int hyp_per_chunk = c->cfg.max_images / per_hyp;
    if (hyp_per_chunk < 1) FAIL(SALVE_BEV_E_CAPACITY, "max_images too small for one hypothesis");
    CU(cudaSetDevice(c->cfg.device));
    c->events_used = 0;
    for (int h = 0; h < n_hyp; h++)
        if (p1[h] < 0 || p1[h] >= c->cfg.max_panos || p2[h] < 0 || p2[h] >= c->cfg.max_panos) FAIL(SALVE_BEV_E_CAPACITY, "pano slot out of range");
    if (c->dedup_unposed && n_hyp > 1) {
        HypPlan plan;
        make_plan(n_hyp, p2, c->cfg.max_panos, plan);
        if ((int)plan.uniq.size() < n_hyp)  // some pano 2 repeats: render each un-posed image once
            return render_hyp_dedup(c, n_hyp, p1, R, t, surfaces, true, host_out, out, nullptr, counts, nullptr, status, nullptr, plan, st);
    }
    std::vector<SplatJob> jobs;
    std::vector<int> slots;
    const size_t n_img_total = (size_t)n_hyp * per_hyp;
    if (host_out && (counts || status)) {
        if (c->h_meta_cap < n_img_total * 9) {
            if (c->h_meta) CU(cudaFreeHost(c->h_meta));
            c->h_meta = nullptr; c->h_meta_cap = 0;
            CU(cudaMallocHost((void**)&c->h_meta, sizeof(int32_t) * n_img_total * 9));
            c->h_meta_cap = n_img_total * 9;
        }
    }
    int32_t* h_counts = c->h_meta;
    int32_t* h_status = c->h_meta ? c->h_meta + n_img_total * 8 : nullptr;
    for (int h0 = 0; h0 < n_hyp; h0 += hyp_per_chunk) {
        const int nh = std::min(hyp_per_chunk, n_hyp - h0);
        const int n_img = nh * per_hyp;
        jobs.clear(); slots.assign(n_img, 0);
        for (int k = 0; k < nh; k++) {
            const int h = h0 + k;
            const int base = k * per_hyp;
            const int f1 = do_f ? base + 0 : -1, f2 = do_f ? base + 1 : -1;
            const int c1 = do_c ? base + (do_f ? 2 : 0) : -1, c2 = do_c ? base + (do_f ? 3 : 1) : -1;
            SplatJob a; a.pano_slot = p1[h]; a.posed = 1;
            memcpy(a.R, R + 4 * (size_t)h, sizeof(float) * 4); memcpy(a.t, t + 2 * (size_t)h, sizeof(float) * 2);
            a.img_floor = f1; a.img_ceil = c1;
            SplatJob b; b.pano_slot = p2[h]; b.posed = 0;
            b.R[0] = 1.f; b.R[1] = 0.f; b.R[2] = 0.f; b.R[3] = 1.f; b.t[0] = 0.f; b.t[1] = 0.f;
            b.img_floor = f2; b.img_ceil = c2;
            if (a.pano_slot < 0 || a.pano_slot >= c->cfg.max_panos || b.pano_slot < 0 || b.pano_slot >= c->cfg.max_panos)
                FAIL(SALVE_BEV_E_CAPACITY, "pano slot out of range");
            jobs.push_back(a); jobs.push_back(b);
            if (f1 >= 0) { slots[f1] = a.pano_slot; slots[f2] = b.pano_slot; }
            if (c1 >= 0) { slots[c1] = a.pano_slot; slots[c2] = b.pano_slot; }
        }
        const size_t img0 = (size_t)h0 * per_hyp;
        const int par = (int)((h0 / hyp_per_chunk) & 1);
        const size_t N = c->cfg.max_images;
        uint8_t* d_out = host_out ? c->out_store + par * N * c->img_bytes : out + img0 * c->img_bytes;
        int32_t* d_counts = host_out ? c->counts + par * N * 8 : (counts ? counts + img0 * 8 : nullptr);
        int32_t* d_status = host_out ? c->status + par * N : (status ? status + img0 : nullptr);
        if (host_out) CU(cudaStreamWaitEvent(st, c->ev_copied[par], 0));  // buffer `par` was drained (chunk k-2)
        int rc = render_chunk(c, n_img, jobs, slots, d_out, d_counts, d_status, st);
        if (rc) return rc;
        if (host_out) {
            // device->host on the copy stream, overlapping the next chunk's kernels
            CU(cudaEventRecord(c->ev_done[par], st));
            CU(cudaStreamWaitEvent(c->copy_stream, c->ev_done[par], 0));
            CU(cudaMemcpyAsync(out + img0 * c->img_bytes, d_out, (size_t)n_img * c->img_bytes, cudaMemcpyDeviceToHost, c->copy_stream));
            if (counts) CU(cudaMemcpyAsync(h_counts + img0 * 8, d_counts, sizeof(int32_t) * 8 * n_img, cudaMemcpyDeviceToHost, c->copy_stream));
            if (status) CU(cudaMemcpyAsync(h_status + img0, d_status, sizeof(int32_t) * n_img, cudaMemcpyDeviceToHost, c->copy_stream));
            CU(cudaEventRecord(c->ev_copied[par], c->copy_stream));
        }
    }
    if (host_out) {
        CU(cudaStreamSynchronize(c->copy_stream));
        CU(cudaStreamSynchronize(st));
        if (counts) memcpy(counts, h_counts, sizeof(int32_t) * 8 * n_img_total);
        if (status) memcpy(status, h_status, sizeof(int32_t) * n_img_total);
    }
    return SALVE_BEV_OK;
}

extern "C" int salve_bev_render_hypotheses(salve_bev_ctx* c, int32_t n_hyp, const int32_t* p1, const int32_t* p2, const float* R,
                                           const float* t, uint32_t surfaces, uint8_t* dev_out, int32_t* dev_counts, int32_t* dev_status,
                                           void* stream) {
    return render_hyp_impl(c, n_hyp, p1, p2, R, t, surfaces, dev_out, dev_counts, dev_status, false, (cudaStream_t)stream);
}
extern "C" int salve_bev_render_hypotheses_host(salve_bev_ctx* c, int32_t n_hyp, const int32_t* p1, const int32_t* p2, const float* R,
                                                const float* t, uint32_t surfaces, uint8_t* host_out, int32_t* host_counts,
                                                int32_t* host_status, void* stream) {
    return render_hyp_impl(c, n_hyp, p1, p2, R, t, surfaces, host_out, host_counts, host_status, true, (cudaStream_t)stream);
}

static int render_compact(salve_bev_ctx* c, int32_t n_hyp, const int32_t* p1, const int32_t* p2, const float* R, const float* t, uint32_t surfaces,
                          bool host_out, uint8_t* out_posed, uint8_t* out_unposed, int32_t* unposed_of_hyp, int32_t* n_unique, int32_t* counts_posed,
                          int32_t* counts_unposed, int32_t* status_posed, int32_t* status_unposed, cudaStream_t st) {
    if (!c || !p1 || !p2 || !R || !t || !out_posed || !out_unposed || !unposed_of_hyp || !n_unique) FAIL(SALVE_BEV_E_INVALID, "null argument");
    if (n_hyp < 0) FAIL(SALVE_BEV_E_INVALID, "negative n_hyp");
    if ((surfaces & 3u) == 0 || (surfaces & ~3u)) FAIL(SALVE_BEV_E_INVALID, "bad surface mask");
    for (int h = 0; h < n_hyp; h++)
        if (p1[h] < 0 || p1[h] >= c->cfg.max_panos || p2[h] < 0 || p2[h] >= c->cfg.max_panos) FAIL(SALVE_BEV_E_CAPACITY, "pano slot out of range");
    CU(cudaSetDevice(c->cfg.device));
    c->events_used = 0;
    HypPlan plan;
    make_plan(n_hyp, p2, c->cfg.max_panos, plan);
    *n_unique = (int32_t)plan.uniq.size();
    for (int h = 0; h < n_hyp; h++) unposed_of_hyp[h] = plan.uniq_of_hyp[h];
    return render_hyp_dedup(c, n_hyp, p1, R, t, surfaces, false, host_out, out_posed, out_unposed, counts_posed, counts_unposed, status_posed,
                            status_unposed, plan, st);
}
extern "C" int salve_bev_render_hypotheses_compact(salve_bev_ctx* c, int32_t n_hyp, const int32_t* p1, const int32_t* p2, const float* R,
                                                   const float* t, uint32_t surfaces, uint8_t* dev_posed, uint8_t* dev_unposed,
                                                   int32_t* host_unposed_of_hyp, int32_t* host_n_unique, int32_t* dev_counts_posed,
                                                   int32_t* dev_counts_unposed, int32_t* dev_status_posed, int32_t* dev_status_unposed,
                                                   void* stream) {
    return render_compact(c, n_hyp, p1, p2, R, t, surfaces, false, dev_posed, dev_unposed, host_unposed_of_hyp, host_n_unique, dev_counts_posed,
                          dev_counts_unposed, dev_status_posed, dev_status_unposed, (cudaStream_t)stream);
}
extern "C" int salve_bev_render_hypotheses_compact_host(salve_bev_ctx* c, int32_t n_hyp, const int32_t* p1, const int32_t* p2, const float* R,
                                                        const float* t, uint32_t surfaces, uint8_t* host_posed, uint8_t* host_unposed,
                                                        int32_t* host_unposed_of_hyp, int32_t* host_n_unique, int32_t* host_counts_posed,
                                                        int32_t* host_counts_unposed, int32_t* host_status_posed, int32_t* host_status_unposed,
                                                        void* stream) {
    return render_compact(c, n_hyp, p1, p2, R, t, surfaces, true, host_posed, host_unposed, host_unposed_of_hyp, host_n_unique, host_counts_posed,
                          host_counts_unposed, host_status_posed, host_status_unposed, (cudaStream_t)stream);
}
// cv2's recipe for the taps of one axis (imgproc/src/resize.cpp): scale in double, position in float, 11-bit weights
static void linear_taps(int src, int dst, int off, int n, std::vector<ResizeTap>& out) {
    const double scale = (double)src / (double)dst;
    out.resize(n);
    for (int i = 0; i < n; i++) {
        const int d = i + off;
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= (float)s;
        if (s < 0) { f = 0.f; s = 0; }
        if (s >= src - 1) { f = 0.f; s = src - 1; }
        ResizeTap t;
        t.s0 = s; t.s1 = std::min(s + 1, src - 1);
        t.w0 = (int)lrintf((1.f - f) * 2048.f); t.w1 = (int)lrintf(f * 2048.f);
        out[i] = t;
    }
}

extern "C" int salve_bev_verifier_preprocess(salve_bev_ctx* c, int32_t n, const uint8_t* const* host_src, int32_t resize_hw, int32_t crop_hw,
                                             float* dev_out, void* stream) {
    if (!c || (n > 0 && (!host_src || !dev_out))) FAIL(SALVE_BEV_E_INVALID, "null argument");
    if (n < 0 || resize_hw < 1 || crop_hw < 1 || crop_hw > resize_hw || crop_hw > 256) FAIL(SALVE_BEV_E_INVALID, "need 1 <= crop <= resize, crop <= 256");
    if (resize_hw > c->G.grid_h || resize_hw > c->G.grid_w) FAIL(SALVE_BEV_E_INVALID, "only down-scaling is pinned against cv2");
    if (n == 0) return SALVE_BEV_OK;
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(c->cfg.device));
    if (!c->ev_pp) CU(cudaEventCreateWithFlags(&c->ev_pp, cudaEventDisableTiming));
    const int key[4] = {c->G.grid_h, c->G.grid_w, resize_hw, crop_hw};
    if (!c->d_taps || memcmp(key, c->taps_key, sizeof(key)) != 0) {
        const int off = (int)((resize_hw - crop_hw) / 2);  // int((h - crop_h) / 2), transform.py:412-413
        std::vector<ResizeTap> tx, ty;
        linear_taps(c->G.grid_w, resize_hw, off, crop_hw, tx);
        linear_taps(c->G.grid_h, resize_hw, off, crop_hw, ty);
        if (!c->d_taps) CU(cudaMalloc((void**)&c->d_taps, sizeof(ResizeTap) * 512));
        CU(cudaStreamSynchronize(st));
        CU(cudaMemcpy(c->d_taps, tx.data(), sizeof(ResizeTap) * crop_hw, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(c->d_taps + 256, ty.data(), sizeof(ResizeTap) * crop_hw, cudaMemcpyHostToDevice));
        memcpy(c->taps_key, key, sizeof(key));
    }
    const size_t np = (size_t)n * 4;
    if (c->pp_cap < np) {
        CU(cudaEventSynchronize(c->ev_pp));
        if (c->h_pp_src) CU(cudaFreeHost(c->h_pp_src));
        if (c->d_pp_src) CU(cudaFree(c->d_pp_src));
        c->h_pp_src = nullptr; c->d_pp_src = nullptr; c->pp_cap = 0;
        CU(cudaMallocHost((void**)&c->h_pp_src, sizeof(void*) * np));
        CU(cudaMalloc((void**)&c->d_pp_src, sizeof(void*) * np));
        c->pp_cap = np;
    }
    CU(cudaEventSynchronize(c->ev_pp));  // the previous call's copy out of the pinned table is done
    memcpy(c->h_pp_src, host_src, sizeof(void*) * np);
    CU(cudaMemcpyAsync(c->d_pp_src, c->h_pp_src, sizeof(void*) * np, cudaMemcpyHostToDevice, st));
    CU(cudaEventRecord(c->ev_pp, st));
    PreprocArgs A;
    A.src = c->d_pp_src; A.xtap = c->d_taps; A.ytap = c->d_taps + 256;
    A.src_w = c->G.grid_w; A.crop_h = crop_hw; A.crop_w = crop_hw;
    const double mean[3] = {0.485, 0.456, 0.406}, stdv[3] = {0.229, 0.224, 0.225};  // normalization_utils.py:21-25
    for (int k = 0; k < 3; k++) { A.mean[k] = (float)(mean[k] * 255); A.stdv[k] = (float)(stdv[k] * 255); }
    A.out = dev_out;
    dim3 grid((crop_hw + PREPROC_ROWS - 1) / PREPROC_ROWS, 4, (unsigned)n);
    verifier_preprocess_kernel<<<grid, 256, 0, st>>>(A);
    c->launches++;
    CU(cudaGetLastError());
    return SALVE_BEV_OK;
}

extern "C" int salve_bev_rasterize_layouts_host(salve_bev_ctx* c, int32_t n_img, const int32_t* host_desc, const int64_t* host_offsets,
                                                const uint8_t* host_init, uint8_t* host_out, void* stream) {
    if (!c || !host_desc || !host_offsets || !host_out) FAIL(SALVE_BEV_E_INVALID, "null argument");
    if (n_img < 0) FAIL(SALVE_BEV_E_INVALID, "negative n_img");
    if (n_img == 0) return SALVE_BEV_OK;
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(c->cfg.device));
    const long long n_words = host_offsets[n_img];
    for (int i = 0; i < n_img; i++) {
        const long long o = host_offsets[i];
        if (o < 0 || o + LD_HDR > host_offsets[i + 1] || host_offsets[i + 1] > n_words) FAIL(SALVE_BEV_E_INVALID, "bad descriptor offsets");
        const int32_t np = host_desc[o + LD_NPOLY], ns = host_desc[o + LD_NSEG];
        if (np < 0 || np > LAYOUT_MAX_POLY || ns < 0 || o + LD_HDR + 2LL * np + 6LL * ns != host_offsets[i + 1])
            FAIL(SALVE_BEV_E_CAPACITY, "layout descriptor: at most 128 polygon vertices; sizes must match the offsets");
    }
    const size_t ib = c->img_bytes;
    void *dd, *doff, *dout, *dinit = nullptr;
    int rc;
    if ((rc = tmp_get(c, 0, sizeof(int32_t) * (size_t)n_words, &dd))) return rc;
    if ((rc = tmp_get(c, 1, sizeof(long long) * (size_t)(n_img + 1), &doff))) return rc;
    if ((rc = tmp_get(c, 2, ib * n_img, &dout))) return rc;
    if (host_init && (rc = tmp_get(c, 3, ib * n_img, &dinit))) return rc;
    CU(cudaMemcpyAsync(dd, host_desc, sizeof(int32_t) * (size_t)n_words, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(doff, host_offsets, sizeof(long long) * (size_t)(n_img + 1), cudaMemcpyHostToDevice, st));
    if (host_init) CU(cudaMemcpyAsync(dinit, host_init, ib * n_img, cudaMemcpyHostToDevice, st));
    LayoutArgs A;
    A.desc = (const int32_t*)dd; A.offset = (const long long*)doff; A.out = (uint8_t*)dout; A.out_stride = ib;
    A.init = (const uint8_t*)dinit; A.h = c->G.grid_h; A.w = c->G.grid_w;
    layout_raster_kernel<<<n_img, LAYOUT_NT, 0, st>>>(A);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(host_out, dout, ib * n_img, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return SALVE_BEV_OK;
}

extern "C" int salve_bev_set_dedup_unposed(salve_bev_ctx* c, int32_t on) {
    if (!c) FAIL(SALVE_BEV_E_INVALID, "null argument");
    c->dedup_unposed = on != 0;
    return SALVE_BEV_OK;
}

extern "C" int salve_bev_render_images_host(salve_bev_ctx* c, int32_t n_img, const int32_t* slot, const int32_t* surface,
                                            const int32_t* posed, const float* R, const float* t, uint8_t* host_out, int32_t* host_counts,
                                            int32_t* host_status, void* stream) {
    if (!c || !slot || !surface || !posed || !R || !t || !host_out) FAIL(SALVE_BEV_E_INVALID, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(c->cfg.device));
    c->events_used = 0;
    std::vector<SplatJob> jobs;
    std::vector<int> slots;
    for (int i0 = 0; i0 < n_img; i0 += c->cfg.max_images) {
        const int n = std::min(c->cfg.max_images, n_img - i0);
        jobs.clear(); slots.assign(n, 0);
        for (int k = 0; k < n; k++) {
            const int i = i0 + k;
            SplatJob j; j.pano_slot = slot[i]; j.posed = posed[i] ? 1 : 0;
            memcpy(j.R, R + 4 * (size_t)i, sizeof(float) * 4); memcpy(j.t, t + 2 * (size_t)i, sizeof(float) * 2);
            if (surface[i] == SALVE_BEV_SURF_FLOOR) { j.img_floor = k; j.img_ceil = -1; }
            else if (surface[i] == SALVE_BEV_SURF_CEILING) { j.img_floor = -1; j.img_ceil = k; }
            else FAIL(SALVE_BEV_E_INVALID, "surface must be SALVE_BEV_SURF_FLOOR or SALVE_BEV_SURF_CEILING");
            if (j.pano_slot < 0 || j.pano_slot >= c->cfg.max_panos) FAIL(SALVE_BEV_E_CAPACITY, "pano slot out of range");
            jobs.push_back(j); slots[k] = j.pano_slot;
        }
        int rc = render_chunk(c, n, jobs, slots, c->out_store, c->counts, c->status, st);
        if (rc) return rc;
        CU(cudaMemcpyAsync(host_out + (size_t)i0 * c->img_bytes, c->out_store, (size_t)n * c->img_bytes, cudaMemcpyDeviceToHost, st));
        if (host_counts) CU(cudaMemcpyAsync(host_counts + (size_t)i0 * 8, c->counts, sizeof(int32_t) * 8 * n, cudaMemcpyDeviceToHost, st));
        if (host_status) CU(cudaMemcpyAsync(host_status + i0, c->status, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    return SALVE_BEV_OK;
}

extern "C" int salve_bev_set_bands(salve_bev_ctx* c, double a_lo, double a_hi, double b_lo, double b_hi) {
    if (!c) FAIL(SALVE_BEV_E_INVALID, "null argument");
    c->band[0] = a_lo; c->band[1] = a_hi; c->band[2] = b_lo; c->band[3] = b_hi;
    return SALVE_BEV_OK;
}

extern "C" int salve_bev_backproject(salve_bev_ctx* c, int32_t slot, double z_lo, double z_hi, int32_t frame, const float* host_R,
                                     const float* host_t, double* host_xyzrgb, int64_t* n_out, void* stream) {
    if (!c || !n_out) FAIL(SALVE_BEV_E_INVALID, "null argument");
    if (frame < 0 || frame > 2 || (frame == 2 && (!host_R || !host_t))) FAIL(SALVE_BEV_E_INVALID, "bad frame / missing pose");
    PoseArg pose = {{1.f, 0.f, 0.f, 1.f}, {0.f, 0.f}};
    if (frame == 2) { memcpy(pose.R, host_R, sizeof(float) * 4); memcpy(pose.t, host_t, sizeof(float) * 2); }
    if (slot < 0 || slot >= c->cfg.max_panos) FAIL(SALVE_BEV_E_CAPACITY, "pano slot out of range");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(c->cfg.device));
    SplatParams P = make_splat_params(c);
    const int rows = P.H - 2 * P.crop_rows;
    const int total = rows * P.W;
    const int nblk = (total + COMPACT_BLOCK - 1) / COMPACT_BLOCK;
    void *bc, *bo;
    int rc = tmp_get(c, 1, sizeof(int32_t) * nblk, &bc); if (rc) return rc;
    rc = tmp_get(c, 2, sizeof(long long) * (nblk + 1), &bo); if (rc) return rc;
    crop_count_kernel<<<nblk, COMPACT_BLOCK, 0, st>>>(P, c->h_depth_ptr[slot], z_lo, z_hi, (int32_t*)bc);
    exclusive_scan_kernel<<<1, 1024, 0, st>>>((const int32_t*)bc, (long long*)bo, nblk);
    c->launches += 2;
    CU(cudaGetLastError());
    long long total_kept = 0;
    CU(cudaMemcpyAsync(&total_kept, (long long*)bo + nblk, sizeof(long long), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    *n_out = total_kept;
    if (!host_xyzrgb || total_kept == 0) return SALVE_BEV_OK;
    void* ob;
    rc = tmp_get(c, 0, sizeof(double) * 6 * (size_t)total_kept, &ob); if (rc) return rc;
    crop_write_kernel<<<nblk, COMPACT_BLOCK, 0, st>>>(P, c->h_depth_ptr[slot], c->h_rgb_ptr[slot], z_lo, z_hi, (const long long*)bo, frame, pose, (double*)ob);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(host_xyzrgb, ob, sizeof(double) * 6 * (size_t)total_kept, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return SALVE_BEV_OK;
}

extern "C" int salve_bev_render_cloud_host(salve_bev_ctx* c, const double* host_xyzrgb, int64_t n, uint8_t* host_out, int32_t* host_counts,
                                           int32_t* host_status, void* stream) {
    if (!c || !host_out || (n > 0 && !host_xyzrgb)) FAIL(SALVE_BEV_E_INVALID, "null argument");
    if (n < 0 || n >= ((int64_t)1 << KEY_IDX_BITS)) FAIL(SALVE_BEV_E_CAPACITY, "cloud too large (n < 2^29)");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(c->cfg.device));
    c->events_used = 0;
    void *dc, *drgb;
    int rc = tmp_get(c, 0, sizeof(double) * 6 * (size_t)std::max<int64_t>(n, 1), &dc); if (rc) return rc;
    rc = tmp_get(c, 3, 3 * (size_t)std::max<int64_t>(n, 1), &drgb); if (rc) return rc;
    if (n) CU(cudaMemcpyAsync(dc, host_xyzrgb, sizeof(double) * 6 * (size_t)n, cudaMemcpyHostToDevice, st));
    const uint8_t* src = (const uint8_t*)drgb;
    CU(cudaMemcpyAsync(c->d_color_src, &src, sizeof(void*), cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaMemsetAsync(c->counts, 0, sizeof(int32_t) * 8, st));
    SplatParams P = make_splat_params(c);
    if (n) {
        splat_cloud_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(P, (const double*)dc, n, c->keygrid, (uint8_t*)drgb, c->counts);
        c->launches++;
        CU(cudaGetLastError());
    }
    c->last_chunk_images = 0;  // no stage taps after this entry point
    c->last_counts = c->counts;
    c->last_jobs.clear();
    rc = run_image_stage(c, 1, c->G, c->keygrid, c->d_color_src, c->out_store, c->counts, c->status, 0, 0, nullptr, nullptr, st, nullptr, nullptr, true);
    if (rc) return rc;
    CU(cudaMemcpyAsync(host_out, c->out_store, c->img_bytes, cudaMemcpyDeviceToHost, st));
    if (host_counts) CU(cudaMemcpyAsync(host_counts, c->counts, sizeof(int32_t) * 8, cudaMemcpyDeviceToHost, st));
    if (host_status) CU(cudaMemcpyAsync(host_status, c->status, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return SALVE_BEV_OK;
}

extern "C" int salve_bev_choose_elevated(salve_bev_ctx* c, const int64_t* hx, const int64_t* hy, const double* hz, int64_t n, double zmin,
                                         double zmax, int32_t num_slices, uint8_t* host_valid, void* stream) {
    if (!c || (n > 0 && (!hx || !hy || !hz || !host_valid))) FAIL(SALVE_BEV_E_INVALID, "null argument");
    if (num_slices < 1 || num_slices > 4096) FAIL(SALVE_BEV_E_INVALID, "bad num_slices");
    if (n == 0) return SALVE_BEV_OK;
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(c->cfg.device));
    int64_t xmax = 0, ymax = 0;
    for (int64_t i = 0; i < n; i++) {
        if (hx[i] < 0 || hy[i] < 0) FAIL(SALVE_BEV_E_INVALID, "negative pixel coordinate");
        xmax = std::max(xmax, hx[i]); ymax = std::max(ymax, hy[i]);
    }
    const int64_t w = xmax + 1, h = ymax + 1;
    if (w * h > ((int64_t)1 << 28)) FAIL(SALVE_BEV_E_CAPACITY, "z-order grid too large");
    // np.linspace(zmin, zmax, num_slices + 1): start + i*step with the last element forced to zmax
    std::vector<double> planes(num_slices + 1);
    const double step = (zmax - zmin) / num_slices;
    for (int i = 0; i <= num_slices; i++) planes[i] = zmin + i * step;
    planes[num_slices] = zmax;
    void *dx, *dy, *dz, *dp, *dg, *dv;
    int rc;
    if ((rc = tmp_get(c, 0, sizeof(int64_t) * n, &dx))) return rc;
    if ((rc = tmp_get(c, 1, sizeof(int64_t) * n, &dy))) return rc;
    if ((rc = tmp_get(c, 2, sizeof(double) * n, &dz))) return rc;
    if ((rc = tmp_get(c, 3, sizeof(double) * (num_slices + 1), &dp))) return rc;
    if ((rc = tmp_get(c, 4, sizeof(unsigned long long) * (size_t)(w * h), &dg))) return rc;
    if ((rc = tmp_get(c, 5, (size_t)n, &dv))) return rc;
    CU(cudaMemcpyAsync(dx, hx, sizeof(int64_t) * n, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(dy, hy, sizeof(int64_t) * n, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(dz, hz, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(dp, planes.data(), sizeof(double) * (num_slices + 1), cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(dg, 0, sizeof(unsigned long long) * (size_t)(w * h), st));
    const unsigned nb = (unsigned)((n + 255) / 256);
    zorder_mark_kernel<<<nb, 256, 0, st>>>((const long long*)dx, (const long long*)dy, (const double*)dz, n, (const double*)dp, num_slices, w,
                                           (unsigned long long*)dg);
    zorder_resolve_kernel<<<nb, 256, 0, st>>>((const long long*)dx, (const long long*)dy, (const double*)dz, n, (const double*)dp, num_slices,
                                              w, (const unsigned long long*)dg, (uint8_t*)dv);
    c->launches += 2;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(host_valid, dv, (size_t)n, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return SALVE_BEV_OK;
}

__global__ void base_raw_kernel(const uint32_t* __restrict__ color, int g, uint8_t* __restrict__ out) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= g) return;
    const uint32_t cw = color[p];
    out[p * 3 + 0] = (uint8_t)(cw & 0xFF); out[p * 3 + 1] = (uint8_t)((cw >> 8) & 0xFF); out[p * 3 + 2] = (uint8_t)((cw >> 16) & 0xFF);
}

extern "C" int salve_bev_interp_dense(salve_bev_ctx* c, const int64_t* host_xy, const double* host_values, int64_t n, int32_t grid_h,
                                      int32_t grid_w, uint8_t* host_img, uint8_t* host_hull, int32_t* status, void* stream) {
    if (!c || !host_img || !status || (n > 0 && (!host_xy || !host_values))) FAIL(SALVE_BEV_E_INVALID, "null argument");
    GridParams G;
    G.grid_h = grid_h; G.grid_w = grid_w; G.wpr = (grid_w + 31) / 32; G.g = grid_h * grid_w; G.K = c->G.K;
    if (grid_h < 1 || grid_w < 1 || grid_h > MAX_GRID_H || grid_w > 2047 || (size_t)G.g > c->g_stride ||
        (size_t)grid_h * G.wpr > c->bits_stride)
        FAIL(SALVE_BEV_E_CAPACITY, "grid exceeds the context's scratch; create a context with this grid size");
    if (n < 0 || n > ((int64_t)1 << 28)) FAIL(SALVE_BEV_E_CAPACITY, "too many points");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(c->cfg.device));
    c->events_used = 0;
    void *dxy, *dval, *drgb, *derr, *dhull = nullptr;
    int rc;
    if ((rc = tmp_get(c, 0, sizeof(int64_t) * 2 * (size_t)std::max<int64_t>(n, 1), &dxy))) return rc;
    if ((rc = tmp_get(c, 1, sizeof(double) * 3 * (size_t)std::max<int64_t>(n, 1), &dval))) return rc;
    if ((rc = tmp_get(c, 3, 3 * (size_t)std::max<int64_t>(n, 1), &drgb))) return rc;
    if ((rc = tmp_get(c, 6, 64, &derr))) return rc;
    if (host_hull) { if ((rc = tmp_get(c, 7, (size_t)G.g, &dhull))) return rc; CU(cudaMemsetAsync(dhull, 0, (size_t)G.g, st)); }
    if (n) {
        CU(cudaMemcpyAsync(dxy, host_xy, sizeof(int64_t) * 2 * (size_t)n, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(dval, host_values, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice, st));
    }
    const uint8_t* src = (const uint8_t*)drgb;
    CU(cudaMemcpyAsync(c->d_color_src, &src, sizeof(void*), cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaMemsetAsync(derr, 0, sizeof(int), st));
    CU(cudaMemsetAsync(c->counts, 0, sizeof(int32_t) * 8, st));
    if (n) {
        points_to_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const long long*)dxy, (const double*)dval, n, grid_h, grid_w,
                                                                          c->keygrid, (uint8_t*)drgb, (int*)derr);
        c->launches++;
        CU(cudaGetLastError());
    }
    c->last_chunk_images = 0;  // no stage taps after this entry point
    c->last_counts = c->counts;
    c->last_jobs.clear();
    if (image_smem_bytes(G.grid_h, G.wpr) <= (size_t)c->max_smem_optin) {
        rc = run_image_stage(c, 1, G, c->keygrid, c->d_color_src, c->out_store, c->counts, c->status, 1, 1, (uint8_t*)dhull, nullptr, st, nullptr, nullptr, true);
    } else {
        rc = run_mesh_stages(c, G, c->keygrid, c->d_color_src, c->out_store, c->counts, c->status, 1, 1, (uint8_t*)dhull, true, st);
        CU(cudaMemsetAsync(c->keygrid, 0, sizeof(uint32_t) * c->g_stride, st));  // the mesh kernels leave the keys in place
    }
    if (rc) return rc;
    int herr = 0, hstatus = 0;
    CU(cudaMemcpyAsync(&herr, derr, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&hstatus, c->status, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (herr) FAIL(SALVE_BEV_E_INVALID, "point outside the grid");
    *status = hstatus;
    if (hstatus == SALVE_BEV_IMG_DEGENERATE) return SALVE_BEV_OK;  // image left untouched, like the reference
    CU(cudaMemcpyAsync(host_img, c->out_store, (size_t)G.g * 3, cudaMemcpyDeviceToHost, st));
    if (host_hull) CU(cudaMemcpyAsync(host_hull, dhull, (size_t)G.g, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return SALVE_BEV_OK;
}

extern "C" int salve_bev_remove_hallucinated(salve_bev_ctx* c, const uint8_t* host_sparse, const uint8_t* host_interp, int32_t h, int32_t w,
                                             int32_t K, uint8_t* host_out, void* stream) {
    if (!c || !host_sparse || !host_interp || !host_out) FAIL(SALVE_BEV_E_INVALID, "null argument");
    if (h < 1 || w < 1 || K < 1 || !(K & 1)) FAIL(SALVE_BEV_E_INVALID, "bad shape or even K");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(c->cfg.device));
    const size_t n = (size_t)h * w;
    void *ds, *di, *dn, *dt, *dout;
    int rc;
    if ((rc = tmp_get(c, 0, n * 3, &ds))) return rc;
    if ((rc = tmp_get(c, 1, n * 3, &di))) return rc;
    if ((rc = tmp_get(c, 2, n, &dn))) return rc;
    if ((rc = tmp_get(c, 3, n, &dt))) return rc;
    if ((rc = tmp_get(c, 4, n * 3, &dout))) return rc;
    CU(cudaMemcpyAsync(ds, host_sparse, n * 3, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(di, host_interp, n * 3, cudaMemcpyHostToDevice, st));
    const unsigned nb = (unsigned)((n + 255) / 256);
    halluc_nonempty_kernel<<<nb, 256, 0, st>>>((const uint8_t*)ds, h, w, (uint8_t*)dn);
    halluc_rowdilate_kernel<<<nb, 256, 0, st>>>((const uint8_t*)dn, h, w, K / 2, (uint8_t*)dt);
    halluc_apply_kernel<<<nb, 256, 0, st>>>((const uint8_t*)dt, (const uint8_t*)di, h, w, K / 2, (uint8_t*)dout);
    c->launches += 3;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(host_out, dout, n * 3, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return SALVE_BEV_OK;
}

__global__ void fill_i32_kernel(int32_t* p, size_t n, int32_t v) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

extern "C" int salve_bev_tap(salve_bev_ctx* c, int32_t image, int32_t what, void* host_buf, int64_t host_buf_bytes, void* stream) {
    if (!c || !host_buf) FAIL(SALVE_BEV_E_INVALID, "null argument");
    if (image < 0 || image >= c->last_chunk_images) FAIL(SALVE_BEV_E_INVALID, "image index outside the last chunk");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(c->cfg.device));
    const GridParams& G = c->G;
    const size_t g = G.g, bits = (size_t)G.grid_h * G.wpr;
    uint32_t* kg = c->keygrid + (size_t)image * c->g_stride;
    const uint8_t* const* csrc = c->d_color_src + image;
    const bool img_ok = image_smem_bytes(G.grid_h, G.wpr) <= (size_t)c->max_smem_optin;
    // Taps recompute.  The render consumed (zeroed) the image's key grid, so the pano pass that produced it is splatted again
    // from the job table of the last chunk (both surfaces of that pass), and the slots are cleared again afterwards.
    int job = -1;
    for (size_t k = 0; k < c->last_jobs.size(); k++)
        if (c->last_jobs[k].img_floor == image || c->last_jobs[k].img_ceil == image) { job = (int)k; break; }
    if (job < 0) FAIL(SALVE_BEV_E_INVALID, "stage taps follow a render of pano images (render_hypotheses / render_images)");
    const SplatJob& J = c->last_jobs[job];
    auto clear_slots = [&]() -> int {
        for (int im : {J.img_floor, J.img_ceil})
            if (im >= 0) CU(cudaMemsetAsync(c->keygrid + (size_t)im * c->g_stride, 0, sizeof(uint32_t) * c->g_stride, st));
        return SALVE_BEV_OK;
    };
    int rc = clear_slots(); if (rc) return rc;
    {
        SplatParams P = make_splat_params(c);
        const int rows = P.H - 2 * P.crop_rows;
        P.rows_per_thread = splat_rows_for(1);
        dim3 grid((unsigned)(((rows + P.rows_per_thread - 1) / P.rows_per_thread) * ((P.W + 1023) >> 10)), 1u);
        splat_pano_kernel<<<grid, 256, 0, st>>>(P, c->d_jobs + job, c->keygrid, c->g_stride, nullptr);
        c->launches++;
        CU(cudaGetLastError());
    }
    // ... on a private copy of the image's counters
    void* dcnt;
    if ((rc = tmp_get(c, 6, 64, &dcnt))) return rc;
    if (c->last_counts) CU(cudaMemcpyAsync(dcnt, c->last_counts + (size_t)image * 8, 32, cudaMemcpyDeviceToDevice, st));
    auto body = [&]() -> int {
    size_t bytes = 0;
    switch (what) {
        case SALVE_BEV_TAP_KEYGRID: {
            bytes = g * 4;
            if ((int64_t)bytes > host_buf_bytes) FAIL(SALVE_BEV_E_CAPACITY, "tap buffer too small");
            CU(cudaMemcpyAsync(host_buf, kg, bytes, cudaMemcpyDeviceToHost, st));
            break;
        }
        case SALVE_BEV_TAP_COLOR: {
            bytes = g * 4;
            if ((int64_t)bytes > host_buf_bytes) FAIL(SALVE_BEV_E_CAPACITY, "tap buffer too small");
            void* d; if ((rc = tmp_get(c, 0, bytes, &d))) return rc;
            const uint8_t* src1 = nullptr;
            CU(cudaMemcpyAsync(&src1, csrc, sizeof(void*), cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            tap_color_kernel<<<(unsigned)((g + 255) / 256), 256, 0, st>>>(kg, src1, (int)g, c->cfg.pano_w, (uint32_t*)d);
            c->launches++;
            CU(cudaGetLastError());
            CU(cudaMemcpyAsync(host_buf, d, bytes, cudaMemcpyDeviceToHost, st));
            break;
        }
        case SALVE_BEV_TAP_OCC:
        case SALVE_BEV_TAP_NONEMPTY:
        case SALVE_BEV_TAP_KEEP:
        case SALVE_BEV_TAP_QTRI: {
            if (!img_ok) FAIL(SALVE_BEV_E_CAPACITY, "tap not available for this grid size");
            bytes = (what == SALVE_BEV_TAP_QTRI) ? g * 3 * 4 : bits * 4;
            if ((int64_t)bytes > host_buf_bytes) FAIL(SALVE_BEV_E_CAPACITY, "tap buffer too small");
            void *dimg, *dq = nullptr;
            if ((rc = tmp_get(c, 0, g * 3, &dimg))) return rc;
            if (what == SALVE_BEV_TAP_QTRI) {
                if ((rc = tmp_get(c, 2, g * 3 * 4, &dq))) return rc;
                fill_i32_kernel<<<(unsigned)((g * 3 + 255) / 256), 256, 0, st>>>((int32_t*)dq, g * 3, -1);
                c->launches++;
            }
            rc = run_image_stage(c, 1, G, kg, csrc, (uint8_t*)dimg, (int32_t*)dcnt, nullptr, 0, 0, nullptr, (int32_t*)dq, st);
            if (rc) return rc;
            // the bit planes are the state the stages keep per image: image 0 of the re-run
            const void* src = what == SALVE_BEV_TAP_QTRI ? dq
                              : (const void*)(c->planes + (what == SALVE_BEV_TAP_OCC ? 0 : what == SALVE_BEV_TAP_NONEMPTY ? 1 : 2) * c->bits_stride);
            CU(cudaMemcpyAsync(host_buf, src, bytes, cudaMemcpyDeviceToHost, st));
            break;
        }
        case SALVE_BEV_TAP_TRIS: {
            // explicit mesh of this image by the mesh path (zipper + parallel Lawson flips)
            rc = run_mesh_stages(c, G, kg, csrc, nullptr, (int32_t*)dcnt, nullptr, 0, 0, nullptr, false, st);
            if (rc) return rc;
            ImgHeader hd;
            CU(cudaMemcpyAsync(&hd, c->headers, sizeof(hd), cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            bytes = (size_t)hd.n_tris * 3 * 4;
            if ((int64_t)bytes > host_buf_bytes) FAIL(SALVE_BEV_E_CAPACITY, "tap buffer too small");
            if (hd.n_tris == 0) return SALVE_BEV_OK;  // (leaves the lambda: the slots are cleared below)
            void* d; if ((rc = tmp_get(c, 0, bytes, &d))) return rc;
            tap_tris_kernel<<<(hd.n_tris + 255) / 256, 256, 0, st>>>(c->tris, hd.n_tris, G.grid_w, (int32_t*)d);
            c->launches++;
            CU(cudaGetLastError());
            CU(cudaMemcpyAsync(host_buf, d, bytes, cudaMemcpyDeviceToHost, st));
            break;
        }
        case SALVE_BEV_TAP_INTERP:
        case SALVE_BEV_TAP_HULL: {
            bytes = (what == SALVE_BEV_TAP_INTERP) ? g * 3 : g;
            if ((int64_t)bytes > host_buf_bytes) FAIL(SALVE_BEV_E_CAPACITY, "tap buffer too small");
            void *dimg, *dhull;
            if ((rc = tmp_get(c, 0, g * 3, &dimg))) return rc;
            if ((rc = tmp_get(c, 7, g, &dhull))) return rc;
            CU(cudaMemsetAsync(dhull, 0, g, st));
            if (img_ok) rc = run_image_stage(c, 1, G, kg, csrc, (uint8_t*)dimg, (int32_t*)dcnt, nullptr, 1, 0, (uint8_t*)dhull, nullptr, st);
            else rc = run_mesh_stages(c, G, kg, csrc, (uint8_t*)dimg, (int32_t*)dcnt, nullptr, 1, 0, (uint8_t*)dhull, true, st);
            if (rc) return rc;
            CU(cudaMemcpyAsync(host_buf, what == SALVE_BEV_TAP_INTERP ? dimg : dhull, bytes, cudaMemcpyDeviceToHost, st));
            break;
        }
        default: FAIL(SALVE_BEV_E_INVALID, "unknown tap");
    }
    return SALVE_BEV_OK;
    };
    const int rc_body = body();
    rc = clear_slots();  // also after a failed tap: the key grid stays all zero between calls
    if (rc_body) return rc_body;
    if (rc) return rc;
    CU(cudaStreamSynchronize(st));
    return SALVE_BEV_OK;
}

extern "C" int salve_bev_enable_timing(salve_bev_ctx* c, int32_t on) {
    if (!c) FAIL(SALVE_BEV_E_INVALID, "null argument");
    c->timing = on != 0; c->events_used = 0;
    return SALVE_BEV_OK;
}

extern "C" int salve_bev_last_timings(salve_bev_ctx* c, float* host_ms) {
    if (!c || !host_ms) FAIL(SALVE_BEV_E_INVALID, "null argument");
    for (int i = 0; i < SALVE_BEV_NTIMINGS; i++) host_ms[i] = 0.f;
    if (!c->timing || c->events_used == 0) return SALVE_BEV_OK;
    CU(cudaSetDevice(c->cfg.device));
    CU(cudaEventSynchronize(c->events[c->events_used - 1]));
    // intervals between the events of a chunk: splat, sites, prep, local, window, shade, finish -> slots 0, 1, 2, 6, 3, 4, 5
    static const int slot[N_STAGE_EVENTS - 1] = {0, 1, 2, 6, 3, 4, 5};
    for (size_t k = 0; k + N_STAGE_EVENTS <= c->events_used; k += N_STAGE_EVENTS) {
        float ms = 0.f;
        for (int s = 0; s + 1 < N_STAGE_EVENTS; s++) {
            CU(cudaEventElapsedTime(&ms, c->events[k + s], c->events[k + s + 1]));
            host_ms[slot[s]] += ms;
        }
        CU(cudaEventElapsedTime(&ms, c->events[k], c->events[k + N_STAGE_EVENTS - 1]));
        host_ms[SALVE_BEV_NTIMINGS - 1] += ms;
    }
    return SALVE_BEV_OK;
}

extern "C" int64_t salve_bev_local_rule_tables(uint32_t* host_words, int64_t n_words) {
    const std::vector<uint32_t>& t = local_rule_tables();
    if (host_words)
        for (int64_t i = 0; i < n_words && i < (int64_t)t.size(); i++) host_words[i] = t[(size_t)i];
    return (int64_t)t.size();
}

extern "C" int64_t salve_bev_launch_count(salve_bev_ctx* c) { return c ? c->launches : 0; }
