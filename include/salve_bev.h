/*
 * salve_bev.h -- C ABI of the B200-native BEV texture-map renderer (libsalve_bev.so).
 *
 * Drop-in boundary for ONE hot path of zillow/salve: equirectangular pano + depth ->
 * Sim(2)-posed, height-cropped points -> BEV splat (z-order rule) -> linear densification ->
 * hallucination mask.  The reference has no FFI for this path (it is pure Python); each entry
 * point below names the reference function(s) it replaces.  Paths are relative to the
 * reference repository root.
 *
 * Conventions
 *   - plain C types only; every function returns 0 on success or a negative SALVE_BEV_E_* code;
 *     no exception crosses the boundary.  salve_bev_last_error() returns a static message.
 *   - pointers named host_* are host memory, dev_* are device memory (cudaMalloc'ed by the caller,
 *     e.g. a torch tensor's data_ptr()).  `stream` is a cudaStream_t passed as void* (NULL = the
 *     legacy default stream).  Calls taking dev_* outputs are asynchronous on `stream`;
 *     calls taking host_* outputs synchronise `stream` before returning.
 *   - the context owns all scratch memory; it is not thread-safe (one context per thread/stream).
 *   - there is NO CPU fallback: without a CUDA device every entry point fails.
 */
#ifndef SALVE_BEV_H
#define SALVE_BEV_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SALVE_BEV_OK 0
#define SALVE_BEV_E_INVALID (-1)  /* bad argument */
#define SALVE_BEV_E_CUDA (-2)     /* CUDA runtime error (see salve_bev_last_error) */
#define SALVE_BEV_E_CAPACITY (-3) /* request exceeds what the context was created for */
#define SALVE_BEV_E_NODEVICE (-4) /* no CUDA device */

/* per-image status written by the render calls */
#define SALVE_BEV_IMG_OK 0
#define SALVE_BEV_IMG_EMPTY 1      /* no point inside the BEV box: reference returns None (bev_rendering_utils.py:279-280) */
#define SALVE_BEV_IMG_DEGENERATE 2 /* <4 sites, all in one row or one column: zero image (interpolation_utils.py:37-42) */
#define SALVE_BEV_IMG_COLLINEAR 3  /* sites on one oblique line: the reference's Qhull call raises */

/* per-image counters, SALVE_BEV_NCOUNTS int32 each */
#define SALVE_BEV_NCOUNTS 8
#define SALVE_BEV_CNT_CROP 0     /* points inside the height band      (bev_rendering_utils.py:408-413) */
#define SALVE_BEV_CNT_BBOX 1     /* ... and inside the BEV box         (bev_rendering_utils.py:38-45)   */
#define SALVE_BEV_CNT_SITES 2    /* distinct pixels after the z-order rule (zorder_utils.py:49-83)      */
#define SALVE_BEV_CNT_NONEMPTY 3 /* sites counted non-empty by the uint8 product (interpolation_utils.py:95) */
#define SALVE_BEV_CNT_KEEP 4     /* pixels of the hallucination keep-mask (interpolation_utils.py:101-115) */
#define SALVE_BEV_CNT_FILLED 5   /* non-site pixels interpolated (inside hull and keep-mask)          */
#define SALVE_BEV_CNT_MAXFLIPS 6 /* longest flip descent of one query pixel (diagnostic)            */
#define SALVE_BEV_CNT_FLIPS 7    /* Lawson flips in total (diagnostic; like MAXFLIPS it depends on how often two warps
                                    reach the same large triangle, i.e. on timing -- images and counters 0..5 do not) */

#define SALVE_BEV_SURF_FLOOR 1   /* z band (-inf, -1.0]   (bev_rendering_utils.py:560-562) */
#define SALVE_BEV_SURF_CEILING 2 /* z band (0.5, +inf)    (bev_rendering_utils.py:564-566) */

typedef struct salve_bev_ctx salve_bev_ctx;

typedef struct salve_bev_config {
    int32_t device;       /* CUDA device ordinal */
    int32_t pano_h;       /* equirect height; reference: 512 (bev_rendering_utils.py:373-375) */
    int32_t pano_w;       /* equirect width;  reference: 1024 */
    int32_t max_panos;    /* resident pano slots */
    int32_t max_images;   /* BEV images rendered per internal chunk (scratch is sized for this) */
    int32_t grid_h;       /* BEV image rows    = BEVParams.img_h + 1 (bev_rendering_utils.py:292): 501 */
    int32_t grid_w;       /* BEV image columns = BEVParams.img_w + 1: 501 */
    int32_t kernel_sz;    /* hallucination box size K (interpolation_utils.py:15): 11 */
    double xmin, ymin;    /* BEV box lower corner, metres (bevparams.py:57-61): -5, -5 */
    double xmax, ymax;    /* upper corner: 5, 5 */
    double px_per_m;      /* Sim(2) scale 1/meters_per_px (bevparams.py:78): 50.0 */
    float depth_scale;    /* float32 multiplier of the uint16 depth (bev_rendering_utils.py:367,611): 0.001f */
    int32_t crop_rows;    /* int(pano_h * crop_ratio) rows dropped top and bottom (bev_rendering_utils.py:397-401): 80 */
} salve_bev_config;

/* Fill `cfg` with the reference defaults for a pano of (pano_h, pano_w). */
void salve_bev_default_config(salve_bev_config* cfg, int32_t pano_h, int32_t pano_w);

const char* salve_bev_last_error(void);

int salve_bev_ctx_create(const salve_bev_config* cfg, salve_bev_ctx** out);
void salve_bev_ctx_destroy(salve_bev_ctx* ctx);

/*
 * Unit-sphere factor tables: x = cos_phi[v]*cos_theta[u], y = cos_phi[v]*sin_theta[u],
 * z = neg_sin_phi[v].  Replaces salve/utils/hohonet_pano_utils.py:10-44 (get_uni_sphere_xyz).
 * The context computes them itself with libm at creation; a caller that must match numpy's
 * trig bit-for-bit (the Python layer does) passes numpy's values here.  Host pointers,
 * pano_h / pano_w doubles.
 */
int salve_bev_set_sphere_tables(salve_bev_ctx* ctx, const double* host_cos_phi, const double* host_neg_sin_phi,
                                const double* host_cos_theta, const double* host_sin_theta);
/* Write the (pano_h, pano_w, 3) float64 unit-sphere table to host memory (get_uni_sphere_xyz). */
int salve_bev_get_uni_sphere_xyz(salve_bev_ctx* ctx, double* host_out);

/*
 * Pano slots.  rgb: (pano_h, pano_w, 3) uint8; depth: (pano_h, pano_w) uint16 millimetres --
 * what imageio.imread returns at bev_rendering_utils.py:367,370.
 */
int salve_bev_upload_pano(salve_bev_ctx* ctx, int32_t slot, const uint8_t* host_rgb, const uint16_t* host_depth, void* stream);
/* Alias device memory owned by the caller instead of copying (must stay valid while used).  rgb 2-byte, depth 8-byte aligned. */
int salve_bev_bind_pano(salve_bev_ctx* ctx, int32_t slot, const uint8_t* dev_rgb, const uint16_t* dev_depth);
/*
 * Full-resolution colour.  ZInD panos are 2048x1024; get_xyzrgb_from_depth brings them to the depth map's
 * 1024x512 with cv2.resize(rgb, (1024, 512), INTER_LINEAR) (bev_rendering_utils.py:373-375), which at a scale of exactly 2 is the
 * rounded 2x2 box mean (sum + 2) >> 2.  rgb_2x: (2*pano_h, 2*pano_w, 3) uint8; depth: (pano_h, pano_w) uint16 as before.  The
 * down-sampled pano is never materialised: the 2x2 mean is taken when a winner's colour is gathered (50 k gathers per image
 * instead of a 6 MB pass per pano).
 */
int salve_bev_upload_pano_fullres(salve_bev_ctx* ctx, int32_t slot, const uint8_t* host_rgb_2x, const uint16_t* host_depth, void* stream);
int salve_bev_bind_pano_fullres(salve_bev_ctx* ctx, int32_t slot, const uint8_t* dev_rgb_2x, const uint16_t* dev_depth);

/*
 * Height bands of the two surfaces, each keeps lo < z <= hi.  Defaults are the reference's
 * floor (-inf, -1.0] and ceiling (0.5, +inf) (bev_rendering_utils.py:560-566); a caller that
 * renders with another args.crop_z_range (render_bev_pair takes it from `args`) sets band A.
 */
int salve_bev_set_bands(salve_bev_ctx* ctx, double a_lo, double a_hi, double b_lo, double b_hi);

/*
 * Render alignment hypotheses.  Replaces render_bev_pair (bev_rendering_utils.py:417-480), once per
 * requested surface, for n_hyp hypotheses in one call.
 *   pano1/pano2 : slot of pano i1 / i2 per hypothesis
 *   R           : n_hyp x 4 float32, i2Ti1 rotation row-major (salve/common/sim2.py:50)
 *   t           : n_hyp x 2 float32, i2Ti1 translation (un-scaled; the x1.5 of
 *                 bev_rendering_utils.py:448-451 is applied inside)
 *   surfaces    : SALVE_BEV_SURF_FLOOR | SALVE_BEV_SURF_CEILING
 * Output image order per hypothesis: for each requested surface (floor first): img1 (pano 1 posed
 * into pano 2's frame), img2 (pano 2).  So n_img = n_hyp * nsurf * 2 images of grid_h*grid_w*3 uint8,
 * rows already flipped (np.flipud, bev_rendering_utils.py:319).
 *   dev_out     : n_img * grid_h * grid_w * 3 bytes
 *   dev_counts  : n_img * SALVE_BEV_NCOUNTS int32 (may be NULL)
 *   dev_status  : n_img int32 SALVE_BEV_IMG_* (may be NULL)
 * Any n_hyp is accepted; the work is cut into chunks of max_images images.
 */
int salve_bev_render_hypotheses(salve_bev_ctx* ctx, int32_t n_hyp, const int32_t* host_pano1, const int32_t* host_pano2,
                                const float* host_R, const float* host_t, uint32_t surfaces, uint8_t* dev_out,
                                int32_t* dev_counts, int32_t* dev_status, void* stream);
/* Same, with host output buffers (device->host copies included; synchronises).  Chunks are double buffered:
 * chunk k+1 renders while chunk k is copied out on a private copy stream; pass page-locked host_out to get the overlap. */
int salve_bev_render_hypotheses_host(salve_bev_ctx* ctx, int32_t n_hyp, const int32_t* host_pano1, const int32_t* host_pano2,
                                     const float* host_R, const float* host_t, uint32_t surfaces, uint8_t* host_out,
                                     int32_t* host_counts, int32_t* host_status, void* stream);

/*
 * De-duplicated layout.  img2 of render_bev_pair does not depend on the hypothesis: only pano 1's cloud is posed
 * (bev_rendering_utils.py:451), pano 2 is rendered in its own frame (:455).  These entry points render every distinct
 * (pano 2, surface) once and return
 *   posed          : n_hyp * nsurf images, hypothesis-major, surface-minor (floor first): img1 of each pair
 *   unposed        : n_unique * nsurf images (n_unique <= min(n_hyp, max_panos)): img2 of the pairs, one per distinct pano 2 in
 *                    order of first appearance; the caller provides room for min(n_hyp, max_panos) * nsurf images
 *   unposed_of_hyp : n_hyp int32 (host), index u of hypothesis h's pano 2: its img2 for surface s is unposed[u * nsurf + s]
 *   n_unique       : 1 int32 (host)
 * counts / status arrays (may be NULL) follow the two image arrays.  salve_bev_render_hypotheses[_host] use the same
 * machinery whenever a pano 2 repeats and then copy the shared image into every hypothesis' slot, so their output is
 * unchanged; salve_bev_set_dedup_unposed(ctx, 0) turns that off (every image rendered from scratch).
 */
int salve_bev_render_hypotheses_compact(salve_bev_ctx* ctx, int32_t n_hyp, const int32_t* host_pano1, const int32_t* host_pano2,
                                        const float* host_R, const float* host_t, uint32_t surfaces, uint8_t* dev_posed,
                                        uint8_t* dev_unposed, int32_t* host_unposed_of_hyp, int32_t* host_n_unique,
                                        int32_t* dev_counts_posed, int32_t* dev_counts_unposed, int32_t* dev_status_posed,
                                        int32_t* dev_status_unposed, void* stream);
int salve_bev_render_hypotheses_compact_host(salve_bev_ctx* ctx, int32_t n_hyp, const int32_t* host_pano1, const int32_t* host_pano2,
                                             const float* host_R, const float* host_t, uint32_t surfaces, uint8_t* host_posed,
                                             uint8_t* host_unposed, int32_t* host_unposed_of_hyp, int32_t* host_n_unique,
                                             int32_t* host_counts_posed, int32_t* host_counts_unposed, int32_t* host_status_posed,
                                             int32_t* host_status_unposed, void* stream);
int salve_bev_set_dedup_unposed(salve_bev_ctx* ctx, int32_t on);

/*
 * Verifier pre-processing, fused: replaces, for a batch of hypotheses kept on the device, the JPEG round trip +
 * ZindData + the val/test transform of the reference: ResizeQuadruplet (cv2.resize INTER_LINEAR, salve/utils/transform.py:256-272) ->
 * CropQuadruplet centre (transform.py:386-420) -> ToTensorQuadruplet (transform.py:105-123) -> NormalizeQuadruplet with the ImageNet
 * mean/std * 255 (transform.py:177-202, normalization_utils.py:13-26) -> torch.cat along channels (models/early_fusion.py:60-61).
 *   host_src : n * 4 DEVICE pointers (host array) to grid_h x grid_w x 3 uint8 renders, per hypothesis in the order the model takes them:
 *              x1c, x2c, x1f, x2f = ceiling img1, ceiling img2, floor img1, floor img2 (dataset/zind_data.py:306-315)
 *   dev_out  : n x 12 x crop_hw x crop_hw float32, bit-identical to the reference chain (cv2's 8-bit bilinear is fixed point)
 * resize_hw <= grid size (down-scaling, the released configs use 234), crop_hw <= min(resize_hw, 256) (224).  Asynchronous on `stream`.
 */
int salve_bev_verifier_preprocess(salve_bev_ctx* ctx, int32_t n, const uint8_t* const* host_src, int32_t resize_hw, int32_t crop_hw,
                                  float* dev_out, void* stream);

/*
 * Render individual images: image k = pano slot[k], surface[k] (SALVE_BEV_SURF_*), posed[k] != 0 ->
 * apply (R[k], t[k]) as for pano 1 of a pair.  Replaces get_xyzrgb_from_depth + the frame change of
 * render_bev_pair + render_bev_image (bev_rendering_utils.py:347-414, 443-451, 254-328).
 */
int salve_bev_render_images_host(salve_bev_ctx* ctx, int32_t n_img, const int32_t* host_slot, const int32_t* host_surface,
                                 const int32_t* host_posed, const float* host_R, const float* host_t, uint8_t* host_out,
                                 int32_t* host_counts, int32_t* host_status, void* stream);

/*
 * Height-crop + stream compaction: replaces get_xyzrgb_from_depth (bev_rendering_utils.py:347-414).
 * Keeps z_lo < z <= z_hi after dropping crop_rows rows top and bottom.  First call with
 * host_xyzrgb == NULL to get *n_out, then with a buffer of n_out*6 doubles (x,y,z,r/255,g/255,b/255),
 * points in pano raster order.
 * frame: 0 = HoHoNet frame, as get_xyzrgb_from_depth returns;
 *        1 = ZInD frame, xy @ rotmat2d(-90).T          (bev_rendering_utils.py:443-446);
 *        2 = additionally xy @ R.T + t*1.5 (pano 1 of a pair; host_R 4 floats, host_t 2 floats)
 *            -- 1 and 2 together replace get_bev_pair_xyzrgb (bev_rendering_utils.py:483-522).
 */
int salve_bev_backproject(salve_bev_ctx* ctx, int32_t slot, double z_lo, double z_hi, int32_t frame, const float* host_R,
                          const float* host_t, double* host_xyzrgb, int64_t* n_out, void* stream);

/*
 * Render an arbitrary coloured cloud: replaces render_bev_image (bev_rendering_utils.py:254-328).
 * host_xyzrgb: n x 6 float64 (rgb in [0,1]).  n < 2^29.  One image to host_out.
 */
int salve_bev_render_cloud_host(salve_bev_ctx* ctx, const double* host_xyzrgb, int64_t n, uint8_t* host_out,
                                int32_t* host_counts, int32_t* host_status, void* stream);

/*
 * Z-order rule on explicit arrays: replaces choose_elevated_repeated_vals (zorder_utils.py:10-83).
 * x, y: pixel coordinates (>= 0); valid[i] = 1 iff point i wins its pixel.
 */
int salve_bev_choose_elevated(salve_bev_ctx* ctx, const int64_t* host_x, const int64_t* host_y, const double* host_z, int64_t n,
                              double zmin, double zmax, int32_t num_slices, uint8_t* host_valid, void* stream);

/*
 * Sparse -> dense: replaces interp_dense_grid_from_sparse (interpolation_utils.py:21-54) for
 * method="linear".  points: n x 2 int64 (x = column, y = row), distinct; values: n x 3 float64
 * (truncated to uint8 like interpolation_utils.py:53).  grid_h*grid_w <= 800000, grid_w <= 2047, grid_h <= 1023; grids whose
 * bit rows fit shared memory (e.g. 501x501) use the staged image pipeline (k_image.cuh), larger ones the explicit-mesh path.  Limits: `values` are
 * taken as the uint8 colours the render path produces (integral, 0..255: the reference interpolates float64 and truncates afterwards, which is the same
 * for such values) and every point must lie inside the grid (the Python mirror raises NotImplementedError otherwise).  host_img: grid_h x grid_w x 3 uint8, fully overwritten unless *status ==
 * SALVE_BEV_IMG_DEGENERATE (then untouched, as in the reference).  host_hull (may be NULL):
 * grid_h x grid_w uint8, 1 inside the closed convex hull.
 */
int salve_bev_interp_dense(salve_bev_ctx* ctx, const int64_t* host_points_xy, const double* host_values, int64_t n,
                           int32_t grid_h, int32_t grid_w, uint8_t* host_img, uint8_t* host_hull, int32_t* status, void* stream);

/*
 * Hallucination mask: replaces remove_hallucinated_content (interpolation_utils.py:74-122).
 * sparse, interp, out: h x w x 3 uint8.  K odd.
 */
int salve_bev_remove_hallucinated(salve_bev_ctx* ctx, const uint8_t* host_sparse, const uint8_t* host_interp, int32_t h, int32_t w,
                                  int32_t K, uint8_t* host_out, void* stream);

/*
 * Layout modality: room polygon + window / door / opening strokes rasterised into grid_h x grid_w x 3 images.
 * Replaces the cv2 calls behind rasterize_single_layout / rasterize_room_layout_pair (bev_rendering_utils.py:48-251): cv2.fillPoly of
 * the room polygon (bit-identical for polygons inside the image), cv2.line(LINE_AA, thickness) strokes (identical in the stroke's
 * interior, toleranced on its anti-aliased rim), np.flipud.
 *   host_desc    : int32 words, one descriptor per image: [n_poly, n_strokes, poly_rgb (r | g<<8 | b<<16), flip,
 *                  n_poly x (x, y) pixel vertices (n_poly <= 128), n_strokes x (x0, y0, x1, y1, rgb, thickness)]
 *   host_offsets : n_img + 1 word offsets of the descriptors in host_desc
 *   host_init    : n_img initial images to draw onto (may be NULL: black)
 *   host_out     : n_img x grid_h x grid_w x 3 uint8
 */
int salve_bev_rasterize_layouts_host(salve_bev_ctx* ctx, int32_t n_img, const int32_t* host_desc, const int64_t* host_offsets,
                                     const uint8_t* host_init, uint8_t* host_out, void* stream);

/*
 * Stage taps of the most recent render of pano images (render_hypotheses* / render_images; parity tests).  image = index within the
 * last internal chunk (for a call that was de-duplicated the chunks hold the unique un-posed images first, then the posed ones).
 * The render consumes (zeroes) its key grids, so a tap re-splats the pano pass that produced the image from the job table of the
 * last chunk and re-runs the stages on it; the panos of that chunk must still be in their slots.
 * what: see SALVE_BEV_TAP_*.  host_buf must be large enough (sizes in comments, g = grid_h*grid_w,
 * wpr = (grid_w+31)/32).
 */
#define SALVE_BEV_TAP_KEYGRID 0  /* g uint32: (slice<<29 | source index) + 1 of the winner, 0 = empty */
#define SALVE_BEV_TAP_COLOR 1    /* g uint32: r | g<<8 | b<<16 | 0xFF<<24 at sites, 0 elsewhere */
#define SALVE_BEV_TAP_OCC 2      /* grid_h*wpr uint32 bit rows: site present */
#define SALVE_BEV_TAP_NONEMPTY 3 /* same layout: uint8 product != 0 */
#define SALVE_BEV_TAP_KEEP 4     /* same layout: hallucination keep-mask */
#define SALVE_BEV_TAP_TRIS 5     /* (2*sites-2) x 3 int32 vertex pixel ids row*grid_w+col, -1 = ghost (hull): the full Delaunay
                                    mesh of the image, built on demand by the explicit-mesh path (zipper + parallel Lawson flips) */
#define SALVE_BEV_TAP_INTERP 6   /* g*3 uint8: interpolated image before mask and flip */
#define SALVE_BEV_TAP_HULL 7     /* g uint8: inside closed convex hull */
#define SALVE_BEV_TAP_QTRI 8     /* g*3 int32: per interpolated pixel the vertex pixel ids (ascending) of the Delaunay triangle its
                                    flip descent ended in; -1 elsewhere */
int salve_bev_tap(salve_bev_ctx* ctx, int32_t image, int32_t what, void* host_buf, int64_t host_buf_bytes, void* stream);

/* Per-stage device time (ms, CUDA events) of the most recent pano render call, summed over its chunks.  host_ms: SALVE_BEV_NTIMINGS floats:
 * [0] splat_pano_kernel  [1] sites stage  [2] prep stage  [3] window stage  [4] shade stage  [5] order + finish stage  [6] local stage
 * [7] total (chunk start to the end of the finish stage). */
#define SALVE_BEV_NTIMINGS 8
int salve_bev_last_timings(salve_bev_ctx* ctx, float* host_ms);
/* Enable/disable per-stage event timing (off by default: events add sync points at read time only). */
int salve_bev_enable_timing(salve_bev_ctx* ctx, int32_t on);

/* The host-built tables of the local rule (k_image.cuh, LocalRule; the prep / local stages' short cut of the Delaunay search of
 * reference interpolation_utils.py:21-54 for queries whose triangle has its vertices among 12 near neighbours): copies up to
 * n_words uint32 words into host_words and returns the table's size in words.  No GPU needed; tests/ check it against an independent
 * enumeration. */
int64_t salve_bev_local_rule_tables(uint32_t* host_words, int64_t n_words);

/* Number of kernel launches issued by this context since creation (bench.py's gpu_launches). */
int64_t salve_bev_launch_count(salve_bev_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* SALVE_BEV_H */
