/*
 * b200geo.h -- C ABI of libb200geo.so: the B200 (sm_100a) implementation of the geometric hot path
 * of choice17/ransac_with_homography (RANSAC homography estimation + backward perspective warp).
 *
 * The reference is pure Python/numpy and has no FFI; these entry points are what a binding for the
 * hot path replaces.  Each declaration cites the reference interface it stands in for
 * (file:line in the reference repository).  The binding a maintainer would add is a ctypes stub:
 * see INTEGRATION.md and ransac_with_homography_b200/_cabi.py.
 *
 * Conventions
 *   - every function returns 0 (B200GEO_OK) or a negative error code; b200geo_last_error() gives the
 *     text of the last failure on the calling thread.  No exceptions cross the boundary.
 *   - "_host" entry points take HOST pointers, perform the host<->device copies themselves and
 *     return when the results are in the caller's buffers (synchronous).
 *   - "_dev" entry points take DEVICE pointers on the context's GPU and are ordered on `stream`
 *     (a cudaStream_t passed as void*; NULL = the legacy default stream).  They do not synchronise.
 *   - a correspondence set is M rows of 4 float32: (x, y, x', y') -- the columns of the reference's
 *     X (2,M) and Y (2,M) arrays (ransac.py:165) interleaved; 16 bytes per match, 16-byte aligned.
 *   - homographies are 9 row-major values.  A sample-index set is K rows of 4 int32.
 *   - a context belongs to one GPU; calls on one context must not overlap (thread-compatible).
 *   - there is no CPU fallback: without a CUDA device b200geo_create fails.
 */
#ifndef B200GEO_H
#define B200GEO_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200GEO_ABI_VERSION 1

/* error codes */
#define B200GEO_OK 0
#define B200GEO_ERR_CUDA (-1)      /* a CUDA runtime call failed (text in last_error) */
#define B200GEO_ERR_ARG (-2)       /* invalid argument */
#define B200GEO_ERR_NODEVICE (-3)  /* no usable sm_100 device */
#define B200GEO_ERR_SINGULAR (-4)  /* degenerate system: the reference raises LinAlgError here */
#define B200GEO_ERR_NOMODEL (-5)   /* every hypothesis scored 0 inliers (reference: np.where(None) fails, ransac.py:206) */

/* HomoModel.computeLoss method (ransac.py:84-98) */
#define B200GEO_FWD 0
#define B200GEO_BACKWARD 1
#define B200GEO_REPROJ 2

/* samplers (homography.py:140  convertfunc) */
#define B200GEO_NN 0
#define B200GEO_BILINEAR 1

/* pixel types */
#define B200GEO_U8 0
#define B200GEO_F32 1
#define B200GEO_F64 2

typedef struct b200geo_ctx b200geo_ctx;

int b200geo_abi_version(void);
const char *b200geo_last_error(void);

/* One context per (process, GPU).  Owns the scratch buffers the entry points need. */
int b200geo_create(int device, b200geo_ctx **out);
int b200geo_destroy(b200geo_ctx *ctx);
/* sm_count, clock (kHz), total memory; any pointer may be NULL */
int b200geo_device_info(b200geo_ctx *ctx, int *sm_count, int *cc, int *clock_khz, size_t *mem_bytes);
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
int64_t b200geo_launch_count(b200geo_ctx *ctx);
/* Number of (hypothesis, correspondence) pairs whose inlier decision fell inside the rounding shell
 * around the threshold and was therefore taken by the reference-order chain (the "reported" band
 * hits); cumulative since creation or the last reset.  Synchronises the device. */
int b200geo_stats(b200geo_ctx *ctx, unsigned long long *exact_path_pairs, int reset);
/* options */
#define B200GEO_OPT_FORCE_EXACT 1   /* value != 0: score every pair with the reference-order chain */
#define B200GEO_OPT_WARP_EXACT 2    /* value != 0: warp uint8 RGB with the float64 kernel (bit-exact bilinear) */
#define B200GEO_OPT_WARP_VARIANT 3  /* fast uint8 warp: 0 = one warp per 32x20 strip where the layout allows, persistent tile kernel for the
                                       rest (default); 1 = one CTA per tile; 2 = persistent tile kernel only */
int b200geo_set_option(b200geo_ctx *ctx, int option, int value);
/* Number of output pixels the fast uint8 warp kernel evaluated with the float64 reference-order
 * path (inside the coordinate-error band around a mask / rounding boundary, or in tiles where
 * the float32 bound does not hold); cumulative.  Synchronises the device. */
int b200geo_warp_stats(b200geo_ctx *ctx, unsigned long long *exact_pixels, int reset);
/* blocks until all work of this context on `stream` is done */
int b200geo_sync(b200geo_ctx *ctx, void *stream);

/* ------------------------------------------------------------------------------------------------
 * (1) 4-point solve -- replaces calcHomography(u, v) on 4 points (homography.py:71-88, matrix of
 *     :4-14) as called once per RANSAC iteration by HomoModel.fit (ransac.py:52, 180), and the
 *     4-point calcHomographyLinear(u, v) of the Scanner (homography.py:90-105, 220).
 *     One homography per sample set: H[k] maps pts[idx[k][j]].(x,y) -> .(x',y'), h22 = 1.
 *     flags[k] != 0: degenerate sample (repeated index or singular system); H[k] is then NaN.
 *     Hinv (optional, may be NULL): inverse of H[k] (np.linalg.inv, ransac.py:74) as float32.
 * ---------------------------------------------------------------------------------------------- */
int b200geo_dlt4_dev(b200geo_ctx *ctx, const float *pts, int M, const int32_t *idx, int K,
                     float *H, float *Hinv, uint8_t *flags, void *stream);
int b200geo_dlt4_host(b200geo_ctx *ctx, const float *pts, int M, const int32_t *idx, int K,
                      float *H, float *Hinv, uint8_t *flags);
/* float64 4-point solve of ONE system given as 4 source and 4 destination points
 * (calcHomographyLinear(u, v), homography.py:90-105; u, v are 4 rows of (x, y)).  H: 9 doubles. */
int b200geo_homography4_host(b200geo_ctx *ctx, const double *u, const double *v, double *H);

/* ------------------------------------------------------------------------------------------------
 * (2a) scoring -- replaces computeLoss + threshold + count, ransac.py:182-184 (HomoModel.fwd /
 *      reproj / dist, ransac.py:55-98), for K hypotheses at once.
 *      counts[k] = #{ j : loss(H[k], pts[j]) < th }.  Hinv is required unless method == FWD.
 *      flags (optional): hypotheses with flags[k] != 0 get count 0.
 *      Results are bit-identical to the reference's float32 numpy chain for the same H.
 * ---------------------------------------------------------------------------------------------- */
int b200geo_score_dev(b200geo_ctx *ctx, const float *pts, int M, const float *H, const float *Hinv,
                      const uint8_t *flags, int K, float th, int method, int32_t *counts, void *stream);
int b200geo_score_host(b200geo_ctx *ctx, const float *pts, int M, const float *H, const float *Hinv,
                       int K, float th, int method, int32_t *counts);
/* per-point loss of ONE hypothesis (HomoModel.computeLoss, ransac.py:84-98); err: M float32 */
int b200geo_loss_host(b200geo_ctx *ctx, const float *pts, int M, const float *H, const float *Hinv,
                      int method, float *err);

/* HomoModel.fwd / HomoModel.reproj as standalone calls (ransac.py:55-76): P is 3 x M row-major
 * float32 (third row 1 for 2-row input), Hm the 3x3 to apply (val, or inv(val) for reproj);
 * out (3 x M) = (Hm @ P) / (row 2 + 1e-10). */
int b200geo_project_host(b200geo_ctx *ctx, const float *Hm, const float *P, int M, float *out);
/* HomoModel.dist (ransac.py:78-82): out[j] = sqrt(sum_r (a[r][j] - b[r][j])^2); a, b rows x M. */
int b200geo_dist_host(b200geo_ctx *ctx, const float *a, const float *b, int rows, int M, float *out);

/* ------------------------------------------------------------------------------------------------
 * (2b) RANSAC search -- replaces the iteration loop of RANSAC.run (ransac.py:176-202) over the
 *      explicit sample sets idx[k_begin .. k_end): solve, score, and reduce to two 64-bit keys
 *        keys[0] = max_k  (count_k << 32) | (0xFFFFFFFF - k)          (best count, lowest index)
 *        keys[1] = max_k  [count_k >= thr] * (0x100000000 - k)        (first k meeting M*d/100+n; 0 = none)
 *      Both reduce across GPUs with one MAX all-reduce of two int64 (hypothesis-block sharding).
 *      keys must be zeroed by the caller before the first shard is accumulated (or pass
 *      zero_keys = 1).  counts (optional): K_total int32, entry k written for k in the range.
 * ---------------------------------------------------------------------------------------------- */
int b200geo_ransac_search_dev(b200geo_ctx *ctx, const float *pts, int M, const int32_t *idx,
                              int k_begin, int k_end, float th, double thr, int method,
                              unsigned long long *keys, int zero_keys, int32_t *counts, void *stream);

/* (2c) finish -- replaces ransac.py:203-211: decode the winner from keys, recompute its H, its
 *      inlier mask (err < th) and the least-squares refit over the inliers (calcHomographyLinear
 *      collective, homography.py:48-69, 90-105 -- solved in float64 on the device).
 *      result (device, 32 doubles):  [0] winner k (-1: none)  [1] inlier count  [2] status
 *      (0 ok, B200GEO_ERR_*)  [3..11] H of the winning sample (float32 values)  [12..20] refit H.
 *      mask: M uint8. */
int b200geo_ransac_finish_dev(b200geo_ctx *ctx, const float *pts, int M, const int32_t *idx,
                              const unsigned long long *keys, float th, int method,
                              double *result, uint8_t *mask, void *stream);

/* (2b', 2c') the same search when every GPU holds only ITS OWN shard of the sample sets (rows k_offset .. k_offset + n_local - 1
 *      of the K x 4 index array, at idx_shard): the 48-byte record `rec` (device) receives this shard's two keys and the four
 *      correspondence indices of the sample sets they name.  The ranks exchange their records (an all-gather of 48 bytes each);
 *      b200geo_ransac_finish_recs_dev then takes the n_recs records, reduces the keys (MAX) and finishes as
 *      b200geo_ransac_finish_dev does -- on every rank, no broadcast.  counts_shard (optional): n_local int32. */
int b200geo_ransac_search_shard_dev(b200geo_ctx *ctx, const float *pts, int M, const int32_t *idx_shard, int k_offset, int n_local,
                                    float th, double thr, int method, void *rec, int32_t *counts_shard, void *stream);
int b200geo_ransac_finish_recs_dev(b200geo_ctx *ctx, const float *pts, int M, const void *recs, int n_recs, float th, int method,
                                   double *result, uint8_t *mask, void *stream);

/* Whole RANSAC.run (ransac.py:159-213) on host buffers: H2D, search, finish, D2H.  The sample sets are uploaded
 * in chunks on a second stream while earlier chunks are being solved and scored.
 *   thr = M*d/100 + n as the reference computes it (ransac.py:169,186).
 *   out_H_best: 9 float32 (winning 4-point hypothesis), out_H: 9 float64 (refit, the reference's
 *   return value), out_mask: M uint8, out_winner / out_count: scalars, out_counts (optional): K inlier counts,
 *   out_last_count (optional): the count of sample set K-1 -- what the reference's closing warning tests when no
 *   hypothesis met the threshold (ransac.py:203-204).
 *   Returns B200GEO_ERR_NOMODEL / B200GEO_ERR_SINGULAR where the reference raises. */
int b200geo_ransac_host(b200geo_ctx *ctx, const float *pts, int M, const int32_t *idx, int K,
                        float th, double thr, int method, float *out_H_best, double *out_H,
                        uint8_t *out_mask, int32_t *out_winner, int32_t *out_count, int32_t *out_counts,
                        int32_t *out_last_count);

/* Batched form (config "batched stitch pipeline"): B independent pairs, each with its own
 * correspondence set (M rows) and K sample sets; one launch sequence, no host round trips.
 * pts: B*M*4, idx: B*K*4, results: B*32 doubles (layout of b200geo_ransac_finish_dev), masks: B*M. */
int b200geo_ransac_batched_dev(b200geo_ctx *ctx, const float *pts, int M, const int32_t *idx, int K,
                               int B, float th, double thr, int method, double *results,
                               uint8_t *masks, void *stream);

/* Least-squares refit alone: calcHomographyLinear(u, v, collective=True) (homography.py:90-105)
 * over the rows with mask[j] != 0 (mask NULL = all rows).  H: 9 doubles. */
int b200geo_refit_host(b200geo_ctx *ctx, const float *pts, int M, const uint8_t *mask, double *H);

/* The homogeneous least-squares homography over M correspondences: calcHomography(u, v, collective=True) (homography.py:71-88
 * with the 2M x 9 matrix of calc_correspCollective, :30-46) -- the right singular vector of the smallest singular value,
 * scaled to h22 = 1.  H: 9 doubles. */
int b200geo_homography_svd_host(b200geo_ctx *ctx, const float *pts, int M, double *H);

/* ------------------------------------------------------------------------------------------------
 * (3) backward warp -- replaces wrapPerspective / wrapPerspectiveScan + the samplers
 *     nearestNeighbor / bilinear (homography.py:108-209) and the astype(uint8) of
 *     transformImage / transformImageH (homography.py:225, 241).
 *
 *     A job describes one image: the output sample grid  gx[c] = x0 + c*dx (last column = x_last),
 *     gy[r] likewise -- exactly np.linspace's values (homography.py:166-167, 197-198) --, the
 *     float64 inverse homography, the bounds the sampler tests against (image h,w for the bbox
 *     form; `res` for the Scan form, homography.py:208), the true image size, and where the
 *     image and its output live.  Source pixel (0,0) reads as 0 (homography.py:112-116).
 * ---------------------------------------------------------------------------------------------- */
typedef struct b200geo_warp_job {
    double invH[9];          /* np.linalg.inv(H), float64 (homography.py:172, 203) */
    double x0, dx, x_last;   /* output grid columns */
    double y0, dy, y_last;   /* output grid rows */
    int32_t ow, oh;          /* output size (max_w, max_h) */
    int32_t bw, bh;          /* bounds given to the sampler (w, h) */
    int32_t iw, ih;          /* true source size */
    int32_t reserved0, reserved1;
    int64_t src_offset;      /* byte offset of this image from the `src` base pointer */
    int64_t src_pitch;       /* bytes per source row */
    int64_t dst_offset;      /* byte offset of this output from the `dst` base pointer */
    int64_t dst_pitch;       /* bytes per output row */
} b200geo_warp_job;

/* n_jobs images in one launch.  channels: 3 or 4.  src_type: B200GEO_U8 | B200GEO_F32.
 * dst_type: B200GEO_U8 (value truncated toward zero, astype(np.uint8)), B200GEO_F32 or B200GEO_F64.
 * jobs: HOST array (copied to the device by the call). */
int b200geo_warp_dev(b200geo_ctx *ctx, const void *src, void *dst, const b200geo_warp_job *jobs,
                     int n_jobs, int channels, int src_type, int dst_type, int sampler, void *stream);
/* jobs already resident on the device (no host work at all; used by the batched pipelines) */
int b200geo_warp_jobs_dev(b200geo_ctx *ctx, const void *src, void *dst, const b200geo_warp_job *jobs_dev,
                          int n_jobs, int max_ow, int max_oh, int channels, int src_type, int dst_type,
                          int sampler, void *stream);
/* Same, with a description of the source buffer: every job's image lies inside one 2-D byte array
 * starting at `src` (16-byte aligned) of `rows` rows, `pitch` bytes apart (pitch % 16 == 0); a job
 * whose src_pitch equals `pitch` and whose src_offset is a multiple of it is then fetched with one
 * 2-D TMA tensor copy per output tile instead of one bulk copy per source row.  Frames stored back
 * to back (the batched configs) are such an array with rows = n_frames * height.  Results do not
 * depend on the layout hint; layout == NULL is b200geo_warp_jobs_dev. */
typedef struct b200geo_src_layout {
    int64_t pitch;       /* bytes between rows */
    int64_t rows;        /* number of rows readable from src */
    /* Optional (0 = unknown): every image occupies image_rows consecutive rows starting at a multiple of
     * image_rows, each row holding row_bytes image bytes (= 3 * iw).  The array is then described to the TMA unit
     * as row_bytes x image_rows x images, and footprints that hang over an image edge are fetched by tensor
     * copy too, with hardware zero fill -- exactly the samplers' view of out-of-image pixels. */
    int64_t image_rows;
    int64_t row_bytes;
} b200geo_src_layout;
int b200geo_warp_batch_dev(b200geo_ctx *ctx, const void *src, void *dst, const b200geo_warp_job *jobs_dev,
                           int n_jobs, int max_ow, int max_oh, int channels, int src_type, int dst_type,
                           int sampler, const b200geo_src_layout *layout, void *stream);
/* A batch of images on HOST buffers (pinned memory gives full overlap; pageable memory works):
 * jobs are processed in chunks of `chunk_jobs` (0 = default: 8, or 16 from 128 jobs) through a three-stage pipeline --
 * host->device copies of chunk i+1, the warp of chunk i and device->host copies of chunk i-1 run
 * concurrently on three streams with double-buffered device staging.  src_offset / dst_offset of
 * the jobs index the host buffers; each job's image (src_pitch * ih bytes) and output
 * (dst_pitch * oh bytes) must be contiguous there.  Synchronous: returns when dst_host is complete. */
int b200geo_warp_batch_host(b200geo_ctx *ctx, const void *src_host, void *dst_host, const b200geo_warp_job *jobs,
                            int n_jobs, int channels, int src_type, int dst_type, int sampler, int chunk_jobs);
/* single image on host buffers: H2D, warp, D2H */
int b200geo_warp_host(b200geo_ctx *ctx, const void *src, void *dst, const b200geo_warp_job *job,
                      int channels, int src_type, int dst_type, int sampler);

/* ------------------------------------------------------------------------------------------------
 * (4) stitch compose, fused with the warp -- replaces the body of stitchPanorama (homography.py:288-338)
 *     for uint8 RGB images: addAlpha('Rate') (:250-286), transformImageH's bilinear bbox warp in float64
 *     (:230-242), the canvas placement, paste or alpha blend, and the final astype(uint8).
 *     `job` describes the warp of imgT exactly as for b200geo_warp_* (its ow x oh output is the warped
 *     rectangle; dst_offset / dst_pitch are ignored), `p` where that rectangle and imgQ sit on the canvas
 *     (homography.py:303-321).  blend = 0: canvas = 0, warped T pasted, Q pasted on top (:335-338).
 *     blend = 1 ('Rate'): out = (aq/(aq+at)) Q + (at/(aq+at)) T in float64, through a float32 canvas
 *     (:322-334); alpha_t / alpha_q / alpha_bg are the float32 values the reference stores:
 *     float32(rate + 1e-10), float32(1 + 1e-10 - rate), float32(1e-10).
 * ---------------------------------------------------------------------------------------------- */
typedef struct b200geo_compose {
    int32_t fw, fh;          /* canvas size */
    int32_t tsx, tsy;        /* canvas position of the warped rectangle (job->ow x job->oh) */
    int32_t qsx, qsy;        /* canvas position of imgQ */
    int32_t qw, qh;          /* imgQ size */
    int64_t q_pitch;         /* bytes per imgQ row */
    int64_t canvas_pitch;    /* bytes per canvas row */
    double alpha_t, alpha_q, alpha_bg;
    int32_t blend;           /* 0 paste, 1 'Rate' blend */
    int32_t reserved;
} b200geo_compose;
/* device buffers (uint8 RGB), asynchronous on `stream` */
int b200geo_stitch_compose_dev(b200geo_ctx *ctx, const void *imgT, const void *imgQ, void *canvas,
                               const b200geo_warp_job *job, const b200geo_compose *p, void *stream);
/* Batched stitch pipeline: the warp jobs of B pairs from their RANSAC results, on the device (no host round trip between
 * b200geo_ransac_batched_dev and b200geo_warp_batch_dev).  Pair b: transformImageH(frame_b, H_b) (homography.py:230-242) with
 * wrapPerspective's output rectangle (:143-163) and inv(H) (:172), frame b at src offset b * ih * iw * 3, output into canvas
 * slot b (canvas_h x canvas_w x 3, back to back), clipped to the slot.  results: B x 32 doubles (b200geo_ransac_finish_dev
 * layout); jobs_dev: B jobs; boxes_dev: B x 5 int32 = (min_x, min_y, out_w, out_h, ok); ok = 0: the search failed or the
 * projected corners are not finite -- the job is the identity warp. */
int b200geo_stitch_jobs_dev(b200geo_ctx *ctx, const double *results, int B, int ih, int iw, int canvas_h, int canvas_w,
                            b200geo_warp_job *jobs_dev, int32_t *boxes_dev, void *stream);
/* host buffers: H2D of both images, one kernel, D2H of the canvas */
int b200geo_stitch_compose_host(b200geo_ctx *ctx, const void *imgT, const void *imgQ, void *canvas,
                                const b200geo_warp_job *job, const b200geo_compose *p);

/* ------------------------------------------------------------------------------------------------
 * (5) brute-force Hamming matcher -- replaces cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck).match(A, B) in
 *     stitching() (ransac.py:258-259), the step between ORB and RANSAC.run.  descA: nA x desc_bytes, descB:
 *     nB x desc_bytes (uint8, desc_bytes a multiple of 4, <= 256; ORB: 32).  train_idx[i] = index in B of the
 *     nearest descriptor of A[i] (first index on ties, as OpenCV) or -1 (nB == 0, or rejected by the
 *     cross-check: A[i] is not the nearest descriptor of that B entry); dist[i] = its Hamming distance.
 *     Host buffers; train_idx and dist hold nA int32 each.
 * ---------------------------------------------------------------------------------------------- */
int b200geo_match_hamming_host(b200geo_ctx *ctx, const uint8_t *descA, int nA, const uint8_t *descB, int nB,
                               int desc_bytes, int cross_check, int32_t *train_idx, int32_t *dist);

/* The samplers called on explicit coordinates: nearestNeighbor / bilinear(z_t, img, h, w, mh, mw)
 * (homography.py:108-138).  zx, zy: N float64 each (rows 0 and 1 of z_t); bh, bw: the bounds the
 * sampler tests against; dst: N*channels values of dst_type.  Host buffers. */
int b200geo_sample_host(b200geo_ctx *ctx, const void *src, int ih, int iw, int channels, int src_type,
                        const double *zx, const double *zy, long long N, int bh, int bw, int sampler,
                        int dst_type, void *dst);

/* ------------------------------------------------------------------------------------------------
 * (6) cv2.remap(img, map_x, map_y, INTER_LINEAR, borderMode=BORDER_TRANSPARENT) for uint8 images of 1, 2 or 4
 *     channels -- the resampling step of cylindricalWarp (homography.py:391-410: BGRA image, float32 maps,
 *     INTER_AREA which cv::remap executes as INTER_LINEAR).  OpenCV's fixed-point algorithm (1/32 px coordinates,
 *     15-bit weights, round half up), bit-exact; destination pixels whose integer source position is outside
 *     the source are left untouched (dst is read and written).  Host buffers; maps and dst are oh x ow.
 * ---------------------------------------------------------------------------------------------- */
int b200geo_remap_linear_host(b200geo_ctx *ctx, const uint8_t *src, int ih, int iw, int channels,
                              const float *map_x, const float *map_y, int oh, int ow, uint8_t *dst);

/* ------------------------------------------------------------------------------------------------
 * (7) device-side JPEG decode / encode -- the GPU counterparts of cv2.imread (app.py:122,154,158;
 *     ransac.py:291-294) and cv2.imwrite (app.py:193-195; ransac.py:301): file bytes <-> interleaved 8-bit
 *     pixels in HBM, so that decoded frames never cross the host link.  Codec: nvJPEG, loaded on first use
 *     (B200GEO_ERR_NODEVICE if libnvjpeg.so.12 is absent).  order: B200GEO_BGR (cv2's) or B200GEO_RGB.
 *     Grayscale files decode to three equal channels (cv2.IMREAD_COLOR).  Decoded values may differ from
 *     libjpeg-turbo's (cv2) by a few grey levels: different IDCT and chroma up-sampling; censused in the tests.
 *     decode: dst holds height rows of dst_pitch >= 3*width bytes, on the device (asynchronous on `stream`)
 *     or on the host (dst_on_device = 0: synchronous).  encode: cv2.imwrite's defaults (pass quality 95;
 *     4:2:0 chroma, baseline Huffman); out = NULL queries *nbytes (the encode is then already done and a
 *     second call with a buffer retrieves nothing new -- call once with a buffer of height*width*3 + 4096).
 * ---------------------------------------------------------------------------------------------- */
#define B200GEO_BGR 0
#define B200GEO_RGB 1
int b200geo_jpeg_info(b200geo_ctx *ctx, const uint8_t *data, size_t nbytes, int *height, int *width, int *components);
int b200geo_jpeg_decode(b200geo_ctx *ctx, const uint8_t *data, size_t nbytes, int order, void *dst,
                        long long dst_pitch, int dst_on_device, void *stream);
int b200geo_jpeg_encode(b200geo_ctx *ctx, const void *src, int height, int width, long long src_pitch, int order,
                        int src_on_device, int quality, uint8_t *out, size_t capacity, size_t *nbytes, void *stream);

/* ------------------------------------------------------------------------------------------------
 * (8) ORB feature front-end -- replaces cv2.ORB_create().detectAndCompute(gray, None) in stitching()
 *     (ransac.py:255-257), the step before the Hamming matcher (5).  OpenCV's defaults as the reference uses
 *     them (8 levels, scale 1.2f, edge threshold 31, patch 31, FAST threshold 20, Harris score, WTA_K 2);
 *     nfeatures is cv2.ORB_create's first argument (500).  img: height x width uint8 with `pitch` bytes per row,
 *     host memory; channels = 1 (grey) or 3 (interleaved RGB: cv2.cvtColor(img, COLOR_RGB2GRAY) of ransac.py:252-253
 *     is then done on the device, bit-exact).  keypoints: capacity x 6 float32 (x, y, size, angle in degrees, Harris response, octave);
 *     descriptors: capacity x 32 bytes; *n_out = number found (a few more than nfeatures when responses tie,
 *     as in OpenCV).  Positions, sizes, angles, responses, octaves, the ORDER of the keypoints and the
 *     descriptors are bit-identical to OpenCV 4.13's (x86-64 build with AVX2 dispatch).  Pyramid, FAST, Harris,
 *     orientation, blur and the 256 pair tests run on the GPU; the two retainBest selections
 *     (std::nth_element order) and cosf / sinf of the angles run on the host.
 * ---------------------------------------------------------------------------------------------- */
int b200geo_orb_host(b200geo_ctx *ctx, const uint8_t *img, int height, int width, long long pitch, int channels, int nfeatures,
                     float *keypoints, uint8_t *descriptors, int capacity, int *n_out);

/* ------------------------------------------------------------------------------------------------
 * (9) The record exchange of a sharded search over peer memory (NVLink / NVSwitch) instead of a collective: one
 *     process per GPU of one node.  b200geo_peer_export allocates this rank's inbox and returns its 64-byte CUDA
 *     IPC handle; the caller gathers the handles of all ranks (any transport: they are host bytes) and passes the
 *     array to b200geo_peer_open, which maps every other rank's inbox; after a barrier,
 *     b200geo_ransac_exchange_peer_dev replaces the all-gather between b200geo_ransac_search_shard_dev and
 *     b200geo_ransac_finish_recs_dev: it stores `rec` (48 bytes, device) into every rank's inbox and fills `recs`
 *     (world x 48 bytes, device) with the records of all ranks -- two single-block kernels on `stream`, no host
 *     synchronisation.  All ranks must call it the same number of times.  A rank that does not answer within a few
 *     seconds contributes an empty record (the search then reports B200GEO_ERR_NOMODEL if no other rank has one).
 * ---------------------------------------------------------------------------------------------- */
int b200geo_peer_export(b200geo_ctx *ctx, int world, void *handle_out);
int b200geo_peer_open(b200geo_ctx *ctx, int rank, int world, const void *handles);
int b200geo_peer_close(b200geo_ctx *ctx);
int b200geo_ransac_exchange_peer_dev(b200geo_ctx *ctx, const void *rec, void *recs, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* B200GEO_H */
