/*
 * include/docscan.h — C ABI of libdocscan.so, the B200 (sm_100a) implementation of the per-pixel
 * document-scan path of Brianlov/Smart-Image-Processing (DocScanner.py + morph_seq).
 *
 * The reference has no native layer: DocScanner.py calls OpenCV (cv2.*) for every pixel operation.
 * Each entry point below replaces one of those cv2 call sites (cited as DocScanner.py:LINE) with a
 * hand-written CUDA kernel that reproduces OpenCV's arithmetic bit for bit, or replaces a whole
 * reference stage function / the per-pixel part of process_document with a fused kernel sequence.
 * INTEGRATION.md shows the ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C, no C++/torch types; every function returns 0 (DOCSCAN_OK) or a negative error code;
 *     docscan_last_error(ctx) gives the message.  There is NO CPU fallback: without a CUDA device
 *     docscan_create fails with DOCSCAN_ERR_NO_DEVICE.
 *   - images are uint8, row-major, interleaved channels, `pitch` in bytes.  `space` says where the
 *     pointer lives: DOCSCAN_HOST (numpy buffers; the library copies in and out through its own
 *     device scratch, pinned fast path when the buffer came from docscan_host_alloc) or
 *     DOCSCAN_DEVICE (CUDA device memory of ctx's device, e.g. a torch tensor's data_ptr()).
 *   - all work is enqueued on the context's stream; host-space calls return after the results are
 *     in the caller's buffer, device-space calls return after enqueueing (docscan_sync to wait).
 *   - a context is not thread-safe; use one per thread (the reference calls the path from a worker
 *     thread, AI_classification.py:855).  No global mutable state.
 */
#ifndef DOCSCAN_H
#define DOCSCAN_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define DOCSCAN_API __attribute__((visibility("default")))
#else
#define DOCSCAN_API
#endif

#define DOCSCAN_OK 0
#define DOCSCAN_ERR_NO_DEVICE (-1)
#define DOCSCAN_ERR_CUDA (-2)
#define DOCSCAN_ERR_BAD_ARG (-3)
#define DOCSCAN_ERR_NOMEM (-4)
#define DOCSCAN_ERR_UNSUPPORTED (-5)

#define DOCSCAN_HOST 0
#define DOCSCAN_DEVICE 1

typedef struct docscan_ctx docscan_ctx;

typedef struct docscan_image {
    void* data;        /* first pixel */
    int32_t width;     /* pixels */
    int32_t height;
    int32_t pitch;     /* bytes between rows */
    int32_t channels;  /* 1 or 3 */
    int32_t space;     /* DOCSCAN_HOST | DOCSCAN_DEVICE */
} docscan_image;

/* ---- context -------------------------------------------------------------------------------- */
DOCSCAN_API int docscan_version(void);
DOCSCAN_API const char* docscan_strerror(int code);
/* `stream` may be NULL (the library creates its own non-blocking stream) or a cudaStream_t /
 * torch.cuda.Stream().cuda_stream of `device`. */
DOCSCAN_API int docscan_create(int device, void* stream, docscan_ctx** out);
DOCSCAN_API int docscan_destroy(docscan_ctx* ctx);
DOCSCAN_API int docscan_sync(docscan_ctx* ctx);
/* the context's cudaStream_t, for callers that enqueue their own work in order with the library's (e.g. the optional
 * nvJPEG decode in control.py, which replaces cv2.imread of DocScanner.py:15-19 when decode="device" is asked for) */
DOCSCAN_API int docscan_get_stream(docscan_ctx* ctx, void** stream);
DOCSCAN_API const char* docscan_last_error(docscan_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
DOCSCAN_API int64_t docscan_launch_count(docscan_ctx* ctx);
/* bytes this context has copied between caller HOST images and the device so far (bench.py's h2d/d2h_bytes_per_step).
 * docscan_process_pages uploads only the part of each HOST photo that lies under its quad. */
DOCSCAN_API int docscan_transfer_bytes(docscan_ctx* ctx, int64_t* h2d, int64_t* d2h);
/* per-kernel timing for bench.py: when enabled every kernel launch is bracketed by CUDA events on the
 * context's stream.  docscan_profile_dump syncs, writes one line per kernel name
 * ("name launches total_ms algorithmic_bytes\n") into buf, clears the records and returns the length. */
DOCSCAN_API int docscan_profile_enable(docscan_ctx* ctx, int on);
DOCSCAN_API int docscan_profile_dump(docscan_ctx* ctx, char* buf, size_t cap);
/* pinned host memory for the fast HOST path */
DOCSCAN_API int docscan_host_alloc(docscan_ctx* ctx, size_t bytes, void** out);
DOCSCAN_API int docscan_host_free(docscan_ctx* ctx, void* p);
/* raw device memory (for callers without torch) */
DOCSCAN_API int docscan_device_alloc(docscan_ctx* ctx, size_t bytes, void** out);
DOCSCAN_API int docscan_device_free(docscan_ctx* ctx, void* p);
DOCSCAN_API int docscan_memcpy_h2d(docscan_ctx* ctx, void* dst, const void* src, size_t bytes);
DOCSCAN_API int docscan_memcpy_d2h(docscan_ctx* ctx, void* dst, const void* src, size_t bytes);

/* ---- parameter preparation (host arithmetic, bit-exact with cv2) ------------------------------ */
/* cv2.getPerspectiveTransform(quad, dst)                                 DocScanner.py:142 */
DOCSCAN_API int docscan_get_perspective_transform(const float quad[8], const float dst[8], double m[9]);
/* cv2.getRotationMatrix2D((cx, cy), angle_deg, 1.0)                      DocScanner.py:234 */
DOCSCAN_API int docscan_get_rotation_matrix(double cx, double cy, double angle_deg, double m[6]);
/* cv2.getGaussianKernel(k, 0, CV_32F) and its 8.8 fixed-point form used for uint8 images */
DOCSCAN_API int docscan_gaussian_kernel_f32(int k, float* out);
DOCSCAN_API int docscan_gaussian_kernel_q8(int k, int32_t* out);
/* cv2.threshold(..., THRESH_OTSU) threshold from a 256-bin histogram     DocScanner.py:187,202 */
DOCSCAN_API int docscan_otsu_from_hist(const int32_t hist[256], int64_t total, double* t);

/* ---- one entry per cv2 call on the path -------------------------------------------------------- */
/* cv2.warpPerspective(img, M, (w,h), INTER_LINEAR) u8c3 (or u8c1), border 0   DocScanner.py:143
 * `gray_out` (may be NULL) additionally receives BGR2GRAY of the result (fused, u8c1). */
DOCSCAN_API int docscan_warp_perspective(docscan_ctx*, const docscan_image* src, const double m_fwd[9],
                             docscan_image* dst, docscan_image* gray_out);
/* cv2.cvtColor(BGR2GRAY) (swap_rb=0) / RGB2GRAY (swap_rb=1)              DocScanner.py:316 */
DOCSCAN_API int docscan_bgr2gray(docscan_ctx*, const docscan_image* src, docscan_image* dst, int swap_rb);
/* cv2.GaussianBlur(gray, (k,k), 0), BORDER_REFLECT_101                   DocScanner.py:153,184 */
DOCSCAN_API int docscan_gaussian_blur(docscan_ctx*, const docscan_image* src, int k, docscan_image* dst);
/* cv2.subtract / cv2.divide(scale=255) / cv2.max / dst = base where mask!=0 else 255
 *                                                                        DocScanner.py:158,155,207,338-339 */
#define DOCSCAN_OP_SUB 0
#define DOCSCAN_OP_DIV255 1
#define DOCSCAN_OP_MAX 2
#define DOCSCAN_OP_MASK_SELECT 3
DOCSCAN_API int docscan_binary_op(docscan_ctx*, int op, const docscan_image* a, const docscan_image* b, docscan_image* dst);
DOCSCAN_API int docscan_minmax(docscan_ctx*, const docscan_image* src, int32_t* mn, int32_t* mx);
DOCSCAN_API int docscan_hist256(docscan_ctx*, const docscan_image* src, int32_t hist[256]);
/* cv2.normalize(src, None, 0, 255, NORM_MINMAX)                          DocScanner.py:156,159,172,186,201 */
DOCSCAN_API int docscan_normalize_minmax(docscan_ctx*, const docscan_image* src, docscan_image* dst);
/* cv2.threshold(src, 0, 255, BINARY+OTSU): returns the Otsu threshold; dst may be NULL */
DOCSCAN_API int docscan_otsu_threshold(docscan_ctx*, const docscan_image* src, double* t, docscan_image* dst);
/* cv2.threshold(src, t, 255, THRESH_BINARY)                              DocScanner.py:189,204 */
DOCSCAN_API int docscan_threshold_binary(docscan_ctx*, const docscan_image* src, int t, docscan_image* dst);
/* cv2.erode / cv2.dilate / cv2.morphologyEx(CLOSE|BLACKHAT), MORPH_RECT kw x kh, default anchor
 *                                                                        DocScanner.py:199-200,211-212,251-254 */
#define DOCSCAN_MORPH_ERODE 0
#define DOCSCAN_MORPH_DILATE 1
#define DOCSCAN_MORPH_CLOSE 2
#define DOCSCAN_MORPH_BLACKHAT 3
#define DOCSCAN_MORPH_OPEN 4
DOCSCAN_API int docscan_morph_rect(docscan_ctx*, int op, const docscan_image* src, int kw, int kh, int iterations,
                       docscan_image* dst);
/* cv2.adaptiveThreshold(gray, 255, MEAN_C|GAUSSIAN_C, THRESH_BINARY, k, C)   DocScanner.py:167
 * Block sizes: GAUSSIAN_C odd 3..257, MEAN_C odd 3..255.
 * cv_tail_compat != 0 reproduces the unfused arithmetic cv2's AVX2 build uses in the last
 * width % 8 columns of GAUSSIAN_C; 0 = fma in every column. */
#define DOCSCAN_ADAPTIVE_MEAN 0
#define DOCSCAN_ADAPTIVE_GAUSSIAN 1
DOCSCAN_API int docscan_adaptive_threshold(docscan_ctx*, const docscan_image* src, int method, int k, int c,
                               int cv_tail_compat, docscan_image* dst);
/* cv2.warpAffine(gray, M, (w,h), INTER_LINEAR, BORDER_REPLICATE) u8c1    DocScanner.py:235 */
DOCSCAN_API int docscan_warp_affine(docscan_ctx*, const docscan_image* src, const double m_fwd[6], docscan_image* dst);

/* ---- the skew estimate of deskew() (DocScanner.py:218-231), on the device ---------------------------- */
/* cv2.Canny(gray, low, high) (aperture 3, L1 gradient) u8c1 -> {0,255}           DocScanner.py:218 (and :78) */
DOCSCAN_API int docscan_canny(docscan_ctx*, const docscan_image* src, double low, double high, docscan_image* dst);
/* cv2.HoughLines(edges, 1, pi/180, threshold): every non-zero pixel votes.  Writes up to max_lines (rho, theta) float
 * pairs in OpenCV's order (votes descending, then accumulator index) and the total number of lines to *n_lines;
 * per_angle (may be NULL) receives the number of lines per angle index 0..179.       DocScanner.py:219 */
DOCSCAN_API int docscan_hough_lines(docscan_ctx*, const docscan_image* edges, int threshold, float* rho_theta, int max_lines,
                                    int32_t* n_lines, int32_t per_angle[180]);
/* np.median of the folded line angles in numpy's float32 arithmetic, 0 beyond max_rotate (host arithmetic)
 *                                                                                   DocScanner.py:221-231 */
DOCSCAN_API int docscan_median_angle(const int32_t per_angle[180], double max_rotate, double* angle_deg);
/* Canny -> HoughLines(1, pi/180, 150) -> median angle: the angle deskew() rotates by      DocScanner.py:218-231 */
DOCSCAN_API int docscan_skew_angle(docscan_ctx*, const docscan_image* gray, double canny_low, double canny_high,
                                   double max_rotate, double* angle_deg);

/* cv2.resize(img, (dst.width, dst.height), interpolation=INTER_AREA | INTER_CUBIC), u8c1 / u8c3: resize_long_side,
 * DocScanner.py:27-36 (the whole-photo fallback taken at :313).  INTER_AREA is implemented for shrinking (bit-exact
 * with cv2); INTER_CUBIC follows OpenCV's own code path — bit-exact with cv2 when IPP is off, within 1 LSB of the
 * IPP-enabled wheels (IPP substitutes its own float cubic there).  cv_tail_compat as in docscan_adaptive_threshold:
 * != 0 reproduces the fixed-point arithmetic cv2's SIMD build uses in the last (width*channels) % 8 elements of a row. */
#define DOCSCAN_INTER_CUBIC 2
#define DOCSCAN_INTER_AREA 3
DOCSCAN_API int docscan_resize(docscan_ctx*, const docscan_image* src, docscan_image* dst, int interpolation, int cv_tail_compat);

/* ---- fused reference stage functions ----------------------------------------------------------- */
/* illumination_correction(gray, method, blur_frac)                       DocScanner.py:147-160
 * method 0 = subtract, 1 = divide; k = the odd kernel size the reference derives from blur_frac. */
DOCSCAN_API int docscan_illumination_correction(docscan_ctx*, const docscan_image* gray, int method, int k, docscan_image* dst);
/* _compute_ink_mask(gray, ...)                                           DocScanner.py:175-214
 * kw_bh x kh_bh is the black-hat rectangle after the reference's odd/min-3 fix-ups. */
DOCSCAN_API int docscan_ink_mask(docscan_ctx*, const docscan_image* gray, int mask_blur_ksize, int kw_bh, int kh_bh,
                     int dilate_iters, int threshold_offset, docscan_image* dst);

/* ---- the whole per-pixel path, batched ---------------------------------------------------------- */
typedef struct docscan_params {      /* process_document tunables that touch pixels (DocScanner.py:262-276) */
    int32_t illum_method;            /* 0 subtract, 1 divide */
    double illum_blur_frac;
    int32_t block_size, C, thresh_method;   /* thresh_method: DOCSCAN_ADAPTIVE_* */
    int32_t mask_blur_ksize, blackhat_ksize;
    double blackhat_vertical_ratio;
    int32_t ink_dilate_iters, mask_thresh_offset;
    int32_t morph_ksize, morph_iters;
    int32_t cv_tail_compat;
    /* deskew()'s own estimate, used for pages whose angle_deg is NaN (DocScanner.py:218-231, :342) */
    double canny_low, canny_high, max_rotate;
} docscan_params;

typedef struct docscan_page {
    docscan_image src;               /* photo, u8c3 BGR */
    float quad[8];                   /* TL,TR,BR,BL from the control path (localize_document) */
    double angle_deg;                /* deskew angle from the control path; NaN = estimate it on the device from the
                                        blended page exactly like deskew() (Canny + HoughLines median); read the
                                        result with docscan_last_angles */
    docscan_image warped;            /* out: u8c3, size = target size of the page (docscan_target_size) */
    docscan_image binary;            /* out: u8c1, same size */
    int32_t use_whole;               /* != 0: no usable quad — `warped` = resize_long_side(src) (DocScanner.py:313): the
                                        photo resized to warped's size, INTER_AREA when that shrinks its long side,
                                        INTER_CUBIC otherwise; `quad` is ignored */
} docscan_page;

DOCSCAN_API void docscan_default_params(docscan_params* p);                 /* CLI defaults, DocScanner.py:262-276 */
/* perspective_warp's target size (DocScanner.py:120-139); page_kind 0 = A-series, 1 = Letter, 2 = quad ratio */
DOCSCAN_API int docscan_target_size(const float quad[8], int page_kind, int scale_long, int32_t* w, int32_t* h);
/* warp -> gray -> illumination -> stretch -> ink mask || adaptive threshold -> blend -> rotate -> close
 * (DocScanner.py:310-346 without the PNG dumps) for n independent pages. */
DOCSCAN_API int docscan_process_pages(docscan_ctx*, int n, docscan_page* pages, const docscan_params* params);
/* The part of a page's photo docscan_process_pages uploads when `src` is a HOST image: a box [x0, x1) x [y0, y1) that
 * holds every source pixel the perspective warp of that page can read (the whole photo when that cannot be bounded, or
 * for use_whole pages).  region = {x0, y0, x1, y1}.  Host arithmetic only. */
DOCSCAN_API int docscan_warp_footprint(const docscan_page* page, int32_t region[4]);
/* the deskew angle each page of the last docscan_process_pages call was rotated by (supplied or estimated); syncs */
DOCSCAN_API int docscan_last_angles(docscan_ctx*, double* angles, int n);

/* ---- bench support: deterministic synthetic page photos rendered on the device ------------------- */
/* Renders a width x height u8c3 photo of a text page (seeded) into `dst` (DEVICE space) and returns
 * the page quad (TL,TR,BR,BL) the control path would find. */
DOCSCAN_API int docscan_synth_page(docscan_ctx*, uint64_t seed, docscan_image* dst, float quad_out[8]);

#ifdef __cplusplus
}
#endif
#endif /* DOCSCAN_H */
