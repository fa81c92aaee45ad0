// Internal declarations shared by the libdocscan.so translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <array>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "docscan.h"

#ifndef __CUDA_ARCH__
#define DS_HOST_ONLY 1
#endif

#define DS_SM_COUNT_FALLBACK 148
#define DS_MAX_STREAMS 4

// ---------------------------------------------------------------------------------------------
// device image view (uint8, interleaved channels)
struct DImg {
    uint8_t* p;
    int w, h, pitch, ch;
};

struct GaussTable;   // blur.cu

struct docscan_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    int64_t launches = 0;
    double* angles_dev = nullptr;             // deskew angles of the last docscan_process_pages call
    int angles_cap = 0, angles_n = 0;
    int64_t h2d_bytes = 0, d2h_bytes = 0;     // bytes this context has moved between host and device buffers
    int sm_count = DS_SM_COUNT_FALLBACK;
    size_t l2_bytes = 0;
    // bump arena for per-call device scratch
    uint8_t* arena = nullptr;
    size_t arena_size = 0, arena_off = 0;
    std::vector<uint8_t*> retired;       // old arena blocks, freed at the next sync point
    // pinned host staging (job arrays, small read-backs)
    uint8_t* pinned = nullptr;
    size_t pinned_size = 0, pinned_off = 0;
    // cached device coefficient tables, keyed by (kind, k, delta)
    std::map<uint64_t, void*> tables;
    // band matrices of the tensor-core stencils (tcblur.cu), keyed by (axis, k, near-border distances, tile geometry)
    std::map<std::array<int, 6>, void*> tc_tables;
    uint32_t* tc_status = nullptr;            // pinned host words the tensor-core kernels report a timed-out wait into
    std::vector<void*> user_allocs;
    // host-buffer pipeline of docscan_process_pages: copy streams + events (created on first use)
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    // pinned mirror of the two staging sets, for callers whose host buffers are pageable (plain numpy arrays)
    uint8_t* pin_mirror = nullptr;
    size_t pin_mirror_size = 0;
    cudaEvent_t pipe_ev[7] = {};
    // second compute stream of the device-resident batch path
    cudaStream_t aux[DS_MAX_STREAMS] = {};      // [0] unused (the context's own stream)
    cudaEvent_t aux_ev[DS_MAX_STREAMS] = {};
    // optional per-kernel timing (docscan_profile_enable): one event pair per launch
    bool prof_on = false;
    struct ProfRec { std::string name; double bytes; cudaEvent_t a, b; };
    std::vector<ProfRec> prof;
};

// Brackets one kernel launch with CUDA events when profiling is on (bench.py's roofline pass).
struct ProfScope {
    docscan_ctx* ctx;
    bool on;
    ProfScope(docscan_ctx* c, const std::string& name, double alg_bytes) : ctx(c), on(c->prof_on) {
        if (!on) return;
        docscan_ctx::ProfRec r{name, alg_bytes, nullptr, nullptr};
        cudaEventCreate(&r.a);
        cudaEventCreate(&r.b);
        cudaEventRecord(r.a, c->stream);
        c->prof.push_back(r);
    }
    ~ProfScope() {
        if (on) cudaEventRecord(ctx->prof.back().b, ctx->stream);
    }
};

int ds_fail(docscan_ctx* ctx, int code, const char* fmt, ...);

#define DS_CUDA(ctx, expr)                                                                         \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            return ds_fail((ctx), DOCSCAN_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,               \
                           cudaGetErrorString(_e), __FILE__, __LINE__);                            \
    } while (0)

#define DS_TRY(expr)                                                                               \
    do {                                                                                           \
        int _rc = (expr);                                                                          \
        if (_rc != DOCSCAN_OK) return _rc;                                                         \
    } while (0)

#define DS_CHECK_LAUNCH(ctx)                                                                       \
    do {                                                                                           \
        (ctx)->launches++;                                                                         \
        cudaError_t _e = cudaGetLastError();                                                       \
        if (_e != cudaSuccess)                                                                     \
            return ds_fail((ctx), DOCSCAN_ERR_CUDA, "kernel launch failed: %s (%s:%d)",           \
                           cudaGetErrorString(_e), __FILE__, __LINE__);                            \
    } while (0)

// ---- arena ------------------------------------------------------------------------------------
struct ArenaScope {     // everything allocated inside the scope is released when it ends
    docscan_ctx* ctx;
    size_t off0, pin0;
    explicit ArenaScope(docscan_ctx* c) : ctx(c), off0(c->arena_off), pin0(c->pinned_off) {}
    ~ArenaScope() { ctx->arena_off = off0; ctx->pinned_off = pin0; }
};
int ds_arena_reserve(docscan_ctx* ctx, size_t total_bytes);   // call before the first alloc of a big call
int ds_arena_alloc(docscan_ctx* ctx, size_t bytes, void** out);
int ds_arena_image(docscan_ctx* ctx, int w, int h, int ch, DImg* out);
int ds_pinned_alloc(docscan_ctx* ctx, size_t bytes, void** out);
static inline size_t ds_image_bytes(int w, int h, int ch) {
    size_t pitch = ((size_t)w * ch + 127) & ~(size_t)127;
    return pitch * (size_t)h + 256;
}

// ---- staging of API images ----------------------------------------------------------------------
int ds_check_image(docscan_ctx* ctx, const docscan_image* im, int channels, const char* what);
// input: returns a device view (copying host images into the arena)
int ds_stage_in(docscan_ctx* ctx, const docscan_image* im, DImg* out);
// output: returns a device view to write into; ds_stage_out copies it back for host images
int ds_stage_out_begin(docscan_ctx* ctx, const docscan_image* im, DImg* out);
int ds_stage_out_end(docscan_ctx* ctx, const docscan_image* im, const DImg& dev);
int ds_finish(docscan_ctx* ctx, bool any_host);   // stream sync when host images were involved

// ---- kernel front-ends (device views, enqueue only) -----------------------------------------------
// pointwise.cu
int k_bgr2gray(docscan_ctx*, const DImg& src, const DImg& dst, int swap_rb);
int k_binary_op(docscan_ctx*, int op, const DImg& a, const DImg& b, const DImg& dst);
int k_apply_lut(docscan_ctx*, const DImg& src, const uint8_t* lut_dev, const DImg& dst);
int k_apply_lut_jobs(docscan_ctx*, const DImg* src, const DImg* dst, const uint8_t* const* luts, int n);
int k_threshold(docscan_ctx*, const DImg& src, const int32_t* t_dev, int t_imm, const DImg& dst);
int k_stats(docscan_ctx*, const DImg& src, uint32_t* minmax_dev, uint32_t* hist_dev);   // either may be null
int k_zero_u32(docscan_ctx*, uint32_t* p, int n, uint32_t value_even, uint32_t value_odd);
// scalars.cu : per-page scalar block (device)
struct PageScalars {
    uint32_t minmax[2];        // running min / max (atomics)
    uint32_t hist_a[256];
    uint32_t hist_b[256];
    uint8_t lut[256];          // normalize LUT (or composed LUT)
    int32_t cut_a, cut_b;      // raw cut-offs: mask = value >= cut
    int32_t otsu_a, otsu_b;    // Otsu thresholds (in normalised units)
    uint32_t pad[2];
};
int k_scalars_reset(docscan_ctx*, PageScalars* s, int n);
// lut[v] = normalize(min,max)(v); when `compose` the contrast-stretch LUT of the result is folded in
int k_build_norm_lut(docscan_ctx*, PageScalars* s, int n, int compose_stretch);
// hist_a / hist_b -> normalise -> Otsu -> (t - offset) -> raw cut-offs
int k_otsu_cuts(docscan_ctx*, PageScalars* s, int n, int threshold_offset, const int32_t* npix_dev);
// plain Otsu of hist_a (no normalisation): result in otsu_a
int k_otsu_plain(docscan_ctx*, PageScalars* s, int n, const int32_t* npix_dev);
// blur.cu
#define DS_EPI_BLUR 0      // dst = blur(src)
#define DS_EPI_SUB 1       // dst = sat(src - blur)
#define DS_EPI_RSUB 2      // dst = sat(blur - src)
#define DS_EPI_DIV 3       // dst = divide(src, blur, 255)
#define DS_EPI_ATHRESH 4   // box mean: dst = src - mean > -C ? 255 : 0
#define DS_EPI_AGAUSS 5    // tensor-core path only: Gaussian local mean in 16.16 fixed point, threshold with a guard band
struct BlurJob {
    const uint8_t* src; uint8_t* dst;
    int src_pitch, dst_pitch, w, h;
    uint32_t* minmax;     // may be null
    uint32_t* hist;       // may be null
};
// gaussian (REFLECT_101, 8.8 fixed point) for kind 0, box (REPLICATE, ones) for kind 1
int k_blur_jobs(docscan_ctx*, int kind, int k, int epi, int c_param, const BlurJob* jobs_host, int n,
                int max_w, int max_h);
// tcblur.cu : the Gaussian on the tensor cores (tcgen05); false = not applicable, run k_blur_jobs' own kernels
bool k_tc_blur_jobs(docscan_ctx*, int kind, int k, int epi, const BlurJob* jobs_host, int n, int* rc);
// morph.cu
struct MorphJob {
    const uint8_t* src; uint8_t* dst;
    const uint8_t* ref;   // blackhat: dst = sat(result - ref); null otherwise
    int src_pitch, dst_pitch, ref_pitch, w, h;
    uint32_t* hist;       // may be null
};
int k_morph_jobs(docscan_ctx*, int is_dilate, int kw, int kh, int ax, int ay, const MorphJob* jobs_host, int n,
                 int max_w, int max_h);
// 3x3 close (open_not_close = 0) / open (1), one iteration, as one fused pass; false when the buffers are not 16-byte aligned
bool k_morph_close3(docscan_ctx*, int open_not_close, const MorphJob* jobs_host, int n, int max_w, int max_h, int* rc);
// adaptive.cu (gaussian, fp32 ordered fma)
struct AdaptJob {
    const uint8_t* src; uint8_t* dst;
    int src_pitch, dst_pitch, w, h;
};
int k_adaptive_gauss_jobs(docscan_ctx*, int k, int c, int cv_tail_compat, const AdaptJob* jobs_host, int n,
                          int max_w, int max_h);
// tcblur.cu : GAUSSIAN_C's local mean on the tensor cores with a guard band (see there); false = not applicable
#define TC_TILE_FLAG_CAP 512      // listed pixels per 128 x 64 tile; a tile that overflows is re-evaluated whole
struct TcFlagLists {              // what the tensor-core pass leaves for the exact re-evaluation
    uint32_t* count; uint16_t* list; int n_tiles, RL, NOUT;
    struct PageTiles { int tile_base, ntx, nty; };
    std::vector<PageTiles> tiles; // per page of the launch
};
bool k_tc_adaptive_jobs(docscan_ctx*, int k, int c_param, const int32_t* w16, int band, const AdaptJob* jobs_host, int n, TcFlagLists* fl, int* rc);
// mask + blend (DocScanner.py:207-212, 338-339)
struct BlendJob {
    const uint8_t* ink_sub; const uint8_t* bh; const uint8_t* base; uint8_t* dst;
    int pitch_sub, pitch_bh, pitch_base, pitch_dst, w, h;
    const PageScalars* sc;
};
int k_mask_blend_jobs(docscan_ctx*, int dilate_iters, int write_mask_only, const BlendJob* jobs_host, int n,
                      int max_w, int max_h);
// warp.cu
struct WarpPJob {
    const uint8_t* src; uint8_t* dst; uint8_t* gray;
    int src_pitch, dst_pitch, gray_pitch, sw, sh, dw, dh, ch;
    int block_w;          // 1024 / min(16, dh): OpenCV's coordinate block width
    // Resident part of the source image [rx0, rx1) x [ry0, ry1): `src` points at pixel (rx0, ry0).  The whole image
    // (0, 0, sw, sh) unless the host pipeline uploaded only the rows / columns under the quad (capi.cu).
    int rx0, ry0, rx1, ry1;
    double m[9];          // inverse matrix (dst -> src)
};
int k_warp_perspective_jobs(docscan_ctx*, const WarpPJob* jobs_host, int n, int max_w, int max_h);
struct WarpAJob {
    const uint8_t* src; uint8_t* dst;
    int src_pitch, dst_pitch, sw, sh, dw, dh;
    double m[6];          // inverse matrix
    int identity;
    int2* delta;          // per-column fixed-point increments (filled by k_warp_affine_jobs)
};
int k_warp_affine_jobs(docscan_ctx*, const WarpAJob* jobs_host, int n, int max_w, int max_h);
// the same in two steps, so that a device kernel (deskew.cu) can fill in the matrices in between
int k_warp_affine_upload(docscan_ctx*, const WarpAJob* jobs_host, int n, WarpAJob** jobs_dev);
int k_warp_affine_launch(docscan_ctx*, const WarpAJob* jobs_dev, const WarpAJob* jobs_host, int n, int max_w, int max_h);
// deskew.cu : cv2.Canny, cv2.HoughLines(1, pi/180), the median line angle and the rotation matrix, on the device
size_t k_skew_scratch_bytes(int w, int h, bool want_list);
int k_skew_estimate(docscan_ctx*, const DImg* gray, int n, double canny_low, double canny_high, int hough_threshold,
                    double max_rotate, const DImg* edges_out, double* const* angles_dev, WarpAJob* const* rot_jobs);
int k_hough_lines(docscan_ctx*, const DImg& edges, int threshold, std::vector<uint2>* lines, int* numrho_out);
// resize.cu : cv2.resize INTER_AREA (shrink) / INTER_CUBIC
int k_resize(docscan_ctx*, const DImg& src, const DImg& dst, int interpolation, int cv_tail_compat);
// synth.cu
int k_synth_page(docscan_ctx*, uint64_t seed, const DImg& dst, float quad_out[8]);

// Segment height for the marching kernels.  A strip is cut into vertical segments, one CTA each; every segment repeats a
// warm-up of ~2r rows, hence `min_rows`.  All CTAs of a launch do the same work, so what matters is how full the waves of
// resident CTAs are: strips * segs CTAs run in ceil(strips * segs / resident) waves, and a partial last wave idles the
// rest of the machine for a whole CTA lifetime.  Pick the segment count (at most 4 waves) that fills its waves best; on
// ties the fewer, longer segments win.
static inline int ds_pick_seg_rows(int resident_ctas, int strips, int max_h, int min_rows, int quantum) {
    if (strips < 1) strips = 1;
    if (resident_ctas < 1) resident_ctas = 1;
    int best_seg = (max_h + quantum - 1) / quantum * quantum;
    double best_fill = -1.0;
    for (int segs = 1; segs <= 64; segs++) {
        int seg = (max_h + segs - 1) / segs;
        if (seg < min_rows) seg = min_rows;
        seg = (seg + quantum - 1) / quantum * quantum;
        const int actual = (max_h + seg - 1) / seg;                       // segments this height really gives
        const long long ctas = (long long)strips * actual;
        const long long waves = (ctas + resident_ctas - 1) / resident_ctas;
        if (waves > 4 && best_fill >= 0) break;
        const double fill = (double)ctas / (double)(waves * resident_ctas);
        if (fill > best_fill + 0.02) { best_fill = fill; best_seg = seg; }
        if (seg <= min_rows) break;
    }
    return best_seg;
}

// upload a host job array into arena memory (stream ordered); returns the device pointer
int ds_upload(docscan_ctx* ctx, const void* host, size_t bytes, void** dev_out);

// host math (hostmath.cpp)
void hm_invert3x3(const double S[9], double T[9]);
void hm_invert_affine(const double Min[6], double M[6]);
void hm_gaussian_kernel_f64(int k, double* c);
void hm_hough_trig_table(float c[180], float s[180]);
void hm_folded_angles(float out[180]);

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// device helpers
__device__ __forceinline__ int ds_reflect101(int p, int len) {
    if ((unsigned)p < (unsigned)len) return p;
    if (len == 1) return 0;
    do {
        if (p < 0) p = -p;
        else p = 2 * len - 2 - p;
    } while ((unsigned)p >= (unsigned)len);
    return p;
}
__device__ __forceinline__ int ds_clamp(int v, int lo, int hi) { return min(max(v, lo), hi); }
// Packed fp32 multiply / add that must not be contracted.  ptxas fuses mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 (also from
// the __fmul2_rn / __fadd2_rn intrinsics under -fmad=false; measured: the Hough vote lost exactness that way), which the
// scalar .rn forms never are.  The sum is therefore taken lane by lane with the scalar add; the products stay packed.
__device__ __forceinline__ float2 ds_mul2_rn(float2 a, float2 b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
    return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float2 ds_add2_unfused(float2 a, float2 b) { return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)); }
__device__ __forceinline__ uint32_t ds_ldg32(const void* p) { return __ldg(reinterpret_cast<const uint32_t*>(p)); }
__device__ __forceinline__ uint8_t ds_div255(uint8_t a, uint8_t b) {
    if (b == 0) return 0;
    float q = __fdiv_rn(__fmul_rn((float)a, 255.0f), (float)b);
    int v = __float2int_rn(q);
    return (uint8_t)min(max(v, 0), 255);
}
#endif
