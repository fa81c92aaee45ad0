// extern "C" entry points of libdocscan.so (see include/docscan.h) and the batched page pipeline.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <thread>

#include <immintrin.h>

#include "common.cuh"

namespace {

size_t host_bytes(const docscan_image* im) {
    return (im && im->space == DOCSCAN_HOST) ? ds_image_bytes(im->width, im->height, im->channels) : 0;
}
size_t plane_bytes(int w, int h) { return ds_image_bytes(w, h, 1); }

int begin_call(docscan_ctx* ctx, size_t scratch) {
    if (!ctx) return DOCSCAN_ERR_BAD_ARG;
    ctx->err.clear();
    DS_CUDA(ctx, cudaSetDevice(ctx->device));
    return ds_arena_reserve(ctx, scratch + (1 << 20));
}

int same_size(docscan_ctx* ctx, const docscan_image* a, const docscan_image* b, const char* what) {
    if (a->width != b->width || a->height != b->height)
        return ds_fail(ctx, DOCSCAN_ERR_BAD_ARG, "%s: size mismatch (%dx%d vs %dx%d)", what, a->width, a->height, b->width, b->height);
    return DOCSCAN_OK;
}

bool is_host(const docscan_image* im) { return im && im->space == DOCSCAN_HOST; }

int read_back(docscan_ctx* ctx, void* host_dst, const void* dev_src, size_t bytes) {
    void* pin = nullptr;
    DS_TRY(ds_pinned_alloc(ctx, bytes, &pin));
    DS_CUDA(ctx, cudaMemcpyAsync(pin, dev_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    DS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(host_dst, pin, bytes);
    return DOCSCAN_OK;
}

int new_scalars(docscan_ctx* ctx, int n, PageScalars** out) {
    void* p = nullptr;
    DS_TRY(ds_arena_alloc(ctx, sizeof(PageScalars) * n, &p));
    *out = (PageScalars*)p;
    return k_scalars_reset(ctx, *out, n);
}

int upload_npix(docscan_ctx* ctx, const std::vector<DImg>& imgs, int32_t** out) {
    std::vector<int32_t> npix(imgs.size());
    for (size_t i = 0; i < imgs.size(); i++) npix[i] = imgs[i].w * imgs[i].h;
    void* p = nullptr;
    DS_TRY(ds_upload(ctx, npix.data(), sizeof(int32_t) * npix.size(), &p));
    *out = (int32_t*)p;
    return DOCSCAN_OK;
}

void max_dims(const std::vector<DImg>& v, int* mw, int* mh) {
    *mw = 0; *mh = 0;
    for (const DImg& d : v) { *mw = std::max(*mw, d.w); *mh = std::max(*mh, d.h); }
}

// ---- batched building blocks (device views) ------------------------------------------------------------
int blur_batch(docscan_ctx* ctx, int kind, int k, int epi, int c, const std::vector<DImg>& src, const std::vector<DImg>& dst,
               PageScalars* sc, bool want_minmax, int hist_sel /*0 none, 1 hist_a, 2 hist_b*/) {
    const int n = (int)src.size();
    std::vector<BlurJob> jobs(n);
    for (int i = 0; i < n; i++) {
        BlurJob& j = jobs[i];
        j.src = src[i].p; j.src_pitch = src[i].pitch; j.dst = dst[i].p; j.dst_pitch = dst[i].pitch; j.w = src[i].w; j.h = src[i].h;
        j.minmax = (sc && want_minmax) ? sc[i].minmax : nullptr;
        j.hist = (sc && hist_sel) ? (hist_sel == 1 ? sc[i].hist_a : sc[i].hist_b) : nullptr;
    }
    int mw, mh; max_dims(src, &mw, &mh);
    return k_blur_jobs(ctx, kind, k, epi, c, jobs.data(), n, mw, mh);
}

int morph_batch(docscan_ctx* ctx, int is_dilate, int kw, int kh, int iterations, const std::vector<DImg>& src,
                const std::vector<DImg>& dst, const std::vector<DImg>* ref, PageScalars* sc, int hist_sel) {
    const int n = (int)src.size();
    // n iterations with a rectangle == one pass with the rectangle grown by (n-1)(k-1) and the anchor scaled
    const int KW = kw + (iterations - 1) * (kw - 1), KH = kh + (iterations - 1) * (kh - 1);
    const int ax = (kw / 2) * iterations, ay = (kh / 2) * iterations;
    std::vector<MorphJob> jobs(n);
    for (int i = 0; i < n; i++) {
        MorphJob& j = jobs[i];
        j.src = src[i].p; j.src_pitch = src[i].pitch; j.dst = dst[i].p; j.dst_pitch = dst[i].pitch; j.w = src[i].w; j.h = src[i].h;
        j.ref = ref ? (*ref)[i].p : nullptr; j.ref_pitch = ref ? (*ref)[i].pitch : 0;
        j.hist = (sc && hist_sel) ? (hist_sel == 1 ? sc[i].hist_a : sc[i].hist_b) : nullptr;
    }
    int mw, mh; max_dims(src, &mw, &mh);
    return k_morph_jobs(ctx, is_dilate, KW, KH, ax, ay, jobs.data(), n, mw, mh);
}

int copy_batch(docscan_ctx* ctx, const std::vector<DImg>& src, const std::vector<DImg>& dst) {
    for (size_t i = 0; i < src.size(); i++)
        DS_CUDA(ctx, cudaMemcpy2DAsync(dst[i].p, dst[i].pitch, src[i].p, src[i].pitch, (size_t)src[i].w * src[i].ch, src[i].h,
                                       cudaMemcpyDeviceToDevice, ctx->stream));
    return DOCSCAN_OK;
}

int alloc_planes(docscan_ctx* ctx, const std::vector<DImg>& like, std::vector<DImg>* out) {
    out->resize(like.size());
    for (size_t i = 0; i < like.size(); i++) DS_TRY(ds_arena_image(ctx, like[i].w, like[i].h, 1, &(*out)[i]));
    return DOCSCAN_OK;
}

// morphologyEx(CLOSE / OPEN / BLACKHAT) and erode / dilate on a batch
int morph_op_batch(docscan_ctx* ctx, int op, int kw, int kh, int iterations, const std::vector<DImg>& src,
                   const std::vector<DImg>& dst, PageScalars* sc, int hist_sel) {
    if (iterations <= 0 || (kw == 1 && kh == 1)) {
        if (op == DOCSCAN_MORPH_BLACKHAT) {     // close(src) == src -> all zeros
            for (size_t i = 0; i < dst.size(); i++)
                DS_CUDA(ctx, cudaMemset2DAsync(dst[i].p, dst[i].pitch, 0, dst[i].w, dst[i].h, ctx->stream));
            if (sc && hist_sel) return ds_fail(ctx, DOCSCAN_ERR_UNSUPPORTED, "black-hat histogram with an empty element");
            return DOCSCAN_OK;
        }
        return copy_batch(ctx, src, dst);
    }
    switch (op) {
        case DOCSCAN_MORPH_ERODE: return morph_batch(ctx, 0, kw, kh, iterations, src, dst, nullptr, sc, hist_sel);
        case DOCSCAN_MORPH_DILATE: return morph_batch(ctx, 1, kw, kh, iterations, src, dst, nullptr, sc, hist_sel);
        case DOCSCAN_MORPH_CLOSE:
        case DOCSCAN_MORPH_BLACKHAT:
        case DOCSCAN_MORPH_OPEN: {
            if (kw == 3 && kh == 3 && iterations == 1 && op != DOCSCAN_MORPH_BLACKHAT && !(sc && hist_sel)) {
                // morph_cleanup's default: one fused register-only pass instead of two marching passes
                const int n = (int)src.size();
                std::vector<MorphJob> jobs(n);
                for (int i = 0; i < n; i++) {
                    MorphJob& j = jobs[i];
                    j = MorphJob{};
                    j.src = src[i].p; j.src_pitch = src[i].pitch; j.dst = dst[i].p; j.dst_pitch = dst[i].pitch; j.w = src[i].w; j.h = src[i].h;
                }
                int mw, mh; max_dims(src, &mw, &mh);
                int rc = DOCSCAN_OK;
                if (k_morph_close3(ctx, op == DOCSCAN_MORPH_OPEN, jobs.data(), n, mw, mh, &rc)) return rc;
            }
            std::vector<DImg> mid;
            DS_TRY(alloc_planes(ctx, src, &mid));
            const int first = op == DOCSCAN_MORPH_OPEN ? 0 : 1;
            DS_TRY(morph_batch(ctx, first, kw, kh, iterations, src, mid, nullptr, nullptr, 0));
            return morph_batch(ctx, 1 - first, kw, kh, iterations, mid, dst, op == DOCSCAN_MORPH_BLACKHAT ? &src : nullptr, sc, hist_sel);
        }
        default: return ds_fail(ctx, DOCSCAN_ERR_BAD_ARG, "bad morphology op %d", op);
    }
}

int adaptive_batch(docscan_ctx* ctx, int method, int k, int c, int tail_compat, const std::vector<DImg>& src,
                   const std::vector<DImg>& dst) {
    if (k % 2 == 0 || k < 3) return ds_fail(ctx, DOCSCAN_ERR_BAD_ARG, "adaptive block size must be odd and >= 3 (got %d)", k);
    const int n = (int)src.size();
    int mw, mh; max_dims(src, &mw, &mh);
    if (method == DOCSCAN_ADAPTIVE_MEAN) return blur_batch(ctx, 1, k, DS_EPI_ATHRESH, c, src, dst, nullptr, false, 0);
    std::vector<AdaptJob> jobs(n);
    for (int i = 0; i < n; i++) {
        AdaptJob& j = jobs[i];
        j.src = src[i].p; j.src_pitch = src[i].pitch; j.dst = dst[i].p; j.dst_pitch = dst[i].pitch; j.w = src[i].w; j.h = src[i].h;
    }
    return k_adaptive_gauss_jobs(ctx, k, c, tail_compat, jobs.data(), n, mw, mh);
}

// _compute_ink_mask up to the raw cut-offs (DocScanner.py:182-204): ink_sub, bh planes + scalars
int ink_branches(docscan_ctx* ctx, const std::vector<DImg>& gray, int blur_k, int kw, int kh, int offset, PageScalars* sc,
                 std::vector<DImg>* ink_sub, std::vector<DImg>* bh) {
    DS_TRY(alloc_planes(ctx, gray, ink_sub));
    DS_TRY(alloc_planes(ctx, gray, bh));
    DS_TRY(blur_batch(ctx, 0, blur_k, DS_EPI_RSUB, 0, gray, *ink_sub, sc, false, 1));
    DS_TRY(morph_op_batch(ctx, DOCSCAN_MORPH_BLACKHAT, kw, kh, 1, gray, *bh, sc, 2));
    int32_t* npix = nullptr;
    DS_TRY(upload_npix(ctx, gray, &npix));
    return k_otsu_cuts(ctx, sc, (int)gray.size(), offset, npix);
}

int blend_batch(docscan_ctx* ctx, int dilate_iters, int mask_only, const std::vector<DImg>& ink_sub, const std::vector<DImg>& bh,
                const std::vector<DImg>* base, const std::vector<DImg>& dst, PageScalars* sc) {
    const int n = (int)dst.size();
    std::vector<BlendJob> jobs(n);
    for (int i = 0; i < n; i++) {
        BlendJob& j = jobs[i];
        j.ink_sub = ink_sub[i].p; j.pitch_sub = ink_sub[i].pitch; j.bh = bh[i].p; j.pitch_bh = bh[i].pitch;
        j.base = base ? (*base)[i].p : nullptr; j.pitch_base = base ? (*base)[i].pitch : 0;
        j.dst = dst[i].p; j.pitch_dst = dst[i].pitch; j.w = dst[i].w; j.h = dst[i].h; j.sc = sc + i;
    }
    int mw, mh; max_dims(dst, &mw, &mh);
    return k_mask_blend_jobs(ctx, dilate_iters, mask_only, jobs.data(), n, mw, mh);
}

int illum_ksize(int w, int h, double frac) {   // DocScanner.py:150-152 (Python round = half to even)
    int base = std::max(15, (int)std::nearbyint((double)std::min(h, w) * frac));
    if (base % 2 == 0) base += 1;
    return base;
}

}  // namespace

// =====================================================================================================
// single-op entry points
// =====================================================================================================
#define DS_ARGS2(src, dst, chs, chd, name)                      \
    if (!ctx) return DOCSCAN_ERR_BAD_ARG;                       \
    DS_TRY(ds_check_image(ctx, src, chs, name ": src"));        \
    DS_TRY(ds_check_image(ctx, dst, chd, name ": dst"));

extern "C" int docscan_bgr2gray(docscan_ctx* ctx, const docscan_image* src, docscan_image* dst, int swap_rb) {
    DS_ARGS2(src, dst, 3, 1, "bgr2gray");
    DS_TRY(same_size(ctx, src, dst, "bgr2gray"));
    DS_TRY(begin_call(ctx, host_bytes(src) + host_bytes(dst)));
    ArenaScope scope(ctx);
    DImg s, d;
    DS_TRY(ds_stage_in(ctx, src, &s));
    DS_TRY(ds_stage_out_begin(ctx, dst, &d));
    DS_TRY(k_bgr2gray(ctx, s, d, swap_rb));
    DS_TRY(ds_stage_out_end(ctx, dst, d));
    return ds_finish(ctx, is_host(src) || is_host(dst));
}

extern "C" int docscan_gaussian_blur(docscan_ctx* ctx, const docscan_image* src, int k, docscan_image* dst) {
    DS_ARGS2(src, dst, 1, 1, "gaussian_blur");
    DS_TRY(same_size(ctx, src, dst, "gaussian_blur"));
    DS_TRY(begin_call(ctx, host_bytes(src) + host_bytes(dst)));
    ArenaScope scope(ctx);
    std::vector<DImg> s(1), d(1);
    DS_TRY(ds_stage_in(ctx, src, &s[0]));
    DS_TRY(ds_stage_out_begin(ctx, dst, &d[0]));
    DS_TRY(blur_batch(ctx, 0, k, DS_EPI_BLUR, 0, s, d, nullptr, false, 0));
    DS_TRY(ds_stage_out_end(ctx, dst, d[0]));
    return ds_finish(ctx, is_host(src) || is_host(dst));
}

extern "C" int docscan_binary_op(docscan_ctx* ctx, int op, const docscan_image* a, const docscan_image* b, docscan_image* dst) {
    if (!ctx) return DOCSCAN_ERR_BAD_ARG;
    if (op < DOCSCAN_OP_SUB || op > DOCSCAN_OP_MASK_SELECT) return ds_fail(ctx, DOCSCAN_ERR_BAD_ARG, "bad binary op %d", op);
    DS_TRY(ds_check_image(ctx, a, 1, "binary_op: a"));
    DS_TRY(ds_check_image(ctx, b, 1, "binary_op: b"));
    DS_TRY(ds_check_image(ctx, dst, 1, "binary_op: dst"));
    DS_TRY(same_size(ctx, a, b, "binary_op"));
    DS_TRY(same_size(ctx, a, dst, "binary_op"));
    DS_TRY(begin_call(ctx, host_bytes(a) + host_bytes(b) + host_bytes(dst)));
    ArenaScope scope(ctx);
    DImg da, db, dd;
    DS_TRY(ds_stage_in(ctx, a, &da));
    DS_TRY(ds_stage_in(ctx, b, &db));
    DS_TRY(ds_stage_out_begin(ctx, dst, &dd));
    DS_TRY(k_binary_op(ctx, op, da, db, dd));
    DS_TRY(ds_stage_out_end(ctx, dst, dd));
    return ds_finish(ctx, is_host(a) || is_host(b) || is_host(dst));
}

extern "C" int docscan_minmax(docscan_ctx* ctx, const docscan_image* src, int32_t* mn, int32_t* mx) {
    if (!ctx || !mn || !mx) return DOCSCAN_ERR_BAD_ARG;
    DS_TRY(ds_check_image(ctx, src, 1, "minmax: src"));
    DS_TRY(begin_call(ctx, host_bytes(src)));
    ArenaScope scope(ctx);
    DImg s;
    DS_TRY(ds_stage_in(ctx, src, &s));
    PageScalars* sc = nullptr;
    DS_TRY(new_scalars(ctx, 1, &sc));
    DS_TRY(k_stats(ctx, s, sc->minmax, nullptr));
    uint32_t res[2];
    DS_TRY(read_back(ctx, res, sc->minmax, sizeof(res)));
    *mn = (int32_t)res[0]; *mx = (int32_t)res[1];
    return DOCSCAN_OK;
}

extern "C" int docscan_hist256(docscan_ctx* ctx, const docscan_image* src, int32_t hist[256]) {
    if (!ctx || !hist) return DOCSCAN_ERR_BAD_ARG;
    DS_TRY(ds_check_image(ctx, src, 1, "hist256: src"));
    DS_TRY(begin_call(ctx, host_bytes(src)));
    ArenaScope scope(ctx);
    DImg s;
    DS_TRY(ds_stage_in(ctx, src, &s));
    PageScalars* sc = nullptr;
    DS_TRY(new_scalars(ctx, 1, &sc));
    DS_TRY(k_stats(ctx, s, nullptr, sc->hist_a));
    return read_back(ctx, hist, sc->hist_a, 256 * sizeof(int32_t));
}

extern "C" int docscan_normalize_minmax(docscan_ctx* ctx, const docscan_image* src, docscan_image* dst) {
    DS_ARGS2(src, dst, 1, 1, "normalize_minmax");
    DS_TRY(same_size(ctx, src, dst, "normalize_minmax"));
    DS_TRY(begin_call(ctx, host_bytes(src) + host_bytes(dst)));
    ArenaScope scope(ctx);
    DImg s, d;
    DS_TRY(ds_stage_in(ctx, src, &s));
    DS_TRY(ds_stage_out_begin(ctx, dst, &d));
    PageScalars* sc = nullptr;
    DS_TRY(new_scalars(ctx, 1, &sc));
    DS_TRY(k_stats(ctx, s, sc->minmax, nullptr));
    DS_TRY(k_build_norm_lut(ctx, sc, 1, 0));
    DS_TRY(k_apply_lut(ctx, s, sc->lut, d));
    DS_TRY(ds_stage_out_end(ctx, dst, d));
    return ds_finish(ctx, is_host(src) || is_host(dst));
}

extern "C" int docscan_otsu_threshold(docscan_ctx* ctx, const docscan_image* src, double* t, docscan_image* dst) {
    if (!ctx || !t) return DOCSCAN_ERR_BAD_ARG;
    DS_TRY(ds_check_image(ctx, src, 1, "otsu_threshold: src"));
    if (dst) {
        DS_TRY(ds_check_image(ctx, dst, 1, "otsu_threshold: dst"));
        DS_TRY(same_size(ctx, src, dst, "otsu_threshold"));
    }
    DS_TRY(begin_call(ctx, host_bytes(src) + host_bytes(dst)));
    ArenaScope scope(ctx);
    std::vector<DImg> s(1);
    DImg d{};
    DS_TRY(ds_stage_in(ctx, src, &s[0]));
    if (dst) DS_TRY(ds_stage_out_begin(ctx, dst, &d));
    PageScalars* sc = nullptr;
    DS_TRY(new_scalars(ctx, 1, &sc));
    DS_TRY(k_stats(ctx, s[0], nullptr, sc->hist_a));
    int32_t* npix = nullptr;
    DS_TRY(upload_npix(ctx, s, &npix));
    DS_TRY(k_otsu_plain(ctx, sc, 1, npix));
    if (dst) {
        DS_TRY(k_threshold(ctx, s[0], &sc->otsu_a, 0, d));
        DS_TRY(ds_stage_out_end(ctx, dst, d));
    }
    int32_t ti = 0;
    DS_TRY(read_back(ctx, &ti, &sc->otsu_a, sizeof(ti)));
    *t = (double)ti;
    return DOCSCAN_OK;
}

extern "C" int docscan_threshold_binary(docscan_ctx* ctx, const docscan_image* src, int t, docscan_image* dst) {
    DS_ARGS2(src, dst, 1, 1, "threshold_binary");
    DS_TRY(same_size(ctx, src, dst, "threshold_binary"));
    DS_TRY(begin_call(ctx, host_bytes(src) + host_bytes(dst)));
    ArenaScope scope(ctx);
    DImg s, d;
    DS_TRY(ds_stage_in(ctx, src, &s));
    DS_TRY(ds_stage_out_begin(ctx, dst, &d));
    DS_TRY(k_threshold(ctx, s, nullptr, t, d));
    DS_TRY(ds_stage_out_end(ctx, dst, d));
    return ds_finish(ctx, is_host(src) || is_host(dst));
}

extern "C" int docscan_morph_rect(docscan_ctx* ctx, int op, const docscan_image* src, int kw, int kh, int iterations,
                                  docscan_image* dst) {
    DS_ARGS2(src, dst, 1, 1, "morph_rect");
    DS_TRY(same_size(ctx, src, dst, "morph_rect"));
    if (kw < 1 || kh < 1) return ds_fail(ctx, DOCSCAN_ERR_BAD_ARG, "morph_rect: bad element %dx%d", kw, kh);
    DS_TRY(begin_call(ctx, host_bytes(src) + host_bytes(dst) + 4 * plane_bytes(src->width, src->height)));
    ArenaScope scope(ctx);
    std::vector<DImg> s(1), d(1);
    DS_TRY(ds_stage_in(ctx, src, &s[0]));
    DS_TRY(ds_stage_out_begin(ctx, dst, &d[0]));
    DS_TRY(morph_op_batch(ctx, op, kw, kh, iterations, s, d, nullptr, 0));
    DS_TRY(ds_stage_out_end(ctx, dst, d[0]));
    return ds_finish(ctx, is_host(src) || is_host(dst));
}

extern "C" int docscan_adaptive_threshold(docscan_ctx* ctx, const docscan_image* src, int method, int k, int c,
                                          int cv_tail_compat, docscan_image* dst) {
    DS_ARGS2(src, dst, 1, 1, "adaptive_threshold");
    DS_TRY(same_size(ctx, src, dst, "adaptive_threshold"));
    DS_TRY(begin_call(ctx, host_bytes(src) + host_bytes(dst) + plane_bytes(src->width, src->height)));   // + the guard-band pixel list
    ArenaScope scope(ctx);
    std::vector<DImg> s(1), d(1);
    DS_TRY(ds_stage_in(ctx, src, &s[0]));
    DS_TRY(ds_stage_out_begin(ctx, dst, &d[0]));
    DS_TRY(adaptive_batch(ctx, method, k, c, cv_tail_compat, s, d));
    DS_TRY(ds_stage_out_end(ctx, dst, d[0]));
    return ds_finish(ctx, is_host(src) || is_host(dst));
}

extern "C" int docscan_warp_perspective(docscan_ctx* ctx, const docscan_image* src, const double m_fwd[9],
                                        docscan_image* dst, docscan_image* gray_out) {
    if (!ctx || !m_fwd) return DOCSCAN_ERR_BAD_ARG;
    DS_TRY(ds_check_image(ctx, src, 0, "warp_perspective: src"));
    DS_TRY(ds_check_image(ctx, dst, src->channels, "warp_perspective: dst"));
    if (gray_out) {
        if (src->channels != 3) return ds_fail(ctx, DOCSCAN_ERR_BAD_ARG, "warp_perspective: gray_out needs a 3-channel source");
        DS_TRY(ds_check_image(ctx, gray_out, 1, "warp_perspective: gray_out"));
        DS_TRY(same_size(ctx, dst, gray_out, "warp_perspective"));
    }
    DS_TRY(begin_call(ctx, host_bytes(src) + host_bytes(dst) + host_bytes(gray_out)));
    ArenaScope scope(ctx);
    DImg s, d, g{};
    DS_TRY(ds_stage_in(ctx, src, &s));
    DS_TRY(ds_stage_out_begin(ctx, dst, &d));
    if (gray_out) DS_TRY(ds_stage_out_begin(ctx, gray_out, &g));
    WarpPJob j{};
    j.src = s.p; j.src_pitch = s.pitch; j.sw = s.w; j.sh = s.h; j.ch = s.ch;
    j.dst = d.p; j.dst_pitch = d.pitch; j.dw = d.w; j.dh = d.h;
    j.gray = gray_out ? g.p : nullptr; j.gray_pitch = g.pitch;
    j.block_w = 1024 / std::min(16, d.h);
    j.rx0 = 0; j.ry0 = 0; j.rx1 = s.w; j.ry1 = s.h;
    hm_invert3x3(m_fwd, j.m);
    DS_TRY(k_warp_perspective_jobs(ctx, &j, 1, d.w, d.h));
    DS_TRY(ds_stage_out_end(ctx, dst, d));
    if (gray_out) DS_TRY(ds_stage_out_end(ctx, gray_out, g));
    return ds_finish(ctx, is_host(src) || is_host(dst) || is_host(gray_out));
}

extern "C" int docscan_warp_affine(docscan_ctx* ctx, const docscan_image* src, const double m_fwd[6], docscan_image* dst) {
    if (!m_fwd) return DOCSCAN_ERR_BAD_ARG;
    DS_ARGS2(src, dst, 1, 1, "warp_affine");
    DS_TRY(begin_call(ctx, host_bytes(src) + host_bytes(dst)));
    ArenaScope scope(ctx);
    DImg s, d;
    DS_TRY(ds_stage_in(ctx, src, &s));
    DS_TRY(ds_stage_out_begin(ctx, dst, &d));
    WarpAJob j{};
    j.src = s.p; j.src_pitch = s.pitch; j.sw = s.w; j.sh = s.h;
    j.dst = d.p; j.dst_pitch = d.pitch; j.dw = d.w; j.dh = d.h;
    hm_invert_affine(m_fwd, j.m);
    DS_TRY(k_warp_affine_jobs(ctx, &j, 1, d.w, d.h));
    DS_TRY(ds_stage_out_end(ctx, dst, d));
    return ds_finish(ctx, is_host(src) || is_host(dst));
}

extern "C" int docscan_canny(docscan_ctx* ctx, const docscan_image* src, double low, double high, docscan_image* dst) {
    DS_ARGS2(src, dst, 1, 1, "canny");
    DS_TRY(same_size(ctx, src, dst, "canny"));
    DS_TRY(begin_call(ctx, host_bytes(src) + host_bytes(dst) + k_skew_scratch_bytes(src->width, src->height, false)));
    ArenaScope scope(ctx);
    DImg s, d;
    DS_TRY(ds_stage_in(ctx, src, &s));
    DS_TRY(ds_stage_out_begin(ctx, dst, &d));
    DS_TRY(k_skew_estimate(ctx, &s, 1, low, high, 0, 0.0, &d, nullptr, nullptr));
    DS_TRY(ds_stage_out_end(ctx, dst, d));
    return ds_finish(ctx, is_host(src) || is_host(dst));
}

extern "C" int docscan_hough_lines(docscan_ctx* ctx, const docscan_image* edges, int threshold, float* rho_theta, int max_lines,
                                   int32_t* n_lines, int32_t per_angle[180]) {
    if (!ctx || !n_lines || max_lines < 0 || (max_lines && !rho_theta)) return DOCSCAN_ERR_BAD_ARG;
    DS_TRY(ds_check_image(ctx, edges, 1, "hough_lines: edges"));
    const size_t numrho = 2 * ((size_t)edges->width + edges->height) + 1;
    DS_TRY(begin_call(ctx, host_bytes(edges) + 4 * (size_t)edges->width * edges->height + 182 * (numrho + 2) * 4 + 180 * numrho * 8 + 8192));
    ArenaScope scope(ctx);
    DImg e;
    DS_TRY(ds_stage_in(ctx, edges, &e));
    std::vector<uint2> lines;
    int nr = 0;
    DS_TRY(k_hough_lines(ctx, e, threshold, &lines, &nr));
    // cv::HoughLinesStandard sorts by votes (descending), ties by accumulator index (ascending)
    std::sort(lines.begin(), lines.end(), [](const uint2& a, const uint2& b) { return a.y > b.y || (a.y == b.y && a.x < b.x); });
    if (per_angle) for (int n = 0; n < 180; n++) per_angle[n] = 0;
    const float theta = (float)(3.1415926535897932384626433832795 / 180);
    for (size_t k = 0; k < lines.size(); k++) {
        const int n = (int)lines[k].x / (nr + 2) - 1, r = (int)lines[k].x - (n + 1) * (nr + 2) - 1;
        if (per_angle) per_angle[n]++;
        if ((int)k < max_lines) {
            rho_theta[2 * k] = ((float)r - (float)(nr - 1) * 0.5f) * 1.0f;
            rho_theta[2 * k + 1] = 0.f + (float)n * theta;
        }
    }
    *n_lines = (int32_t)lines.size();
    return DOCSCAN_OK;
}

extern "C" int docscan_skew_angle(docscan_ctx* ctx, const docscan_image* gray, double canny_low, double canny_high,
                                  double max_rotate, double* angle_deg) {
    if (!ctx || !angle_deg) return DOCSCAN_ERR_BAD_ARG;
    DS_TRY(ds_check_image(ctx, gray, 1, "skew_angle: gray"));
    DS_TRY(begin_call(ctx, host_bytes(gray) + k_skew_scratch_bytes(gray->width, gray->height, true)));
    ArenaScope scope(ctx);
    DImg g;
    DS_TRY(ds_stage_in(ctx, gray, &g));
    void* a = nullptr;
    DS_TRY(ds_arena_alloc(ctx, sizeof(double), &a));
    double* ad = (double*)a;
    DS_TRY(k_skew_estimate(ctx, &g, 1, canny_low, canny_high, 150, max_rotate, nullptr, &ad, nullptr));
    return read_back(ctx, angle_deg, ad, sizeof(double));
}

extern "C" int docscan_last_angles(docscan_ctx* ctx, double* angles, int n) {
    if (!ctx || !angles || n < 0) return DOCSCAN_ERR_BAD_ARG;
    if (n > ctx->angles_n) return ds_fail(ctx, DOCSCAN_ERR_BAD_ARG, "last_angles: the last batch had %d pages", ctx->angles_n);
    DS_CUDA(ctx, cudaSetDevice(ctx->device));
    DS_CUDA(ctx, cudaMemcpyAsync(angles, ctx->angles_dev, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    DS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DOCSCAN_OK;
}

extern "C" int docscan_resize(docscan_ctx* ctx, const docscan_image* src, docscan_image* dst, int interpolation, int cv_tail_compat) {
    if (!ctx) return DOCSCAN_ERR_BAD_ARG;
    DS_TRY(ds_check_image(ctx, src, 0, "resize: src"));
    DS_TRY(ds_check_image(ctx, dst, src->channels, "resize: dst"));
    DS_TRY(begin_call(ctx, host_bytes(src) + host_bytes(dst) + 64 * ((size_t)src->width + src->height + dst->width + dst->height) + 4096));
    ArenaScope scope(ctx);
    DImg s, d;
    DS_TRY(ds_stage_in(ctx, src, &s));
    DS_TRY(ds_stage_out_begin(ctx, dst, &d));
    DS_TRY(k_resize(ctx, s, d, interpolation, cv_tail_compat));
    DS_TRY(ds_stage_out_end(ctx, dst, d));
    return ds_finish(ctx, is_host(src) || is_host(dst));
}

// =====================================================================================================
// fused reference stage functions
// =====================================================================================================
extern "C" int docscan_illumination_correction(docscan_ctx* ctx, const docscan_image* gray, int method, int k, docscan_image* dst) {
    DS_ARGS2(gray, dst, 1, 1, "illumination_correction");
    DS_TRY(same_size(ctx, gray, dst, "illumination_correction"));
    DS_TRY(begin_call(ctx, host_bytes(gray) + host_bytes(dst) + plane_bytes(gray->width, gray->height)));
    ArenaScope scope(ctx);
    std::vector<DImg> s(1), tmp, d(1);
    DS_TRY(ds_stage_in(ctx, gray, &s[0]));
    DS_TRY(ds_stage_out_begin(ctx, dst, &d[0]));
    DS_TRY(alloc_planes(ctx, s, &tmp));
    PageScalars* sc = nullptr;
    DS_TRY(new_scalars(ctx, 1, &sc));
    DS_TRY(blur_batch(ctx, 0, k, method == 1 ? DS_EPI_DIV : DS_EPI_SUB, 0, s, tmp, sc, true, 0));
    DS_TRY(k_build_norm_lut(ctx, sc, 1, 0));
    DS_TRY(k_apply_lut(ctx, tmp[0], sc->lut, d[0]));
    DS_TRY(ds_stage_out_end(ctx, dst, d[0]));
    return ds_finish(ctx, is_host(gray) || is_host(dst));
}

extern "C" int docscan_ink_mask(docscan_ctx* ctx, const docscan_image* gray, int mask_blur_ksize, int kw_bh, int kh_bh,
                                int dilate_iters, int threshold_offset, docscan_image* dst) {
    DS_ARGS2(gray, dst, 1, 1, "ink_mask");
    DS_TRY(same_size(ctx, gray, dst, "ink_mask"));
    DS_TRY(begin_call(ctx, host_bytes(gray) + host_bytes(dst) + 6 * plane_bytes(gray->width, gray->height)));
    ArenaScope scope(ctx);
    std::vector<DImg> s(1), d(1), ink_sub, bh;
    DS_TRY(ds_stage_in(ctx, gray, &s[0]));
    DS_TRY(ds_stage_out_begin(ctx, dst, &d[0]));
    PageScalars* sc = nullptr;
    DS_TRY(new_scalars(ctx, 1, &sc));
    DS_TRY(ink_branches(ctx, s, mask_blur_ksize, kw_bh, kh_bh, threshold_offset, sc, &ink_sub, &bh));
    DS_TRY(blend_batch(ctx, std::max(dilate_iters, 0), 1, ink_sub, bh, nullptr, d, sc));
    DS_TRY(ds_stage_out_end(ctx, dst, d[0]));
    return ds_finish(ctx, is_host(gray) || is_host(dst));
}

// =====================================================================================================
// the whole per-pixel path for a batch of pages
// =====================================================================================================
namespace {

size_t page_scratch(const docscan_page& pg) {
    const int w = pg.binary.width, h = pg.binary.height;
    if (std::isnan(pg.angle_deg)) return 17 * plane_bytes(w, h) + 8192 + k_skew_scratch_bytes(w, h, true) + 4096 +
                                         (pg.use_whole ? 64 * ((size_t)pg.src.width + pg.src.height + w + h) + 4096 : 0);
    size_t tables = pg.use_whole ? 64 * ((size_t)pg.src.width + pg.src.height + w + h) + 4096 : 0;    // resize tables
    return 17 * plane_bytes(w, h) + 65536 + tables;
}

size_t page_staging(const docscan_page& pg) {
    return host_bytes(&pg.src) + host_bytes(&pg.warped) + host_bytes(&pg.binary);
}

// The chain for n pages whose images are already device views (src read-only; warped, binary written).
// Resident region of a page photo: the host pipeline uploads only the rows / columns the warp can touch.
struct SrcRegion { int x0, y0, x1, y1; };

// Source pixels the perspective warp of `pg` can read, as a box in the photo (with a margin for the rounding to
// 1/32 px, the second bilinear tap and the 16-byte load window of the kernel).  The destination rectangle is convex and
// a projective map whose denominator keeps one sign over it maps it onto the convex quadrilateral spanned by the images
// of its corners; anything else (denominator changing sign, non-finite corners) keeps the whole photo.
SrcRegion warp_footprint(const docscan_page& pg) {
    const int sw = pg.src.width, sh = pg.src.height, dw = pg.warped.width, dh = pg.warped.height;
    const SrcRegion whole{0, 0, sw, sh};
    const float dstq[8] = {0, 0, (float)(dw - 1), 0, (float)(dw - 1), (float)(dh - 1), 0, (float)(dh - 1)};
    double m[9], inv[9];
    if (docscan_get_perspective_transform(pg.quad, dstq, m) != DOCSCAN_OK) return whole;
    hm_invert3x3(m, inv);
    double lo_x = 1e300, hi_x = -1e300, lo_y = 1e300, hi_y = -1e300, w_min = 1e300, w_max = -1e300;
    for (int c = 0; c < 4; c++) {
        const double x = dstq[2 * c], y = dstq[2 * c + 1];
        const double w = inv[6] * x + inv[7] * y + inv[8];
        const double X = (inv[0] * x + inv[1] * y + inv[2]) / w, Y = (inv[3] * x + inv[4] * y + inv[5]) / w;
        if (!std::isfinite(w) || !std::isfinite(X) || !std::isfinite(Y)) return whole;
        w_min = std::min(w_min, w); w_max = std::max(w_max, w);
        lo_x = std::min(lo_x, X); hi_x = std::max(hi_x, X); lo_y = std::min(lo_y, Y); hi_y = std::max(hi_y, Y);
    }
    if (!(w_min > 0.0 || w_max < 0.0)) return whole;                                   // denominator changes sign
    if (std::min(std::fabs(w_min), std::fabs(w_max)) < 1e-9 * std::max(std::fabs(w_min), std::fabs(w_max))) return whole;
    if (hi_x - lo_x > 1e6 || hi_y - lo_y > 1e6) return whole;
    SrcRegion r;
    r.x0 = std::max(0, (int)std::floor(lo_x) - 8); r.x1 = std::min(sw, (int)std::ceil(hi_x) + 10);
    r.y0 = std::max(0, (int)std::floor(lo_y) - 2); r.y1 = std::min(sh, (int)std::ceil(hi_y) + 4);
    if (r.x1 - r.x0 < 16 || r.y1 - r.y0 < 2) return whole;
    return r;
}

// `regions` (may be null = whole photos): src[i] then views only that part of page i's photo.
int run_group(docscan_ctx* ctx, int n, docscan_page* pages, const docscan_params& P, const std::vector<DImg>& src,
              const std::vector<DImg>& warped, const std::vector<DImg>& binary, int page0, const SrcRegion* regions = nullptr) {
    ArenaScope scope(ctx);
    std::vector<DImg> gray;
    std::vector<WarpPJob> wj(n);
    DS_TRY(alloc_planes(ctx, binary, &gray));
    int mw, mh; max_dims(binary, &mw, &mh);
    // a1 + a2: perspective warp with fused BGR2GRAY; pages without a usable quad take resize_long_side + BGR2GRAY
    std::vector<WarpPJob> wj_quad;
    int qw = 0, qh = 0;
    for (int i = 0; i < n; i++) {
        if (pages[i].use_whole) {
            if (regions && (regions[i].x0 != 0 || regions[i].y0 != 0 || regions[i].x1 != pages[i].src.width || regions[i].y1 != pages[i].src.height))
                return ds_fail(ctx, DOCSCAN_ERR_BAD_ARG, "internal: whole-photo page with a partial upload");
            const int long_src = std::max(src[i].w, src[i].h), long_dst = std::max(warped[i].w, warped[i].h);
            DS_TRY(k_resize(ctx, src[i], warped[i], long_dst < long_src ? DOCSCAN_INTER_AREA : DOCSCAN_INTER_CUBIC, P.cv_tail_compat));
            DS_TRY(k_bgr2gray(ctx, warped[i], gray[i], 0));
            continue;
        }
        WarpPJob& j = wj[i];
        j = WarpPJob{};
        j.src = src[i].p; j.src_pitch = src[i].pitch; j.sw = pages[i].src.width; j.sh = pages[i].src.height; j.ch = 3;
        if (regions) { j.rx0 = regions[i].x0; j.ry0 = regions[i].y0; j.rx1 = regions[i].x1; j.ry1 = regions[i].y1; }
        else { j.rx0 = 0; j.ry0 = 0; j.rx1 = j.sw; j.ry1 = j.sh; }
        j.dst = warped[i].p; j.dst_pitch = warped[i].pitch; j.dw = warped[i].w; j.dh = warped[i].h;
        j.gray = gray[i].p; j.gray_pitch = gray[i].pitch;
        j.block_w = 1024 / std::min(16, warped[i].h);
        const float dstq[8] = {0, 0, (float)(warped[i].w - 1), 0, (float)(warped[i].w - 1), (float)(warped[i].h - 1), 0, (float)(warped[i].h - 1)};
        double m[9];
        DS_TRY(docscan_get_perspective_transform(pages[i].quad, dstq, m));
        hm_invert3x3(m, j.m);
        wj_quad.push_back(j);
        qw = std::max(qw, j.dw); qh = std::max(qh, j.dh);
    }
    if (!wj_quad.empty()) DS_TRY(k_warp_perspective_jobs(ctx, wj_quad.data(), (int)wj_quad.size(), qw, qh));

    // a3 + a4: illumination correction; its MINMAX LUT and contrast_stretch's fold into one LUT
    PageScalars* sc = nullptr;
    DS_TRY(new_scalars(ctx, n, &sc));
    std::vector<DImg> stretched;
    DS_TRY(alloc_planes(ctx, binary, &stretched));
    // blur kernel size depends on the page size (DocScanner.py:150): group pages by k
    {
        std::vector<int> ks(n);
        for (int i = 0; i < n; i++) ks[i] = illum_ksize(gray[i].w, gray[i].h, P.illum_blur_frac);
        std::vector<char> done(n, 0);
        for (int i = 0; i < n; i++) {
            if (done[i]) continue;
            std::vector<DImg> gs, gd;
            std::vector<PageScalars*> idx;
            // pages with the same k must be contiguous in the scalar array: launch per run of equal k
            int e = i;
            while (e < n && ks[e] == ks[i]) { gs.push_back(gray[e]); gd.push_back(stretched[e]); done[e] = 1; e++; }
            DS_TRY(blur_batch(ctx, 0, ks[i], P.illum_method == 1 ? DS_EPI_DIV : DS_EPI_SUB, 0, gs, gd, sc + i, true, 0));
        }
    }
    DS_TRY(k_build_norm_lut(ctx, sc, n, 1));
    {
        std::vector<const uint8_t*> luts(n);
        for (int i = 0; i < n; i++) luts[i] = sc[i].lut;
        DS_TRY(k_apply_lut_jobs(ctx, stretched.data(), stretched.data(), luts.data(), n));
    }
    // a5: ink mask branches -> raw cut-offs
    int blur_k = P.mask_blur_ksize;
    if (blur_k % 2 == 0) blur_k += 1;
    int bk = P.blackhat_ksize;
    if (bk < 3) bk = 3;
    if (bk % 2 == 0) bk += 1;
    int bh_h = std::max(3, (int)std::nearbyint((double)bk * P.blackhat_vertical_ratio));
    if (bh_h % 2 == 0) bh_h += 1;
    std::vector<DImg> ink_sub, bh;
    DS_TRY(ink_branches(ctx, stretched, blur_k, bk, bh_h, P.mask_thresh_offset, sc, &ink_sub, &bh));
    // a6: adaptive threshold
    int block = P.block_size;
    if (block % 2 == 0) block += 1;
    std::vector<DImg> base;
    DS_TRY(alloc_planes(ctx, binary, &base));
    DS_TRY(adaptive_batch(ctx, P.thresh_method, block, P.C, P.cv_tail_compat, stretched, base));
    // what follows: rotate (a8) unless every angle is 0, close (a9) unless ksize <= 1
    const bool do_close = P.morph_ksize > 1;
    std::vector<DImg> blend_dst, rot_dst;
    if (do_close) {
        DS_TRY(alloc_planes(ctx, binary, &blend_dst));
        DS_TRY(alloc_planes(ctx, binary, &rot_dst));
    } else {
        DS_TRY(alloc_planes(ctx, binary, &blend_dst));
        rot_dst = binary;
    }
    // a7: mask combine + 2x2 dilate + masked blend
    DS_TRY(blend_batch(ctx, std::max(P.ink_dilate_iters, 0), 0, ink_sub, bh, &base, blend_dst, sc));
    // a8: deskew.  Pages without a supplied angle get deskew()'s own estimate (Canny + HoughLines median on the blended
    // page, DocScanner.py:218-231) on the device; the finishing kernel writes their rotation matrices into the jobs.
    {
        std::vector<WarpAJob> aj(n);
        std::vector<double> given(n);
        std::vector<DImg> est_gray;
        std::vector<int> est_idx;
        for (int i = 0; i < n; i++) {
            WarpAJob& j = aj[i];
            j = WarpAJob{};
            j.src = blend_dst[i].p; j.src_pitch = blend_dst[i].pitch; j.sw = blend_dst[i].w; j.sh = blend_dst[i].h;
            j.dst = rot_dst[i].p; j.dst_pitch = rot_dst[i].pitch; j.dw = rot_dst[i].w; j.dh = rot_dst[i].h;
            given[i] = pages[i].angle_deg;
            if (std::isnan(pages[i].angle_deg)) { est_gray.push_back(blend_dst[i]); est_idx.push_back(i); given[i] = 0.0; continue; }
            double m[6];
            DS_TRY(docscan_get_rotation_matrix(blend_dst[i].w / 2.0, blend_dst[i].h / 2.0, pages[i].angle_deg, m));
            hm_invert_affine(m, j.m);
        }
        DS_CUDA(ctx, cudaMemcpyAsync(ctx->angles_dev + page0, given.data(), sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
        WarpAJob* aj_dev = nullptr;
        DS_TRY(k_warp_affine_upload(ctx, aj.data(), n, &aj_dev));
        if (!est_idx.empty()) {
            std::vector<double*> ang(est_idx.size());
            std::vector<WarpAJob*> rot(est_idx.size());
            for (size_t k = 0; k < est_idx.size(); k++) { ang[k] = ctx->angles_dev + page0 + est_idx[k]; rot[k] = aj_dev + est_idx[k]; }
            DS_TRY(k_skew_estimate(ctx, est_gray.data(), (int)est_gray.size(), P.canny_low, P.canny_high, 150, P.max_rotate, nullptr,
                                   ang.data(), rot.data()));
        }
        DS_TRY(k_warp_affine_launch(ctx, aj_dev, aj.data(), n, mw, mh));
    }
    // a9: morph_cleanup (close)
    if (do_close) DS_TRY(morph_op_batch(ctx, DOCSCAN_MORPH_CLOSE, P.morph_ksize, P.morph_ksize, P.morph_iters, rot_dst, binary, nullptr, 0));
    return DOCSCAN_OK;
}

DImg view_of(const docscan_image& im) {
    DImg d;
    d.p = (uint8_t*)im.data; d.w = im.width; d.h = im.height; d.pitch = im.pitch; d.ch = im.channels;
    return d;
}

int copy_2d(docscan_ctx* ctx, void* dst, size_t dpitch, const void* src, size_t spitch, size_t row_bytes, int rows,
            cudaMemcpyKind kind, cudaStream_t st) {
    (kind == cudaMemcpyHostToDevice ? ctx->h2d_bytes : ctx->d2h_bytes) += (int64_t)row_bytes * rows;
    if (dpitch == row_bytes && spitch == row_bytes)
        DS_CUDA(ctx, cudaMemcpyAsync(dst, src, row_bytes * (size_t)rows, kind, st));
    else
        DS_CUDA(ctx, cudaMemcpy2DAsync(dst, dpitch, src, spitch, row_bytes, rows, kind, st));
    return DOCSCAN_OK;
}

// ---- pageable host buffers ------------------------------------------------------------------------------------------------
// cudaMemcpyAsync from pageable memory is staged by the driver through a small buffer and blocks the calling thread (about
// 5 GB/s measured on the B200 box, a tenth of the link).  For callers that hand over plain numpy arrays the library therefore
// keeps a pinned mirror of its staging sets and moves the bytes between the caller's buffers and the mirror itself, with
// several threads, while the previous group's DMA and kernels run.
struct HostCopy { uint8_t* dst; size_t dpitch; const uint8_t* src; size_t spitch; size_t row_bytes; int rows; };

bool is_pageable(const void* p) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}

// Large host-to-host copies with streaming stores: a plain memcpy of a few MB reads the destination lines before it
// overwrites them (three bytes of memory traffic per byte copied instead of two), and the pageable leg is bound by exactly
// that traffic.  Falls back to memcpy on CPUs without AVX2 and for the unaligned head / tail of a row.
__attribute__((target("avx2"))) void stream_copy_avx2(uint8_t* dst, const uint8_t* src, size_t n) {
    const size_t head = std::min(n, (size_t)((32 - (reinterpret_cast<uintptr_t>(dst) & 31)) & 31));
    if (head) { memcpy(dst, src, head); dst += head; src += head; n -= head; }
    size_t i = 0;
    for (; i + 128 <= n; i += 128) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 32));
        const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 64));
        const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 96));
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), a);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 32), b);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 64), c);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 96), d);
    }
    if (i < n) memcpy(dst + i, src + i, n - i);
}

void host_copy(uint8_t* dst, const uint8_t* src, size_t n) {
    static const bool avx2 = __builtin_cpu_supports("avx2") && !getenv("DOCSCAN_NO_STREAM_COPY");
    if (avx2 && n >= 4096) stream_copy_avx2(dst, src, n);
    else memcpy(dst, src, n);
}

void parallel_copy(const std::vector<HostCopy>& jobs) {
    struct Chunk { int job, row0, rows; };
    std::vector<Chunk> chunks;
    for (size_t j = 0; j < jobs.size(); j++) {
        const int per = std::max(1, (int)((size_t)(1 << 20) / std::max<size_t>(jobs[j].row_bytes, 1)));      // ~1 MB per chunk
        for (int r0 = 0; r0 < jobs[j].rows; r0 += per) chunks.push_back({(int)j, r0, std::min(per, jobs[j].rows - r0)});
    }
    if (chunks.empty()) return;
    int nt = (int)std::thread::hardware_concurrency() * 3 / 4;      // measured on the 16-core box: 8 -> 14.4 k, 12 -> 15.8 k, 16 -> 15.7 k MP/s
    if (const char* e = getenv("DOCSCAN_COPY_THREADS")) nt = atoi(e);
    nt = std::max(1, std::min(std::min(nt, 32), (int)chunks.size()));
    std::atomic<size_t> next{0};
    auto work = [&]() {
        for (size_t c = next.fetch_add(1); c < chunks.size(); c = next.fetch_add(1)) {
            const HostCopy& J = jobs[chunks[c].job];
            if (J.dpitch == J.row_bytes && J.spitch == J.row_bytes)
                host_copy(J.dst + (size_t)chunks[c].row0 * J.dpitch, J.src + (size_t)chunks[c].row0 * J.spitch, J.row_bytes * (size_t)chunks[c].rows);
            else
                for (int r = chunks[c].row0; r < chunks[c].row0 + chunks[c].rows; r++) host_copy(J.dst + (size_t)r * J.dpitch, J.src + (size_t)r * J.spitch, J.row_bytes);
        }
        _mm_sfence();
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nt; t++) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
}

}  // namespace

extern "C" int docscan_warp_footprint(const docscan_page* page, int32_t region[4]) {
    if (!page || !region || page->src.width <= 0 || page->src.height <= 0 || page->warped.width <= 0 || page->warped.height <= 0)
        return DOCSCAN_ERR_BAD_ARG;
    const SrcRegion r = page->use_whole ? SrcRegion{0, 0, page->src.width, page->src.height} : warp_footprint(*page);
    region[0] = r.x0; region[1] = r.y0; region[2] = r.x1; region[3] = r.y1;
    return DOCSCAN_OK;
}

extern "C" int docscan_process_pages(docscan_ctx* ctx, int n, docscan_page* pages, const docscan_params* params) {
    if (!ctx || n < 0 || (n && !pages) || !params) return DOCSCAN_ERR_BAD_ARG;
    if (n == 0) return DOCSCAN_OK;
    bool any_host = false;
    size_t max_page = 0;
    for (int i = 0; i < n; i++) {
        docscan_page& pg = pages[i];
        DS_TRY(ds_check_image(ctx, &pg.src, 3, "process_pages: src"));
        DS_TRY(ds_check_image(ctx, &pg.warped, 3, "process_pages: warped"));
        DS_TRY(ds_check_image(ctx, &pg.binary, 1, "process_pages: binary"));
        DS_TRY(same_size(ctx, &pg.warped, &pg.binary, "process_pages"));
        any_host = any_host || is_host(&pg.src) || is_host(&pg.warped) || is_host(&pg.binary);
        max_page = std::max(max_page, page_scratch(pg));
    }
    DS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->angles_cap < n) {
        // an earlier batch may still be reading the old buffer: it is tiny, so simply keep it until the context dies
        if (ctx->angles_dev) ctx->user_allocs.push_back(ctx->angles_dev);
        DS_CUDA(ctx, cudaMalloc((void**)&ctx->angles_dev, sizeof(double) * (size_t)n));
        ctx->angles_cap = n;
    }
    ctx->angles_n = n;
    // pages per launch group: enough 128-column strips to give every SM a few CTAs without cutting pages
    // into short vertical segments (each segment repeats a 2r-row warm-up in the stencil kernels)
    const int strips_per_page = (pages[0].binary.width + 127) / 128;
    int group = (2 * ctx->sm_count + strips_per_page - 1) / strips_per_page;
    group = std::max(4, std::min(group, 32));
    if (n >= 4 * 64 && strips_per_page <= 12) group = 64;   // big resident batches of ordinary pages: one 64-page group per stream measured best
    if (const char* e = getenv("DOCSCAN_GROUP")) group = std::max(1, atoi(e));
    if (any_host) group = std::min(group, 8);          // finer pipeline granularity: copies overlap compute
    bool any_pageable = false;
    if (any_host)
        for (int i = 0; i < n && !any_pageable; i++)
            any_pageable = (is_host(&pages[i].src) && is_pageable(pages[i].src.data)) || (is_host(&pages[i].warped) && is_pageable(pages[i].warped.data)) ||
                           (is_host(&pages[i].binary) && is_pageable(pages[i].binary.data));
    if (any_pageable) group = std::min(group, 4);      // the pinned mirror is 2 x group x (photo + results)
    group = std::min(group, n);
    if (!any_host) {
        // Device-resident batch.  Consecutive groups rotate over the context's stream and up to DS_MAX_STREAMS-1 extra
        // compute streams (own scratch region each): the small serial kernels of one group (Otsu scan, LUTs) and its wave tails
        // overlap the big kernels of the other.  With per-kernel profiling on, a single stream keeps timings clean.
        int ns = ctx->prof_on ? 1 : std::min(DS_MAX_STREAMS, (n + group - 1) / group);   // 4 measured best on B200 (profiles/README.md)
        if (const char* e = getenv("DOCSCAN_STREAMS")) ns = std::max(1, std::min(std::min(DS_MAX_STREAMS, (n + group - 1) / group), atoi(e)));
        DS_TRY(begin_call(ctx, max_page * group * ns));
        for (int k = 1; k < ns; k++)
            if (!ctx->aux[k]) {
                DS_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->aux[k], cudaStreamNonBlocking));
                DS_CUDA(ctx, cudaEventCreateWithFlags(&ctx->aux_ev[k], cudaEventDisableTiming));
            }
        cudaStream_t main_stream = ctx->stream;
        if (ns > 1) {
            if (!ctx->aux_ev[0]) DS_CUDA(ctx, cudaEventCreateWithFlags(&ctx->aux_ev[0], cudaEventDisableTiming));
            DS_CUDA(ctx, cudaEventRecord(ctx->aux_ev[0], main_stream));
            for (int k = 1; k < ns; k++) DS_CUDA(ctx, cudaStreamWaitEvent(ctx->aux[k], ctx->aux_ev[0], 0));
        }
        const size_t base = ctx->arena_off, region = (max_page * group + 255) & ~(size_t)255;
        int rc = DOCSCAN_OK, g = 0;
        for (int i = 0; i < n && rc == DOCSCAN_OK; i += group, g++) {
            const int m = std::min(group, n - i);
            std::vector<DImg> src(m), warped(m), binary(m);
            for (int j = 0; j < m; j++) {
                src[j] = view_of(pages[i + j].src); warped[j] = view_of(pages[i + j].warped); binary[j] = view_of(pages[i + j].binary);
            }
            const int k = g % ns;
            ctx->stream = k ? ctx->aux[k] : main_stream;
            ctx->arena_off = base + (size_t)k * region;
            rc = run_group(ctx, m, pages + i, *params, src, warped, binary, i);
        }
        ctx->stream = main_stream;
        ctx->arena_off = base;
        // join the extra streams first, also when a group failed: their queued kernels still use the arena regions
        for (int k = 1; k < ns; k++) {
            DS_CUDA(ctx, cudaEventRecord(ctx->aux_ev[k], ctx->aux[k]));
            DS_CUDA(ctx, cudaStreamWaitEvent(main_stream, ctx->aux_ev[k], 0));
        }
        return rc;
    }

    // ---- host buffers: three-stage pipeline over groups (H2D | kernels | D2H) on three streams with two
    // staging sets, so the PCIe copies of neighbouring groups overlap the compute of the current one.
    // only the part of a photo under its quad is uploaded (the warp reads nothing else)
    std::vector<SrcRegion> regions(n);
    size_t max_stage = 0;
    for (int i = 0; i < n; i++) {
        regions[i] = (is_host(&pages[i].src) && !pages[i].use_whole) ? warp_footprint(pages[i])
                                                                      : SrcRegion{0, 0, pages[i].src.width, pages[i].src.height};
        const size_t src_stage = is_host(&pages[i].src) ? ds_image_bytes(regions[i].x1 - regions[i].x0, regions[i].y1 - regions[i].y0, 3) : 0;
        max_stage = std::max(max_stage, src_stage + host_bytes(&pages[i].warped) + host_bytes(&pages[i].binary) + 1024);
    }
    DS_TRY(begin_call(ctx, (max_page + 2 * max_stage) * group));
    ArenaScope scope(ctx);
    if (!ctx->copy_in) {
        DS_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
        DS_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
        for (int e = 0; e < 7; e++) DS_CUDA(ctx, cudaEventCreateWithFlags(&ctx->pipe_ev[e], cudaEventDisableTiming));
    }
    uint8_t* stage_base[2];
    for (int sset = 0; sset < 2; sset++) {
        void* p = nullptr;
        DS_TRY(ds_arena_alloc(ctx, max_stage * group, &p));
        stage_base[sset] = (uint8_t*)p;
    }
    uint8_t* pin_base[2] = {nullptr, nullptr};
    if (any_pageable) {
        const size_t need = 2 * max_stage * group;
        if (ctx->pin_mirror_size < need) {
            if (ctx->pin_mirror) { DS_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); cudaFreeHost(ctx->pin_mirror); ctx->pin_mirror = nullptr; ctx->pin_mirror_size = 0; }
            DS_CUDA(ctx, cudaHostAlloc((void**)&ctx->pin_mirror, need, cudaHostAllocDefault));
            ctx->pin_mirror_size = need;
        }
        pin_base[0] = ctx->pin_mirror; pin_base[1] = ctx->pin_mirror + max_stage * group;
    }
    std::vector<HostCopy> out_jobs[2];               // results of a group waiting in the pinned mirror for their trip to the caller
    cudaEvent_t* in_done = &ctx->pipe_ev[0];     // [2]
    cudaEvent_t* comp_done = &ctx->pipe_ev[2];   // [2]
    cudaEvent_t* out_done = &ctx->pipe_ev[4];    // [2]
    cudaEvent_t start = ctx->pipe_ev[6];
    DS_CUDA(ctx, cudaEventRecord(start, ctx->stream));
    DS_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_in, start, 0));
    DS_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_out, start, 0));
    int g = 0;
    for (int i = 0; i < n; i += group, g++) {
        const int m = std::min(group, n - i);
        const int sset = g & 1;
        std::vector<DImg> src(m), warped(m), binary(m);
        if (g >= 2) {
            // The set is carved per page as [src][warped][binary] with page-dependent sizes, so group g's uploads may land on
            // bytes group g-2's results are still being read back from: wait for its kernels AND its D2H copies.
            DS_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_in, comp_done[sset], 0));
            DS_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_in, out_done[sset], 0));
        }
        uint8_t* cur = stage_base[sset];
        auto mirror_of = [&](const uint8_t* dev) { return pin_base[sset] + (dev - stage_base[sset]); };
        std::vector<HostCopy> in_jobs;
        struct Pending { uint8_t* dev; size_t bytes; };
        std::vector<Pending> in_dma;
        if (any_pageable && g >= 2) DS_CUDA(ctx, cudaEventSynchronize(in_done[sset]));    // the mirror set's last uploads have left it
        auto carve = [&](const docscan_image& im, bool dense) {
            DImg d;
            // inputs are staged densely when the caller's rows are dense: one contiguous DMA per page instead of a
            // row-by-row 2-D copy (the gathering warp kernel does not need aligned rows)
            const size_t row = (size_t)im.width * im.channels;
            const size_t pitch = (dense && (size_t)im.pitch == row && (row & 3) == 0) ? row : ((row + 127) & ~(size_t)127);
            d.p = cur; d.w = im.width; d.h = im.height; d.pitch = (int)pitch; d.ch = im.channels;
            cur += ds_image_bytes(im.width, im.height, im.channels);
            cur = (uint8_t*)(((uintptr_t)cur + 255) & ~(uintptr_t)255);
            return d;
        };
        for (int j = 0; j < m; j++) {
            docscan_page& pg = pages[i + j];
            if (is_host(&pg.src)) {
                const SrcRegion& rg = regions[i + j];
                docscan_image part = pg.src;
                part.width = rg.x1 - rg.x0; part.height = rg.y1 - rg.y0;
                part.data = (uint8_t*)pg.src.data + (size_t)rg.y0 * pg.src.pitch + (size_t)rg.x0 * 3;
                src[j] = carve(part, part.width == pg.src.width);
                if (any_pageable && is_pageable(pg.src.data)) {
                    // caller's rows -> mirror (host threads, below), mirror -> device as one contiguous DMA
                    in_jobs.push_back({mirror_of(src[j].p), (size_t)src[j].pitch, (const uint8_t*)part.data, (size_t)part.pitch, (size_t)part.width * 3, part.height});
                    in_dma.push_back({src[j].p, (size_t)src[j].pitch * part.height});
                    ctx->h2d_bytes += (int64_t)part.width * 3 * part.height;
                } else
                DS_TRY(copy_2d(ctx, src[j].p, src[j].pitch, part.data, part.pitch, (size_t)part.width * 3, part.height,
                               cudaMemcpyHostToDevice, ctx->copy_in));
            } else src[j] = view_of(pg.src);
            warped[j] = is_host(&pg.warped) ? carve(pg.warped, false) : view_of(pg.warped);
            binary[j] = is_host(&pg.binary) ? carve(pg.binary, false) : view_of(pg.binary);
        }
        if (!in_jobs.empty()) {
            parallel_copy(in_jobs);
            for (const Pending& d : in_dma) DS_CUDA(ctx, cudaMemcpyAsync(d.dev, mirror_of(d.dev), d.bytes, cudaMemcpyHostToDevice, ctx->copy_in));
        }
        DS_CUDA(ctx, cudaEventRecord(in_done[sset], ctx->copy_in));
        DS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, in_done[sset], 0));
        if (g >= 2) DS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, out_done[sset], 0));       // set's outputs drained
        DS_TRY(run_group(ctx, m, pages + i, *params, src, warped, binary, i, regions.data() + i));
        DS_CUDA(ctx, cudaEventRecord(comp_done[sset], ctx->stream));
        DS_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_out, comp_done[sset], 0));
        for (int j = 0; j < m; j++) {
            docscan_page& pg = pages[i + j];
            if (is_host(&pg.warped)) {
                if (any_pageable && is_pageable(pg.warped.data)) {
                    DS_CUDA(ctx, cudaMemcpyAsync(mirror_of(warped[j].p), warped[j].p, (size_t)warped[j].pitch * pg.warped.height, cudaMemcpyDeviceToHost, ctx->copy_out));
                    out_jobs[sset].push_back({(uint8_t*)pg.warped.data, (size_t)pg.warped.pitch, mirror_of(warped[j].p), (size_t)warped[j].pitch,
                                              (size_t)pg.warped.width * 3, pg.warped.height});
                    ctx->d2h_bytes += (int64_t)pg.warped.width * 3 * pg.warped.height;
                } else
                DS_TRY(copy_2d(ctx, pg.warped.data, pg.warped.pitch, warped[j].p, warped[j].pitch, (size_t)pg.warped.width * 3,
                               pg.warped.height, cudaMemcpyDeviceToHost, ctx->copy_out));
            }
            if (is_host(&pg.binary)) {
                if (any_pageable && is_pageable(pg.binary.data)) {
                    DS_CUDA(ctx, cudaMemcpyAsync(mirror_of(binary[j].p), binary[j].p, (size_t)binary[j].pitch * pg.binary.height, cudaMemcpyDeviceToHost, ctx->copy_out));
                    out_jobs[sset].push_back({(uint8_t*)pg.binary.data, (size_t)pg.binary.pitch, mirror_of(binary[j].p), (size_t)binary[j].pitch,
                                              (size_t)pg.binary.width, pg.binary.height});
                    ctx->d2h_bytes += (int64_t)pg.binary.width * pg.binary.height;
                } else
                DS_TRY(copy_2d(ctx, pg.binary.data, pg.binary.pitch, binary[j].p, binary[j].pitch, (size_t)pg.binary.width,
                               pg.binary.height, cudaMemcpyDeviceToHost, ctx->copy_out));
            }
        }
        DS_CUDA(ctx, cudaEventRecord(out_done[sset], ctx->copy_out));
        // the previous group's results have reached the mirror by now (or soon): hand them to the caller while this group runs
        if (g >= 1 && !out_jobs[sset ^ 1].empty()) {
            DS_CUDA(ctx, cudaEventSynchronize(out_done[sset ^ 1]));
            parallel_copy(out_jobs[sset ^ 1]);
            out_jobs[sset ^ 1].clear();
        }
    }
    for (int k = 0; k < 2; k++) {
        const int sset = (g + k) & 1;                  // older group first
        if (out_jobs[sset].empty()) continue;
        DS_CUDA(ctx, cudaEventSynchronize(out_done[sset]));
        parallel_copy(out_jobs[sset]);
        out_jobs[sset].clear();
    }
    for (int sset = 0; sset < std::min(g, 2); sset++) DS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, out_done[sset], 0));
    return ds_finish(ctx, true);
}

extern "C" int docscan_synth_page(docscan_ctx* ctx, uint64_t seed, docscan_image* dst, float quad_out[8]) {
    if (!ctx) return DOCSCAN_ERR_BAD_ARG;
    DS_TRY(ds_check_image(ctx, dst, 3, "synth_page: dst"));
    DS_TRY(begin_call(ctx, host_bytes(dst)));
    ArenaScope scope(ctx);
    DImg d;
    DS_TRY(ds_stage_out_begin(ctx, dst, &d));
    DS_TRY(k_synth_page(ctx, seed, d, quad_out));
    DS_TRY(ds_stage_out_end(ctx, dst, d));
    return ds_finish(ctx, is_host(dst));
}
