// Per-page scalar work that sits between the image passes: min-max normalisation LUTs
// (cv2.normalize NORM_MINMAX), Otsu thresholds from 256-bin histograms, and the raw cut-off values the
// mask kernel compares against.  Everything stays on the device; one small block per page.
#include <cfloat>

#include "common.cuh"

__global__ void scalars_reset_kernel(PageScalars* s, int n) {
    const int page = blockIdx.x;
    if (page >= n) return;
    PageScalars& S = s[page];
    const int t = threadIdx.x;
    S.hist_a[t] = 0;
    S.hist_b[t] = 0;
    S.lut[t] = (uint8_t)t;
    if (t == 0) {
        S.minmax[0] = 255; S.minmax[1] = 0;
        S.cut_a = 256; S.cut_b = 256; S.otsu_a = 0; S.otsu_b = 0;
    }
}

int k_scalars_reset(docscan_ctx* ctx, PageScalars* s, int n) {
    ProfScope prof(ctx, "scalars_reset", 0);
    scalars_reset_kernel<<<n, 256, 0, ctx->stream>>>(s, n);
    DS_CHECK_LAUNCH(ctx);
    return DOCSCAN_OK;
}

// cv2.normalize(src, None, 0, 255, NORM_MINMAX) for uint8 as a LUT entry (DocScanner.py:156,159,172,186,201):
// scale = 255 * (1/(max-min)) in fp64 (0 when max == min), shift = -min*scale, both cast to fp32, then one
// fused multiply-add and round-half-even.
__device__ __forceinline__ uint8_t norm_entry(int v, int smin, int smax) {
    const double d = (double)smax - (double)smin;
    const double scale = __dmul_rn(255.0, d > DBL_EPSILON ? __ddiv_rn(1.0, d) : 0.0);
    const double shift = __dsub_rn(0.0, __dmul_rn((double)smin, scale));
    const float sf = (float)scale, bf = (float)shift;
    const int r = __float2int_rn(__fmaf_rn((float)v, sf, bf));
    return (uint8_t)min(max(r, 0), 255);
}

__global__ void build_norm_lut_kernel(PageScalars* s, int n, int compose_stretch) {
    const int page = blockIdx.x;
    if (page >= n) return;
    PageScalars& S = s[page];
    const int v = threadIdx.x;
    int mn = (int)S.minmax[0], mx = (int)S.minmax[1];
    if (mn > mx) { mn = 0; mx = 0; }
    uint8_t e = norm_entry(v, mn, mx);
    if (compose_stretch) {
        // contrast_stretch(illum) (DocScanner.py:171-172) normalises the already normalised image again:
        // its min / max are the images of min / max because the LUT is monotone.
        const int mn2 = norm_entry(mn, mn, mx), mx2 = norm_entry(mx, mn, mx);
        e = norm_entry(e, mn2, mx2);
    }
    S.lut[v] = e;
}

int k_build_norm_lut(docscan_ctx* ctx, PageScalars* s, int n, int compose_stretch) {
    ProfScope prof(ctx, "scalars_norm_lut", 0);
    build_norm_lut_kernel<<<n, 256, 0, ctx->stream>>>(s, n, compose_stretch);
    DS_CHECK_LAUNCH(ctx);
    return DOCSCAN_OK;
}

// cv::threshold(THRESH_OTSU): the between-class variance scan is an ordered fp64 recurrence
// (getThreshVal_Otsu_8u), so one lane walks the 256 bins and stores (w_lo, m_lo) per bin; the variances
// and the first-maximum search are then done by the whole warp.
__device__ int otsu_warp(const uint32_t* hist /*smem, 256*/, double total, double* s_w, double* s_m, uint8_t* s_ok) {
    const int lane = threadIdx.x & 31;
    // sum i*h[i] is an exact integer in fp64 in any order
    unsigned long long part = 0;
    for (int i = lane; i < 256; i += 32) part += (unsigned long long)i * hist[i];
    for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    const double norm = __ddiv_rn(1.0, total);
    const double mean_all = __dmul_rn((double)part, norm);
    if (lane == 0) {
        double m_lo = 0.0, w_lo = 0.0;
        for (int i = 0; i < 256; i++) {
            const double p = __dmul_rn((double)hist[i], norm);
            m_lo = __dmul_rn(m_lo, w_lo);
            w_lo = __dadd_rn(w_lo, p);
            const double w_hi = __dsub_rn(1.0, w_lo);
            const bool skip = fmin(w_lo, w_hi) < (double)FLT_EPSILON || fmax(w_lo, w_hi) > 1.0 - (double)FLT_EPSILON;
            if (!skip) m_lo = __ddiv_rn(__dadd_rn(m_lo, __dmul_rn((double)i, p)), w_lo);
            s_w[i] = w_lo; s_m[i] = m_lo; s_ok[i] = skip ? 0 : 1;
        }
    }
    __syncwarp();
    double best = 0.0;
    int best_i = 0;
    for (int i = lane; i < 256; i += 32) {
        if (!s_ok[i]) continue;
        const double w_lo = s_w[i], m_lo = s_m[i];
        const double w_hi = __dsub_rn(1.0, w_lo);
        const double m_hi = __ddiv_rn(__dsub_rn(mean_all, __dmul_rn(w_lo, m_lo)), w_hi);
        const double diff = __dsub_rn(m_lo, m_hi);
        const double between = __dmul_rn(__dmul_rn(__dmul_rn(w_lo, w_hi), diff), diff);
        if (between > best) { best = between; best_i = i; }   // per lane: increasing i, strict > keeps the first
    }
    // first index attaining the global maximum (the sequential loop's answer)
    for (int o = 16; o; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
        if (ob > best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
    }
    return best > 0.0 ? best_i : 0;
}

// mode bit0: normalise the histogram first (the reference thresholds a MINMAX-normalised image whose
// histogram is the LUT-remapped histogram of the raw one — SURVEY A.6)
__global__ void __launch_bounds__(64) otsu_cuts_kernel(PageScalars* s, int n, int threshold_offset,
                                                       const int32_t* __restrict__ npix, int normalise) {
    const int page = blockIdx.x;
    if (page >= n) return;
    PageScalars& S = s[page];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __shared__ uint32_t s_hist[2][256];
    __shared__ uint8_t s_lut[2][256];
    __shared__ double s_w[2][256], s_m[2][256];
    __shared__ uint8_t s_ok[2][256];
    const uint32_t* raw = warp == 0 ? S.hist_a : S.hist_b;
    // min / max of the raw image from its histogram
    int lo = 256, hi = -1;
    for (int i = lane; i < 256; i += 32) {
        s_hist[warp][i] = 0;
        if (raw[i]) { lo = min(lo, i); hi = max(hi, i); }
    }
    for (int o = 16; o; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (hi < 0) { lo = 0; hi = 0; }
    __syncwarp();
    for (int i = lane; i < 256; i += 32) {
        const uint8_t e = normalise ? norm_entry(i, lo, hi) : (uint8_t)i;
        s_lut[warp][i] = e;
        if (raw[i]) atomicAdd(&s_hist[warp][e], raw[i]);
    }
    __syncwarp();
    const int t = otsu_warp(s_hist[warp], (double)npix[page], s_w[warp], s_m[warp], s_ok[warp]);
    // t' = max(0, int(round(t - offset)))  (DocScanner.py:188,203); mask = normalised > t'
    const int tq = max(0, t - threshold_offset);
    int cut = 256;
    for (int i = lane; i < 256; i += 32)
        if ((int)s_lut[warp][i] > tq) cut = min(cut, i);
    for (int o = 16; o; o >>= 1) cut = min(cut, __shfl_xor_sync(0xffffffffu, cut, o));
    if (lane == 0) {
        if (warp == 0) { S.otsu_a = t; S.cut_a = cut; }
        else { S.otsu_b = t; S.cut_b = cut; }
    }
}

int k_otsu_cuts(docscan_ctx* ctx, PageScalars* s, int n, int threshold_offset, const int32_t* npix_dev) {
    ProfScope prof(ctx, "scalars_otsu", 0);
    otsu_cuts_kernel<<<n, 64, 0, ctx->stream>>>(s, n, threshold_offset, npix_dev, 1);
    DS_CHECK_LAUNCH(ctx);
    return DOCSCAN_OK;
}

int k_otsu_plain(docscan_ctx* ctx, PageScalars* s, int n, const int32_t* npix_dev) {
    ProfScope prof(ctx, "scalars_otsu", 0);
    otsu_cuts_kernel<<<n, 64, 0, ctx->stream>>>(s, n, 0, npix_dev, 0);
    DS_CHECK_LAUNCH(ctx);
    return DOCSCAN_OK;
}
