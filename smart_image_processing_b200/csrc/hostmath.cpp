// Host-side parameter preparation, bit-exact with the OpenCV helpers the reference calls.
// Compiled with -ffp-contract=off: every rounding below is where OpenCV's (non-FMA baseline) code rounds.
#include <cfloat>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>

#include "docscan.h"

// cv::getGaussianKernel(k, sigma <= 0) as called by GaussianBlur(..., 0) — DocScanner.py:153,184 and
// inside adaptiveThreshold(GAUSSIAN_C) — DocScanner.py:167.
void hm_gaussian_kernel_f64(int k, double* c) {
    static const double t1[] = {1.0};
    static const double t3[] = {0.25, 0.5, 0.25};
    static const double t5[] = {0.0625, 0.25, 0.375, 0.25, 0.0625};
    static const double t7[] = {0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125};
    static const double t9[] = {4.0 / 256, 13.0 / 256, 30.0 / 256, 51.0 / 256, 60.0 / 256,
                                51.0 / 256, 30.0 / 256, 13.0 / 256, 4.0 / 256};
    const double* fixed = nullptr;
    switch (k) {
        case 1: fixed = t1; break;
        case 3: fixed = t3; break;
        case 5: fixed = t5; break;
        case 7: fixed = t7; break;
        case 9: fixed = t9; break;
        default: break;
    }
    if (fixed) {
        for (int i = 0; i < k; i++) c[i] = fixed[i];
        return;
    }
    const double sigma = std::fma((double)k, 0.15, 0.35);
    const double scale2 = -0.125 / (sigma * sigma);
    const int half = (k - 1) / 2;
    double total = 0.0;
    int x = 1 - k;
    for (int i = 0; i < half; i++, x += 2) {
        c[i] = std::exp((double)(x * x) * scale2);
        total += c[i];
    }
    total *= 2.0;
    total += 1.0;
    const double inv = 1.0 / total;
    for (int i = 0; i < half; i++) {
        c[i] = c[i] * inv;
        c[k - 1 - i] = c[i];
    }
    c[half] = inv;
}

extern "C" int docscan_gaussian_kernel_f32(int k, float* out) {
    if (k < 1 || (k & 1) == 0 || !out) return DOCSCAN_ERR_BAD_ARG;
    double* c = new double[k];
    hm_gaussian_kernel_f64(k, c);
    for (int i = 0; i < k; i++) out[i] = (float)c[i];
    delete[] c;
    return DOCSCAN_OK;
}

// 8.8 fixed point with error diffusion; the centre tap absorbs the remainder so the taps sum to 256.
extern "C" int docscan_gaussian_kernel_q8(int k, int32_t* out) {
    if (k < 1 || (k & 1) == 0 || !out) return DOCSCAN_ERR_BAD_ARG;
    double* c = new double[k];
    hm_gaussian_kernel_f64(k, c);
    const int half = k / 2;
    double carry = 0.0;
    int64_t side = 0;
    for (int i = 0; i < half; i++) {
        const double want = c[i] * 256.0 + carry;
        const int64_t q = (int64_t)std::lrint(want);
        carry = want - (double)q;
        out[i] = out[k - 1 - i] = (int32_t)q;
        side += q;
    }
    out[half] = (int32_t)(256 - 2 * side);
    delete[] c;
    return DOCSCAN_OK;
}

// cv::getPerspectiveTransform (DocScanner.py:142): the 8x8 system is solved with partial-pivot LU in
// fp64; the four products -src*dst are formed in float first, exactly like the Point2f expression.
extern "C" int docscan_get_perspective_transform(const float quad[8], const float dst[8], double m[9]) {
    if (!quad || !dst || !m) return DOCSCAN_ERR_BAD_ARG;
    double A[8][8];
    double b[8];
    std::memset(A, 0, sizeof(A));
    for (int i = 0; i < 4; i++) {
        const float sx = quad[2 * i], sy = quad[2 * i + 1];
        const float dx = dst[2 * i], dy = dst[2 * i + 1];
        volatile float nxx = -sx * dx, nyx = -sy * dx, nxy = -sx * dy, nyy = -sy * dy;
        double* r0 = A[i];
        double* r1 = A[i + 4];
        r0[0] = sx; r0[1] = sy; r0[2] = 1.0; r0[6] = nxx; r0[7] = nyx;
        r1[3] = sx; r1[4] = sy; r1[5] = 1.0; r1[6] = nxy; r1[7] = nyy;
        b[i] = dx;
        b[i + 4] = dy;
    }
    const double tiny = DBL_EPSILON * 100;
    for (int col = 0; col < 8; col++) {
        int piv = col;
        for (int row = col + 1; row < 8; row++)
            if (std::fabs(A[row][col]) > std::fabs(A[piv][col])) piv = row;
        if (std::fabs(A[piv][col]) < tiny) {
            std::memset(m, 0, sizeof(double) * 9);
            m[8] = 1.0;
            return DOCSCAN_OK;   // cv::solve leaves X = 0 for a singular system; caller gets a degenerate M
        }
        if (piv != col) {
            for (int j = col; j < 8; j++) { const double t = A[col][j]; A[col][j] = A[piv][j]; A[piv][j] = t; }
            const double t = b[col]; b[col] = b[piv]; b[piv] = t;
        }
        const double neg_inv = -1 / A[col][col];
        for (int row = col + 1; row < 8; row++) {
            const double f = A[row][col] * neg_inv;
            for (int j = col + 1; j < 8; j++) A[row][j] += f * A[col][j];
            b[row] += f * b[col];
        }
    }
    for (int row = 7; row >= 0; row--) {
        double acc = b[row];
        for (int j = row + 1; j < 8; j++) acc -= A[row][j] * b[j];
        b[row] = acc / A[row][row];
    }
    for (int i = 0; i < 8; i++) m[i] = b[i];
    m[8] = 1.0;
    return DOCSCAN_OK;
}

// cv::invert(3x3 fp64) as cv::warpPerspective applies it before remapping.
void hm_invert3x3(const double S[9], double T[9]) {
    double det = S[0] * (S[4] * S[8] - S[5] * S[7]) - S[1] * (S[3] * S[8] - S[5] * S[6]) +
                 S[2] * (S[3] * S[7] - S[4] * S[6]);
    if (det == 0.0) {
        std::memset(T, 0, sizeof(double) * 9);
        return;
    }
    det = 1.0 / det;
    T[0] = (S[4] * S[8] - S[5] * S[7]) * det;
    T[1] = (S[2] * S[7] - S[1] * S[8]) * det;
    T[2] = (S[1] * S[5] - S[2] * S[4]) * det;
    T[3] = (S[5] * S[6] - S[3] * S[8]) * det;
    T[4] = (S[0] * S[8] - S[2] * S[6]) * det;
    T[5] = (S[2] * S[3] - S[0] * S[5]) * det;
    T[6] = (S[3] * S[7] - S[4] * S[6]) * det;
    T[7] = (S[1] * S[6] - S[0] * S[7]) * det;
    T[8] = (S[0] * S[4] - S[1] * S[3]) * det;
}

// cv::getRotationMatrix2D (DocScanner.py:234); the centre is a Point2f.
extern "C" int docscan_get_rotation_matrix(double cx, double cy, double angle_deg, double m[6]) {
    if (!m) return DOCSCAN_ERR_BAD_ARG;
    const float fx = (float)cx, fy = (float)cy;
    const double rad = angle_deg * (3.1415926535897932384626433832795 / 180);
    const double ca = std::cos(rad), sa = std::sin(rad);
    m[0] = ca;  m[1] = sa; m[2] = (1 - ca) * fx - sa * fy;
    m[3] = -sa; m[4] = ca; m[5] = sa * fx + (1 - ca) * fy;
    return DOCSCAN_OK;
}

// The inverse cv::warpAffine derives from the forward 2x3 matrix.
void hm_invert_affine(const double F[6], double I[6]) {
    double det = F[0] * F[4] - F[1] * F[3];
    det = det != 0 ? 1. / det : 0;
    const double a11 = F[4] * det, a22 = F[0] * det;
    I[0] = a11;
    I[1] = F[1] * -det;
    I[3] = F[3] * -det;
    I[4] = a22;
    I[2] = -I[0] * F[2] - I[1] * F[5];
    I[5] = -I[3] * F[2] - I[4] * F[5];
}

// cv::threshold(THRESH_OTSU) on a 256-bin histogram (ordered fp64 scan) — DocScanner.py:187,202.
extern "C" int docscan_otsu_from_hist(const int32_t hist[256], int64_t total, double* t) {
    if (!hist || !t || total <= 0) return DOCSCAN_ERR_BAD_ARG;
    const double norm = 1.0 / (double)total;
    double mean_all = 0.0;
    for (int i = 0; i < 256; i++) mean_all += (double)i * (double)hist[i];
    mean_all *= norm;
    double m_lo = 0.0, w_lo = 0.0, best = 0.0, best_t = 0.0;
    for (int i = 0; i < 256; i++) {
        const double p = (double)hist[i] * norm;
        m_lo *= w_lo;
        w_lo += p;
        const double w_hi = 1.0 - w_lo;
        if (std::fmin(w_lo, w_hi) < FLT_EPSILON || std::fmax(w_lo, w_hi) > 1.0 - FLT_EPSILON) continue;
        m_lo = (m_lo + (double)i * p) / w_lo;
        const double m_hi = (mean_all - w_lo * m_lo) / w_hi;
        const double between = w_lo * w_hi * (m_lo - m_hi) * (m_lo - m_hi);
        if (between > best) { best = between; best_t = (double)i; }
    }
    *t = best_t;
    return DOCSCAN_OK;
}

extern "C" void docscan_default_params(docscan_params* p) {
    if (!p) return;
    p->illum_method = 0;            // "subtract"
    p->illum_blur_frac = 0.02;
    p->block_size = 35;
    p->C = 10;
    p->thresh_method = DOCSCAN_ADAPTIVE_GAUSSIAN;
    p->mask_blur_ksize = 51;
    p->blackhat_ksize = 9;
    p->blackhat_vertical_ratio = 2.0;
    p->ink_dilate_iters = 1;
    p->mask_thresh_offset = 8;
    p->canny_low = 50; p->canny_high = 150; p->max_rotate = 10.0;
    p->morph_ksize = 3;
    p->morph_iters = 1;
    p->cv_tail_compat = 1;
}

static double py_round(double v) { return std::nearbyint(v); }   // Python round(): half to even

// perspective_warp's target size (DocScanner.py:120-139).  np.linalg.norm on float32 vectors is
// evaluated in float32 (sqrt of the float32 dot product).
extern "C" int docscan_target_size(const float quad[8], int page_kind, int scale_long, int32_t* w, int32_t* h) {
    if (!quad || !w || !h) return DOCSCAN_ERR_BAD_ARG;
    auto norm2 = [](float ax, float ay, float bx, float by) -> float {
        const float dx = ax - bx, dy = ay - by;
        volatile float xx = dx * dx, yy = dy * dy;
        volatile float s = xx + yy;
        return std::sqrt((float)s);
    };
    const float* tl = quad; const float* tr = quad + 2; const float* br = quad + 4; const float* bl = quad + 6;
    const float w_top = norm2(tr[0], tr[1], tl[0], tl[1]);
    const float w_bot = norm2(br[0], br[1], bl[0], bl[1]);
    const float h_left = norm2(bl[0], bl[1], tl[0], tl[1]);
    const float h_right = norm2(br[0], br[1], tr[0], tr[1]);
    const int width = (int)w_top > (int)w_bot ? (int)w_top : (int)w_bot;
    const int height = (int)h_left > (int)h_right ? (int)h_left : (int)h_right;
    const bool portrait = height >= width;
    double ratio;
    if (page_kind == 0) ratio = std::sqrt(2.0);
    else if (page_kind == 1) ratio = 11.0 / 8.5;
    else ratio = (double)height / (double)(width > 1 ? width : 1);
    if (portrait) {
        *h = scale_long;
        *w = (int)py_round((double)scale_long / ratio);
    } else {
        *w = scale_long;
        *h = (int)py_round((double)scale_long * ratio);
    }
    return DOCSCAN_OK;
}

// cv::resize INTER_CUBIC tables: fx in fp32, cv::interpolateCubic (A = -0.75) in fp32, taps quantised to 11 bits with
// cvRound, source indices clamped to the image (DocScanner.py:35-36 through resize_long_side).
void hm_cubic_taps(int ssize, int dsize, int d, int idx[4], short w[4]) {
    const double scale = (double)ssize / dsize;
    float fx = (float)((d + 0.5) * scale - 0.5);
    const int sx = (int)std::floor(fx);
    fx -= (float)sx;
    const float A = -0.75f;
    float c[4];
    c[0] = ((A * (fx + 1) - 5 * A) * (fx + 1) + 8 * A) * (fx + 1) - 4 * A;
    c[1] = ((A + 2) * fx - (A + 3)) * fx * fx + 1;
    c[2] = ((A + 2) * (1 - fx) - (A + 3)) * (1 - fx) * (1 - fx) + 1;
    c[3] = 1.f - c[0] - c[1] - c[2];
    for (int k = 0; k < 4; k++) {
        long v = std::lrintf(c[k] * 2048.f);
        w[k] = (short)(v < -32768 ? -32768 : (v > 32767 ? 32767 : v));
        const int s = sx - 1 + k;
        idx[k] = s < 0 ? 0 : (s > ssize - 1 ? ssize - 1 : s);
    }
}

// cv::HoughLines' trig tables for rho = 1, theta = pi/180 (cv::createTrigTable): the angle is accumulated in fp32.
void hm_hough_trig_table(float c[180], float s[180]) {
    float ang = 0.f;
    const double step = 3.1415926535897932384626433832795 / 180;
    for (int n = 0; n < 180; n++) {
        s[n] = (float)std::sin((double)ang);
        c[n] = (float)std::cos((double)ang);
        ang += (float)step;
    }
}

// DocScanner.py:223-226 for the line angle of Hough index n, in the arithmetic numpy >= 2 gives it: theta is np.float32
// (cv::HoughLines returns min_theta + n * theta in fp32), so theta*180.0/np.pi and (ang + 90.0) % 180.0 - 90.0 stay fp32.
void hm_folded_angles(float out[180]) {
    const float step = (float)(3.1415926535897932384626433832795 / 180), pi_f = (float)3.1415926535897932384626433832795;
    for (int n = 0; n < 180; n++) {
        const float theta = 0.f + (float)n * step;
        float ang = theta * 180.0f;
        ang = ang / pi_f;
        out[n] = std::fmod(ang + 90.0f, 180.0f) - 90.0f;       // the operand is >= 0, where Python's % is fmod
    }
}

// np.median of the folded angles (mean of the middle two in fp32 for an even count), 0 beyond max_rotate or without
// lines — DocScanner.py:227-231.  per_angle[n] = number of Hough lines with angle index n.
extern "C" int docscan_median_angle(const int32_t per_angle[180], double max_rotate, double* angle_deg) {
    if (!per_angle || !angle_deg) return DOCSCAN_ERR_BAD_ARG;
    float folded[180];
    hm_folded_angles(folded);
    int order[180];
    for (int i = 0; i < 180; i++) order[i] = i;
    std::stable_sort(order, order + 180, [&](int a, int b) { return folded[a] < folded[b]; });
    int64_t total = 0;
    for (int n = 0; n < 180; n++) {
        if (per_angle[n] < 0) return DOCSCAN_ERR_BAD_ARG;
        total += per_angle[n];
    }
    *angle_deg = 0.0;
    if (total == 0) return DOCSCAN_OK;
    const int64_t k_lo = (total - 1) / 2, k_hi = total / 2;
    int i_lo = -1, i_hi = -1;
    int64_t run = 0;
    for (int i = 0; i < 180; i++) {
        run += per_angle[order[i]];
        if (i_lo < 0 && run > k_lo) i_lo = i;
        if (i_hi < 0 && run > k_hi) { i_hi = i; break; }
    }
    const float a = folded[order[i_lo]], b = folded[order[i_hi]];
    const float med = (total & 1) ? a : (a + b) / 2.0f;
    if (!(std::fabs((double)med) > max_rotate)) *angle_deg = (double)med;
    return DOCSCAN_OK;
}
