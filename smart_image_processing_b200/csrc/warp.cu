// Geometric resampling stages.
//   cv2.warpPerspective(u8, M, INTER_LINEAR, BORDER_CONSTANT 0)   DocScanner.py:143  (+ fused BGR2GRAY, :316)
//   cv2.warpAffine(u8c1, M, INTER_LINEAR, BORDER_REPLICATE)        DocScanner.py:235
// Both reproduce OpenCV's fixed-point pipeline: source coordinates are rounded (half-to-even) to 1/32 px
// (INTER_BITS = 5), the four bilinear weights are the integers (32-ax)(32-ay)*32 ... that sum to 2^15, and
// the result is (sum + 2^14) >> 15.  Texture units are not used: their 9-bit filtering cannot give these
// roundings.  The perspective coordinates are evaluated in fp64 in the same operation order as OpenCV
// (per 64-pixel block origin, no fma contraction), the affine ones in OpenCV's 10-bit fixed point.
#include <climits>
#include <cmath>
#include <cstdlib>

#include "common.cuh"

#ifndef WP3_MINB
#define WP3_MINB 8      // resident CTAs per SM the 3-channel perspective kernel is compiled for: 64 registers (measured: 7 -> 2.81 ms, 8 -> 2.65, 9 -> 2.92, 10 -> 3.09)
#endif

namespace {

__device__ __forceinline__ uint8_t gray15(int b, int g, int r) {
    return (uint8_t)((3735 * b + 19235 * g + 9798 * r + 16384) >> 15);
}

// saturate_cast<int>(max(INT_MIN, min(INT_MAX, v))): the conversion instruction saturates on its own; std::min/max
// as OpenCV writes them turn a NaN into INT_MAX
__device__ __forceinline__ int round_clamped(double v) {
    return v != v ? INT_MAX : __double2int_rn(v);
}

template <int CH>
__global__ void __launch_bounds__(128) warp_perspective_kernel(const WarpPJob* __restrict__ jobs) {
    const WarpPJob& J = jobs[blockIdx.z];
    const int y = blockIdx.y * 4 + threadIdx.y;
    const int xt = blockIdx.x * 128;                 // a warp covers 128 consecutive destination pixels of one row
    if (y >= J.dh || xt >= J.dw) return;
    const double m0 = J.m[0], m1 = J.m[1], m2 = J.m[2], m3 = J.m[3], m4 = J.m[4], m5 = J.m[5], m6 = J.m[6], m7 = J.m[7], m8 = J.m[8];
    const double dy = (double)y;
    const uint8_t* __restrict__ src = J.src;
    const int sw = J.sw, sh = J.sh, sp = J.src_pitch;
    int xb = -1;
    double X0 = 0, Y0 = 0, W0 = 0;
    uint8_t* drow = J.dst + (size_t)y * J.dst_pitch;
    uint8_t* grow = (CH == 3 && J.gray) ? J.gray + (size_t)y * J.gray_pitch : nullptr;
    // lane l handles pixels xt + l, xt + 32 + l, xt + 64 + l, xt + 96 + l: every gather instruction of the warp then
    // covers 32 ADJACENT destination pixels, i.e. a compact ~200-byte source span (2-3 cache lines) instead of the
    // 7 lines a 4-pixels-per-lane layout touches.  The kernel is bound by L1 wavefronts, so this is the lever.
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int x = xt + 32 * i + threadIdx.x;
        if (x >= J.dw) break;
        // OpenCV evaluates the row terms at the origin of a block that is 64 px wide (1024 / min(16, rows))
        const int xbi = J.block_w == 64 ? (x & ~63) : x - x % J.block_w;
        if (xbi != xb) {
            xb = xbi;
            const double dxb = (double)xb;
            X0 = __dadd_rn(__dadd_rn(__dmul_rn(m0, dxb), __dmul_rn(m1, dy)), m2);
            Y0 = __dadd_rn(__dadd_rn(__dmul_rn(m3, dxb), __dmul_rn(m4, dy)), m5);
            W0 = __dadd_rn(__dadd_rn(__dmul_rn(m6, dxb), __dmul_rn(m7, dy)), m8);
        }
        const double x1 = (double)(x - xb);
        double W = __dadd_rn(W0, __dmul_rn(m6, x1));
        W = W != 0.0 ? __ddiv_rn(32.0, W) : 0.0;
        const double fX = __dmul_rn(__dadd_rn(X0, __dmul_rn(m0, x1)), W);
        const double fY = __dmul_rn(__dadd_rn(Y0, __dmul_rn(m3, x1)), W);
        const int X = round_clamped(fX), Y = round_clamped(fY);
        // OpenCV keeps sx, sy as saturated shorts; it also requires images below 32767 px, so a coordinate beyond that range is
        // outside the image with or without the saturation
        const int sx = X >> 5, sy = Y >> 5;
        const int ax = X & 31, ay = Y & 31;
        const bool y0in = (unsigned)sy < (unsigned)sh, y1in = (unsigned)(sy + 1) < (unsigned)sh;
        const int cy0 = ds_clamp(sy, J.ry0, J.ry1 - 1) - J.ry0, cy1 = ds_clamp(sy + 1, J.ry0, J.ry1 - 1) - J.ry0;
        const uint8_t* r0 = src + (size_t)cy0 * sp - (size_t)J.rx0 * CH;
        const uint8_t* r1 = src + (size_t)cy1 * sp - (size_t)J.rx0 * CH;
        int acc[CH];
        if (CH == 3 && sx >= J.rx0 + 3 && sx <= J.rx1 - 6) {
            // Interior fast path.  The two taps of a source row are 6 consecutive bytes: fetch the 16-byte aligned-8 window
            // around them with two 64-bit loads (4 load instructions per pixel instead of 12 byte gathers: the kernel is
            // bound by L1 wavefronts), shift the 6 bytes down, and filter horizontally with dp4a on (p0, p1) x (32-ax, ax).
            // (32-ax)(32-ay)32 p00 + ... == 32 * [(32-ay) h0 + ay h1] exactly, so (.. + 2^14) >> 15 == (v + 512) >> 10.
            // The window stays inside the resident row: 3*(sx-rx0) - 7 >= 2 and 3*(sx-rx0) + 15 <= 3*(rx1-rx0) - 3.
            const uint32_t wx = (uint32_t)(32 - ax) | ((uint32_t)ax << 8);
            const int wy0 = y0in ? 32 - ay : 0, wy1 = y1in ? ay : 0;         // rows outside the image: BORDER_CONSTANT 0
            int h[2][3];
#pragma unroll
            for (int rr = 0; rr < 2; rr++) {
                const uintptr_t A = reinterpret_cast<uintptr_t>((rr ? r1 : r0) + 3 * sx);
                const uint2* base = reinterpret_cast<const uint2*>(A & ~(uintptr_t)7);
                const uint2 lo = __ldg(base), hi = __ldg(base + 1);
                const uint32_t sft = (uint32_t)(A & 7);
                const bool up = sft >= 4;
                const uint32_t wa = up ? lo.y : lo.x, wb = up ? hi.x : lo.y, wc = up ? hi.y : hi.x;
                const uint32_t b0 = __funnelshift_r(wa, wb, 8 * sft), b1 = __funnelshift_r(wb, wc, 8 * sft);   // shift mod 32
                h[rr][0] = __dp4a(__byte_perm(b0, b1, 0x0030), wx, 0u);      // (B0, B1)
                h[rr][1] = __dp4a(__byte_perm(b0, b1, 0x0041), wx, 0u);      // (G0, G1)
                h[rr][2] = __dp4a(__byte_perm(b0, b1, 0x0052), wx, 0u);      // (R0, R1)
            }
#pragma unroll
            for (int c = 0; c < 3; c++) if (c < CH) acc[c] = (wy0 * h[0][c] + wy1 * h[1][c] + 512) << 5;
        } else {
            // branch-free taps: out-of-image taps get weight 0 (BORDER_CONSTANT 0) and a clamped, always-valid address
            const int w00 = (32 - ax) * (32 - ay) * 32, w01 = ax * (32 - ay) * 32, w10 = (32 - ax) * ay * 32, w11 = ax * ay * 32;
            const bool x0in = (unsigned)sx < (unsigned)sw, x1in = (unsigned)(sx + 1) < (unsigned)sw;
            const int cx0 = ds_clamp(sx, J.rx0, J.rx1 - 1), cx1 = ds_clamp(sx + 1, J.rx0, J.rx1 - 1);
            const int v00 = (x0in && y0in) ? w00 : 0, v01 = (x1in && y0in) ? w01 : 0;
            const int v10 = (x0in && y1in) ? w10 : 0, v11 = (x1in && y1in) ? w11 : 0;
#pragma unroll
            for (int c = 0; c < CH; c++)
                acc[c] = 16384 + v00 * __ldg(r0 + cx0 * CH + c) + v01 * __ldg(r0 + cx1 * CH + c) +
                         v10 * __ldg(r1 + cx0 * CH + c) + v11 * __ldg(r1 + cx1 * CH + c);
        }
        uint8_t* dp = drow + (size_t)x * CH;
#pragma unroll
        for (int c = 0; c < CH; c++) dp[c] = (uint8_t)(acc[c] >> 15);      // <= 255 by construction
        if (CH == 3 && grow) grow[x] = gray15(acc[0] >> 15, acc[1] >> 15, acc[2] >> 15);
    }
}

// 3-channel variant used by the page pipeline.  Same arithmetic as warp_perspective_kernel<3>, organised in three phases
// so that all 16 loads of a thread's 4 pixels are in flight together (the gathers are latency-bound otherwise):
// coordinates -> loads -> filter + stores.  Interior pixels fetch the 6 bytes of a row's two taps through a 16-byte
// window (two 64-bit loads at the enclosing 8-byte boundary); pixels whose taps touch the left/right image border take the
// per-byte path of the generic kernel afterwards.  Requires a resident width of at least 9 px.
// SAFE_RCP: the host has checked that the denominator W keeps one sign and stays within [1e-290, 1e290] over the whole
// destination rectangle (it is affine in (x, y), so its extremes sit at the corners).  32 is a power of two, so the
// correctly rounded quotient 32 / W is then exactly 32 * RN(1 / W) — the reciprocal sequence is half as long as the division.
// P8: every source pitch is a multiple of 8 bytes, so the two rows of a pixel share their offset inside the 8-byte window.
template <bool SAFE_RCP, bool P8>
__global__ void __launch_bounds__(128, WP3_MINB) warp_perspective3_kernel(const WarpPJob* __restrict__ jobs) {
    const WarpPJob& J = jobs[blockIdx.z];
    const int y = blockIdx.y * 4 + threadIdx.y;
    const int xt = blockIdx.x * 128;                 // a warp covers 128 consecutive destination pixels of one row
    if (y >= J.dh || xt >= J.dw) return;
    // every job field the loops use, read once: the byte stores below may alias anything, so fields read through the
    // reference would be fetched again after each of them
    const int sw = J.sw, sh = J.sh, sp = J.src_pitch;
    const uint8_t* __restrict__ src = J.src;
    const int rx0 = J.rx0, rx1 = J.rx1, ry0 = J.ry0, ry1 = J.ry1, dw = J.dw, block_w = J.block_w;
    uint8_t* const drow = J.dst + (size_t)y * J.dst_pitch;
    uint8_t* const grow = J.gray ? J.gray + (size_t)y * J.gray_pitch : nullptr;
    int Xs[4], Ys[4];
    {
        const double m0 = J.m[0], m1 = J.m[1], m2 = J.m[2], m3 = J.m[3], m4 = J.m[4], m5 = J.m[5], m6 = J.m[6], m7 = J.m[7], m8 = J.m[8];
        const double dy = (double)y;
        int xb = -1;
        double X0 = 0, Y0 = 0, W0 = 0;
        // SAFE_RCP: the warp's 128 pixels share one row and two block origins, i.e. six row terms.  Lane k (k < 6) evaluates
        // term k (one mul-mul-add-add chain per warp instead of six per thread) and the others fetch them with shuffles.
        double T6 = 0;
        if (SAFE_RCP) {
            const int k = threadIdx.x % 6, blk = k / 3, row = k - 3 * blk;
            const double a = row == 0 ? m0 : row == 1 ? m3 : m6, b = row == 0 ? m1 : row == 1 ? m4 : m7, c = row == 0 ? m2 : row == 1 ? m5 : m8;
            T6 = __dadd_rn(__dadd_rn(__dmul_rn(a, (double)(xt + 64 * blk)), __dmul_rn(b, dy)), c);
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int x = xt + 32 * i + threadIdx.x;
            // OpenCV evaluates the row terms at the origin of a block that is 64 px wide (1024 / min(16, rows)); the
            // SAFE_RCP instances are only launched for 64-px blocks, where a thread's pixels 0,1 / 2,3 share an origin
            const int xbi = SAFE_RCP ? xt + 64 * (i >> 1) : (block_w == 64 ? (x & ~63) : x - x % block_w);
            if (SAFE_RCP) {
                if ((i & 1) == 0) {
                    xb = xbi;
                    X0 = __shfl_sync(0xffffffffu, T6, 3 * (i >> 1));
                    Y0 = __shfl_sync(0xffffffffu, T6, 3 * (i >> 1) + 1);
                    W0 = __shfl_sync(0xffffffffu, T6, 3 * (i >> 1) + 2);
                }
            } else if (xbi != xb) {
                xb = xbi;
                const double dxb = (double)xb;
                X0 = __dadd_rn(__dadd_rn(__dmul_rn(m0, dxb), __dmul_rn(m1, dy)), m2);
                Y0 = __dadd_rn(__dadd_rn(__dmul_rn(m3, dxb), __dmul_rn(m4, dy)), m5);
                W0 = __dadd_rn(__dadd_rn(__dmul_rn(m6, dxb), __dmul_rn(m7, dy)), m8);
            }
            const double x1 = (double)(x - xb);
            double W = __dadd_rn(W0, __dmul_rn(m6, x1));
            if (SAFE_RCP) W = __dmul_rn(__drcp_rn(W), 32.0);
            else W = W != 0.0 ? __ddiv_rn(32.0, W) : 0.0;
            const double fX = __dmul_rn(__dadd_rn(X0, __dmul_rn(m0, x1)), W), fY = __dmul_rn(__dadd_rn(Y0, __dmul_rn(m3, x1)), W);
            // SAFE_RCP: finite matrix and a bounded, non-zero denominator — the products cannot be NaN, and the conversion
            // instruction saturates like OpenCV's clamp
            Xs[i] = SAFE_RCP ? __double2int_rn(fX) : round_clamped(fX);
            Ys[i] = SAFE_RCP ? __double2int_rn(fY) : round_clamped(fY);
        }
    }
    uint2 lo[4][2], hi[4][2];
    uint32_t sft[4][2];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int sx = ds_clamp(Xs[i] >> 5, rx0 + 3, rx1 - 6) - rx0, sy = Ys[i] >> 5;
        const int cy0 = ds_clamp(sy, ry0, ry1 - 1);
        const uintptr_t A0 = reinterpret_cast<uintptr_t>(src + ((uint32_t)(cy0 - ry0) * (uint32_t)sp + 3u * (uint32_t)sx));      // < 2^32: checked by the host
        const uint32_t step = (sy >= ry0 && sy + 1 < ry1) ? (uint32_t)sp : 0u;               // row clamp(sy + 1)
        const uintptr_t B0 = A0 & ~(uintptr_t)7, B1 = P8 ? B0 + step : (A0 + step) & ~(uintptr_t)7;
        lo[i][0] = __ldg(reinterpret_cast<const uint2*>(B0)); hi[i][0] = __ldg(reinterpret_cast<const uint2*>(B0) + 1);
        lo[i][1] = __ldg(reinterpret_cast<const uint2*>(B1)); hi[i][1] = __ldg(reinterpret_cast<const uint2*>(B1) + 1);
        sft[i][0] = (uint32_t)(A0 & 7);
        sft[i][1] = P8 ? sft[i][0] : (uint32_t)((A0 + step) & 7);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int x = xt + 32 * i + threadIdx.x;
        if (x >= dw) break;
        const int sx = Xs[i] >> 5, sy = Ys[i] >> 5, ax = Xs[i] & 31, ay = Ys[i] & 31;
        const bool y0in = (unsigned)sy < (unsigned)sh, y1in = (unsigned)(sy + 1) < (unsigned)sh;
        int acc[3];
        if (sx >= rx0 + 3 && sx <= rx1 - 6) {
            // (32-ax)(32-ay)32 p00 + ... == 32 * [(32-ay) h0 + ay h1] exactly, so (.. + 2^14) >> 15 == (v + 512) >> 10
            // The four weights / 32 are at most 1024: they ride as 16-bit pairs (left, right) of one word per source row and
            // dp2a applies a pair to two pixel bytes — 6 dot products and 3 byte permutes per pixel, no separate vertical pass.
            const uint32_t wxx = 32u + 0xffffu * (uint32_t)ax;              // (32 - ax) | ax << 16
            const uint32_t wt = (y0in ? 32u - (uint32_t)ay : 0u) * wxx;      // rows outside the image: BORDER_CONSTANT 0
            const uint32_t wb = (y1in ? (uint32_t)ay : 0u) * wxx;
            uint32_t b0[2], gr[2];
#pragma unroll
            for (int rr = 0; rr < 2; rr++) {
                const bool up = sft[i][rr] >= 4;
                const uint32_t wa = up ? lo[i][rr].y : lo[i][rr].x, wm = up ? hi[i][rr].x : lo[i][rr].y, wc = up ? hi[i][rr].y : hi[i][rr].x;
                b0[rr] = __funnelshift_r(wa, wm, 8 * sft[i][rr]);                                     // B0 G0 R0 B1   (shift mod 32)
                gr[rr] = __byte_perm(b0[rr], __funnelshift_r(wm, wc, 8 * sft[i][rr]), 0x5241);       // G0 G1 R0 R1
            }
            const uint32_t bb = __byte_perm(b0[0], b0[1], 0x7430);                                    // B0 B1 of row 0, B0 B1 of row 1
            acc[0] = (int)(__dp2a_hi(wb, bb, __dp2a_lo(wt, bb, 512u)) >> 10);
            acc[1] = (int)(__dp2a_lo(wb, gr[1], __dp2a_lo(wt, gr[0], 512u)) >> 10);
            acc[2] = (int)(__dp2a_hi(wb, gr[1], __dp2a_hi(wt, gr[0], 512u)) >> 10);
        } else {
            const int w00 = (32 - ax) * (32 - ay) * 32, w01 = ax * (32 - ay) * 32, w10 = (32 - ax) * ay * 32, w11 = ax * ay * 32;
            const bool x0in = (unsigned)sx < (unsigned)sw, x1in = (unsigned)(sx + 1) < (unsigned)sw;
            const int cx0 = ds_clamp(sx, rx0, rx1 - 1) - rx0, cx1 = ds_clamp(sx + 1, rx0, rx1 - 1) - rx0;
            const int v00 = (x0in && y0in) ? w00 : 0, v01 = (x1in && y0in) ? w01 : 0;
            const int v10 = (x0in && y1in) ? w10 : 0, v11 = (x1in && y1in) ? w11 : 0;
            const uint8_t* r0 = src + (size_t)(ds_clamp(sy, ry0, ry1 - 1) - ry0) * sp;
            const uint8_t* r1 = src + (size_t)(ds_clamp(sy + 1, ry0, ry1 - 1) - ry0) * sp;
#pragma unroll
            for (int c = 0; c < 3; c++)
                acc[c] = (16384 + v00 * __ldg(r0 + cx0 * 3 + c) + v01 * __ldg(r0 + cx1 * 3 + c) +
                          v10 * __ldg(r1 + cx0 * 3 + c) + v11 * __ldg(r1 + cx1 * 3 + c)) >> 15;
        }
        uint8_t* dp = drow + (size_t)x * 3;
        dp[0] = (uint8_t)acc[0]; dp[1] = (uint8_t)acc[1]; dp[2] = (uint8_t)acc[2];      // <= 255 by construction
        if (grow) grow[x] = gray15(acc[0], acc[1], acc[2]);
    }
}

// ---- 1-D bulk copies global -> shared (the TMA unit's cp.async.bulk) completing on an mbarrier ----
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_addr(bar)), "r"(parity) : "memory");
    } while (!done);
}

// Tile-staged variant of warp_perspective3_kernel<true, true> for 16-byte aligned sources with 8-byte-multiple pitches: a CTA owns a 64 x 8 destination
// tile (one OpenCV coordinate block wide).  After the coordinates are known the CTA reduces the bounding box of its source
// taps, warp 0 requests the box row by row with bulk copies (no per-byte instructions, full-line requests instead of
// 8-byte gathers) and every thread filters its 4 pixels out of shared memory with the same 16-byte-window arithmetic.  The
// gathers then see shared-memory latency instead of L2 / HBM latency.  A box that does not fit `cap` bytes (steep
// perspective) or would run past the last resident row is read straight from global memory with the same code: the
// loads are generic, only their base and pitch change.  Arithmetic identical to the other instances.
constexpr int TILE_W = 64, TILE_H = 8;
__global__ void __launch_bounds__(128, 8) warp_perspective3_tile_kernel(const WarpPJob* __restrict__ jobs, const int cap) {
    extern __shared__ __align__(128) uint8_t s_tile[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ int s_box[4][4];
    const WarpPJob& J = jobs[blockIdx.z];
    const int x0 = blockIdx.x * TILE_W, y0 = blockIdx.y * TILE_H;
    if (x0 >= J.dw || y0 >= J.dh) return;
    const int lane = threadIdx.x, wp = threadIdx.y;
    if (lane == 0 && wp == 0) mbar_init(&s_bar, 1);
    const int sh = J.sh, sp = J.src_pitch;
    const uint8_t* __restrict__ src = J.src;
    // pixel k of a thread: row y0 + wp + 4 (k >> 1), column x0 + lane + 32 (k & 1)
    int Xs[4], Ys[4];
    {
        const double m0 = J.m[0], m1 = J.m[1], m2 = J.m[2], m3 = J.m[3], m4 = J.m[4], m5 = J.m[5], m6 = J.m[6], m7 = J.m[7], m8 = J.m[8];
        const double dxb = (double)x0;
        const double xa = (double)lane, xb = (double)(lane + 32);
        const double ax0 = __dmul_rn(m0, xa), ax1 = __dmul_rn(m0, xb), ay0 = __dmul_rn(m3, xa), ay1 = __dmul_rn(m3, xb);
        const double aw0 = __dmul_rn(m6, xa), aw1 = __dmul_rn(m6, xb);
#pragma unroll
        for (int rr = 0; rr < 2; rr++) {
            const double dy = (double)(y0 + wp + 4 * rr);
            const double X0 = __dadd_rn(__dadd_rn(__dmul_rn(m0, dxb), __dmul_rn(m1, dy)), m2);
            const double Y0 = __dadd_rn(__dadd_rn(__dmul_rn(m3, dxb), __dmul_rn(m4, dy)), m5);
            const double W0 = __dadd_rn(__dadd_rn(__dmul_rn(m6, dxb), __dmul_rn(m7, dy)), m8);
#pragma unroll
            for (int c = 0; c < 2; c++) {
                const double W = __dmul_rn(__drcp_rn(__dadd_rn(W0, c ? aw1 : aw0)), 32.0);      // == 32 / W (host-proved safe)
                Xs[2 * rr + c] = __double2int_rn(__dmul_rn(__dadd_rn(X0, c ? ax1 : ax0), W));
                Ys[2 * rr + c] = __double2int_rn(__dmul_rn(__dadd_rn(Y0, c ? ay1 : ay0), W));
            }
        }
    }
    // clamped tap origin of every pixel (byte offset inside the resident row, resident row indices) and their bounding box
    int ao[4], c0[4], c1[4];
    int amin = INT_MAX, amax = INT_MIN, rmin = INT_MAX, rmax = INT_MIN;
    const int nres = J.ry1 - J.ry0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const bool valid = y0 + wp + 4 * (k >> 1) < J.dh && x0 + lane + 32 * (k & 1) < J.dw;
        ao[k] = 3 * (ds_clamp(Xs[k] >> 5, J.rx0 + 3, J.rx1 - 6) - J.rx0);
        const int sy = Ys[k] >> 5;
        c0[k] = ds_clamp(sy, J.ry0, J.ry1 - 1) - J.ry0;
        c1[k] = ds_clamp(sy + 1, J.ry0, J.ry1 - 1) - J.ry0;
        if (valid) { amin = min(amin, ao[k]); amax = max(amax, ao[k]); rmin = min(rmin, c0[k]); rmax = max(rmax, c1[k]); }
    }
    amin = __reduce_min_sync(0xffffffffu, amin); amax = __reduce_max_sync(0xffffffffu, amax);
    rmin = __reduce_min_sync(0xffffffffu, rmin); rmax = __reduce_max_sync(0xffffffffu, rmax);
    if (lane == 0) { s_box[wp][0] = amin; s_box[wp][1] = amax; s_box[wp][2] = rmin; s_box[wp][3] = rmax; }
    __syncthreads();                                   // also publishes the initialised barrier
#pragma unroll
    for (int w = 0; w < 4; w++) {
        amin = min(amin, s_box[w][0]); amax = max(amax, s_box[w][1]); rmin = min(rmin, s_box[w][2]); rmax = max(rmax, s_box[w][3]);
    }
    // box in bytes [bx0, bx1) x rows [rmin, rmax]: every 16-byte load window [(a & ~7), +16) of the tile lies inside.
    // A bulk copy needs a 16-byte aligned source: the source base is (host-checked), its pitch only a multiple of 8, so row
    // r starts `(src + r * pitch) & 15` (0 or 8) bytes into its shared-memory row and 16 bytes more are copied per row.
    const int bx0 = amin & ~15, bx1 = ((amax & ~7) + 16 + 15) & ~15;
    const int pitch_s = bx1 - bx0 + 16, nrows = rmax - rmin + 1;
    const bool staged = (long long)nrows * pitch_s <= cap && (bx1 + 16 <= 3 * (J.rx1 - J.rx0) || rmax + 1 < nres);
    const uint32_t src_lo = (uint32_t)reinterpret_cast<uintptr_t>(src);
    if (staged) {
        if (wp == 0) {
            if (lane == 0) mbar_arrive_expect_tx(&s_bar, (uint32_t)(nrows * pitch_s));
            __syncwarp();
            for (int r = lane; r < nrows; r += 32) {
                const uintptr_t g = reinterpret_cast<uintptr_t>(src + (size_t)(rmin + r) * sp + bx0) & ~(uintptr_t)15;
                bulk_g2s(s_tile + r * pitch_s, reinterpret_cast<const void*>(g), (uint32_t)pitch_s, &s_bar);
            }
        }
        mbar_wait(&s_bar, 0);
    }
    const uintptr_t gbase = staged ? reinterpret_cast<uintptr_t>(s_tile) - (uintptr_t)bx0 : reinterpret_cast<uintptr_t>(src);
    // byte offset of resident row r from gbase
    auto row_off = [&](int r) -> uint32_t {
        return staged ? (uint32_t)((r - rmin) * pitch_s) + ((src_lo + (uint32_t)r * (uint32_t)sp) & 15u) : (uint32_t)r * (uint32_t)sp;
    };
    uint2 lo[4][2], hi[4][2];
    uint32_t sft[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const bool valid = y0 + wp + 4 * (k >> 1) < J.dh && x0 + lane + 32 * (k & 1) < J.dw;
        const int a = valid ? ao[k] : amin, r0 = valid ? c0[k] : rmin, r1 = valid ? c1[k] : rmin;      // lanes past the page read a harmless tap
        const uintptr_t A0 = gbase + (row_off(r0) + (uint32_t)a);
        const uintptr_t B0 = A0 & ~(uintptr_t)7, B1 = (gbase + (row_off(r1) + (uint32_t)a)) & ~(uintptr_t)7;   // same offset mod 8 in both rows
        lo[k][0] = *reinterpret_cast<const uint2*>(B0); hi[k][0] = *(reinterpret_cast<const uint2*>(B0) + 1);
        lo[k][1] = *reinterpret_cast<const uint2*>(B1); hi[k][1] = *(reinterpret_cast<const uint2*>(B1) + 1);
        sft[k] = (uint32_t)(A0 & 7);
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int y = y0 + wp + 4 * (k >> 1), x = x0 + lane + 32 * (k & 1);
        if (y >= J.dh || x >= J.dw) continue;
        const int sx = Xs[k] >> 5, sy = Ys[k] >> 5, ax = Xs[k] & 31, ay = Ys[k] & 31;
        const bool y0in = (unsigned)sy < (unsigned)sh, y1in = (unsigned)(sy + 1) < (unsigned)sh;
        int acc[3];
        if (sx >= J.rx0 + 3 && sx <= J.rx1 - 6) {
            const uint32_t wxx = 32u + 0xffffu * (uint32_t)ax;              // (32 - ax) | ax << 16
            const uint32_t wt = (y0in ? 32u - (uint32_t)ay : 0u) * wxx;      // rows outside the image: BORDER_CONSTANT 0
            const uint32_t wb = (y1in ? (uint32_t)ay : 0u) * wxx;
            uint32_t b0[2], gr[2];
            const bool up = sft[k] >= 4;
#pragma unroll
            for (int rr = 0; rr < 2; rr++) {
                const uint32_t wa = up ? lo[k][rr].y : lo[k][rr].x, wm = up ? hi[k][rr].x : lo[k][rr].y, wc = up ? hi[k][rr].y : hi[k][rr].x;
                b0[rr] = __funnelshift_r(wa, wm, 8 * sft[k]);                                     // B0 G0 R0 B1   (shift mod 32)
                gr[rr] = __byte_perm(b0[rr], __funnelshift_r(wm, wc, 8 * sft[k]), 0x5241);       // G0 G1 R0 R1
            }
            const uint32_t bb = __byte_perm(b0[0], b0[1], 0x7430);                                // B0 B1 of row 0, B0 B1 of row 1
            acc[0] = (int)(__dp2a_hi(wb, bb, __dp2a_lo(wt, bb, 512u)) >> 10);
            acc[1] = (int)(__dp2a_lo(wb, gr[1], __dp2a_lo(wt, gr[0], 512u)) >> 10);
            acc[2] = (int)(__dp2a_hi(wb, gr[1], __dp2a_hi(wt, gr[0], 512u)) >> 10);
        } else {
            const int sw = J.sw;
            const int w00 = (32 - ax) * (32 - ay) * 32, w01 = ax * (32 - ay) * 32, w10 = (32 - ax) * ay * 32, w11 = ax * ay * 32;
            const bool x0in = (unsigned)sx < (unsigned)sw, x1in = (unsigned)(sx + 1) < (unsigned)sw;
            const int cx0 = ds_clamp(sx, J.rx0, J.rx1 - 1) - J.rx0, cx1 = ds_clamp(sx + 1, J.rx0, J.rx1 - 1) - J.rx0;
            const int v00 = (x0in && y0in) ? w00 : 0, v01 = (x1in && y0in) ? w01 : 0;
            const int v10 = (x0in && y1in) ? w10 : 0, v11 = (x1in && y1in) ? w11 : 0;
            const uint8_t* r0 = src + (size_t)c0[k] * sp;
            const uint8_t* r1 = src + (size_t)c1[k] * sp;
#pragma unroll
            for (int c = 0; c < 3; c++)
                acc[c] = (16384 + v00 * __ldg(r0 + cx0 * 3 + c) + v01 * __ldg(r0 + cx1 * 3 + c) +
                          v10 * __ldg(r1 + cx0 * 3 + c) + v11 * __ldg(r1 + cx1 * 3 + c)) >> 15;
        }
        uint8_t* dp = J.dst + (size_t)y * J.dst_pitch + (size_t)x * 3;
        dp[0] = (uint8_t)acc[0]; dp[1] = (uint8_t)acc[1]; dp[2] = (uint8_t)acc[2];      // <= 255 by construction
        if (J.gray) J.gray[(size_t)y * J.gray_pitch + x] = gray15(acc[0], acc[1], acc[2]);
    }
}

// cv::warpAffine precomputes adelta[x] = saturate_cast<int>(M[0]*x*1024), bdelta[x] = saturate_cast<int>(M[3]*x*1024)
// once per call; so do we (one tiny kernel), which keeps fp64 out of the per-pixel loop.  The same kernel tabulates the
// row origins X0[y] = saturate_cast<int>((M[1]*y + M[2])*1024) + 16 (and Y0) behind the column table: rows start at
// delta[dw + 4].
__global__ void affine_delta_kernel(const WarpAJob* __restrict__ jobs) {
    const WarpAJob& J = jobs[blockIdx.y];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const double dt = (double)t;
    if (t < J.dw)
        J.delta[t] = make_int2(round_clamped(__dmul_rn(__dmul_rn(J.m[0], dt), 1024.0)),
                               round_clamped(__dmul_rn(__dmul_rn(J.m[3], dt), 1024.0)));
    if (t < J.dh)
        J.delta[J.dw + 4 + t] = make_int2(round_clamped(__dmul_rn(__dadd_rn(__dmul_rn(J.m[1], dt), J.m[2]), 1024.0)) + 16,
                                          round_clamped(__dmul_rn(__dadd_rn(__dmul_rn(J.m[4], dt), J.m[5]), 1024.0)) + 16);
}

// One warp = 128 destination pixels of a row, lane-interleaved (see warp_perspective_kernel): compact gathers, coalesced
// byte stores.  The kernel runs at 90 % occupancy, so the per-pixel chain table -> address -> taps -> filter is hidden by
// the other warps; batching a thread's 16 gathers ahead of the filter (as warp_perspective3_kernel does) only costs
// registers here (measured: 1.05 -> 1.10 ms).  Interior pixels skip the replicate clamps.
__global__ void __launch_bounds__(128) warp_affine_kernel(const WarpAJob* __restrict__ jobs) {
    const WarpAJob& J = jobs[blockIdx.z];
    const int y = blockIdx.y * 4 + threadIdx.y;
    const int xt = blockIdx.x * 128;
    const int dw = J.dw;
    if (y >= J.dh || xt >= dw) return;
    const int2* __restrict__ tab = J.delta;
    const int2 org = __ldg(tab + dw + 4 + y);          // cv::warpAffine's X0, Y0 of this row
    const uint8_t* __restrict__ src = J.src;
    const int sw = J.sw, sh = J.sh, sp = J.src_pitch;
    const int xl = xt + threadIdx.x;
    uint8_t* const dpx = J.dst + (size_t)y * J.dst_pitch + xl;
    const int2* const dlp = tab + xl;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if (xl + 32 * i >= dw) break;
        const int2 dl = __ldg(dlp + 32 * i);
        const int X = (org.x + dl.x) >> 5, Y = (org.y + dl.y) >> 5;
        // (OpenCV keeps sx, sy as saturated shorts; with sources below 32767 px the replicate clamp gives the same taps)
        const int sx = X >> 5, sy = Y >> 5;
        const int ax = X & 31, ay = Y & 31;
        int h0, h1;
        // (32-ax)(32-ay)32 p00 + ... == 32 * [(32-ay) * ((32-ax) p00 + ax p01) + ay * ((32-ax) p10 + ax p11)] exactly
        if ((unsigned)sx < (unsigned)(sw - 1) && (unsigned)sy < (unsigned)(sh - 1)) {      // interior: no clamping
            const uint8_t* r0 = src + ((uint32_t)sy * (uint32_t)sp + (uint32_t)sx);          // < 2^32: checked by the host
            const uint8_t* r1 = r0 + (uint32_t)sp;
            h0 = (32 - ax) * __ldg(r0) + ax * __ldg(r0 + 1);
            h1 = (32 - ax) * __ldg(r1) + ax * __ldg(r1 + 1);
        } else {
            const int xa = ds_clamp(sx, 0, sw - 1), xb = ds_clamp(sx + 1, 0, sw - 1);
            const int ya = ds_clamp(sy, 0, sh - 1), yb = ds_clamp(sy + 1, 0, sh - 1);
            const uint8_t* r0 = src + (size_t)ya * sp;
            const uint8_t* r1 = src + (size_t)yb * sp;
            h0 = (32 - ax) * __ldg(r0 + xa) + ax * __ldg(r0 + xb);
            h1 = (32 - ax) * __ldg(r1 + xa) + ax * __ldg(r1 + xb);
        }
        dpx[32 * i] = (uint8_t)(((32 - ay) * h0 + ay * h1 + 512) >> 10);
    }
}

}  // namespace

int k_warp_perspective_jobs(docscan_ctx* ctx, const WarpPJob* jobs_host, int n, int max_w, int max_h) {
    void* dev = nullptr;
    DS_TRY(ds_upload(ctx, jobs_host, sizeof(WarpPJob) * n, &dev));
    const int ch = jobs_host[0].ch;
    for (int i = 0; i < n; i++) {
        if (jobs_host[i].ch != ch) return ds_fail(ctx, DOCSCAN_ERR_BAD_ARG, "mixed channel counts in one warp batch");
        const WarpPJob& j = jobs_host[i];
        if (j.sw >= 32767 || j.sh >= 32767)      // cv::remap's own limit (coordinates are shorts)
            return ds_fail(ctx, DOCSCAN_ERR_UNSUPPORTED, "warp source larger than 32766 px");
        if (j.rx0 < 0 || j.ry0 < 0 || j.rx1 > j.sw || j.ry1 > j.sh || j.rx1 <= j.rx0 || j.ry1 <= j.ry0)
            return ds_fail(ctx, DOCSCAN_ERR_BAD_ARG, "warp: bad resident region");
    }
    dim3 grid((max_w + 127) / 128, (max_h + 3) / 4, n), block(32, 4);
    // algorithmic bytes: source pixels under the quad, at most 4 taps per output pixel, read once;
    // destination (and the fused gray plane) written once
    double bytes = 0;
    for (int i = 0; i < n; i++) {
        const WarpPJob& j = jobs_host[i];
        const double cx[4] = {0, (double)j.dw, (double)j.dw, 0}, cy[4] = {0, 0, (double)j.dh, (double)j.dh};
        double sx[4], sy[4];
        for (int c = 0; c < 4; c++) {
            const double w = j.m[6] * cx[c] + j.m[7] * cy[c] + j.m[8];
            sx[c] = (j.m[0] * cx[c] + j.m[1] * cy[c] + j.m[2]) / w;
            sy[c] = (j.m[3] * cx[c] + j.m[4] * cy[c] + j.m[5]) / w;
        }
        double area = 0;
        for (int c = 0; c < 4; c++) area += sx[c] * sy[(c + 1) & 3] - sx[(c + 1) & 3] * sy[c];
        area = fabs(area) * 0.5;
        const double np = (double)j.dw * j.dh;
        const double src_px = fmin(fmin(area, 4.0 * np), (double)j.sw * j.sh);
        bytes += j.ch * (src_px + np) + (j.gray ? np : 0.0);
    }
    ProfScope prof(ctx, ch == 3 ? "warp_perspective_c3" : "warp_perspective_c1", bytes);
    bool wide = ch == 3;
    for (int i = 0; i < n; i++)      // the fast kernel keeps byte offsets inside the resident region in 32 bits
        wide = wide && jobs_host[i].rx1 - jobs_host[i].rx0 >= 9 && jobs_host[i].src_pitch > 0 &&
               (unsigned long long)(jobs_host[i].ry1 - jobs_host[i].ry0) * (unsigned long long)jobs_host[i].src_pitch < 0xffff0000ull;
    bool safe_rcp = wide;
    for (int i = 0; i < n && safe_rcp; i++) {
        const WarpPJob& j = jobs_host[i];
        double lo = 1e308, hi = -1e308;
        const double cx[2] = {0.0, (double)(((j.dw + 127) / 128) * 128)}, cy[2] = {0.0, (double)(((j.dh + 7) / 8) * 8)};     // lanes past dw / dh compute too
        for (int a = 0; a < 2; a++)
            for (int b = 0; b < 2; b++) {
                const double w = j.m[6] * cx[a] + j.m[7] * cy[b] + j.m[8];
                lo = fmin(lo, w); hi = fmax(hi, w);
            }
        bool finite = std::isfinite(lo) && std::isfinite(hi);
        for (int e = 0; e < 9; e++) finite = finite && std::isfinite(j.m[e]) && std::fabs(j.m[e]) < 1e150;
        safe_rcp = finite && j.block_w == 64 && ((lo > 1e-290 && hi < 1e290) || (hi < -1e-290 && lo > -1e290));
    }
    bool p8 = true;
    for (int i = 0; i < n; i++) p8 = p8 && jobs_host[i].src_pitch % 8 == 0;
    bool a16 = p8;                                   // bulk copies: 16-byte aligned source base (rows may start 8 bytes off)
    for (int i = 0; i < n; i++) a16 = a16 && (reinterpret_cast<uintptr_t>(jobs_host[i].src) & 15) == 0;
    // Opt-in (DOCSCAN_WARP_TILE=1): measured 3.58 ms against 2.65 ms for the gather kernel on the benchmark batch.  Same DRAM
    // traffic, higher issue rate, but 47 % more instructions (bounding-box reduction, per-row offsets, the barrier wait) than
    // the gathers' latency costs, which 32 resident warps already hide (profiles/r1_ncu_tile_warp.txt).  Kept, parity-tested.
    const char* tile_env = getenv("DOCSCAN_WARP_TILE");
    const int tile_mode = tile_env ? atoi(tile_env) : 0;
    if (wide && safe_rcp && a16 && tile_mode) {
        constexpr int cap = 24 * 1024;
        dim3 tgrid((max_w + TILE_W - 1) / TILE_W, (max_h + TILE_H - 1) / TILE_H, n);
        warp_perspective3_tile_kernel<<<tgrid, block, cap, ctx->stream>>>((const WarpPJob*)dev, cap);
    } else if (wide && safe_rcp && p8) warp_perspective3_kernel<true, true><<<grid, block, 0, ctx->stream>>>((const WarpPJob*)dev);
    else if (wide && safe_rcp) warp_perspective3_kernel<true, false><<<grid, block, 0, ctx->stream>>>((const WarpPJob*)dev);
    else if (wide) warp_perspective3_kernel<false, false><<<grid, block, 0, ctx->stream>>>((const WarpPJob*)dev);
    else if (ch == 3) warp_perspective_kernel<3><<<grid, block, 0, ctx->stream>>>((const WarpPJob*)dev);
    else warp_perspective_kernel<1><<<grid, block, 0, ctx->stream>>>((const WarpPJob*)dev);
    DS_CHECK_LAUNCH(ctx);
    return DOCSCAN_OK;
}

int k_warp_affine_upload(docscan_ctx* ctx, const WarpAJob* jobs_in, int n, WarpAJob** jobs_dev) {
    std::vector<WarpAJob> jobs(jobs_in, jobs_in + n);
    for (int i = 0; i < n; i++) {
        if (jobs[i].sw >= 32767 || jobs[i].sh >= 32767)      // cv::remap's own limit (coordinates are shorts)
            return ds_fail(ctx, DOCSCAN_ERR_UNSUPPORTED, "warp source larger than 32766 px");
        if (jobs[i].src_pitch <= 0 || (unsigned long long)jobs[i].sh * (unsigned long long)jobs[i].src_pitch >= 0xffff0000ull)
            return ds_fail(ctx, DOCSCAN_ERR_UNSUPPORTED, "warp source plane of 4 GiB or more");
        void* t = nullptr;
        DS_TRY(ds_arena_alloc(ctx, sizeof(int2) * ((size_t)jobs[i].dw + 4 + (size_t)jobs[i].dh), &t));      // columns, then rows
        jobs[i].delta = (int2*)t;
    }
    void* dev = nullptr;
    DS_TRY(ds_upload(ctx, jobs.data(), sizeof(WarpAJob) * n, &dev));
    *jobs_dev = (WarpAJob*)dev;
    return DOCSCAN_OK;
}

int k_warp_affine_launch(docscan_ctx* ctx, const WarpAJob* jobs_dev, const WarpAJob* jobs_host, int n, int max_w, int max_h) {
    {
        ProfScope prof(ctx, "warp_affine_deltas", 0);
        affine_delta_kernel<<<dim3((std::max(max_w, max_h) + 127) / 128, n), 128, 0, ctx->stream>>>(jobs_dev);
        DS_CHECK_LAUNCH(ctx);
    }
    dim3 grid((max_w + 127) / 128, (max_h + 3) / 4, n), block(32, 4);
    double px = 0;
    for (int i = 0; i < n; i++) px += (double)jobs_host[i].dw * jobs_host[i].dh;
    ProfScope prof(ctx, "warp_affine", 2.0 * px);
    warp_affine_kernel<<<grid, block, 0, ctx->stream>>>(jobs_dev);
    DS_CHECK_LAUNCH(ctx);
    return DOCSCAN_OK;
}

int k_warp_affine_jobs(docscan_ctx* ctx, const WarpAJob* jobs_in, int n, int max_w, int max_h) {
    WarpAJob* dev = nullptr;
    DS_TRY(k_warp_affine_upload(ctx, jobs_in, n, &dev));
    return k_warp_affine_launch(ctx, dev, jobs_in, n, max_w, max_h);
}
