// cv2.adaptiveThreshold(..., ADAPTIVE_THRESH_GAUSSIAN_C, THRESH_BINARY, k, C) (DocScanner.py:167) and the
// ink-mask combine + masked blend that follows it (DocScanner.py:207-212, 338-339).
//
// GAUSSIAN_C is the one floating-point stage of the path: OpenCV converts the page to fp32, runs a
// separable fp32 Gaussian (BORDER_REPLICATE) and rounds the mean to uint8.  The result depends on the
// order of the fp32 operations, so the kernel reproduces it: row taps left to right with one fma per
// tap, column taps as symmetric pairs (a rounded add, then an fma), outermost pair last.
// Same marching layout as blur.cu: 128-column strips, 16 rows per step, fp32 ring of row-filtered
// rows in shared memory; row pass register-blocked 16 outputs/thread, column pass one column and 16
// rows per thread with the whole column window held in registers.
#include <cmath>

#include "common.cuh"

namespace {

constexpr int TW = 128, BR = 16, NT = 128;
constexpr int RPF = 132;      // ring row pitch in floats
constexpr int GMAX = 129;     // half-kernel table size (k <= 257)

// The half kernel travels in the launch parameters: every tap index below is a compile-time constant, so the
// coefficient of each fma is a constant-bank operand (no register, no shared-memory load).
struct AdaptLaunch {
    int k, r, c_param, seg_rows, spf, tail_compat;
    float gh[GMAX];           // gh[j] = g[r + j] for j <= r, 0 beyond (zero taps leave an fp32 accumulator unchanged)
};

// rare path of the staging load (strip edges, unaligned caller buffers): kept out of line
__device__ __noinline__ uint32_t fetch_word_clamped(const uint8_t* rowp, int gx, int w) {
    uint32_t word = 0;
    for (int b = 0; b < 4; b++) word |= (uint32_t)rowp[ds_clamp(gx + b, 0, w - 1)] << (8 * b);
    return word;
}

// Radii up to 17 (every preset) are compiled for five CTAs per SM: 96 registers without a spill, and 5 x (stage + ring) =
// 5 x 44.6 KB of shared memory just fit beside the 1 KB the system keeps per CTA.  The kernel waits on latencies (two barriers
// per 16-row step, shared-memory round trips), so the fifth CTA's four warps are worth more than the registers.
template <int RMAX, int MINB>
__global__ void __launch_bounds__(NT, MINB) adaptive_gauss_kernel(const AdaptJob* __restrict__ jobs, const __grid_constant__ AdaptLaunch L) {
    constexpr int DELTA = (4 - (RMAX & 3)) & 3;                    // staged column 0 <-> x0 - RMAX - DELTA (word aligned)
    constexpr int STAGE_WORDS = (TW + 2 * RMAX + DELTA + 3) >> 2;
    constexpr int RR = ((2 * RMAX + BR - 1) / BR + 1) * BR;         // ring rows
    constexpr int D = (2 * RMAX + BR - 1) / BR;                     // the column pass lags the row pass by D steps
    const AdaptJob J = jobs[blockIdx.z];
    const int x0 = blockIdx.x * TW;
    const int y_begin = blockIdx.y * L.seg_rows;
    if (x0 >= J.w || y_begin >= J.h) return;
    const int y_end = min(J.h, y_begin + L.seg_rows);
    const int rows_out = y_end - y_begin;
    const int tid = threadIdx.x;

    extern __shared__ __align__(16) uint8_t smem_raw[];
    float* s_stage = reinterpret_cast<float*>(smem_raw);           // BR * spf, spf odd
    // the ring's rows are written as float4: its base is rounded up to 16 bytes
    float* s_ring = s_stage + ((BR * L.spf + 3) & ~3);             // RR * RPF

    const bool src_al = ((reinterpret_cast<uintptr_t>(J.src) | (uintptr_t)J.src_pitch) & 3) == 0;
    const int n_vb = (rows_out + BR - 1) / BR;
    const bool row_identity = J.w == 1, col_identity = J.h == 1;   // cv::GaussianBlur shrinks the kernel on 1-px axes
    // columns cv2's AVX2 build evaluates without fma (k <= 9: the taps are dyadic, every order is exact)
    const int tail = (L.tail_compat && L.k >= 11) ? (J.w & 7) : 0;
    const int xt_col = J.w - tail;                                  // column filter: mul+add from here on

    for (int hb = 0; hb < n_vb + D; hb++) {
        {
            // virtual row v of the segment <-> source row clamp(y_begin - RMAX + v): taps beyond r carry zero weight
            const int srow_id = tid >> 3;              // 16 rows x 8 lanes; a lane strides along its row
            const uint8_t* rowp = J.src + (size_t)ds_clamp(y_begin - RMAX + hb * BR + srow_id, 0, J.h - 1) * J.src_pitch;
            float* sp = s_stage + srow_id * L.spf;
            // interior strips (all but the first and last of a page) skip the per-word tests
            const int gx0 = x0 - RMAX - DELTA;
            const bool interior = src_al && gx0 >= 0 && gx0 + 4 * STAGE_WORDS <= J.w;
            constexpr int NJ = (STAGE_WORDS + 7) / 8;
            const int l8 = tid & 7;
            uint32_t wv[NJ];                               // issue the loads first, convert and store afterwards
            if (interior) {
                const uint32_t* rp32 = reinterpret_cast<const uint32_t*>(rowp + gx0) + l8;
#pragma unroll
                for (int j = 0; j < NJ; j++) wv[j] = (8 * j + 7 < STAGE_WORDS || l8 + 8 * j < STAGE_WORDS) ? __ldg(rp32 + 8 * j) : 0u;
            } else {
#pragma unroll
                for (int j = 0; j < NJ; j++) {
                    const int wi = l8 + 8 * j;
                    const int gx = gx0 + 4 * wi;
                    wv[j] = 0;
                    if (wi < STAGE_WORDS) wv[j] = (src_al && gx >= 0 && gx + 3 < J.w) ? ds_ldg32(rowp + gx) : fetch_word_clamped(rowp, gx, J.w);
                }
            }
#pragma unroll
            for (int j = 0; j < NJ; j++) {
                const int wi = l8 + 8 * j;
                if (8 * j + 7 >= STAGE_WORDS && wi >= STAGE_WORDS) continue;
                // u8 -> fp32 without the slow I2F unit: 2^23 + v is exact in fp32, subtract 2^23 again
#pragma unroll
                for (int b = 0; b < 4; b++) sp[4 * wi + b] = __fsub_rn(__uint_as_float(__byte_perm(wv[j], 0x4B000000u, 0x7440 + b)), 8388608.0f);
            }
        }
        __syncthreads();
        {   // ---- row pass: out[c] = sum_i g[i] * f[c + i - r], taps in increasing i, one fma each.
            // thread = (staged row, 16 consecutive outputs); lane <-> row (16 rows x 2 column groups per warp): with the
            // row pitch == 1 (mod 32) a warp's loads hit 32 different banks.  Each staged value is loaded once and fed to
            // every output it belongs to; inputs ascend, so each accumulator sees its taps in increasing order.
            const int lane = tid & 31, wrp = tid >> 5;
            const int hr = lane & 15;
            const int c0 = 16 * (2 * wrp + (lane >> 4));
            const float* sp = s_stage + hr * L.spf + DELTA + c0;      // sp[u] <-> column x0 + c0 + u - RMAX
            float acc[16];
#pragma unroll
            for (int o = 0; o < 16; o++) acc[o] = 0.0f;
#pragma unroll
            for (int u = 0; u < 16 + 2 * RMAX; u++) {
                const float f = sp[u];
#pragma unroll
                for (int o = 0; o < 16; o++) {
                    const int i = u - o;                              // tap index in the padded kernel
                    if (i >= 0 && i <= 2 * RMAX) acc[o] = __fmaf_rn(f, L.gh[i < RMAX ? RMAX - i : i - RMAX], acc[o]);
                }
            }
            if (row_identity) {
#pragma unroll
                for (int o = 0; o < 16; o++) acc[o] = sp[RMAX + o];
            }
            // (the few columns cv2's row filter evaluates with scalar code are redone by adaptive_tail_kernel afterwards:
            // their serial tap chains would otherwise hold up the whole CTA of the last strip)
            float4* dst = reinterpret_cast<float4*>(s_ring + ((hb * BR + hr) % RR) * RPF + c0);
            dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
            dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
            dst[2] = make_float4(acc[8], acc[9], acc[10], acc[11]);
            dst[3] = make_float4(acc[12], acc[13], acc[14], acc[15]);
        }
        __syncthreads();
        if (hb < D) continue;
        // ---- column pass: thread = one column, 16 rows; centre of output o sits at window index o + RMAX.
        // The ring size is a compile-time constant and a window starts at a multiple of 16 rows, so every ring offset
        // (wrap included) is an immediate.
        const int vb = hb - D;
        const int col = tid;
        const int x = x0 + col;
        float Wn[BR + 2 * RMAX];
        const int slot0 = (vb * BR) % RR;
        {
            const float* colp = s_ring + col;
#define DS_LOAD_WINDOW(S0)                                                                      \
    _Pragma("unroll") for (int i = 0; i < BR + 2 * RMAX; i++) Wn[i] = colp[(((S0) + i) % RR) * RPF];
            switch (slot0 / BR) {
                case 0: DS_LOAD_WINDOW(0) break;
                case 1: DS_LOAD_WINDOW(16) break;
                case 2: DS_LOAD_WINDOW(32) break;
                case 3: DS_LOAD_WINDOW(48) break;
                case 4: DS_LOAD_WINDOW(64) break;
                case 5: DS_LOAD_WINDOW(80) break;
                case 6: DS_LOAD_WINDOW(96) break;
                default: DS_LOAD_WINDOW(112) break;
            }
#undef DS_LOAD_WINDOW
        }
        float acc[BR];
#pragma unroll
        for (int o = 0; o < BR; o++) acc[o] = __fmul_rn(L.gh[0], Wn[o + RMAX]);
        if (x < xt_col) {
#pragma unroll
            for (int j = 1; j <= RMAX; j++) {
#pragma unroll
                for (int o = 0; o < BR; o++) acc[o] = __fmaf_rn(__fadd_rn(Wn[o + RMAX + j], Wn[o + RMAX - j]), L.gh[j], acc[o]);
            }
        } else {
#pragma unroll
            for (int j = 1; j <= RMAX; j++) {
#pragma unroll
                for (int o = 0; o < BR; o++) acc[o] = __fadd_rn(acc[o], __fmul_rn(L.gh[j], __fadd_rn(Wn[o + RMAX + j], Wn[o + RMAX - j])));
            }
        }
        if (x < J.w) {
            const int y0 = y_begin + vb * BR;
            const int rows = min(BR, y_end - y0);
            if (col_identity) {
#pragma unroll
                for (int o = 0; o < BR; o++) acc[o] = Wn[o + RMAX];
            }
            // the mean of uint8 data under a kernel that sums to 1 lies in [0, 255.0001]: adding 1.5 * 2^23 leaves the mean rounded
            // to nearest even in the mantissa, and  src - mean > -C  <=>  src + C + bits(1.5 * 2^23) > bits(mean + 1.5 * 2^23)
            const int cbias = L.c_param + 0x4B400000;
            const uint8_t* sp_ = J.src + (size_t)y0 * J.src_pitch + x;
            uint8_t* dp_ = J.dst + (size_t)y0 * J.dst_pitch + x;
            int cpx[BR];                                   // centre pixels, loaded up front
            if (rows == BR) {
#pragma unroll
                for (int o = 0; o < BR; o++) { cpx[o] = __ldg(sp_); sp_ += J.src_pitch; }
#pragma unroll
                for (int o = 0; o < BR; o++) {
                    *dp_ = (cpx[o] + cbias > __float_as_int(__fadd_rn(acc[o], 12582912.0f))) ? 255 : 0;
                    dp_ += J.dst_pitch;
                }
            } else {
#pragma unroll
                for (int o = 0; o < BR; o++) {
                    if (o >= rows) break;
                    dp_[(size_t)o * J.dst_pitch] = ((int)sp_[(size_t)o * J.src_pitch] + cbias > __float_as_int(__fadd_rn(acc[o], 12582912.0f))) ? 255 : 0;
                }
            }
        }
    }
}

// The same kernel on packed pairs (FFMA2 / FADD2 / FMUL2, sm_100): a packed instruction still occupies the fp32 pipe for
// two cycles (tests/tools/ffma2_probe.cu: 128 lanes per clock and SM either way) but takes one issue slot instead of two,
// and issue slots are what the scalar kernel runs out of (115 instructions per pixel, 71 of them the parity-fixed fp32
// operations).  Every lane of a packed operation is an IEEE fma / add / mul of its own, so the results stay bit-identical.
// Row pass: a thread's 16 outputs are 8 pairs two columns apart, (o, o + 2); the staged row is stored as pairs
// G[u] = (f[u], f[u + 2]) for EVERY u, so the operand of tap i of pair o is the aligned 8-byte entry G[o + i] as it comes out
// of shared memory (assembling pairs from neighbouring registers costs a MOV per operation: measured, 87 instructions
// per pixel).  Every accumulator still sees its taps in increasing order.  Column pass: thread = two adjacent columns x
// 8 rows, the window held as pairs.
template <int RMAX, int MINB>
__global__ void __launch_bounds__(NT, MINB) adaptive_gauss2_kernel(const AdaptJob* __restrict__ jobs, const __grid_constant__ AdaptLaunch L) {
    constexpr int DELTA = (4 - (RMAX & 3)) & 3;                    // staged column 0 <-> x0 - RMAX - DELTA (word aligned)
    constexpr int OFFS = DELTA & 1;                                 // entries start one late so that the row pass reads 16-byte aligned
    constexpr int STAGE_WORDS = (TW + 2 * RMAX + DELTA + 3) >> 2;
    constexpr int RR = ((2 * RMAX + BR - 1) / BR + 1) * BR;         // ring rows
    constexpr int D = (2 * RMAX + BR - 1) / BR;                     // the column pass lags the row pass by D steps
    constexpr int NG = 14 + 2 * RMAX;                               // entries G[0 .. NG - 1] of a thread's 16 outputs (an even count)
    constexpr int CR = 8;                                           // rows per thread in the column pass
    const AdaptJob J = jobs[blockIdx.z];
    const int x0 = blockIdx.x * TW;
    const int y_begin = blockIdx.y * L.seg_rows;
    if (x0 >= J.w || y_begin >= J.h) return;
    const int y_end = min(J.h, y_begin + L.seg_rows);
    const int rows_out = y_end - y_begin;
    const int tid = threadIdx.x;

    extern __shared__ __align__(16) uint8_t smem_raw[];
    float2* s_stage = reinterpret_cast<float2*>(smem_raw);         // BR rows of spf / 2 entries; spf / 4 odd: 128-bit loads of 8 rows hit 8 x 16 bytes
    float* s_ring = reinterpret_cast<float*>(smem_raw) + BR * L.spf;   // RR * RPF
    const int spe = L.spf >> 1;

    const bool src_al = ((reinterpret_cast<uintptr_t>(J.src) | (uintptr_t)J.src_pitch) & 3) == 0;
    const bool io16 = ((reinterpret_cast<uintptr_t>(J.src) | (uintptr_t)J.src_pitch | reinterpret_cast<uintptr_t>(J.dst) | (uintptr_t)J.dst_pitch) & 1) == 0;
    const int n_vb = (rows_out + BR - 1) / BR;
    const bool row_identity = J.w == 1, col_identity = J.h == 1;   // cv::GaussianBlur shrinks the kernel on 1-px axes
    const int tail = (L.tail_compat && L.k >= 11) ? (J.w & 7) : 0;
    const int xt_col = J.w - tail;                                  // column filter: mul+add from here on (even: pairs never straddle it)
    auto gpair = [&](int j) { const float c = L.gh[j]; return make_float2(c, c); };

    for (int hb = 0; hb < n_vb + D; hb++) {
        {
            const int srow_id = tid >> 3;              // 16 rows x 8 lanes; a lane strides along its row
            const uint8_t* rowp = J.src + (size_t)ds_clamp(y_begin - RMAX + hb * BR + srow_id, 0, J.h - 1) * J.src_pitch;
            float2* sp = s_stage + srow_id * spe + OFFS;
            const int gx0 = x0 - RMAX - DELTA;
            auto fetch = [&](int wi) -> uint32_t {
                const int gx = gx0 + 4 * wi;
                return (src_al && gx >= 0 && gx + 3 < J.w) ? ds_ldg32(rowp + gx) : fetch_word_clamped(rowp, gx, J.w);
            };
            // interior strips (all but the first and last of a page): no per-word tests, the right neighbour of a word comes from
            // the next lane (from the first lane's next word for the last of the eight)
            const bool interior = src_al && gx0 >= 0 && gx0 + 4 * STAGE_WORDS + 4 <= J.w;
            constexpr int NJ = (STAGE_WORDS + 7) / 8;
            const int l8 = tid & 7;
            uint32_t wv[NJ + 1], wn[NJ];
            if (interior) {
                const uint32_t* rp32 = reinterpret_cast<const uint32_t*>(rowp + gx0) + l8;
#pragma unroll
                for (int j = 0; j <= NJ; j++) wv[j] = (8 * j < STAGE_WORDS + 1 && l8 + 8 * j < STAGE_WORDS + 1) ? __ldg(rp32 + 8 * j) : 0u;
#pragma unroll
                for (int j = 0; j < NJ; j++) {
                    const uint32_t nx = __shfl_down_sync(0xffffffffu, wv[j], 1), wrap = __shfl_up_sync(0xffffffffu, wv[j + 1], 7);
                    wn[j] = l8 == 7 ? wrap : nx;
                }
            } else {
#pragma unroll
                for (int j = 0; j < NJ; j++) {
                    const int wi = l8 + 8 * j;
                    wv[j] = 0; wn[j] = 0;
                    if (wi < STAGE_WORDS) { wv[j] = fetch(wi); wn[j] = fetch(wi + 1); }
                }
            }
            // next step's rows towards L1 while this step computes (the loads above then wait tens of cycles, not hundreds)
            if (hb + 1 < n_vb + D && tid < 2 * BR) {
                const uint8_t* np = J.src + (size_t)ds_clamp(y_begin - RMAX + (hb + 1) * BR + (tid >> 1), 0, J.h - 1) * J.src_pitch
                                    + ds_clamp(gx0 + (tid & 1) * 128, 0, J.w - 1);
                asm volatile("prefetch.global.L1 [%0];" ::"l"(np));
            }
#pragma unroll
            for (int j = 0; j < NJ; j++) {
                const int wi = l8 + 8 * j;
                if (wi >= STAGE_WORDS) continue;
                // u8 -> fp32 without the conversion unit: 2^23 + v is exact in fp32, subtract 2^23 again (two lanes at a time)
                const float2 m23 = make_float2(-8388608.0f, -8388608.0f);
                const uint32_t a = wv[j], b = wn[j];
                auto cv = [&](uint32_t lo, uint32_t hi) { return __fadd2_rn(make_float2(__uint_as_float(lo), __uint_as_float(hi)), m23); };
                const uint32_t f0 = __byte_perm(a, 0x4B000000u, 0x7440), f1 = __byte_perm(a, 0x4B000000u, 0x7441);
                const uint32_t f2 = __byte_perm(a, 0x4B000000u, 0x7442), f3 = __byte_perm(a, 0x4B000000u, 0x7443);
                const uint32_t f4 = __byte_perm(b, 0x4B000000u, 0x7440), f5 = __byte_perm(b, 0x4B000000u, 0x7441);
                sp[4 * wi] = cv(f0, f2); sp[4 * wi + 1] = cv(f1, f3); sp[4 * wi + 2] = cv(f2, f4); sp[4 * wi + 3] = cv(f3, f5);
            }
        }
        __syncthreads();
        {   // ---- row pass: thread = (staged row, 16 consecutive outputs); lane <-> row as in the scalar kernel
            const int lane = tid & 31, wrp = tid >> 5;
            const int hr = lane & 15;
            const int c0 = 16 * (2 * wrp + (lane >> 4));
            // G[u] = (f[u], f[u + 2]), f[u] <-> column x0 + c0 + u - RMAX; the entry index OFFS + DELTA + c0 is even
            const float2* G = s_stage + hr * spe + OFFS + DELTA + c0;
            float2 acc[8];                                                    // pair j: outputs o_j and o_j + 2, o_j = 4 (j / 2) + (j & 1)
#pragma unroll
            for (int j = 0; j < 8; j++) acc[j] = make_float2(0.0f, 0.0f);
#pragma unroll
            for (int u2 = 0; u2 < NG / 2; u2++) {
                // (a compiler fence every six loads: ptxas otherwise requests all 24 at once and spills 16 loop invariants around them)
                if (u2 % 6 == 0 && u2) asm volatile("" ::: "memory");
                const float4 gg = *reinterpret_cast<const float4*>(G + 2 * u2);
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int u = 2 * u2 + h;
                    const float2 in = h ? make_float2(gg.z, gg.w) : make_float2(gg.x, gg.y);
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const int i = u - (4 * (j >> 1) + (j & 1));          // tap index in the padded kernel
                        if (i >= 0 && i <= 2 * RMAX) acc[j] = __ffma2_rn(in, gpair(i < RMAX ? RMAX - i : i - RMAX), acc[j]);
                    }
                }
            }
            if (row_identity) {
#pragma unroll
                for (int j = 0; j < 8; j++) acc[j] = G[RMAX + 4 * (j >> 1) + (j & 1)];
            }
            float4* dst = reinterpret_cast<float4*>(s_ring + ((hb * BR + hr) % RR) * RPF + c0);
            dst[0] = make_float4(acc[0].x, acc[1].x, acc[0].y, acc[1].y);
            dst[1] = make_float4(acc[2].x, acc[3].x, acc[2].y, acc[3].y);
            dst[2] = make_float4(acc[4].x, acc[5].x, acc[4].y, acc[5].y);
            dst[3] = make_float4(acc[6].x, acc[7].x, acc[6].y, acc[7].y);
        }
        __syncthreads();
        if (hb < D) continue;
        // ---- column pass: thread = (pair of columns, 8 rows); centre of output o sits at window index o + RMAX.  A window
        // starts at a multiple of 8 rows and the ring size is a compile-time constant: every ring offset is an immediate.
        const int vb = hb - D;
        const int cp = tid & 63, rh = tid >> 6;
        const int x = x0 + 2 * cp;
        // two halves of four rows: the second half's last four window rows are requested after the first half's arithmetic,
        // which keeps the live window at 38 pairs (the whole kernel within 128 registers)
        constexpr int HR = CR / 2, NW1 = HR + 2 * RMAX;
        float2 Wn[CR + 2 * RMAX];
        const int start = ((vb * BR) % RR + CR * rh) % RR;
        const float2* colp = reinterpret_cast<const float2*>(s_ring) + cp;
#define DS_LOAD_WINDOW2(S0, I0, I1) \
    case (S0) / CR: _Pragma("unroll") for (int i = (I0); i < (I1); i++) Wn[i] = colp[(((S0) + i) % RR) * (RPF / 2)]; break;
#define DS_LOAD_WINDOW2_ALL(I0, I1)                                                                                              \
    switch (start / CR) {                                                                                                        \
        DS_LOAD_WINDOW2(0, I0, I1) DS_LOAD_WINDOW2(8, I0, I1) DS_LOAD_WINDOW2(16, I0, I1) DS_LOAD_WINDOW2(24, I0, I1)             \
        DS_LOAD_WINDOW2(32, I0, I1) DS_LOAD_WINDOW2(40, I0, I1) DS_LOAD_WINDOW2(48, I0, I1) DS_LOAD_WINDOW2(56, I0, I1)           \
        DS_LOAD_WINDOW2(64, I0, I1) DS_LOAD_WINDOW2(72, I0, I1) DS_LOAD_WINDOW2(80, I0, I1) DS_LOAD_WINDOW2(88, I0, I1)           \
        default: break;                                                                                                          \
    }
        float2 acc[CR];
        const bool fused = x < xt_col;
        const float2 g0 = gpair(0);
        DS_LOAD_WINDOW2_ALL(0, NW1)
#pragma unroll
        for (int half = 0; half < 2; half++) {
            if (half == 1) { DS_LOAD_WINDOW2_ALL(NW1, CR + 2 * RMAX) }
#pragma unroll
            for (int o = half * HR; o < (half + 1) * HR; o++) acc[o] = __fmul2_rn(g0, Wn[o + RMAX]);
            if (fused) {
#pragma unroll
                for (int j = 1; j <= RMAX; j++) {
                    const float2 gj = gpair(j);
#pragma unroll
                    for (int o = half * HR; o < (half + 1) * HR; o++) acc[o] = __ffma2_rn(__fadd2_rn(Wn[o + RMAX + j], Wn[o + RMAX - j]), gj, acc[o]);
                }
            } else {
#pragma unroll
                for (int j = 1; j <= RMAX; j++) {
                    const float2 gj = gpair(j);
#pragma unroll
                    for (int o = half * HR; o < (half + 1) * HR; o++)
                        acc[o] = ds_add2_unfused(acc[o], ds_mul2_rn(gj, __fadd2_rn(Wn[o + RMAX + j], Wn[o + RMAX - j])));
                }
            }
            if (col_identity) {
#pragma unroll
                for (int o = half * HR; o < (half + 1) * HR; o++) acc[o] = Wn[o + RMAX];
            }
        }
#undef DS_LOAD_WINDOW2_ALL
#undef DS_LOAD_WINDOW2
        const int y0 = y_begin + vb * BR + CR * rh;
        const int rows = min(CR, y_end - y0);
        if (x < J.w && rows > 0) {
            // the mean of uint8 data under a kernel that sums to 1 lies in [0, 255.0001]: adding 1.5 * 2^23 leaves the mean rounded
            // to nearest even in the mantissa, and  src - mean > -C  <=>  src + C + bits(1.5 * 2^23) > bits(mean + 1.5 * 2^23)
            const float2 magic = make_float2(12582912.0f, 12582912.0f);
            const int cbias = L.c_param + 0x4B400000;
            const uint8_t* sp_ = J.src + (size_t)y0 * J.src_pitch + x;
            uint8_t* dp_ = J.dst + (size_t)y0 * J.dst_pitch + x;
            if (io16 && x + 1 < J.w && rows == CR) {
                uint32_t cpx[CR];                              // centre pixel pairs, loaded up front
#pragma unroll
                for (int o = 0; o < CR; o++) cpx[o] = __ldg(reinterpret_cast<const uint16_t*>(sp_ + (size_t)o * J.src_pitch));
#pragma unroll
                for (int o = 0; o < CR; o++) {
                    const float2 mb = __fadd2_rn(acc[o], magic);
                    const uint32_t r0 = ((int)(cpx[o] & 255u) + cbias > __float_as_int(mb.x)) ? 0x00ffu : 0u;
                    const uint32_t r1 = ((int)(cpx[o] >> 8) + cbias > __float_as_int(mb.y)) ? 0xff00u : 0u;
                    *reinterpret_cast<uint16_t*>(dp_ + (size_t)o * J.dst_pitch) = (uint16_t)(r0 | r1);
                }
            } else {
                const bool two = x + 1 < J.w;
#pragma unroll
                for (int o = 0; o < CR; o++) {
                    if (o >= rows) break;
                    const float2 mb = __fadd2_rn(acc[o], magic);
                    const uint8_t* s1 = sp_ + (size_t)o * J.src_pitch;
                    uint8_t* d1 = dp_ + (size_t)o * J.dst_pitch;
                    d1[0] = ((int)s1[0] + cbias > __float_as_int(mb.x)) ? 255 : 0;
                    if (two) d1[1] = ((int)s1[1] + cbias > __float_as_int(mb.y)) ? 255 : 0;
                }
            }
        }
    }
}

// Block sizes beyond the unrolled kernels (radius 33 .. GMAX - 1): same marching layout and the same operation order, but
// with run-time loops — the taps come from shared memory and the column window is read from the ring instead of living in
// registers.  About three times the instructions per pixel of the unrolled kernels; it exists so that every block size
// cv2 accepts up to 257 works, not for speed.
__global__ void __launch_bounds__(NT) adaptive_gauss_generic_kernel(const AdaptJob* __restrict__ jobs, const __grid_constant__ AdaptLaunch L,
                                                                    int ring_rows) {
    const AdaptJob J = jobs[blockIdx.z];
    const int x0 = blockIdx.x * TW;
    const int y_begin = blockIdx.y * L.seg_rows;
    if (x0 >= J.w || y_begin >= J.h) return;
    const int y_end = min(J.h, y_begin + L.seg_rows);
    const int tid = threadIdx.x, r = L.r, k = L.k;
    const int delta = (4 - (r & 3)) & 3;
    const int stage_words = (TW + 2 * r + delta + 3) >> 2;
    const int D = (2 * r + BR - 1) / BR;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    float* s_g = reinterpret_cast<float*>(smem_raw);               // k taps in order, padded to a multiple of 4
    float* s_stage = s_g + ((k + 3) & ~3);                         // BR * spf
    float* s_ring = s_stage + BR * L.spf;                          // ring_rows * RPF
    for (int i = tid; i < k; i += NT) s_g[i] = L.gh[abs(i - r)];
    const bool src_al = ((reinterpret_cast<uintptr_t>(J.src) | (uintptr_t)J.src_pitch) & 3) == 0;
    const int n_vb = (y_end - y_begin + BR - 1) / BR;
    const bool row_identity = J.w == 1, col_identity = J.h == 1;
    const int tail = L.tail_compat ? (J.w & 7) : 0;                // k >= 67 here: cv2's unfused tail columns always apply
    const int xt_col = J.w - tail;
    for (int hb = 0; hb < n_vb + D; hb++) {
        {
            const int srow_id = tid >> 3;
            const uint8_t* rowp = J.src + (size_t)ds_clamp(y_begin - r + hb * BR + srow_id, 0, J.h - 1) * J.src_pitch;
            float* sp = s_stage + srow_id * L.spf;
            for (int wi = tid & 7; wi < stage_words; wi += 8) {
                const int gx = x0 - r - delta + 4 * wi;
                const uint32_t wv = (src_al && gx >= 0 && gx + 3 < J.w) ? ds_ldg32(rowp + gx) : fetch_word_clamped(rowp, gx, J.w);
#pragma unroll
                for (int b = 0; b < 4; b++) sp[4 * wi + b] = __fsub_rn(__uint_as_float(__byte_perm(wv, 0x4B000000u, 0x7440 + b)), 8388608.0f);
            }
        }
        __syncthreads();
        {   // row pass: taps in increasing order, one fma each (the first tap starts from 0: fma(f, g0, 0) == f * g0)
            const int lane = tid & 31, wrp = tid >> 5;
            const int hr = lane & 15;
            const int c0 = 16 * (2 * wrp + (lane >> 4));
            const float* sp = s_stage + hr * L.spf + delta + c0;      // sp[u] <-> column x0 + c0 + u - r
            float acc[16];
#pragma unroll
            for (int o = 0; o < 16; o++) acc[o] = 0.0f;
            if (row_identity) {
#pragma unroll
                for (int o = 0; o < 16; o++) acc[o] = sp[r + o];
            } else {
                for (int i = 0; i < k; i++) {
                    const float gi = s_g[i];
#pragma unroll
                    for (int o = 0; o < 16; o++) acc[o] = __fmaf_rn(sp[o + i], gi, acc[o]);
                }
            }
            float4* dst = reinterpret_cast<float4*>(s_ring + ((hb * BR + hr) % ring_rows) * RPF + c0);
            dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
            dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
            dst[2] = make_float4(acc[8], acc[9], acc[10], acc[11]);
            dst[3] = make_float4(acc[12], acc[13], acc[14], acc[15]);
        }
        __syncthreads();
        if (hb < D) continue;
        const int vb = hb - D;
        const int col = tid, x = x0 + col;
        const float* colp = s_ring + col;
        const int base = vb * BR;                                      // ring row of window index 0 (virtual row vb * BR)
        float acc[BR];
        const float g0 = s_g[r];
#pragma unroll
        for (int o = 0; o < BR; o++) acc[o] = colp[((base + o + r) % ring_rows) * RPF];       // centre values
        if (!col_identity) {
#pragma unroll
            for (int o = 0; o < BR; o++) acc[o] = __fmul_rn(g0, acc[o]);
            const bool fused = x < xt_col;
            for (int j = 1; j <= r; j++) {
                const float gj = s_g[r + j];
                int up = (base + r + j) % ring_rows, dn = (base + r - j) % ring_rows;
#pragma unroll
                for (int o = 0; o < BR; o++) {
                    const float pair = __fadd_rn(colp[up * RPF], colp[dn * RPF]);
                    acc[o] = fused ? __fmaf_rn(pair, gj, acc[o]) : __fadd_rn(acc[o], __fmul_rn(gj, pair));
                    if (++up == ring_rows) up = 0;
                    if (++dn == ring_rows) dn = 0;
                }
            }
        }
        if (x < J.w) {
            const int y0 = y_begin + vb * BR;
            const int rows = min(BR, y_end - y0);
#pragma unroll
            for (int o = 0; o < BR; o++) {
                if (o >= rows) break;
                const int mean = min(max(__float2int_rn(acc[o]), 0), 255);
                const int c = J.src[(size_t)(y0 + o) * J.src_pitch + x];
                J.dst[(size_t)(y0 + o) * J.dst_pitch + x] = (c - mean > -L.c_param) ? 255 : 0;
            }
        }
    }
}

// cv2's row filter leaves the last w % 4 columns (w % 8 >= 4: the last w % 8 - 4) to scalar code: mul+add per tap,
// except that the (k-1) % 4 remainder taps are fma (oracle/docscan_oracle.c, A.9); its column filter is unfused from
// column w - w % 8 on.  This kernel recomputes those <= 3 columns of every page after the main kernel: a CTA takes
// TAIL_ROWS output rows, evaluates the row-filter chains of the rows it needs into shared memory, then the columns.
constexpr int TAIL_ROWS = 64;
__global__ void __launch_bounds__(128) adaptive_tail_kernel(const AdaptJob* __restrict__ jobs, const __grid_constant__ AdaptLaunch L) {
    const AdaptJob J = jobs[blockIdx.y];
    const int tail = J.w & 7;
    const int nt = tail >= 4 ? tail - 4 : tail;                     // scalar-row columns
    const int y0 = blockIdx.x * TAIL_ROWS;
    if (nt == 0 || y0 >= J.h) return;
    const int xt = J.w - nt, r = L.r, k = L.k;
    __shared__ float s_row[(TAIL_ROWS + 2 * (GMAX - 1)) * 3];
    const int nrows = min(TAIL_ROWS, J.h - y0), nv = nrows + 2 * r;
    const bool row_identity = J.w == 1, col_identity = J.h == 1;
    const int first_fused = k - ((k - 1) & 3);
    for (int t = threadIdx.x; t < nv * nt; t += 128) {
        const int v = t / nt, c = t - v * nt;
        const uint8_t* rowp = J.src + (size_t)ds_clamp(y0 - r + v, 0, J.h - 1) * J.src_pitch;
        const int xc = xt + c;
        float a;
        if (row_identity) a = (float)rowp[xc];
        else {
            a = __fmul_rn(L.gh[r], (float)rowp[ds_clamp(xc - r, 0, J.w - 1)]);
            int i = 1;
            for (; i < first_fused; i++) a = __fadd_rn(a, __fmul_rn(L.gh[abs(i - r)], (float)rowp[ds_clamp(xc - r + i, 0, J.w - 1)]));
            for (; i < k; i++) a = __fmaf_rn((float)rowp[ds_clamp(xc - r + i, 0, J.w - 1)], L.gh[abs(i - r)], a);
        }
        s_row[v * 3 + c] = a;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < nrows * nt; t += 128) {
        const int o = t / nt, c = t - o * nt;
        float m = s_row[(o + r) * 3 + c];
        if (!col_identity) {
            m = __fmul_rn(L.gh[0], m);
            for (int j = 1; j <= r; j++) m = __fadd_rn(m, __fmul_rn(L.gh[j], __fadd_rn(s_row[(o + r + j) * 3 + c], s_row[(o + r - j) * 3 + c])));
        }
        const int mean = min(max(__float2int_rn(m), 0), 255);
        const int y = y0 + o, x = xt + c;
        J.dst[(size_t)y * J.dst_pitch + x] = ((int)J.src[(size_t)y * J.src_pitch + x] - mean > -L.c_param) ? 255 : 0;
    }
}

// ---- exact evaluation of single pixels -------------------------------------------------------------------------------------
// The tensor-core path (tcblur.cu) decides every pixel whose fixed-point mean is further from the rounding boundary than its
// error bound and lists the others (about one in a thousand).  Those get cv2's exact fp32 value here, in cv2's operation
// order: the same row chains (fma per tap; the scalar variant in the last w % 4 columns), the same column chain (symmetric pairs,
// fused before column w - w % 8, mul+add from there on), the same kernel shrinking on 1-pixel axes.
// One CTA per 128 x 64 tile that has listed pixels: the tile's source window (BORDER_REPLICATE resolved while loading) goes to
// shared memory once; then one warp per listed pixel: lane l runs the row chains of window rows l and l + 32 (all taps fetched
// before the chain starts), the row values meet in shared memory, and the column chain — 2r dependent operations — runs once.
constexpr int FIX_WARPS = 4, FIX_TM = 128, FIX_WIN_W = 128;
struct FixPage { int tile_base, ntx, nty; };
__global__ void __launch_bounds__(FIX_WARPS * 32) adaptive_fix_tiles_kernel(const AdaptJob* __restrict__ jobs, const __grid_constant__ AdaptLaunch L,
                                                                            const FixPage* __restrict__ pages, int n_pages,
                                                                            const uint32_t* __restrict__ counts, const uint16_t* __restrict__ lists,
                                                                            int RL, int NOUT) {
    const int t = blockIdx.x;
    const uint32_t cnt = counts[t];
    if (cnt == 0) return;
    extern __shared__ __align__(16) uint8_t s_win[];              // (128 + 2r) rows x 128 bytes, then FIX_WARPS x 2 * GMAX floats
    int pg = 0;
    while (pg + 1 < n_pages && t >= pages[pg + 1].tile_base) pg++;
    const AdaptJob J = jobs[pg];
    const int idx = t - pages[pg].tile_base, ty = idx / pages[pg].ntx, tx = idx - ty * pages[pg].ntx;
    const int x0 = tx * NOUT, y0 = ty * FIX_TM;
    const int r = L.r, k = L.k;
    const int rows = FIX_TM + 2 * r;
    float* s_rows = reinterpret_cast<float*>(s_win + rows * FIX_WIN_W);
    // window byte (j, c) = source pixel (clamp(y0 - r + j), clamp(x0 - RL + c))
    for (int q = threadIdx.x; q < rows * (FIX_WIN_W / 16); q += FIX_WARPS * 32) {
        const int j = q >> 3, ch = q & 7;
        const uint8_t* rowp = J.src + (size_t)ds_clamp(y0 - r + j, 0, J.h - 1) * J.src_pitch;
        const int gx = x0 - RL + ch * 16;
        uint4 v;
        if (gx >= 0 && gx + 16 <= J.w) v = __ldg(reinterpret_cast<const uint4*>(rowp + gx));
        else {
            uint32_t wv[4] = {0, 0, 0, 0};
            for (int b2 = 0; b2 < 16; b2++) wv[b2 >> 2] |= (uint32_t)rowp[ds_clamp(gx + b2, 0, J.w - 1)] << (8 * (b2 & 3));
            v = make_uint4(wv[0], wv[1], wv[2], wv[3]);
        }
        *reinterpret_cast<uint4*>(s_win + j * FIX_WIN_W + ch * 16) = v;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* my_rows = s_rows + warp * 2 * GMAX;
    const int tail = (L.tail_compat && k >= 11) ? (J.w & 7) : 0;
    const int xt_col = J.w - tail;
    const int nt = tail >= 4 ? tail - 4 : tail;
    const int first_fused = k - ((k - 1) & 3);
    const bool all = cnt > TC_TILE_FLAG_CAP;                       // the list overflowed: every pixel of the tile
    const uint32_t n = all ? (uint32_t)(FIX_TM * NOUT) : cnt;
    for (uint32_t e = warp; e < n; e += FIX_WARPS) {
        int lr, lc;
        if (all) { lr = (int)(e / (uint32_t)NOUT); lc = (int)(e % (uint32_t)NOUT); }
        else { const uint32_t code = lists[(size_t)t * TC_TILE_FLAG_CAP + e]; lr = (int)(code >> 6); lc = (int)(code & 63u); }
        const int x = x0 + lc, y = y0 + lr;
        if (x >= J.w || y >= J.h) continue;                        // (only on the overflow path)
        const bool scalar_row = x >= J.w - nt;
        for (int v = lane; v < k; v += 32) {                       // window row lr + v <-> source row clamp(y - r + v)
            const uint8_t* wp = s_win + (lr + v) * FIX_WIN_W + (lc + RL - r);
            float a = __fmul_rn(L.gh[r], (float)wp[0]);
            if (scalar_row) {
                int i2 = 1;
                for (; i2 < first_fused; i2++) a = __fadd_rn(a, __fmul_rn(L.gh[abs(i2 - r)], (float)wp[i2]));
                for (; i2 < k; i2++) a = __fmaf_rn((float)wp[i2], L.gh[abs(i2 - r)], a);
            } else {
                for (int i0 = 1; i0 < k; i0 += 8) {
                    float f[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) f[u] = (float)wp[min(i0 + u, k - 1)];
#pragma unroll
                    for (int u = 0; u < 8; u++)
                        if (i0 + u < k) a = __fmaf_rn(f[u], L.gh[abs(i0 + u - r)], a);
                }
            }
            my_rows[v] = a;
        }
        __syncwarp();
        if (lane == 0) {
            float m = __fmul_rn(L.gh[0], my_rows[r]);
            const bool fused = x < xt_col;
            for (int j = 1; j <= r; j++) {
                const float pair = __fadd_rn(my_rows[r + j], my_rows[r - j]);
                m = fused ? __fmaf_rn(pair, L.gh[j], m) : __fadd_rn(m, __fmul_rn(L.gh[j], pair));
            }
            const int mean = min(max(__float2int_rn(m), 0), 255);
            J.dst[(size_t)y * J.dst_pitch + x] = ((int)s_win[(lr + r) * FIX_WIN_W + lc + RL] - mean > -L.c_param) ? 255 : 0;
        }
        __syncwarp();
    }
}

// combined = max(ink_sub_n > t_sub, bh_n > t_bh) -> dilate rect 2x2 x iters (window {x-n..x} x {y-n..y},
// out-of-image ignored) -> bin = base where combined else 255.  The normalise + threshold of both
// branches is folded into the raw cut-offs computed by scalars.cu.
__global__ void __launch_bounds__(128) mask_blend_kernel(const BlendJob* __restrict__ jobs, int n_dil, int mask_only) {
    const BlendJob J = jobs[blockIdx.z];
    const int y = blockIdx.y;
    const int x = (blockIdx.x * 128 + threadIdx.x) * 4;
    if (y >= J.h || x >= J.w) return;
    const int cut_a = J.sc->cut_a, cut_b = J.sc->cut_b;
    const bool fast = n_dil <= 4 && x + 4 <= J.w &&
                      ((reinterpret_cast<uintptr_t>(J.ink_sub) | reinterpret_cast<uintptr_t>(J.bh) | reinterpret_cast<uintptr_t>(J.dst) |
                        (mask_only ? 0 : reinterpret_cast<uintptr_t>(J.base)) | (uintptr_t)J.pitch_sub | (uintptr_t)J.pitch_bh |
                        (uintptr_t)J.pitch_dst | (uintptr_t)(mask_only ? 0 : J.pitch_base)) & 3) == 0;
    if (fast) {
        // packed path: 4 px per thread, 32-bit loads, byte-wise >= through the carry-free compare trick
        const uint32_t ca = min(cut_a, 256), cb = min(cut_b, 256);
        auto ge4 = [](uint32_t w, uint32_t cut) -> uint32_t {       // 0xff in every byte of w that is >= cut
            if (cut > 255) return 0u;
            if (cut == 0) return 0xffffffffu;
            return __vcmpgeu4(w, cut * 0x01010101u);
        };
        uint32_t ink = 0;
        for (int dy = 0; dy <= n_dil; dy++) {
            const int yy = y - dy;
            if (yy < 0) break;
            const uint8_t* ra = J.ink_sub + (size_t)yy * J.pitch_sub + x;
            const uint8_t* rb = J.bh + (size_t)yy * J.pitch_bh + x;
            const uint32_t cur = ge4(ds_ldg32(ra), ca) | ge4(ds_ldg32(rb), cb);
            const uint32_t prev = (n_dil && x >= 4) ? (ge4(ds_ldg32(ra - 4), ca) | ge4(ds_ldg32(rb - 4), cb)) : 0u;
            ink |= cur;
            for (int sft = 1; sft <= n_dil; sft++) ink |= sft == 4 ? prev : __funnelshift_l(prev, cur, 8 * sft);
        }
        uint32_t out = ink;
        if (!mask_only) out = (ds_ldg32(J.base + (size_t)y * J.pitch_base + x) & ink) | ~ink;
        *reinterpret_cast<uint32_t*>(J.dst + (size_t)y * J.pitch_dst + x) = out;
        return;
    }
    uint32_t hit = 0;   // bit i: pixel x+i has ink in its window
    for (int dy = 0; dy <= n_dil; dy++) {
        const int yy = y - dy;
        if (yy < 0) break;
        const uint8_t* ra = J.ink_sub + (size_t)yy * J.pitch_sub;
        const uint8_t* rb = J.bh + (size_t)yy * J.pitch_bh;
        uint32_t rowbits = 0;   // bit j <-> column x - n_dil + j
        for (int j = 0; j < 4 + n_dil; j++) {
            const int xx = x - n_dil + j;
            if (xx < 0 || xx >= J.w) continue;
            if ((int)ra[xx] >= cut_a || (int)rb[xx] >= cut_b) rowbits |= 1u << j;
        }
        for (int i = 0; i < 4; i++) {
            const uint32_t win = (rowbits >> i) & ((2u << n_dil) - 1u);
            if (win) hit |= 1u << i;
        }
    }
    uint8_t* rd = J.dst + (size_t)y * J.pitch_dst;
    const uint8_t* rbase = mask_only ? nullptr : J.base + (size_t)y * J.pitch_base;
    for (int i = 0; i < 4 && x + i < J.w; i++) {
        const bool ink = (hit >> i) & 1u;
        rd[x + i] = mask_only ? (ink ? 255 : 0) : (ink ? rbase[x + i] : 255);
    }
}

// Fast variant for the library's own planes (16-byte aligned rows): a thread owns a 16-pixel column chunk and marches
// down BLEND_ROWS rows, so the ink decision of row y-1 is reused for row y and every access is 128 bits wide.
constexpr int BLEND_ROWS = 8;
template <int NDIL>
__global__ void __launch_bounds__(128) mask_blend16_kernel(const BlendJob* __restrict__ jobs, int mask_only, int chunks, int row_groups) {
    const BlendJob J = jobs[blockIdx.y];
    const int id = blockIdx.x * 128 + threadIdx.x;
    const int cpr = (J.w + 15) >> 4;                       // chunks per row of this page
    const int rg = id / chunks, xc = id - rg * chunks;
    if (rg >= row_groups || xc >= cpr) return;
    const int x = xc * 16, y0 = rg * BLEND_ROWS;
    if (y0 >= J.h) return;
    const int cut_a = J.sc->cut_a, cut_b = J.sc->cut_b;
    const uint32_t ca = cut_a > 255 ? 0u : (uint32_t)max(cut_a, 0) * 0x01010101u, cb = cut_b > 255 ? 0u : (uint32_t)max(cut_b, 0) * 0x01010101u;
    const uint32_t all_a = cut_a <= 0 ? 0xffffffffu : 0u, none_a = cut_a > 255 ? 0u : 0xffffffffu;
    const uint32_t all_b = cut_b <= 0 ? 0xffffffffu : 0u, none_b = cut_b > 255 ? 0u : 0xffffffffu;
    auto ink4 = [&](uint32_t wa, uint32_t wb) -> uint32_t {     // 0xff where either branch is at / above its cut-off
        return ((__vcmpgeu4(wa, ca) | all_a) & none_a) | ((__vcmpgeu4(wb, cb) | all_b) & none_b);
    };
    // horizontally dilated ink of one row: byte i <- OR of raw ink at columns x+i-NDIL .. x+i
    auto row_ink = [&](int y, uint32_t (&m)[4]) {
        if (y < 0) { m[0] = m[1] = m[2] = m[3] = 0u; return; }
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(J.ink_sub + (size_t)y * J.pitch_sub + x));
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(J.bh + (size_t)y * J.pitch_bh + x));
        uint32_t r[5];
        r[0] = 0u;
        if (NDIL > 0 && x > 0) r[0] = ink4(ds_ldg32(J.ink_sub + (size_t)y * J.pitch_sub + x - 4), ds_ldg32(J.bh + (size_t)y * J.pitch_bh + x - 4));
        r[1] = ink4(a.x, b.x); r[2] = ink4(a.y, b.y); r[3] = ink4(a.z, b.z); r[4] = ink4(a.w, b.w);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t v = r[j + 1];
#pragma unroll
            for (int sft = 1; sft <= NDIL; sft++) v |= __funnelshift_l(r[j], r[j + 1], 8 * sft);
            m[j] = v;
        }
    };
    uint32_t prev[NDIL > 0 ? NDIL : 1][4];
#pragma unroll
    for (int d = 0; d < NDIL; d++) row_ink(y0 - NDIL + d, prev[d]);
    const int rows = min(BLEND_ROWS, J.h - y0);
    const bool full = x + 16 <= J.w;
#pragma unroll
    for (int i = 0; i < BLEND_ROWS; i++) {
        if (i >= rows) break;
        const int y = y0 + i;
        uint32_t cur[4], ink[4];
        row_ink(y, cur);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            ink[j] = cur[j];
#pragma unroll
            for (int d = 0; d < NDIL; d++) ink[j] |= prev[d][j];
        }
#pragma unroll
        for (int d = 0; d + 1 < NDIL; d++)
#pragma unroll
            for (int j = 0; j < 4; j++) prev[d][j] = prev[d + 1][j];
        if (NDIL > 0) {
#pragma unroll
            for (int j = 0; j < 4; j++) prev[NDIL - 1][j] = cur[j];
        }
        uint4 out = make_uint4(ink[0], ink[1], ink[2], ink[3]);
        if (!mask_only) {
            const uint4 bs = __ldg(reinterpret_cast<const uint4*>(J.base + (size_t)y * J.pitch_base + x));
            out = make_uint4((bs.x & ink[0]) | ~ink[0], (bs.y & ink[1]) | ~ink[1], (bs.z & ink[2]) | ~ink[2], (bs.w & ink[3]) | ~ink[3]);
        }
        uint8_t* dp = J.dst + (size_t)y * J.pitch_dst + x;
        if (full) *reinterpret_cast<uint4*>(dp) = out;
        else {
            const uint32_t ow[4] = {out.x, out.y, out.z, out.w};
            for (int b = 0; b < J.w - x; b++) dp[b] = (uint8_t)(ow[b >> 2] >> (8 * (b & 3)));
        }
    }
}

struct AdaptGridInfo { int strips, max_w, max_h, n, seg_min; };

template <int RMAX, int MINB = 4>
int launch_adaptive2(docscan_ctx* ctx, const AdaptJob* jd, AdaptLaunch L, const AdaptGridInfo& G, size_t smem) {
    if (smem > 48 * 1024)
        DS_CUDA(ctx, cudaFuncSetAttribute(adaptive_gauss2_kernel<RMAX, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    DS_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, adaptive_gauss2_kernel<RMAX, MINB>, NT, smem));
    L.seg_rows = ds_pick_seg_rows(per_sm * ctx->sm_count, G.strips, G.max_h, G.seg_min, BR);
    dim3 grid((G.max_w + TW - 1) / TW, (G.max_h + L.seg_rows - 1) / L.seg_rows, G.n);
    adaptive_gauss2_kernel<RMAX, MINB><<<grid, NT, smem, ctx->stream>>>(jd, L);
    DS_CHECK_LAUNCH(ctx);
    return DOCSCAN_OK;
}

template <int RMAX, int MINB>
int launch_adaptive_t(docscan_ctx* ctx, const AdaptJob* jd, AdaptLaunch L, const AdaptGridInfo& G, size_t smem) {
    if (smem > 48 * 1024)
        DS_CUDA(ctx, cudaFuncSetAttribute(adaptive_gauss_kernel<RMAX, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    DS_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, adaptive_gauss_kernel<RMAX, MINB>, NT, smem));
    L.seg_rows = ds_pick_seg_rows(per_sm * ctx->sm_count, G.strips, G.max_h, G.seg_min, BR);
    dim3 grid((G.max_w + TW - 1) / TW, (G.max_h + L.seg_rows - 1) / L.seg_rows, G.n);
    adaptive_gauss_kernel<RMAX, MINB><<<grid, NT, smem, ctx->stream>>>(jd, L);
    DS_CHECK_LAUNCH(ctx);
    return DOCSCAN_OK;
}

// radii up to 17: five CTAs per SM (DOCSCAN_ADAPT_MINB=4 selects the 118-register, four-CTA instance for measurements)
template <int RMAX>
int launch_adaptive(docscan_ctx* ctx, const AdaptJob* jd, const AdaptLaunch& L, const AdaptGridInfo& G, size_t smem) {
    if (RMAX > 17) return launch_adaptive_t<RMAX, 4>(ctx, jd, L, G, smem);
    const char* e = getenv("DOCSCAN_ADAPT_MINB");
    if (e && atoi(e) == 4) return launch_adaptive_t<RMAX, 4>(ctx, jd, L, G, smem);
    return launch_adaptive_t<RMAX, (RMAX > 17 ? 4 : 5)>(ctx, jd, L, G, smem);
}

}  // namespace

int k_adaptive_gauss_jobs(docscan_ctx* ctx, int k, int c, int cv_tail_compat, const AdaptJob* jobs_host, int n,
                          int max_w, int max_h) {
    if (k < 3 || (k & 1) == 0 || k > 2 * (GMAX - 1) + 1)
        return ds_fail(ctx, DOCSCAN_ERR_UNSUPPORTED, "adaptive GAUSSIAN_C block size must be odd and in 3..%d (got %d)", 2 * (GMAX - 1) + 1, k);
    AdaptLaunch L{};
    L.k = k; L.r = k / 2; L.c_param = c; L.tail_compat = cv_tail_compat;
    const bool generic = L.r > 32;
    const int rmax = L.r <= 5 ? 5 : L.r <= 9 ? 9 : L.r <= 13 ? 13 : L.r <= 15 ? 15 : L.r <= 17 ? 17 : L.r <= 25 ? 25 : L.r <= 32 ? 32 : L.r;
    const int delta = (4 - (rmax & 3)) & 3;
    // Opt-in (DOCSCAN_ADAPT_PACKED=1, radii up to 17): the packed-pair kernel.  Bit-exact (same GPU tests), 25 % fewer warp
    // instructions, but not faster: at 16 warps per SM the kernel waits on its barriers and shared-memory round trips, not
    // on issue slots (2.74 ms against 2.47 ms per 256 pages at 4 CTAs per SM; 2.34 ms at 3 CTAs and 160 registers, where the
    // whole step loses 0.2 ms because the other streams' kernels find less room) - profiles/README.md.
    bool packed = false;
    if (const char* e = getenv("DOCSCAN_ADAPT_PACKED")) packed = !generic && rmax <= 17 && atoi(e) != 0;
    if (packed) {
        // floats per staged row: (f[u], f[u + 2]) pairs for every staged column (+ 1 entry of offset); pitch / 4 odd, so that the
        // 128-bit loads of 8 rows hit 8 different 16-byte bank groups
        L.spf = 2 * (4 * ((TW + 2 * rmax + delta + 3) >> 2) + 2);
        L.spf = (L.spf + 3) & ~3;
        if (((L.spf >> 2) & 1) == 0) L.spf += 4;
    } else {
        // any odd pitch: the 16 staged rows a half-warp reads fall into 16 different banks, and the other half-warp (16 columns
        // further) into the other 16 (r * p = r' * p + 16 (mod 32) would need |r - r'| = 16)
        L.spf = (TW + 2 * rmax + delta + 4) | 1;
    }
    const int ring_rows = ((2 * rmax + BR - 1) / BR + 1) * BR;
    std::vector<float> g(k);
    docscan_gaussian_kernel_f32(k, g.data());
    for (int j = 0; j <= L.r; j++) L.gh[j] = g[L.r + j];     // the half kernel rides in the launch parameters

    // ---- the tensor-core path: fixed-point mean + guard band, exact evaluation of the listed pixels (see tcblur.cu)
    // Opt-in (DOCSCAN_TC_ADAPTIVE=1): bit-exact, but on pipeline pages 0.23 % of the pixels fall inside the guard band and their
    // exact re-evaluation (about 1000 warp instructions each) costs more than the contraction saves (profiles/README.md).
    const char* tc_env = getenv("DOCSCAN_TC_ADAPTIVE");
    if (tc_env && atoi(tc_env) != 0 && k <= 65 && n <= 64) {
        std::vector<int32_t> w16(k);
        double err = 0;                                      // sum |w / 65536 - g|: bound of the weight quantisation, per pass
        for (int i = 0; i < k; i++) {
            w16[i] = (int32_t)std::lround((double)g[i] * 65536.0);
            err += std::fabs((double)w16[i] / 65536.0 - (double)g[i]);
        }
        // |fixed-point mean - cv2's fp32 mean| <= 255 * (2 err + err^2)   weights, both passes
        //                                        + 2^-9                      row means kept in 8.8
        //                                        + (2k + 4) * 255 * 2^-23    cv2's own fp32 roundings (k fma + k/2 add + k/2 fma)
        //                                        + 2^-8                      margin (truncated low product, tie direction)
        const double eps = 255.0 * (2.0 * err + err * err) + 1.0 / 512 + (2.0 * k + 4) * 255.0 / 8388608.0 + 1.0 / 256;
        const int band = (int)std::ceil(eps * 65536.0);
        double px_total = 0;
        for (int i = 0; i < n; i++) px_total += (double)jobs_host[i].w * jobs_host[i].h;
        TcFlagLists fl;
        int rc2 = DOCSCAN_OK;
        if (k_tc_adaptive_jobs(ctx, k, c, w16.data(), band, jobs_host, n, &fl, &rc2)) {
            DS_TRY(rc2);
            std::vector<FixPage> fp(n);
            for (int i = 0; i < n; i++) { fp[i].tile_base = fl.tiles[i].tile_base; fp[i].ntx = fl.tiles[i].ntx; fp[i].nty = fl.tiles[i].nty; }
            void* devj = nullptr; void* devp = nullptr;
            DS_TRY(ds_upload(ctx, jobs_host, sizeof(AdaptJob) * n, &devj));
            DS_TRY(ds_upload(ctx, fp.data(), sizeof(FixPage) * n, &devp));
            const size_t fsmem = (size_t)(FIX_TM + 2 * L.r) * FIX_WIN_W + sizeof(float) * FIX_WARPS * 2 * GMAX;
            {
                ProfScope prof(ctx, "adaptive_gauss_fix", 0);
                DS_CUDA(ctx, cudaFuncSetAttribute(adaptive_fix_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
                adaptive_fix_tiles_kernel<<<fl.n_tiles, FIX_WARPS * 32, fsmem, ctx->stream>>>((const AdaptJob*)devj, L, (const FixPage*)devp, n, fl.count,
                                                                                               fl.list, fl.RL, fl.NOUT);
                DS_CHECK_LAUNCH(ctx);
            }
            if (getenv("DOCSCAN_TC_DEBUG")) {
                std::vector<uint32_t> cnt(fl.n_tiles);
                cudaMemcpyAsync(cnt.data(), fl.count, 4 * (size_t)fl.n_tiles, cudaMemcpyDeviceToHost, ctx->stream);
                cudaStreamSynchronize(ctx->stream);
                unsigned long long tot = 0; uint32_t mx = 0, over = 0;
                for (uint32_t v : cnt) { tot += v; mx = std::max(mx, v); over += v > TC_TILE_FLAG_CAP; }
                fprintf(stderr, "[tc debug] adaptive k=%d band=%d/65536 (eps %.4f) listed %llu of %.0f px, %d tiles, max %u per tile, %u overflowed\n", k, band,
                        eps, tot, px_total, fl.n_tiles, mx, over);
            }
            return DOCSCAN_OK;
        }
    }
    const AdaptGridInfo G{n * ((max_w + TW - 1) / TW), max_w, max_h, n, max(64, 4 * L.r)};
    const size_t smem = sizeof(float) * ((size_t)BR * L.spf + (size_t)ring_rows * RPF);
    void* dev = nullptr;
    DS_TRY(ds_upload(ctx, jobs_host, sizeof(AdaptJob) * n, &dev));
    const AdaptJob* jd = (const AdaptJob*)dev;
    double px = 0;
    for (int i = 0; i < n; i++) px += (double)jobs_host[i].w * jobs_host[i].h;
    int rc;
    {
    ProfScope prof(ctx, "adaptive_gauss_k" + std::to_string(k), 2.0 * px);
    if (generic) {
        const size_t gsmem = smem + sizeof(float) * ((k + 3) & ~3);
        if (gsmem > 220 * 1024) return ds_fail(ctx, DOCSCAN_ERR_UNSUPPORTED, "adaptive GAUSSIAN_C block size %d needs too much shared memory", k);
        DS_CUDA(ctx, cudaFuncSetAttribute(adaptive_gauss_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem));
        int per_sm = 1;
        DS_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, adaptive_gauss_generic_kernel, NT, gsmem));
        L.seg_rows = ds_pick_seg_rows(per_sm * ctx->sm_count, G.strips, G.max_h, G.seg_min, BR);
        dim3 grid((G.max_w + TW - 1) / TW, (G.max_h + L.seg_rows - 1) / L.seg_rows, G.n);
        adaptive_gauss_generic_kernel<<<grid, NT, gsmem, ctx->stream>>>(jd, L, ring_rows);
        ctx->launches++;
        rc = cudaGetLastError() == cudaSuccess ? DOCSCAN_OK : ds_fail(ctx, DOCSCAN_ERR_CUDA, "adaptive generic kernel launch failed");
    }
    else if (packed && L.r <= 5) rc = launch_adaptive2<5>(ctx, jd, L, G, smem);
    else if (packed && L.r <= 9) rc = launch_adaptive2<9>(ctx, jd, L, G, smem);
    else if (packed && L.r <= 13) rc = launch_adaptive2<13>(ctx, jd, L, G, smem);
    else if (packed && L.r <= 15) rc = launch_adaptive2<15>(ctx, jd, L, G, smem);
    else if (packed && L.r <= 17) rc = (getenv("DOCSCAN_ADAPT_MINB3") ? launch_adaptive2<17, 3>(ctx, jd, L, G, smem) : launch_adaptive2<17>(ctx, jd, L, G, smem));
    else if (L.r <= 5) rc = launch_adaptive<5>(ctx, jd, L, G, smem);
    else if (L.r <= 9) rc = launch_adaptive<9>(ctx, jd, L, G, smem);
    else if (L.r <= 13) rc = launch_adaptive<13>(ctx, jd, L, G, smem);
    else if (L.r <= 15) rc = launch_adaptive<15>(ctx, jd, L, G, smem);      // k = 31 (GUI preset), exact
    else if (L.r <= 17) rc = launch_adaptive<17>(ctx, jd, L, G, smem);      // k = 35 (CLI default), exact
    else if (L.r <= 25) rc = launch_adaptive<25>(ctx, jd, L, G, smem);
    else rc = launch_adaptive<32>(ctx, jd, L, G, smem);
    }
    DS_TRY(rc);
    if (cv_tail_compat && k >= 11) {
        bool any = false;
        for (int i = 0; i < n; i++) any = any || (jobs_host[i].w & 3) != 0;
        if (any) {
            ProfScope prof(ctx, "adaptive_gauss_tail", 0);
            adaptive_tail_kernel<<<dim3((max_h + TAIL_ROWS - 1) / TAIL_ROWS, n), 128, 0, ctx->stream>>>(jd, L);
            DS_CHECK_LAUNCH(ctx);
        }
    }
    return DOCSCAN_OK;
}

int k_mask_blend_jobs(docscan_ctx* ctx, int dilate_iters, int write_mask_only, const BlendJob* jobs_host, int n,
                      int max_w, int max_h) {
    if (dilate_iters < 0 || dilate_iters > 24) return ds_fail(ctx, DOCSCAN_ERR_UNSUPPORTED, "ink dilate iterations must be in 0..24");
    void* dev = nullptr;
    DS_TRY(ds_upload(ctx, jobs_host, sizeof(BlendJob) * n, &dev));
    dim3 grid((max_w + 511) / 512, max_h, n);
    double px = 0;
    for (int i = 0; i < n; i++) px += (double)jobs_host[i].w * jobs_host[i].h;
    ProfScope prof(ctx, "mask_blend", (write_mask_only ? 3.0 : 4.0) * px);
    bool aligned16 = dilate_iters <= 2;
    for (int i = 0; i < n && aligned16; i++) {
        const BlendJob& j = jobs_host[i];
        uintptr_t bits = reinterpret_cast<uintptr_t>(j.ink_sub) | reinterpret_cast<uintptr_t>(j.bh) | reinterpret_cast<uintptr_t>(j.dst) |
                         (uintptr_t)j.pitch_sub | (uintptr_t)j.pitch_bh | (uintptr_t)j.pitch_dst;
        if (!write_mask_only) bits |= reinterpret_cast<uintptr_t>(j.base) | (uintptr_t)j.pitch_base;
        // whole 16-byte chunks are read up to the end of the last chunk: the row pitch must cover them
        const int need = ((j.w + 15) >> 4) << 4;
        aligned16 = (bits & 15) == 0 && j.pitch_sub >= need && j.pitch_bh >= need && (write_mask_only || j.pitch_base >= need);
    }
    if (aligned16) {
        const int chunks = (max_w + 15) >> 4, row_groups = (max_h + BLEND_ROWS - 1) / BLEND_ROWS;
        dim3 g16((chunks * row_groups + 127) / 128, n);
        if (dilate_iters == 0) mask_blend16_kernel<0><<<g16, 128, 0, ctx->stream>>>((const BlendJob*)dev, write_mask_only, chunks, row_groups);
        else if (dilate_iters == 1) mask_blend16_kernel<1><<<g16, 128, 0, ctx->stream>>>((const BlendJob*)dev, write_mask_only, chunks, row_groups);
        else mask_blend16_kernel<2><<<g16, 128, 0, ctx->stream>>>((const BlendJob*)dev, write_mask_only, chunks, row_groups);
        DS_CHECK_LAUNCH(ctx);
        return DOCSCAN_OK;
    }
    mask_blend_kernel<<<grid, 128, 0, ctx->stream>>>((const BlendJob*)dev, dilate_iters, write_mask_only);
    DS_CHECK_LAUNCH(ctx);
    return DOCSCAN_OK;
}
