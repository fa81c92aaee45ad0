// Separable integer blur of uint8 pages with a fused per-pixel epilogue.
//
//   kind 0  cv2.GaussianBlur(u8, (k,k), 0), BORDER_REFLECT_101 (DocScanner.py:153,184): OpenCV's 8.8
//           fixed-point kernel; H pass exact in u16, V pass exact in u32, dst = (V + 32768) >> 16.
//   kind 1  k x k box sum with BORDER_REPLICATE, mean = round(V / k^2): the local mean of
//           cv2.adaptiveThreshold(ADAPTIVE_THRESH_MEAN_C) (DocScanner.py:167).
//
// One CTA owns a 128-column strip of one page segment and marches down it.  Per step it stages 16
// source rows (halo columns resolved through the border rule) in shared memory, H-filters them with
// dp4a (4 taps per instruction, coefficient words pre-shifted on the host so that all loads are
// 32-bit aligned) into a ring of u16 rows, then V-filters 16 output rows out of the ring with a
// register-blocked 8-row x 2-column accumulator tile per thread.  Every source byte is read from
// HBM/L2 once per strip (plus the halo), the H pass is never recomputed inside a segment, and the
// epilogue (subtract / divide / threshold, min-max, histogram) is applied before the only store.
// All sums are exact integers, so the regrouping is bit-exact with OpenCV.
#include "common.cuh"

namespace {

constexpr int TW = 128;       // output columns per strip
constexpr int BR = 16;        // rows per march step
constexpr int NT = 128;       // threads per CTA
constexpr int RP2 = 132;      // ring pitch in 32-bit words per ROW PAIR (one word = rows 2P, 2P+1 of a column; 128 + 4 pad)

struct BlurTable {
    int k_eff, r_eff, delta, M, nb;
    uint32_t kk;              // k*k (box) or 0
    float inv_kk;             // (float)(1.0 / (k*k)) for the box mean
    const uint4* qH;          // M + 6 entries (3 zero entries each side)
    const uint32_t* qV;       // 2 * (4*nb + 8) entries: even-aligned tap pairs, then odd-aligned tap pairs (dp2a operands)
};

struct BlurLaunch {
    BlurTable t;
    int seg_rows, border, c_param;
    int spw;                  // staging row pitch in words (odd)
    int ring_rows;
};

__device__ __forceinline__ int border_map(int p, int len, int border) {
    return border == 0 ? ds_reflect101(p, len) : ds_clamp(p, 0, len - 1);
}

// rare path of the staging load (strip edges, unaligned caller buffers): kept out of line
__device__ __noinline__ uint32_t fetch_word_slow(const uint8_t* rowp, int gx, int w, int border) {
    uint32_t word = 0;
    for (int b = 0; b < 4; b++) word |= (uint32_t)rowp[border_map(gx + b, w, border)] << (8 * b);
    return word;
}

constexpr int SB = 7;         // staging loads a thread keeps in flight (7 x 8 words covers k <= 63 in one go)

// KEFF > 0: a kernel instance for one effective tap count (the sizes the pipeline uses all day).  Loop bounds, the
// coefficient windows and — above all — which (input word, output) combinations carry only zero taps are then known at
// compile time: both passes are fully unrolled and issue only the dot products that can contribute (k = 23: 18 instead of
// 26 per pixel).  KEFF == 0 is the general kernel with run-time loops.
template <int EPI, bool STATS, int KEFF>
__global__ void __launch_bounds__(NT) blur_march_kernel(const BlurJob* __restrict__ jobs, const BlurLaunch L) {
    constexpr bool FIXED = KEFF > 0;
    constexpr int FR = KEFF / 2, FDELTA = (4 - (FR & 3)) & 3, FM = (FDELTA + KEFF + 2) / 4 + 1;      // as get_table() derives them
    constexpr int FNP = (KEFF + 6) / 2 + 1;                                                          // row pairs 8 outputs reach over
    const BlurJob J = jobs[blockIdx.z];
    const int x0 = blockIdx.x * TW;
    const int y_begin = blockIdx.y * L.seg_rows;
    if (x0 >= J.w || y_begin >= J.h) return;
    const int y_end = min(J.h, y_begin + L.seg_rows);
    const int rows_out = y_end - y_begin;
    const int tid = threadIdx.x;
    const BlurTable& T = L.t;
    const int r = T.r_eff;
    const int r4 = r + T.delta;

    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint4* s_qH = reinterpret_cast<uint4*>(smem_raw);
    uint32_t* s_qV = reinterpret_cast<uint32_t*>(s_qH + (T.M + 6));
    uint32_t* s_stage = s_qV + 2 * (4 * T.nb + 8);
    // (on the next 16-byte boundary, rounded as an index: a pointer rounded through an integer cast becomes generic and every
    // access through it an LD / ST / ATOM instead of LDS / STS / ATOMS; everything before the ring is a whole number of words)
    const int ring_off = (int)(s_stage - reinterpret_cast<uint32_t*>(smem_raw)) + BR * L.spw;
    uint32_t* s_ring = reinterpret_cast<uint32_t*>(smem_raw) + ((ring_off + 3) & ~3);
    uint32_t* s_hist = s_ring + (L.ring_rows / 2) * RP2;   // 4 x 256, only when STATS && J.hist

    for (int i = tid; i < T.M + 6; i += NT) s_qH[i] = T.qH[i];
    for (int i = tid; i < 2 * (4 * T.nb + 8); i += NT) s_qV[i] = T.qV[i];
    if (STATS && J.hist)
        for (int i = tid; i < 4 * 256; i += NT) s_hist[i] = 0;

    const bool src_al = ((reinterpret_cast<uintptr_t>(J.src) | (uintptr_t)J.src_pitch) & 3) == 0;
    const int D = (2 * r + BR - 1) / BR;                       // V batch lags the H batch by D steps
    const int n_vb = (rows_out + BR - 1) / BR;
    uint32_t st_lo = 255, st_hi = 0, zero_count = 0;

    for (int hb = 0; hb < n_vb + D; hb++) {
        // ---- stage BR source rows: virtual row v <-> source row border(y_begin - r + v)
        {
            const int srow_id = tid >> 3;              // 16 rows x 8 lanes; a lane strides along its row
            const uint8_t* rowp = J.src + (size_t)border_map(y_begin - r + hb * BR + srow_id, J.h, L.border) * J.src_pitch;
            uint32_t* srow_w = s_stage + srow_id * L.spw;
            for (int w0 = tid & 7; w0 < L.spw; w0 += 8 * SB) {
                uint32_t wv[SB];                       // issue SB independent loads, then store them
#pragma unroll
                for (int j = 0; j < SB; j++) {
                    const int wi = w0 + 8 * j;
                    const int gx = x0 - r4 + 4 * wi;
                    wv[j] = 0;
                    if (wi < L.spw) wv[j] = (src_al && gx >= 0 && gx + 3 < J.w) ? ds_ldg32(rowp + gx) : fetch_word_slow(rowp, gx, J.w, L.border);
                }
#pragma unroll
                for (int j = 0; j < SB; j++) if (w0 + 8 * j < L.spw) srow_w[w0 + 8 * j] = wv[j];
            }
        }
        __syncthreads();
        // ---- H pass: thread = (row, 16 consecutive columns)
        {
            const int hr = tid >> 3, cg = tid & 7;
            const uint32_t* srow = s_stage + hr * L.spw + cg * 4;
            uint32_t acc[16];
#pragma unroll
            for (int i = 0; i < 16; i++) acc[i] = 0;
            if (FIXED) {
                // acc[4g + s] += dp4a(word wi, coefficient word (m = wi - g, shift s)); word (m, s) holds taps 4m + b - s - delta
#pragma unroll
                for (int wi = 0; wi < FM + 3; wi++) {
                    const uint32_t W = srow[wi];
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        const int m = wi - g;
                        if (m < 0 || m >= FM) continue;
                        const uint4 Q = s_qH[m + 3];
                        const uint32_t qs[4] = {Q.x, Q.y, Q.z, Q.w};
#pragma unroll
                        for (int sft = 0; sft < 4; sft++) {
                            const int lo = 4 * m - sft - FDELTA, hi = lo + 3;               // tap range of this word
                            if (hi >= 0 && lo < KEFF) acc[4 * g + sft] = __dp4a(W, qs[sft], acc[4 * g + sft]);
                        }
                    }
                }
            }
            uint4 q0 = s_qH[0], q1 = s_qH[1], q2 = s_qH[2], q3;    // zero entries: q[wi + 3 - g], g = 3,2,1
            for (int wi = 0; !FIXED && wi < T.M + 3; wi++) {
                q3 = s_qH[wi + 3];
                const uint32_t W = srow[wi];
                // group g uses coefficient word m = wi - g  -> table entry m + 3
                acc[0] = __dp4a(W, q3.x, acc[0]);  acc[1] = __dp4a(W, q3.y, acc[1]);
                acc[2] = __dp4a(W, q3.z, acc[2]);  acc[3] = __dp4a(W, q3.w, acc[3]);
                acc[4] = __dp4a(W, q2.x, acc[4]);  acc[5] = __dp4a(W, q2.y, acc[5]);
                acc[6] = __dp4a(W, q2.z, acc[6]);  acc[7] = __dp4a(W, q2.w, acc[7]);
                acc[8] = __dp4a(W, q1.x, acc[8]);  acc[9] = __dp4a(W, q1.y, acc[9]);
                acc[10] = __dp4a(W, q1.z, acc[10]); acc[11] = __dp4a(W, q1.w, acc[11]);
                acc[12] = __dp4a(W, q0.x, acc[12]); acc[13] = __dp4a(W, q0.y, acc[13]);
                acc[14] = __dp4a(W, q0.z, acc[14]); acc[15] = __dp4a(W, q0.w, acc[15]);
                q0 = q1; q1 = q2; q2 = q3;
            }
            // The V pass consumes dp2a operands: one word = (row 2P, row 2P+1) of one column.  Rows hr and hr^1 live in
            // lanes tid and tid^8 of the same warp: swap halves, so the even row's thread packs columns 0..7 and the odd
            // row's thread packs columns 8..15 of the pair.
            const bool odd = hr & 1;
            uint32_t packed[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const uint32_t send = odd ? acc[i] : acc[8 + i];
                const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 8);
                const uint32_t mine = odd ? acc[8 + i] : acc[i];
                packed[i] = odd ? (recv | (mine << 16)) : (mine | (recv << 16));
            }
            const int pslot = ((hb * BR + hr) >> 1) % (L.ring_rows >> 1);
            uint4* dst = reinterpret_cast<uint4*>(s_ring + pslot * RP2 + cg * 16 + (odd ? 8 : 0));
            dst[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
            dst[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
        }
        __syncthreads();
        if (hb < D) continue;
        // ---- V pass: thread = (column pair, 8 rows)
        const int vb = hb - D;
        const int cp = tid & 63, rg = tid >> 6;
        const int vbase = vb * BR + rg * 8;
        uint32_t a0[8], a1[8];
#pragma unroll
        for (int o = 0; o < 8; o++) { a0[o] = 0; a1[o] = 0; }
        // dp2a: acc += row(2P) * q[t] + row(2P+1) * q[t+1].  Output o at pair step P needs taps (2P - o, 2P + 1 - o):
        // even o -> even-aligned pair E[P - o/2], odd o -> odd-aligned pair O[P - (o+1)/2]; 4 pair steps per block,
        // sliding windows of 8 coefficient words each, all indices static.
        const uint32_t* s_qE = s_qV;
        const uint32_t* s_qO = s_qV + (4 * T.nb + 8);
        uint32_t E[8], O[8];
        {
            const uint4 e = *reinterpret_cast<const uint4*>(s_qE), o4 = *reinterpret_cast<const uint4*>(s_qO);
            E[4] = e.x; E[5] = e.y; E[6] = e.z; E[7] = e.w; O[4] = o4.x; O[5] = o4.y; O[6] = o4.z; O[7] = o4.w;
        }
        // a block of 4 row pairs starts at a multiple of 4 pairs and the ring holds a multiple of 8: no wrap inside a block
        const int ring_pairs = L.ring_rows >> 1;
        const uint32_t* rp = s_ring + ((vbase >> 1) % ring_pairs) * RP2 + 2 * cp;
        const uint32_t* const rend = s_ring + ring_pairs * RP2 + 2 * cp;
        if (FIXED) {
            // pair step p feeds output o through the even-aligned coefficient pair m = p - o/2 (even o) or the odd-aligned
            // pair m = p - (o+1)/2 (odd o); pairs whose two taps both fall outside [0, KEFF) are skipped
#pragma unroll
            for (int p = 0; p < FNP; p++) {
                if (p && (p & 3) == 0) {
                    rp += 4 * RP2;
                    if (rp >= rend) rp -= ring_pairs * RP2;
                }
                const uint2 w = *reinterpret_cast<const uint2*>(rp + (p & 3) * RP2);
#pragma unroll
                for (int o = 0; o < 8; o++) {
                    const int m = (o & 1) ? p - (o + 1) / 2 : p - o / 2;
                    const bool nz = (o & 1) ? (m >= -1 && 2 * m + 1 < KEFF) : (m >= 0 && 2 * m < KEFF);
                    if (!nz) continue;
                    const uint32_t c = (o & 1) ? s_qO[m + 4] : s_qE[m + 4];
                    a0[o] = __dp2a_lo(w.x, c, a0[o]);
                    a1[o] = __dp2a_lo(w.y, c, a1[o]);
                }
            }
        }
        for (int b = 0; !FIXED && b < T.nb; b++) {
#pragma unroll
            for (int i = 0; i < 4; i++) { E[i] = E[i + 4]; O[i] = O[i + 4]; }
            {
                const uint4 e = *reinterpret_cast<const uint4*>(s_qE + 4 * b + 4), o4 = *reinterpret_cast<const uint4*>(s_qO + 4 * b + 4);
                E[4] = e.x; E[5] = e.y; E[6] = e.z; E[7] = e.w; O[4] = o4.x; O[5] = o4.y; O[6] = o4.z; O[7] = o4.w;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint2 w = *reinterpret_cast<const uint2*>(rp + u * RP2);   // two columns of one row pair
#pragma unroll
                for (int o = 0; o < 8; o++) {
                    const uint32_t c = (o & 1) ? O[u - (o + 1) / 2 + 4] : E[u - o / 2 + 4];
                    a0[o] = __dp2a_lo(w.x, c, a0[o]);
                    a1[o] = __dp2a_lo(w.y, c, a1[o]);
                }
            }
            rp += 4 * RP2;
            if (rp >= rend) rp -= ring_pairs * RP2;
        }
        // ---- epilogue
        const int x = x0 + 2 * cp;
        if (x < J.w) {
            const bool two = x + 1 < J.w;
            const int rows_here = min(8, y_end - (y_begin + vbase));
            int c0[8], c1[8];                          // centre pixels of the 8 rows, loaded up front
            if (EPI != DS_EPI_BLUR) {
                const uint8_t* cp0 = J.src + (size_t)(y_begin + vbase) * J.src_pitch + x;
#pragma unroll
                for (int o = 0; o < 8; o++) {
                    c0[o] = 0; c1[o] = 0;
                    if (o < rows_here) { c0[o] = cp0[0]; c1[o] = two ? cp0[1] : 0; }
                    cp0 += J.src_pitch;
                }
            }
#pragma unroll
            for (int o = 0; o < 8; o++) {
                const int y = y_begin + vbase + o;
                if (o >= rows_here) break;
                uint32_t b0, b1;
                // cv::boxFilter scales the integer box sum in fp32: cvRound((float)sum * (float)(1.0 / k^2))
                if (T.kk) { b0 = (uint32_t)__float2int_rn(__fmul_rn((float)a0[o], T.inv_kk)); b1 = (uint32_t)__float2int_rn(__fmul_rn((float)a1[o], T.inv_kk)); }
                else { b0 = (a0[o] + 32768u) >> 16; b1 = (a1[o] + 32768u) >> 16; }
                uint32_t v0 = b0, v1 = b1;
                if (EPI != DS_EPI_BLUR) {
                    const int s0 = c0[o], s1 = c1[o];
                    if (EPI == DS_EPI_SUB) { v0 = max(s0 - (int)b0, 0); v1 = max(s1 - (int)b1, 0); }
                    else if (EPI == DS_EPI_RSUB) { v0 = max((int)b0 - s0, 0); v1 = max((int)b1 - s1, 0); }
                    else if (EPI == DS_EPI_DIV) { v0 = ds_div255((uint8_t)s0, (uint8_t)b0); v1 = ds_div255((uint8_t)s1, (uint8_t)b1); }
                    else if (EPI == DS_EPI_ATHRESH) { v0 = (s0 - (int)b0 > -L.c_param) ? 255 : 0; v1 = (s1 - (int)b1 > -L.c_param) ? 255 : 0; }
                }
                uint8_t* dp = J.dst + (size_t)y * J.dst_pitch + x;
                dp[0] = (uint8_t)v0;
                if (two) dp[1] = (uint8_t)v1;
                if (STATS) {
                    st_lo = min(st_lo, v0); st_hi = max(st_hi, v0);
                    if (two) { st_lo = min(st_lo, v1); st_hi = max(st_hi, v1); }
                    if (J.hist) {
                        uint32_t* hw = s_hist + (tid >> 5) * 256;
                        if (v0) atomicAdd(&hw[v0], 1u); else zero_count++;
                        if (two) { if (v1) atomicAdd(&hw[v1], 1u); else zero_count++; }
                    }
                }
            }
        }
    }
    if (STATS) {
        if (J.minmax) {
            for (int o = 16; o; o >>= 1) {
                st_lo = min(st_lo, __shfl_xor_sync(0xffffffffu, st_lo, o));
                st_hi = max(st_hi, __shfl_xor_sync(0xffffffffu, st_hi, o));
            }
            if ((tid & 31) == 0) { atomicMin(&J.minmax[0], st_lo); atomicMax(&J.minmax[1], st_hi); }
        }
        if (J.hist) {
            for (int o = 16; o; o >>= 1) zero_count += __shfl_xor_sync(0xffffffffu, zero_count, o);
            if ((tid & 31) == 0 && zero_count) atomicAdd(&s_hist[(tid >> 5) * 256], zero_count);
            __syncthreads();
            for (int i = tid; i < 256; i += NT) {
                const uint32_t s = s_hist[i] + s_hist[256 + i] + s_hist[512 + i] + s_hist[768 + i];
                if (s) atomicAdd(&J.hist[i], s);
            }
        }
    }
}

// ---- host: coefficient tables ----------------------------------------------------------------------
int get_table(docscan_ctx* ctx, int kind, int k, BlurTable* out) {
    std::vector<int32_t> q(k);
    if (kind == 0) {
        if (docscan_gaussian_kernel_q8(k, q.data()) != DOCSCAN_OK) return ds_fail(ctx, DOCSCAN_ERR_BAD_ARG, "bad blur ksize %d", k);
    } else {
        for (int i = 0; i < k; i++) q[i] = 1;
    }
    int z = 0;
    while (z < k / 2 && q[z] == 0) z++;            // the quantised Gaussian has zero tails: trim them
    const int k_eff = k - 2 * z, r_eff = k_eff / 2;
    const int delta = (4 - (r_eff & 3)) & 3;
    const int M = (delta + k_eff + 2) / 4 + 1;
    const int nb = ((k_eff + 7 + 1) / 2 + 3) / 4;          // V pass: blocks of 4 row pairs covering 8 outputs + k_eff - 1 rows
    out->k_eff = k_eff; out->r_eff = r_eff; out->delta = delta; out->M = M; out->nb = nb;
    out->kk = kind == 1 ? (uint32_t)k * (uint32_t)k : 0;
    out->inv_kk = (float)(1.0 / ((double)k * (double)k));
    const uint64_t key = ((uint64_t)(kind + 1) << 32) | (uint32_t)k;
    const int nqv = 4 * nb + 8;
    const size_t qh_bytes = sizeof(uint4) * (M + 6), qv_bytes = sizeof(uint32_t) * 2 * nqv;
    auto it = ctx->tables.find(key);
    if (it == ctx->tables.end()) {
        std::vector<uint8_t> host(qh_bytes + qv_bytes, 0);
        uint8_t* qh = host.data();
        uint32_t* qv = reinterpret_cast<uint32_t*>(host.data() + qh_bytes);
        auto tap = [&](int jp) -> int {            // q'[j'] : taps shifted right by delta
            const int j = jp - delta;
            return (j >= 0 && j < k_eff) ? q[z + j] : 0;
        };
        for (int m = 0; m < M; m++)
            for (int s = 0; s < 4; s++)
                for (int b = 0; b < 4; b++)
                    qh[(size_t)(m + 3) * 16 + s * 4 + b] = (uint8_t)tap(4 * m + b - s);
        auto qt = [&](int t) -> uint32_t { return (t >= 0 && t < k_eff) ? (uint32_t)q[z + t] : 0u; };
        for (int m = -4; m < nqv - 4; m++) {
            qv[m + 4] = qt(2 * m) | (qt(2 * m + 1) << 8);               // even-aligned pair (q[2m], q[2m+1])
            qv[nqv + m + 4] = qt(2 * m + 1) | (qt(2 * m + 2) << 8);     // odd-aligned pair (q[2m+1], q[2m+2])
        }
        void* dev = nullptr;
        DS_CUDA(ctx, cudaMalloc(&dev, host.size()));
        DS_CUDA(ctx, cudaMemcpyAsync(dev, host.data(), host.size(), cudaMemcpyHostToDevice, ctx->stream));
        DS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        it = ctx->tables.emplace(key, dev).first;
    }
    out->qH = reinterpret_cast<const uint4*>(it->second);
    out->qV = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(it->second) + qh_bytes);
    return DOCSCAN_OK;
}

struct BlurGridInfo { int strips, max_w, max_h, n, seg_min; };

template <int EPI, bool STATS, int KEFF>
int launch_k(docscan_ctx* ctx, const BlurJob* jobs_dev, BlurLaunch L, const BlurGridInfo& G, size_t smem) {
    if (smem > 48 * 1024)
        DS_CUDA(ctx, cudaFuncSetAttribute(blur_march_kernel<EPI, STATS, KEFF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    DS_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, blur_march_kernel<EPI, STATS, KEFF>, NT, smem));
    L.seg_rows = ds_pick_seg_rows(per_sm * ctx->sm_count, G.strips, G.max_h, G.seg_min, BR);
    dim3 grid((G.max_w + TW - 1) / TW, (G.max_h + L.seg_rows - 1) / L.seg_rows, G.n);
    blur_march_kernel<EPI, STATS, KEFF><<<grid, NT, smem, ctx->stream>>>(jobs_dev, L);
    DS_CHECK_LAUNCH(ctx);
    return DOCSCAN_OK;
}

template <int EPI, bool STATS>
int launch(docscan_ctx* ctx, const BlurJob* jobs_dev, BlurLaunch L, const BlurGridInfo& G, size_t smem) {
    // fixed-size instances for the epilogues of the page pipeline (background subtract / divide with min-max, ink branch
    // with histogram) at the tap counts its default and GUI presets produce: k = 23, 43, 51, 57
    if (STATS && (EPI == DS_EPI_SUB || EPI == DS_EPI_RSUB || EPI == DS_EPI_DIV)) {
        switch (L.t.k_eff) {
            case 21: return launch_k<EPI, STATS, 21>(ctx, jobs_dev, L, G, smem);
            case 39: return launch_k<EPI, STATS, 39>(ctx, jobs_dev, L, G, smem);
            case 45: return launch_k<EPI, STATS, 45>(ctx, jobs_dev, L, G, smem);
            case 51: return launch_k<EPI, STATS, 51>(ctx, jobs_dev, L, G, smem);
            default: break;
        }
    }
    return launch_k<EPI, STATS, 0>(ctx, jobs_dev, L, G, smem);
}

}  // namespace

int k_blur_jobs(docscan_ctx* ctx, int kind, int k, int epi, int c_param, const BlurJob* jobs_host, int n,
                int max_w, int max_h) {
    if (k < 1 || (k & 1) == 0) return ds_fail(ctx, DOCSCAN_ERR_BAD_ARG, "blur ksize must be odd and positive (got %d)", k);
    // box sums travel through the ring as 16-bit values: 255 * k must fit (the 8.8 Gaussian sums to 256 for every k)
    if (kind == 1 && k > 255) return ds_fail(ctx, DOCSCAN_ERR_UNSUPPORTED, "box mean block size must be at most 255 (got %d)", k);
    if (kind == 0 && k == 1) kind = 1;             // 1x1 Gaussian is the identity; so is the 1x1 box mean
    {
        int rc = DOCSCAN_OK;                       // the Gaussian as two banded contractions on the tensor cores (tcblur.cu)
        if (k_tc_blur_jobs(ctx, kind, k, epi, jobs_host, n, &rc)) return rc;
    }
    BlurLaunch L{};
    DS_TRY(get_table(ctx, kind, k, &L.t));
    L.border = kind == 0 ? 0 : 1;
    L.c_param = c_param;
    L.spw = (TW / 4 + L.t.M + 3) | 1;
    L.ring_rows = ((2 * L.t.r_eff + BR - 1) / BR + 1) * BR;
    bool stats = false;
    for (int i = 0; i < n; i++) stats = stats || jobs_host[i].minmax || jobs_host[i].hist;
    const BlurGridInfo G{n * ((max_w + TW - 1) / TW), max_w, max_h, n, max(64, 4 * L.t.r_eff)};
    const size_t smem = sizeof(uint4) * (L.t.M + 6) + sizeof(uint32_t) * 2 * (4 * L.t.nb + 8) +
                        sizeof(uint32_t) * BR * L.spw + 16 + sizeof(uint32_t) * (L.ring_rows / 2) * RP2 +
                        (stats ? 4 * 256 * sizeof(uint32_t) : 0);
    if (smem > 220 * 1024)
        return ds_fail(ctx, DOCSCAN_ERR_UNSUPPORTED, "blur ksize %d needs %zu bytes of shared memory for its row ring (limit 220 KB: k up to about 900)", k, smem);
    void* dev = nullptr;
    DS_TRY(ds_upload(ctx, jobs_host, sizeof(BlurJob) * n, &dev));
    const BlurJob* jd = (const BlurJob*)dev;
    double px = 0;
    for (int i = 0; i < n; i++) px += (double)jobs_host[i].w * jobs_host[i].h;
    ProfScope prof(ctx, std::string(kind == 0 ? "blur_gauss_k" : "blur_box_k") + std::to_string(k), 2.0 * px);
#define DS_BLUR_CASE(E)                                                            \
    case E:                                                                        \
        return stats ? launch<E, true>(ctx, jd, L, G, smem) : launch<E, false>(ctx, jd, L, G, smem);
    switch (epi) {
        DS_BLUR_CASE(DS_EPI_BLUR)
        DS_BLUR_CASE(DS_EPI_SUB)
        DS_BLUR_CASE(DS_EPI_RSUB)
        DS_BLUR_CASE(DS_EPI_DIV)
        DS_BLUR_CASE(DS_EPI_ATHRESH)
        default: return ds_fail(ctx, DOCSCAN_ERR_BAD_ARG, "bad blur epilogue %d", epi);
    }
#undef DS_BLUR_CASE
}
