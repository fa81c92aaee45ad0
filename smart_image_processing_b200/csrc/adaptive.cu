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
#include "common.cuh"

namespace {

constexpr int TW = 128, BR = 16, NT = 128;
constexpr int RPF = 132;      // ring row pitch in floats

struct AdaptLaunch {
    int k, r, delta, c_param, seg_rows, spf, ring_rows, nblk, tail_compat;
    const float* g_row;       // 8 zeros + k taps + zeros up to 8*nblk + 16
    const float* g_col;       // g_col[j] = g[r + j] for j <= r, zero up to 64
};

// rare path of the staging load (strip edges, unaligned caller buffers): kept out of line
__device__ __noinline__ uint32_t fetch_word_clamped(const uint8_t* rowp, int gx, int w) {
    uint32_t word = 0;
    for (int b = 0; b < 4; b++) word |= (uint32_t)rowp[ds_clamp(gx + b, 0, w - 1)] << (8 * b);
    return word;
}

template <int RMAX>
__global__ void __launch_bounds__(NT, 4) adaptive_gauss_kernel(const AdaptJob* __restrict__ jobs, const AdaptLaunch L) {
    const AdaptJob J = jobs[blockIdx.z];
    const int x0 = blockIdx.x * TW;
    const int y_begin = blockIdx.y * L.seg_rows;
    if (x0 >= J.w || y_begin >= J.h) return;
    const int y_end = min(J.h, y_begin + L.seg_rows);
    const int rows_out = y_end - y_begin;
    const int tid = threadIdx.x;
    const int r = L.r, r4 = L.r + L.delta;

    extern __shared__ __align__(16) uint8_t smem_raw[];
    float* s_grow = reinterpret_cast<float*>(smem_raw);            // 8*nblk + 16: 8 zeros, k taps, zeros
    float* s_gcol = s_grow + 8 * L.nblk + 16;                      // 64
    float* s_stage = s_gcol + 64;                                  // BR * spf, spf == 1 (mod 32)
    float* s_ring = s_stage + BR * L.spf;                          // ring_rows * RPF
    for (int i = tid; i < 8 * L.nblk + 16; i += NT) s_grow[i] = L.g_row[i];
    if (tid < 64) s_gcol[tid] = L.g_col[tid];

    const bool src_al = ((reinterpret_cast<uintptr_t>(J.src) | (uintptr_t)J.src_pitch) & 3) == 0;
    const int stage_words = (TW + 8 * L.nblk + 4) >> 2;   // every column the row pass can touch holds a finite value
    const int D = (2 * r + BR - 1) / BR;
    const int n_vb = (rows_out + BR - 1) / BR;
    const bool row_identity = J.w == 1, col_identity = J.h == 1;   // cv::GaussianBlur shrinks the kernel on 1-px axes
    // columns cv2's AVX2 build evaluates without fma (k <= 9: the taps are dyadic, every order is exact)
    const int tail = (L.tail_compat && L.k >= 11) ? (J.w & 7) : 0;
    const int xt_col = J.w - tail;                                  // column filter: mul+add from here on
    const int xt_row = xt_col + (tail >= 4 ? 4 : 0);                // row filter: scalar code from here on

    for (int hb = 0; hb < n_vb + D; hb++) {
        {
            const int srow_id = tid >> 3;              // 16 rows x 8 lanes; a lane strides along its row
            const uint8_t* rowp = J.src + (size_t)ds_clamp(y_begin - r + hb * BR + srow_id, 0, J.h - 1) * J.src_pitch;
            float* sp = s_stage + srow_id * L.spf;
            for (int w0 = tid & 7; w0 < stage_words; w0 += 8 * 7) {
                uint32_t wv[7];                            // issue the loads first, convert and store afterwards
#pragma unroll
                for (int j = 0; j < 7; j++) {
                    const int wi = w0 + 8 * j;
                    const int gx = x0 - r4 + 4 * wi;
                    wv[j] = 0;
                    if (wi < stage_words) wv[j] = (src_al && gx >= 0 && gx + 3 < J.w) ? ds_ldg32(rowp + gx) : fetch_word_clamped(rowp, gx, J.w);
                }
#pragma unroll
                for (int j = 0; j < 7; j++) {
                    const int wi = w0 + 8 * j;
                    if (wi >= stage_words) continue;
                    // u8 -> fp32 without the slow I2F unit: 2^23 + v is exact in fp32, subtract 2^23 again
#pragma unroll
                    for (int b = 0; b < 4; b++) sp[4 * wi + b] = __fsub_rn(__uint_as_float(0x4B000000u | ((wv[j] >> (8 * b)) & 255u)), 8388608.0f);
                }
            }
        }
        __syncthreads();
        {   // ---- row pass: out[c] = sum_i g[i] * f[c + i - r], taps in increasing i, one fma each.
            // lane <-> staged row (16 rows x 2 column groups per warp): with the row pitch == 1 (mod 32) a warp's loads
            // hit 32 different banks.  8 outputs per group, sliding window of 16 coefficients in registers.
            const int lane = tid & 31, wrp = tid >> 5;
            const int hr = lane & 15, hh = lane >> 4;
            const float* srow = s_stage + hr * L.spf + L.delta;
            const int slot = (hb * BR + hr) % L.ring_rows;
#pragma unroll 1
            for (int it = 0; it < 2; it++) {
                const int pr = wrp * 2 + it;
                const int c0 = 8 * ((pr >> 1) * 4 + (pr & 1) + 2 * hh);
                const float* sp = srow + c0;
                float acc[8];
#pragma unroll
                for (int i = 0; i < 8; i++) acc[i] = 0.0f;
                float G[16];
                {
                    const float4 t0 = *reinterpret_cast<const float4*>(s_grow), t1 = *reinterpret_cast<const float4*>(s_grow + 4);
                    G[8] = t0.x; G[9] = t0.y; G[10] = t0.z; G[11] = t0.w; G[12] = t1.x; G[13] = t1.y; G[14] = t1.z; G[15] = t1.w;
                }
                for (int b = 0; b < L.nblk; b++) {
#pragma unroll
                    for (int i = 0; i < 8; i++) G[i] = G[i + 8];
                    const float4 t0 = *reinterpret_cast<const float4*>(s_grow + 8 * b + 8), t1 = *reinterpret_cast<const float4*>(s_grow + 8 * b + 12);
                    G[8] = t0.x; G[9] = t0.y; G[10] = t0.z; G[11] = t0.w; G[12] = t1.x; G[13] = t1.y; G[14] = t1.z; G[15] = t1.w;
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const float f = sp[8 * b + u];
#pragma unroll
                        for (int o = 0; o < 8; o++) acc[o] = __fmaf_rn(f, G[u - o + 8], acc[o]);
                    }
                }
                if (row_identity) {
#pragma unroll
                    for (int o = 0; o < 8; o++) acc[o] = sp[r + o];
                } else if (xt_row < J.w && x0 + c0 + 7 >= xt_row) {
                    // cv2's row filter leaves the last w % 4 columns to scalar code: mul+add per tap, except that the
                    // (k-1) % 4 remainder taps are fma (oracle/docscan_oracle.c, A.9).  At most 3 columns per row.
                    const int first_fused = L.k - ((L.k - 1) & 3);
#pragma unroll
                    for (int o = 0; o < 8; o++) {
                        const int x = x0 + c0 + o;
                        if (x < xt_row || x >= J.w) continue;
                        float a = __fmul_rn(s_grow[8], sp[o]);
                        for (int i = 1; i < L.k; i++) {
                            const float f = sp[o + i];
                            a = i >= first_fused ? __fmaf_rn(f, s_grow[8 + i], a) : __fadd_rn(a, __fmul_rn(s_grow[8 + i], f));
                        }
                        acc[o] = a;
                    }
                }
                float4* dst = reinterpret_cast<float4*>(s_ring + slot * RPF + c0);
                dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
                dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
            }
        }
        __syncthreads();
        if (hb < D) continue;
        // ---- column pass: thread = one column, 16 rows; centre of output o sits at window index o + RMAX
        const int vb = hb - D;
        const int col = tid;
        const int x = x0 + col;
        float Wn[BR + 2 * RMAX];
        const int shift = RMAX - r;                 // window index i <-> virtual row vb*BR + i - shift
        const int slot0 = (vb * BR) % L.ring_rows;
        if (shift == 0) {
            // exact instantiation (k = 2*RMAX+1): the ring size is a compile-time constant and a window starts at a
            // multiple of 16 rows, so every ring offset (wrap included) is an immediate
            constexpr int RR = ((2 * RMAX + BR - 1) / BR + 1) * BR;
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
        } else {
#pragma unroll
            for (int i = 0; i < BR + 2 * RMAX; i++) {
                const int rel = i - shift;
                float v = 0.0f;
                if (rel >= 0 && rel <= BR - 1 + 2 * r) {
                    int slot = slot0 + rel;
                    if (slot >= L.ring_rows) slot -= L.ring_rows;
                    v = s_ring[slot * RPF + col];
                }
                Wn[i] = v;
            }
        }
        float acc[BR];
        const float gc0 = s_gcol[0];
#pragma unroll
        for (int o = 0; o < BR; o++) acc[o] = __fmul_rn(gc0, Wn[o + RMAX]);
        if (x < xt_col) {
#pragma unroll
            for (int j = 1; j <= RMAX; j++) {
                const float gj = s_gcol[j];
#pragma unroll
                for (int o = 0; o < BR; o++) acc[o] = __fmaf_rn(__fadd_rn(Wn[o + RMAX + j], Wn[o + RMAX - j]), gj, acc[o]);
            }
        } else {
#pragma unroll
            for (int j = 1; j <= RMAX; j++) {
                const float gj = s_gcol[j];
#pragma unroll
                for (int o = 0; o < BR; o++) acc[o] = __fadd_rn(acc[o], __fmul_rn(gj, __fadd_rn(Wn[o + RMAX + j], Wn[o + RMAX - j])));
            }
        }
        if (x < J.w) {
            const int y0 = y_begin + vb * BR;
            const uint8_t* sp = J.src + (size_t)y0 * J.src_pitch + x;
            uint8_t* dp = J.dst + (size_t)y0 * J.dst_pitch + x;
            const int rows = min(BR, y_end - y0);
            int cpx[BR];                                   // centre pixels, loaded up front
#pragma unroll
            for (int o = 0; o < BR; o++) { cpx[o] = o < rows ? (int)*sp : 0; sp += J.src_pitch; }
#pragma unroll
            for (int o = 0; o < BR; o++) {
                if (o >= rows) break;
                const float m = col_identity ? Wn[o + RMAX] : acc[o];
                const int mean = min(max(__float2int_rn(m), 0), 255);
                *dp = (cpx[o] - mean > -L.c_param) ? 255 : 0;
                dp += J.dst_pitch;
            }
        }
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

struct AdaptGridInfo { int strips, max_w, max_h, n, seg_min; };

template <int RMAX>
int launch_adaptive(docscan_ctx* ctx, const AdaptJob* jd, AdaptLaunch L, const AdaptGridInfo& G, size_t smem) {
    if (smem > 48 * 1024)
        DS_CUDA(ctx, cudaFuncSetAttribute(adaptive_gauss_kernel<RMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    DS_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, adaptive_gauss_kernel<RMAX>, NT, smem));
    L.seg_rows = ds_pick_seg_rows(per_sm * ctx->sm_count, G.strips, G.max_h, G.seg_min, BR);
    dim3 grid((G.max_w + TW - 1) / TW, (G.max_h + L.seg_rows - 1) / L.seg_rows, G.n);
    adaptive_gauss_kernel<RMAX><<<grid, NT, smem, ctx->stream>>>(jd, L);
    DS_CHECK_LAUNCH(ctx);
    return DOCSCAN_OK;
}

}  // namespace

int k_adaptive_gauss_jobs(docscan_ctx* ctx, int k, int c, int cv_tail_compat, const AdaptJob* jobs_host, int n,
                          int max_w, int max_h) {
    if (k < 3 || (k & 1) == 0 || k > 65)
        return ds_fail(ctx, DOCSCAN_ERR_UNSUPPORTED, "adaptive GAUSSIAN_C block size must be odd and in 3..65 (got %d)", k);
    AdaptLaunch L{};
    L.k = k; L.r = k / 2; L.delta = (4 - (L.r & 3)) & 3; L.c_param = c; L.tail_compat = cv_tail_compat;
    L.nblk = (k + 7 + 7) / 8;
    L.spf = TW + 8 * L.nblk + 8;
    L.spf += (33 - (L.spf & 31)) & 31;                 // pitch == 1 (mod 32): lanes of a warp read different banks
    L.ring_rows = ((2 * L.r + BR - 1) / BR + 1) * BR;
    // coefficient tables (device, cached per k)
    const uint64_t key = ((uint64_t)7 << 32) | (uint32_t)k;
    const int n_row = 8 * L.nblk + 16;
    auto it = ctx->tables.find(key);
    if (it == ctx->tables.end()) {
        std::vector<float> g(k), host(n_row + 64 + k, 0.0f);
        docscan_gaussian_kernel_f32(k, g.data());
        for (int i = 0; i < k; i++) host[8 + i] = g[i];
        for (int j = 0; j <= L.r; j++) host[n_row + j] = g[L.r + j];
        for (int i = 0; i < k; i++) host[n_row + 64 + i] = g[i];
        void* dev = nullptr;
        DS_CUDA(ctx, cudaMalloc(&dev, host.size() * sizeof(float)));
        DS_CUDA(ctx, cudaMemcpyAsync(dev, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        DS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        it = ctx->tables.emplace(key, dev).first;
    }
    const float* tab = (const float*)it->second;
    L.g_row = tab; L.g_col = tab + n_row;

    const AdaptGridInfo G{n * ((max_w + TW - 1) / TW), max_w, max_h, n, max(64, 4 * L.r)};
    const size_t smem = sizeof(float) * ((size_t)n_row + 64 + (size_t)BR * L.spf + (size_t)L.ring_rows * RPF);
    void* dev = nullptr;
    DS_TRY(ds_upload(ctx, jobs_host, sizeof(AdaptJob) * n, &dev));
    const AdaptJob* jd = (const AdaptJob*)dev;
    double px = 0;
    for (int i = 0; i < n; i++) px += (double)jobs_host[i].w * jobs_host[i].h;
    int rc;
    {
    ProfScope prof(ctx, "adaptive_gauss_k" + std::to_string(k), 2.0 * px);
    if (L.r <= 5) rc = launch_adaptive<5>(ctx, jd, L, G, smem);
    else if (L.r <= 9) rc = launch_adaptive<9>(ctx, jd, L, G, smem);
    else if (L.r <= 13) rc = launch_adaptive<13>(ctx, jd, L, G, smem);
    else if (L.r == 15) rc = launch_adaptive<15>(ctx, jd, L, G, smem);      // k = 31 (GUI preset), exact
    else if (L.r <= 17) rc = launch_adaptive<17>(ctx, jd, L, G, smem);      // k = 35 (CLI default), exact
    else if (L.r <= 25) rc = launch_adaptive<25>(ctx, jd, L, G, smem);
    else rc = launch_adaptive<32>(ctx, jd, L, G, smem);
    }
    DS_TRY(rc);
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
    mask_blend_kernel<<<grid, 128, 0, ctx->stream>>>((const BlendJob*)dev, dilate_iters, write_mask_only);
    DS_CHECK_LAUNCH(ctx);
    return DOCSCAN_OK;
}
