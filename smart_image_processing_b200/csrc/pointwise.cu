// Pointwise stages: BGR2GRAY, subtract / divide / max / masked select, LUT, threshold, min-max + histogram.
// All are pure streaming kernels (HBM-bound): 4 pixels per thread through aligned 32-bit accesses when
// the buffers allow it, byte accesses otherwise.
#include "common.cuh"

enum PwOp { PW_SUB = 0, PW_DIV = 1, PW_MAX = 2, PW_SELECT = 3, PW_LUT = 4, PW_THRESH = 5, PW_GRAY_BGR = 6, PW_GRAY_RGB = 7 };

struct PwJob {
    const uint8_t* a; const uint8_t* b; uint8_t* dst;
    int pa, pb, pd, w, h;
    const uint8_t* lut;       // PW_LUT
    const int32_t* t_dev;     // PW_THRESH (device scalar) or null -> t
    int t;
};

template <int OP>
__device__ __forceinline__ uint8_t pw_px(uint8_t a, uint8_t b, const uint8_t* lut, int t) {
    if (OP == PW_SUB) return a > b ? a - b : 0;
    if (OP == PW_DIV) return ds_div255(a, b);
    if (OP == PW_MAX) return a > b ? a : b;
    if (OP == PW_SELECT) return b == 0 ? 255 : a;
    if (OP == PW_LUT) return lut[a];
    if (OP == PW_THRESH) return a > t ? 255 : 0;
    return 0;
}

__device__ __forceinline__ uint8_t gray_px(int c0, int c1, int c2, bool rgb) {
    // cv::cvtColor BGR2GRAY 15-bit fixed point (DocScanner.py:316)
    const int k0 = rgb ? 9798 : 3735, k2 = rgb ? 3735 : 9798;
    return (uint8_t)((k0 * c0 + 19235 * c1 + k2 * c2 + 16384) >> 15);
}

template <int OP>
__global__ void __launch_bounds__(128) pw_kernel(const PwJob* __restrict__ jobs) {
    const PwJob J = jobs[blockIdx.z];
    const int y = blockIdx.y;
    if (y >= J.h) return;
    const int x = (blockIdx.x * 128 + threadIdx.x) * 4;
    if (x >= J.w) return;
    int t = J.t;
    if (OP == PW_THRESH && J.t_dev) t = *J.t_dev;
    const uint8_t* ra = J.a + (size_t)y * J.pa;
    uint8_t* rd = J.dst + (size_t)y * J.pd;
    constexpr bool TWO = (OP == PW_SUB || OP == PW_DIV || OP == PW_MAX || OP == PW_SELECT);
    constexpr bool GRAY = (OP == PW_GRAY_BGR || OP == PW_GRAY_RGB);
    const uint8_t* rb = TWO ? J.b + (size_t)y * J.pb : nullptr;
    const bool full = x + 4 <= J.w;
    if (GRAY) {
        const uint8_t* s = ra + (size_t)x * 3;
        uint8_t o[4];
        if (full && ((reinterpret_cast<uintptr_t>(s) & 3) == 0)) {
            uint32_t w0 = ds_ldg32(s), w1 = ds_ldg32(s + 4), w2 = ds_ldg32(s + 8);
            o[0] = gray_px(w0 & 255, (w0 >> 8) & 255, (w0 >> 16) & 255, OP == PW_GRAY_RGB);
            o[1] = gray_px(w0 >> 24, w1 & 255, (w1 >> 8) & 255, OP == PW_GRAY_RGB);
            o[2] = gray_px((w1 >> 16) & 255, w1 >> 24, w2 & 255, OP == PW_GRAY_RGB);
            o[3] = gray_px((w2 >> 8) & 255, (w2 >> 16) & 255, w2 >> 24, OP == PW_GRAY_RGB);
        } else {
            for (int i = 0; i < 4 && x + i < J.w; i++)
                o[i] = gray_px(s[3 * i], s[3 * i + 1], s[3 * i + 2], OP == PW_GRAY_RGB);
        }
        if (full && ((reinterpret_cast<uintptr_t>(rd + x) & 3) == 0))
            *reinterpret_cast<uint32_t*>(rd + x) = o[0] | (o[1] << 8) | (o[2] << 16) | ((uint32_t)o[3] << 24);
        else
            for (int i = 0; i < 4 && x + i < J.w; i++) rd[x + i] = o[i];
        return;
    }
    const bool al = full && (((reinterpret_cast<uintptr_t>(ra + x) | reinterpret_cast<uintptr_t>(rd + x) |
                               (TWO ? reinterpret_cast<uintptr_t>(rb + x) : 0)) & 3) == 0);
    if (al) {
        uint32_t wa = ds_ldg32(ra + x), wb = TWO ? ds_ldg32(rb + x) : 0, o = 0;
#pragma unroll
        for (int i = 0; i < 4; i++)
            o |= (uint32_t)pw_px<OP>((wa >> (8 * i)) & 255, (wb >> (8 * i)) & 255, J.lut, t) << (8 * i);
        *reinterpret_cast<uint32_t*>(rd + x) = o;
    } else {
        for (int i = 0; i < 4 && x + i < J.w; i++) rd[x + i] = pw_px<OP>(ra[x + i], TWO ? rb[x + i] : 0, J.lut, t);
    }
}

// LUT for the library's own planes (16-byte aligned rows): the 256-entry table sits in shared memory, a thread owns a
// 16-pixel column chunk and walks LUT_ROWS rows, every global access is 128 bits wide.
constexpr int LUT_ROWS = 8;
__global__ void __launch_bounds__(256) lut16_kernel(const PwJob* __restrict__ jobs, int chunks, int row_groups) {
    __shared__ uint32_t s_lut_w[64];
    const PwJob J = jobs[blockIdx.y];
    if (threadIdx.x < 64) s_lut_w[threadIdx.x] = ds_ldg32(J.lut + 4 * threadIdx.x);
    __syncthreads();
    const uint8_t* s_lut = reinterpret_cast<const uint8_t*>(s_lut_w);
    const int id = blockIdx.x * 256 + threadIdx.x;
    const int rg = id / chunks, xc = id - rg * chunks;
    const int x = xc * 16, y0 = rg * LUT_ROWS;
    if (rg >= row_groups || x >= J.w || y0 >= J.h) return;
    const int rows = min(LUT_ROWS, J.h - y0);
    const bool full = x + 16 <= J.w;
    uint4 v[LUT_ROWS];
#pragma unroll
    for (int i = 0; i < LUT_ROWS; i++)
        if (i < rows) v[i] = __ldg(reinterpret_cast<const uint4*>(J.a + (size_t)(y0 + i) * J.pa + x));
#pragma unroll
    for (int i = 0; i < LUT_ROWS; i++) {
        if (i >= rows) break;
        uint32_t in[4] = {v[i].x, v[i].y, v[i].z, v[i].w}, out[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t b0 = s_lut[in[j] & 255u], b1 = s_lut[(in[j] >> 8) & 255u], b2 = s_lut[(in[j] >> 16) & 255u], b3 = s_lut[in[j] >> 24];
            out[j] = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
        }
        uint8_t* dp = J.dst + (size_t)(y0 + i) * J.pd + x;
        if (full) *reinterpret_cast<uint4*>(dp) = make_uint4(out[0], out[1], out[2], out[3]);
        else for (int b = 0; b < J.w - x; b++) dp[b] = (uint8_t)(out[b >> 2] >> (8 * (b & 3)));
    }
}

static int pw_launch(docscan_ctx* ctx, int op, const PwJob* jobs_host, int n, int max_w, int max_h) {
    void* dev = nullptr;
    DS_TRY(ds_upload(ctx, jobs_host, sizeof(PwJob) * n, &dev));
    dim3 grid((max_w + 511) / 512, max_h, n), block(128);
    const PwJob* j = (const PwJob*)dev;
    static const char* names[] = {"pw_subtract", "pw_divide255", "pw_max", "pw_mask_select", "pw_lut", "pw_threshold", "pw_bgr2gray", "pw_rgb2gray"};
    double px = 0;
    for (int i = 0; i < n; i++) px += (double)jobs_host[i].w * jobs_host[i].h;
    ProfScope prof(ctx, names[op], px * (op <= PW_SELECT ? 3 : (op >= PW_GRAY_BGR ? 4 : 2)));
    switch (op) {
        case PW_SUB: pw_kernel<PW_SUB><<<grid, block, 0, ctx->stream>>>(j); break;
        case PW_DIV: pw_kernel<PW_DIV><<<grid, block, 0, ctx->stream>>>(j); break;
        case PW_MAX: pw_kernel<PW_MAX><<<grid, block, 0, ctx->stream>>>(j); break;
        case PW_SELECT: pw_kernel<PW_SELECT><<<grid, block, 0, ctx->stream>>>(j); break;
        case PW_LUT: pw_kernel<PW_LUT><<<grid, block, 0, ctx->stream>>>(j); break;
        case PW_THRESH: pw_kernel<PW_THRESH><<<grid, block, 0, ctx->stream>>>(j); break;
        case PW_GRAY_BGR: pw_kernel<PW_GRAY_BGR><<<grid, block, 0, ctx->stream>>>(j); break;
        case PW_GRAY_RGB: pw_kernel<PW_GRAY_RGB><<<grid, block, 0, ctx->stream>>>(j); break;
        default: return ds_fail(ctx, DOCSCAN_ERR_BAD_ARG, "bad pointwise op %d", op);
    }
    DS_CHECK_LAUNCH(ctx);
    return DOCSCAN_OK;
}

int k_bgr2gray(docscan_ctx* ctx, const DImg& src, const DImg& dst, int swap_rb) {
    PwJob j{}; j.a = src.p; j.pa = src.pitch; j.dst = dst.p; j.pd = dst.pitch; j.w = dst.w; j.h = dst.h;
    return pw_launch(ctx, swap_rb ? PW_GRAY_RGB : PW_GRAY_BGR, &j, 1, j.w, j.h);
}
int k_binary_op(docscan_ctx* ctx, int op, const DImg& a, const DImg& b, const DImg& dst) {
    PwJob j{}; j.a = a.p; j.pa = a.pitch; j.b = b.p; j.pb = b.pitch; j.dst = dst.p; j.pd = dst.pitch; j.w = dst.w; j.h = dst.h;
    return pw_launch(ctx, op, &j, 1, j.w, j.h);
}
int k_apply_lut(docscan_ctx* ctx, const DImg& src, const uint8_t* lut_dev, const DImg& dst) {
    PwJob j{}; j.a = src.p; j.pa = src.pitch; j.dst = dst.p; j.pd = dst.pitch; j.w = dst.w; j.h = dst.h; j.lut = lut_dev;
    return pw_launch(ctx, PW_LUT, &j, 1, j.w, j.h);
}
int k_threshold(docscan_ctx* ctx, const DImg& src, const int32_t* t_dev, int t_imm, const DImg& dst) {
    PwJob j{}; j.a = src.p; j.pa = src.pitch; j.dst = dst.p; j.pd = dst.pitch; j.w = dst.w; j.h = dst.h; j.t_dev = t_dev; j.t = t_imm;
    return pw_launch(ctx, PW_THRESH, &j, 1, j.w, j.h);
}

// LUT applied to many pages in one launch (pipeline).  jobs: (src, dst, lut) triples.
int k_apply_lut_jobs(docscan_ctx* ctx, const DImg* src, const DImg* dst, const uint8_t* const* luts, int n) {
    std::vector<PwJob> jobs(n);
    int mw = 0, mh = 0;
    for (int i = 0; i < n; i++) {
        PwJob& j = jobs[i];
        j = PwJob{};
        j.a = src[i].p; j.pa = src[i].pitch; j.dst = dst[i].p; j.pd = dst[i].pitch; j.w = dst[i].w; j.h = dst[i].h;
        j.lut = luts[i];
        mw = max(mw, j.w); mh = max(mh, j.h);
    }
    bool aligned16 = true;
    for (int i = 0; i < n && aligned16; i++) {
        const PwJob& j = jobs[i];
        const int need = ((j.w + 15) >> 4) << 4;      // whole chunks are read: the source pitch must cover them
        aligned16 = ((reinterpret_cast<uintptr_t>(j.a) | reinterpret_cast<uintptr_t>(j.dst) | (uintptr_t)j.pa | (uintptr_t)j.pd) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(j.lut) & 3) == 0 && j.pa >= need;
    }
    if (!aligned16) return pw_launch(ctx, PW_LUT, jobs.data(), n, mw, mh);
    void* dev = nullptr;
    DS_TRY(ds_upload(ctx, jobs.data(), sizeof(PwJob) * n, &dev));
    const int chunks = (mw + 15) >> 4, row_groups = (mh + LUT_ROWS - 1) / LUT_ROWS;
    double px = 0;
    for (int i = 0; i < n; i++) px += (double)jobs[i].w * jobs[i].h;
    ProfScope prof(ctx, "pw_lut", 2.0 * px);
    lut16_kernel<<<dim3((chunks * row_groups + 255) / 256, n), 256, 0, ctx->stream>>>((const PwJob*)dev, chunks, row_groups);
    DS_CHECK_LAUNCH(ctx);
    return DOCSCAN_OK;
}

// ---- min/max + 256-bin histogram ---------------------------------------------------------------------
__global__ void __launch_bounds__(256) stats_kernel(const uint8_t* __restrict__ src, int pitch, int w, int h,
                                                    uint32_t* __restrict__ minmax, uint32_t* __restrict__ hist) {
    __shared__ uint32_t s_hist[8][256];   // one private histogram per warp
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 8 * 256; i += 256) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    uint32_t lo = 255, hi = 0;
    const int words = (w + 3) >> 2;
    const bool al = ((reinterpret_cast<uintptr_t>(src) | (uintptr_t)pitch) & 3) == 0;
    for (int y = blockIdx.x; y < h; y += gridDim.x) {
        const uint8_t* row = src + (size_t)y * pitch;
        for (int wi = threadIdx.x; wi < words; wi += 256) {
            const int x = wi * 4;
            uint32_t v;
            int nvalid = min(4, w - x);
            if (al && nvalid == 4) v = ds_ldg32(row + x);
            else {
                v = 0;
                for (int i = 0; i < nvalid; i++) v |= (uint32_t)row[x + i] << (8 * i);
            }
            for (int i = 0; i < nvalid; i++) {
                uint32_t b = (v >> (8 * i)) & 255;
                lo = min(lo, b); hi = max(hi, b);
                if (hist) atomicAdd(&s_hist[warp][b], 1u);
            }
        }
    }
    if (minmax) {
        for (int o = 16; o; o >>= 1) {
            lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if ((threadIdx.x & 31) == 0) { atomicMin(&minmax[0], lo); atomicMax(&minmax[1], hi); }
    }
    if (hist) {
        __syncthreads();
        uint32_t s = 0;
        for (int k = 0; k < 8; k++) s += s_hist[k][threadIdx.x];
        if (s) atomicAdd(&hist[threadIdx.x], s);
    }
}

__global__ void fill_u32_kernel(uint32_t* p, int n, uint32_t even, uint32_t odd) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (i & 1) ? odd : even;
}

int k_zero_u32(docscan_ctx* ctx, uint32_t* p, int n, uint32_t value_even, uint32_t value_odd) {
    fill_u32_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(p, n, value_even, value_odd);
    DS_CHECK_LAUNCH(ctx);
    return DOCSCAN_OK;
}

int k_stats(docscan_ctx* ctx, const DImg& src, uint32_t* minmax_dev, uint32_t* hist_dev) {
    int blocks = min(src.h, ctx->sm_count * 4);
    ProfScope prof(ctx, "stats_minmax_hist", (double)src.w * src.h);
    stats_kernel<<<blocks, 256, 0, ctx->stream>>>(src.p, src.pitch, src.w, src.h, minmax_dev, hist_dev);
    DS_CHECK_LAUNCH(ctx);
    return DOCSCAN_OK;
}
