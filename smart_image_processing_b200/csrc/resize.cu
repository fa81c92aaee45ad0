// cv2.resize(img, (new_w, new_h), INTER_AREA | INTER_CUBIC): the whole-photo fallback of process_document
// (resize_long_side, DocScanner.py:27-36, taken at :313 when no usable quad was found).
//
//   INTER_AREA (shrink)   integer scale factors on both axes: box sums, (s+2)>>2 for 2x2, cvRound(s * (1.f/area))
//                         otherwise; any other factor: OpenCV's decimation tables — per axis the source cells a
//                         destination cell covers with fp32 weights — accumulated in fp32 in table order, unfused
//                         (horizontal first, then the rows), cvRound at the end.
//   INTER_CUBIC (enlarge) OpenCV's own code path (the one cv2 takes without IPP): cubic taps (A = -0.75) evaluated
//                         in fp32 and quantised to 11 bits, exact integer horizontal pass, vertical pass in fp32
//                         (acc = S3*b3; acc += S2*b2; acc += S1*b1; acc += S0*b0, b = beta * 2^-22) for the
//                         elements cv2's 8-wide SIMD loop covers and in 22-bit fixed point for the last
//                         (width*channels) % 8 elements of a row when cv_tail_compat is set.
// The tables are a few kilobytes, built on the host with the same double / float operations as OpenCV
// (hostmath.cpp is compiled without fp contraction) and uploaded with the launch.  One thread per destination pixel;
// neighbouring threads read neighbouring source cells, so the gathers are served by L1.
#include <cmath>

#include "common.cuh"

namespace {

struct AreaEntry { int si; float alpha; };

struct ResizeJob {
    const uint8_t* src; uint8_t* dst;
    int src_pitch, dst_pitch, sw, sh, dw, dh, cn;
    // INTER_AREA, general: per destination index the [begin, end) range into the entry arrays
    const int* x_begin; const AreaEntry* x_ent;
    const int* y_begin; const AreaEntry* y_ent;
    // INTER_AREA, integer factors
    int ix, iy;
    // INTER_CUBIC: 4 source indices + 4 taps per destination index
    const int4* x_idx; const short4* x_w;
    const int4* y_idx; const short4* y_w;
    int body;            // elements of a row evaluated in fp32 (the rest in fixed point)
};

template <int CN>
__global__ void __launch_bounds__(256) resize_area_kernel(const ResizeJob J) {
    const int dx = blockIdx.x * 64 + (threadIdx.x & 63), dy = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (dx >= J.dw || dy >= J.dh) return;
    const int xb = J.x_begin[dx], xe = J.x_begin[dx + 1];
    float sum[CN];
    bool first = true;
    for (int j = J.y_begin[dy]; j < J.y_begin[dy + 1]; j++) {
        const AreaEntry ye = J.y_ent[j];
        const uint8_t* S = J.src + (size_t)ye.si * J.src_pitch;
        float buf[CN];
#pragma unroll
        for (int c = 0; c < CN; c++) buf[c] = 0.0f;
        for (int k = xb; k < xe; k++) {
            const AreaEntry xe_ = J.x_ent[k];
            const uint8_t* p = S + xe_.si * CN;
#pragma unroll
            for (int c = 0; c < CN; c++) buf[c] = __fadd_rn(buf[c], __fmul_rn((float)p[c], xe_.alpha));
        }
#pragma unroll
        for (int c = 0; c < CN; c++) sum[c] = first ? __fmul_rn(ye.alpha, buf[c]) : __fadd_rn(sum[c], __fmul_rn(ye.alpha, buf[c]));
        first = false;
    }
    uint8_t* d = J.dst + (size_t)dy * J.dst_pitch + dx * CN;
#pragma unroll
    for (int c = 0; c < CN; c++) d[c] = (uint8_t)min(max(__float2int_rn(first ? 0.0f : sum[c]), 0), 255);
}

template <int CN>
__global__ void __launch_bounds__(256) resize_area_fast_kernel(const ResizeJob J) {
    const int dx = blockIdx.x * 64 + (threadIdx.x & 63), dy = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (dx >= J.dw || dy >= J.dh) return;
    int sum[CN];
#pragma unroll
    for (int c = 0; c < CN; c++) sum[c] = 0;
    for (int j = 0; j < J.iy; j++) {
        const uint8_t* p = J.src + (size_t)(dy * J.iy + j) * J.src_pitch + (size_t)dx * J.ix * CN;
        for (int i = 0; i < J.ix; i++, p += CN) {
#pragma unroll
            for (int c = 0; c < CN; c++) sum[c] += p[c];
        }
    }
    const bool two = J.ix == 2 && J.iy == 2;
    const float scale = __fdiv_rn(1.0f, (float)(J.ix * J.iy));
    uint8_t* d = J.dst + (size_t)dy * J.dst_pitch + dx * CN;
#pragma unroll
    for (int c = 0; c < CN; c++) d[c] = two ? (uint8_t)((sum[c] + 2) >> 2) : (uint8_t)min(max(__float2int_rn(__fmul_rn((float)sum[c], scale)), 0), 255);
}

template <int CN>
__global__ void __launch_bounds__(256) resize_cubic_kernel(const ResizeJob J) {
    const int dx = blockIdx.x * 64 + (threadIdx.x & 63), dy = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (dx >= J.dw || dy >= J.dh) return;
    const int4 xi = J.x_idx[dx], yi = J.y_idx[dy];
    const short4 xw = J.x_w[dx], yw = J.y_w[dy];
    const int yidx[4] = {yi.x, yi.y, yi.z, yi.w};
    int h[4][CN];
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const uint8_t* S = J.src + (size_t)yidx[r] * J.src_pitch;
#pragma unroll
        for (int c = 0; c < CN; c++)
            h[r][c] = (int)S[xi.x * CN + c] * xw.x + (int)S[xi.y * CN + c] * xw.y + (int)S[xi.z * CN + c] * xw.z + (int)S[xi.w * CN + c] * xw.w;
    }
    const float scale = 1.0f / (2048.0f * 2048.0f);
    const float b0 = __fmul_rn((float)yw.x, scale), b1 = __fmul_rn((float)yw.y, scale), b2 = __fmul_rn((float)yw.z, scale), b3 = __fmul_rn((float)yw.w, scale);
    uint8_t* d = J.dst + (size_t)dy * J.dst_pitch + dx * CN;
#pragma unroll
    for (int c = 0; c < CN; c++) {
        int v;
        if (dx * CN + c < J.body) {
            float acc = __fmul_rn((float)h[3][c], b3);
            acc = __fadd_rn(acc, __fmul_rn((float)h[2][c], b2));
            acc = __fadd_rn(acc, __fmul_rn((float)h[1][c], b1));
            acc = __fadd_rn(acc, __fmul_rn((float)h[0][c], b0));
            v = __float2int_rn(acc);
        } else {
            v = (h[0][c] * yw.x + h[1][c] * yw.y + h[2][c] * yw.z + h[3][c] * yw.w + (1 << 21)) >> 22;
        }
        d[c] = (uint8_t)min(max(v, 0), 255);
    }
}

// cv::computeResizeAreaTab
void area_table(int ssize, int dsize, double scale, std::vector<int>* begin, std::vector<AreaEntry>* ent) {
    begin->assign(dsize + 1, 0);
    ent->clear();
    for (int dx = 0; dx < dsize; dx++) {
        (*begin)[dx] = (int)ent->size();
        const double fsx1 = dx * scale, fsx2 = fsx1 + scale;
        const double cell = std::min(scale, ssize - fsx1);
        int sx1 = (int)std::ceil(fsx1), sx2 = (int)std::floor(fsx2);
        sx2 = std::min(sx2, ssize - 1);
        sx1 = std::min(sx1, sx2);
        if (sx1 - fsx1 > 1e-3) ent->push_back({sx1 - 1, (float)((sx1 - fsx1) / cell)});
        for (int sx = sx1; sx < sx2; sx++) ent->push_back({sx, (float)(1.0 / cell)});
        if (fsx2 - sx2 > 1e-3) ent->push_back({sx2, (float)(std::min(std::min(fsx2 - sx2, 1.0), cell) / cell)});
    }
    (*begin)[dsize] = (int)ent->size();
}

template <typename T>
int upload_vec(docscan_ctx* ctx, const std::vector<T>& v, const T** out) {
    void* dev = nullptr;
    DS_TRY(ds_upload(ctx, v.data(), sizeof(T) * std::max<size_t>(v.size(), 1), &dev));
    *out = (const T*)dev;
    return DOCSCAN_OK;
}

}  // namespace

// hostmath.cpp: OpenCV's cubic taps for one destination index (fp32 arithmetic, 11-bit quantisation)
void hm_cubic_taps(int ssize, int dsize, int d, int idx[4], short w[4]);

int k_resize(docscan_ctx* ctx, const DImg& src, const DImg& dst, int interpolation, int cv_tail_compat) {
    if (src.ch != dst.ch || (src.ch != 1 && src.ch != 3)) return ds_fail(ctx, DOCSCAN_ERR_BAD_ARG, "resize: channel mismatch");
    ResizeJob J{};
    J.src = src.p; J.dst = dst.p; J.src_pitch = src.pitch; J.dst_pitch = dst.pitch;
    J.sw = src.w; J.sh = src.h; J.dw = dst.w; J.dh = dst.h; J.cn = src.ch;
    const dim3 grid((dst.w + 63) / 64, (dst.h + 3) / 4), block(256);
    const double px_in = (double)src.w * src.h * src.ch, px_out = (double)dst.w * dst.h * dst.ch;
    if (interpolation == DOCSCAN_INTER_AREA) {
        if (dst.w > src.w || dst.h > src.h)
            return ds_fail(ctx, DOCSCAN_ERR_UNSUPPORTED, "resize: INTER_AREA is implemented for shrinking only (%dx%d -> %dx%d)", src.w, src.h, dst.w, dst.h);
        const double scale_x = (double)src.w / dst.w, scale_y = (double)src.h / dst.h;
        const int ix = (int)std::nearbyint(scale_x), iy = (int)std::nearbyint(scale_y);
        if (std::fabs(scale_x - ix) < 2.220446049250313e-16 && std::fabs(scale_y - iy) < 2.220446049250313e-16) {
            J.ix = ix; J.iy = iy;
            ProfScope prof(ctx, "resize_area_fast", px_in + px_out);
            if (src.ch == 3) resize_area_fast_kernel<3><<<grid, block, 0, ctx->stream>>>(J);
            else resize_area_fast_kernel<1><<<grid, block, 0, ctx->stream>>>(J);
            DS_CHECK_LAUNCH(ctx);
            return DOCSCAN_OK;
        }
        std::vector<int> xb, yb;
        std::vector<AreaEntry> xe, ye;
        area_table(src.w, dst.w, scale_x, &xb, &xe);
        area_table(src.h, dst.h, scale_y, &yb, &ye);
        DS_TRY(upload_vec(ctx, xb, &J.x_begin)); DS_TRY(upload_vec(ctx, xe, &J.x_ent));
        DS_TRY(upload_vec(ctx, yb, &J.y_begin)); DS_TRY(upload_vec(ctx, ye, &J.y_ent));
        ProfScope prof(ctx, "resize_area", px_in + px_out);
        if (src.ch == 3) resize_area_kernel<3><<<grid, block, 0, ctx->stream>>>(J);
        else resize_area_kernel<1><<<grid, block, 0, ctx->stream>>>(J);
        DS_CHECK_LAUNCH(ctx);
        return DOCSCAN_OK;
    }
    if (interpolation == DOCSCAN_INTER_CUBIC) {
        std::vector<int4> xi(dst.w), yi(dst.h);
        std::vector<short4> xw(dst.w), yw(dst.h);
        for (int x = 0; x < dst.w; x++) hm_cubic_taps(src.w, dst.w, x, &xi[x].x, &xw[x].x);
        for (int y = 0; y < dst.h; y++) hm_cubic_taps(src.h, dst.h, y, &yi[y].x, &yw[y].x);
        DS_TRY(upload_vec(ctx, xi, &J.x_idx)); DS_TRY(upload_vec(ctx, xw, &J.x_w));
        DS_TRY(upload_vec(ctx, yi, &J.y_idx)); DS_TRY(upload_vec(ctx, yw, &J.y_w));
        const int n = dst.w * dst.ch;
        J.body = cv_tail_compat ? n - n % 8 : n;
        ProfScope prof(ctx, "resize_cubic", px_in + px_out);
        if (src.ch == 3) resize_cubic_kernel<3><<<grid, block, 0, ctx->stream>>>(J);
        else resize_cubic_kernel<1><<<grid, block, 0, ctx->stream>>>(J);
        DS_CHECK_LAUNCH(ctx);
        return DOCSCAN_OK;
    }
    return ds_fail(ctx, DOCSCAN_ERR_UNSUPPORTED, "resize: interpolation %d (only INTER_CUBIC = 2 and INTER_AREA = 3)", interpolation);
}
