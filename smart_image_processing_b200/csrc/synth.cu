// Bench support: deterministic synthetic page photos rendered on the device, so that batches of
// 12 MP inputs (BASELINE.json configs 2 and 3) never have to cross PCIe or be stored on the host.
// A page (A-series portrait, light paper, ~70 lines of dark word boxes) is placed in the photo through a
// per-seed homography, lit by a diagonal illumination gradient, tinted per channel and given +-3 noise;
// everything outside the page is a dark desk.  This is input generation, not part of the measured path.
#include "common.cuh"

namespace {

__device__ __host__ inline uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

struct SynthParams {
    float hinv[9];       // photo pixel -> page coordinates
    float page_w, page_h;
    uint32_t seed;
};

__device__ float page_value(float u, float v, const SynthParams& P) {
    if (u < 0.f || v < 0.f || u >= P.page_w || v >= P.page_h) return -1.f;     // off the page
    const float line_pitch = P.page_h / 75.5f;
    const float top = 2.6f * line_pitch;
    const float lf = (v - top) / line_pitch;
    const int line = (int)floorf(lf);
    if (lf < 0.f || line >= 70) return 235.f;
    const uint32_t hl = mix32(P.seed * 2654435761u + (uint32_t)line * 97u + 13u);
    const float text_h = line_pitch * (0.24f + 0.22f * (float)(hl & 255u) / 255.f);
    if ((lf - (float)line) * line_pitch > text_h) return 235.f;
    const float margin = 0.05f * P.page_w;
    const float cell = P.page_w / 26.f;
    const float uf = (u - margin) / cell;
    const int word = (int)floorf(uf);
    if (uf < 0.f || u > P.page_w - margin) return 235.f;
    const uint32_t hw = mix32(hl + (uint32_t)word * 7919u);
    if ((hw & 127u) > 108u) return 235.f;                                       // ~15 % of the cells stay empty
    const float gap = cell * (0.12f + 0.2f * (float)((hw >> 8) & 255u) / 255.f);
    const float inside = (uf - (float)word) * cell;
    if (inside < gap) return 235.f;
    return 20.f + (float)((hw >> 16) % 70u);
}

__global__ void __launch_bounds__(256) synth_page_kernel(uint8_t* __restrict__ dst, int pitch, int w, int h, SynthParams P) {
    const int x = blockIdx.x * 64 + (threadIdx.x & 63);
    const int y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= w || y >= h) return;
    float acc = 0.f;
#pragma unroll
    for (int s = 0; s < 4; s++) {                                               // 2x2 supersampling: soft edges
        const float fx = (float)x + ((s & 1) ? 0.25f : -0.25f), fy = (float)y + ((s & 2) ? 0.25f : -0.25f);
        const float ww = P.hinv[6] * fx + P.hinv[7] * fy + P.hinv[8];
        const float u = (P.hinv[0] * fx + P.hinv[1] * fy + P.hinv[2]) / ww;
        const float v = (P.hinv[3] * fx + P.hinv[4] * fy + P.hinv[5]) / ww;
        const float pv = page_value(u, v, P);
        acc += pv < 0.f ? 40.f : pv;
    }
    float val = acc * 0.25f;
    val *= 0.55f + 0.45f * (0.6f * (float)x / (float)w + 0.4f * (float)y / (float)h);
    const uint32_t n = mix32(P.seed ^ mix32((uint32_t)y * 65537u + (uint32_t)x));
    const float noise = ((float)(n & 255u) + (float)((n >> 8) & 255u) + (float)((n >> 16) & 255u) + (float)(n >> 24) - 510.f) * (3.f / 147.8f);
    const float gains[3] = {0.97f, 1.0f, 1.02f};
    uint8_t* p = dst + (size_t)y * pitch + (size_t)x * 3;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float o = val * gains[c] + noise;
        p[c] = (uint8_t)fminf(fmaxf(rintf(o), 0.f), 255.f);
    }
}

}  // namespace

int k_synth_page(docscan_ctx* ctx, uint64_t seed, const DImg& dst, float quad_out[8]) {
    const float W = (float)dst.w, H = (float)dst.h;
    const uint32_t s32 = mix32((uint32_t)seed ^ mix32((uint32_t)(seed >> 32) + 0x9e3779b9u));
    const float base[8] = {0.10f * W, 0.07f * H, 0.90f * W, 0.09f * H, 0.93f * W, 0.93f * H, 0.07f * W, 0.91f * H};
    const float jitter = 0.02f * (W < H ? W : H);
    float quad[8];
    for (int i = 0; i < 8; i++) {
        const uint32_t r = mix32(s32 + 101u * (uint32_t)(i + 1));
        quad[i] = base[i] + jitter * ((float)(r & 0xffffu) / 32767.5f - 1.0f);
    }
    // page rectangle (A-series portrait) sized like the photo's page
    const float ph = 0.85f * H, pw = ph / 1.41421356f;
    const float rect[8] = {0, 0, pw - 1, 0, pw - 1, ph - 1, 0, ph - 1};
    double m[9];
    DS_TRY(docscan_get_perspective_transform(quad, rect, m));     // photo -> page
    SynthParams P;
    for (int i = 0; i < 9; i++) P.hinv[i] = (float)m[i];
    P.page_w = pw; P.page_h = ph; P.seed = s32;
    dim3 grid((dst.w + 63) / 64, (dst.h + 3) / 4);
    ProfScope prof(ctx, "synth_page", 3.0 * dst.w * dst.h);
    synth_page_kernel<<<grid, 256, 0, ctx->stream>>>(dst.p, dst.pitch, dst.w, dst.h, P);
    DS_CHECK_LAUNCH(ctx);
    if (quad_out)
        for (int i = 0; i < 8; i++) quad_out[i] = quad[i];
    return DOCSCAN_OK;
}
