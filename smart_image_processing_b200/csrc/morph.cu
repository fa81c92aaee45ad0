// cv2.erode / cv2.dilate with a MORPH_RECT kw x kh element (DocScanner.py:199-200,211-212,251-254;
// morph_seq grayscale_erosion / binary_closing).  A rectangle is separable, so one 2-D pass is a
// horizontal 1-D min/max followed by a vertical one.  Each 1-D pass is O(log k) per pixel: the tile is
// held in shared memory as packed bytes (4 pixels per word), window minima of length 1,2,4,..,P are
// built by doubling (A_2p[i] = op(A_p[i], A_p[i+p])) with ping-pong buffers, and the final window of
// length k is op(A_P[i], A_P[i+k-P]).  Pixels outside the image are ignored, exactly like OpenCV's
// default morphology border: they are loaded as the neutral element (255 for erode, 0 for dilate).
// The last pass can fuse the black-hat subtraction (close(src) - src) and a 256-bin histogram.
#include "common.cuh"

namespace {

constexpr int NT = 256;

struct MorphLaunch {
    int k, a, is_dilate;
    int len, cnt;        // outputs along the axis per tile / lines across the axis per tile (words for the V pass)
    int nw;              // shared row pitch in words (H pass) or rows per tile incl. halo (V pass)
};

__device__ __forceinline__ uint32_t op4(uint32_t a, uint32_t b, int is_dilate) {
    return is_dilate ? __vmaxu4(a, b) : __vminu4(a, b);
}

__device__ __forceinline__ uint32_t load_word(const MorphJob& J, int gy, int gx, uint32_t neutral, bool al) {
    if (gy < 0 || gy >= J.h || gx + 3 < 0 || gx >= J.w) return neutral;
    const uint8_t* rowp = J.src + (size_t)gy * J.src_pitch;
    if (al && gx >= 0 && gx + 3 < J.w) return ds_ldg32(rowp + gx);
    uint32_t word = 0;
#pragma unroll
    for (int b = 0; b < 4; b++) {
        const int x = gx + b;
        const uint32_t v = (x >= 0 && x < J.w) ? rowp[x] : (neutral & 255u);
        word |= v << (8 * b);
    }
    return word;
}

__device__ __forceinline__ void store_word(const MorphJob& J, int y, int x, uint32_t res, uint32_t* s_hist, bool dst_al) {
    if (y >= J.h || x >= J.w) return;
    const int nvalid = min(4, J.w - x);
    if (J.ref) {
        const uint8_t* rp = J.ref + (size_t)y * J.ref_pitch + x;
        uint32_t out = 0;
        for (int b = 0; b < nvalid; b++) {
            const int v = (int)((res >> (8 * b)) & 255u) - (int)rp[b];
            out |= (uint32_t)max(v, 0) << (8 * b);
        }
        res = out;
    }
    if (s_hist)
        for (int b = 0; b < nvalid; b++) atomicAdd(&s_hist[(threadIdx.x >> 5) * 256 + ((res >> (8 * b)) & 255u)], 1u);
    uint8_t* dp = J.dst + (size_t)y * J.dst_pitch + x;
    if (dst_al && nvalid == 4) *reinterpret_cast<uint32_t*>(dp) = res;
    else
        for (int b = 0; b < nvalid; b++) dp[b] = (uint8_t)(res >> (8 * b));
}

// byte-granular read of 4 consecutive bytes starting at byte column c of a packed row
__device__ __forceinline__ uint32_t read_bytes(const uint32_t* row, int c, int last_word) {
    const int w = c >> 2, sh = (c & 3) * 8;
    const uint32_t lo = row[min(w, last_word)], hi = row[min(w + 1, last_word)];
    return __funnelshift_r(lo, hi, sh);
}

template <int AXIS>
__global__ void __launch_bounds__(NT) morph_1d_kernel(const MorphJob* __restrict__ jobs, const MorphLaunch L) {
    const MorphJob J = jobs[blockIdx.z];
    extern __shared__ __align__(16) uint32_t smem_u32[];
    const int tid = threadIdx.x;
    const uint32_t neutral = L.is_dilate ? 0u : 0xffffffffu;
    const bool src_al = ((reinterpret_cast<uintptr_t>(J.src) | (uintptr_t)J.src_pitch) & 3) == 0;
    const bool dst_al = ((reinterpret_cast<uintptr_t>(J.dst) | (uintptr_t)J.dst_pitch) & 3) == 0;
    const int k = L.k;
    int P = 1;
    while (P * 2 <= k) P *= 2;

    // thread layout: 16 lines x 16 lanes; a lane strides along its line, so no index division anywhere
    const int ty = tid >> 4, tx = tid & 15;
    if (AXIS == 0) {
        const int tx0 = blockIdx.x * L.len, ty0 = blockIdx.y * L.cnt;       // L.cnt == 16 rows
        if (tx0 >= J.w || ty0 >= J.h) return;
        const int nw = L.nw, total = nw * L.cnt;
        uint32_t* bufA = smem_u32;
        uint32_t* bufB = smem_u32 + total;
        uint32_t* s_hist = J.hist ? smem_u32 + 2 * total : nullptr;
        if (s_hist) for (int i = tid; i < 8 * 256; i += NT) s_hist[i] = 0;
        const int gx0 = (tx0 - L.a) & ~3;            // floor to a multiple of 4 (also for negatives)
        const int delta = (tx0 - L.a) - gx0;
        {
            const int gy = ty0 + ty;
            uint32_t* r = bufA + ty * nw;
            for (int wi = tx; wi < nw; wi += 16) r[wi] = load_word(J, gy, gx0 + 4 * wi, neutral, src_al);
        }
        __syncthreads();
        uint32_t* cur = bufA;
        uint32_t* nxt = bufB;
        for (int p = 1; p < P; p *= 2) {
            const uint32_t* r = cur + ty * nw;
            uint32_t* w = nxt + ty * nw;
            if (p < 4) {
                for (int wi = tx; wi < nw; wi += 16) w[wi] = op4(r[wi], __funnelshift_r(r[wi], r[min(wi + 1, nw - 1)], 8 * p), L.is_dilate);
            } else {
                const int off = p >> 2;
                for (int wi = tx; wi < nw; wi += 16) w[wi] = op4(r[wi], r[min(wi + off, nw - 1)], L.is_dilate);
            }
            __syncthreads();
            uint32_t* t = cur; cur = nxt; nxt = t;
        }
        {
            const int out_words = L.len >> 2;
            const uint32_t* r = cur + ty * nw;
            for (int wo = tx; wo < out_words; wo += 16) {
                uint32_t res = read_bytes(r, delta + 4 * wo, nw - 1);
                if (k > P) res = op4(res, read_bytes(r, delta + 4 * wo + (k - P), nw - 1), L.is_dilate);
                store_word(J, ty0 + ty, tx0 + 4 * wo, res, s_hist, dst_al);
            }
        }
        if (s_hist) {
            __syncthreads();
            uint32_t s = 0;
            for (int w = 0; w < 8; w++) s += s_hist[w * 256 + tid];
            if (s) atomicAdd(&J.hist[tid], s);
        }
    } else {
        const int cw = L.cnt;                         // 16 word columns per tile
        const int tx0 = blockIdx.x * cw * 4, ty0 = blockIdx.y * L.len;
        if (tx0 >= J.w || ty0 >= J.h) return;
        const int nr = L.nw, total = nr * cw;
        uint32_t* bufA = smem_u32;
        uint32_t* bufB = smem_u32 + total;
        uint32_t* s_hist = J.hist ? smem_u32 + 2 * total : nullptr;
        if (s_hist) for (int i = tid; i < 8 * 256; i += NT) s_hist[i] = 0;
        for (int row = ty; row < nr; row += 16) bufA[row * cw + tx] = load_word(J, ty0 - L.a + row, tx0 + 4 * tx, neutral, src_al);
        __syncthreads();
        uint32_t* cur = bufA;
        uint32_t* nxt = bufB;
        for (int p = 1; p < P; p *= 2) {
            for (int row = ty; row < nr; row += 16)
                nxt[row * cw + tx] = op4(cur[row * cw + tx], cur[min(row + p, nr - 1) * cw + tx], L.is_dilate);
            __syncthreads();
            uint32_t* t = cur; cur = nxt; nxt = t;
        }
        for (int row = ty; row < L.len; row += 16) {
            uint32_t res = cur[row * cw + tx];
            if (k > P) res = op4(res, cur[min(row + (k - P), nr - 1) * cw + tx], L.is_dilate);
            store_word(J, ty0 + row, tx0 + 4 * tx, res, s_hist, dst_al);
        }
        if (s_hist) {
            __syncthreads();
            uint32_t s = 0;
            for (int w = 0; w < 8; w++) s += s_hist[w * 256 + tid];
            if (s) atomicAdd(&J.hist[tid], s);
        }
    }
}

int launch_axis(docscan_ctx* ctx, int axis, int is_dilate, int k, int a, const MorphJob* jobs_dev, int n, int max_w,
                int max_h, bool hist, double alg_bytes) {
    MorphLaunch L{};
    L.k = k; L.a = a; L.is_dilate = is_dilate;
    dim3 grid;
    size_t words;
    if (axis == 0) {
        L.len = 512; L.cnt = 16;
        L.nw = ((3 + L.len + k - 1 + 3) >> 2) + 1;
        L.nw |= 1;
        words = (size_t)2 * L.nw * L.cnt;
        grid = dim3((max_w + L.len - 1) / L.len, (max_h + L.cnt - 1) / L.cnt, n);
    } else {
        L.cnt = 16;
        L.len = 128;
        while (L.len < 2 * k) L.len *= 2;
        L.nw = L.len + k - 1;
        words = (size_t)2 * L.nw * L.cnt;
        grid = dim3((max_w + 4 * L.cnt - 1) / (4 * L.cnt), (max_h + L.len - 1) / L.len, n);
    }
    const size_t smem = (words + (hist ? 8 * 256 : 0)) * sizeof(uint32_t);
    ProfScope prof(ctx, std::string(axis == 0 ? "morph_h_k" : "morph_v_k") + std::to_string(k), alg_bytes);
    if (smem > 200 * 1024) return ds_fail(ctx, DOCSCAN_ERR_UNSUPPORTED, "structuring element %d too large", k);
    if (axis == 0) {
        if (smem > 48 * 1024) DS_CUDA(ctx, cudaFuncSetAttribute(morph_1d_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        morph_1d_kernel<0><<<grid, NT, smem, ctx->stream>>>(jobs_dev, L);
    } else {
        if (smem > 48 * 1024) DS_CUDA(ctx, cudaFuncSetAttribute(morph_1d_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        morph_1d_kernel<1><<<grid, NT, smem, ctx->stream>>>(jobs_dev, L);
    }
    DS_CHECK_LAUNCH(ctx);
    return DOCSCAN_OK;
}

}  // namespace

// One 2-D erode/dilate pass for every job: H pass src -> tmp (arena), V pass tmp -> dst (+ epilogue).
// (kw, kh, ax, ay) already include the `iterations` enlargement.
int k_morph_jobs(docscan_ctx* ctx, int is_dilate, int kw, int kh, int ax, int ay, const MorphJob* jobs_host, int n,
                 int max_w, int max_h) {
    if (kw < 1 || kh < 1) return ds_fail(ctx, DOCSCAN_ERR_BAD_ARG, "bad structuring element %dx%d", kw, kh);
    std::vector<MorphJob> hjobs(jobs_host, jobs_host + n), vjobs(jobs_host, jobs_host + n);
    bool hist = false;
    for (int i = 0; i < n; i++) {
        DImg tmp;
        DS_TRY(ds_arena_image(ctx, jobs_host[i].w, jobs_host[i].h, 1, &tmp));
        hjobs[i].dst = tmp.p; hjobs[i].dst_pitch = tmp.pitch; hjobs[i].ref = nullptr; hjobs[i].hist = nullptr;
        vjobs[i].src = tmp.p; vjobs[i].src_pitch = tmp.pitch;
        hist = hist || jobs_host[i].hist;
    }
    void *dh = nullptr, *dv = nullptr;
    DS_TRY(ds_upload(ctx, hjobs.data(), sizeof(MorphJob) * n, &dh));
    DS_TRY(ds_upload(ctx, vjobs.data(), sizeof(MorphJob) * n, &dv));
    double px = 0, refpx = 0;
    for (int i = 0; i < n; i++) {
        px += (double)jobs_host[i].w * jobs_host[i].h;
        if (jobs_host[i].ref) refpx += (double)jobs_host[i].w * jobs_host[i].h;
    }
    DS_TRY(launch_axis(ctx, 0, is_dilate, kw, ax, (const MorphJob*)dh, n, max_w, max_h, false, 2.0 * px));
    DS_TRY(launch_axis(ctx, 1, is_dilate, kh, ay, (const MorphJob*)dv, n, max_w, max_h, hist, 2.0 * px + refpx));
    return DOCSCAN_OK;
}
