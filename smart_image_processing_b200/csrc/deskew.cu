// The skew estimate of deskew() (DocScanner.py:218-231) on the device, so that the deskew rotation no longer needs a
// host round trip between the blend and the rotate kernels:
//
//   edges = cv2.Canny(gray, low, high)                      canny_nms_kernel + union-find hysteresis (ccl_* kernels)
//   lines = cv2.HoughLines(edges, 1, pi/180, 150)           hough_vote_kernel + hough_peaks_kernel
//   angle = median of the folded line angles, 0 beyond max_rotate; getRotationMatrix2D   skew_finish_kernel
//
// Everything is exact: Canny is integer arithmetic (Sobel 3x3 with replicated borders, |dx|+|dy|, 15-bit fixed-point
// direction test); its hysteresis result is "the candidates 8-connected to a candidate above `high`", which does not
// depend on traversal order, so it is computed as connected components with an atomic union-find instead of OpenCV's
// stack flood fill.  The Hough accumulator is integer votes at r = cvRound(x*cos + y*sin) with OpenCV's fp32 tables
// (built on the host exactly like cv::createTrigTable).  The median only depends on how many lines each of the 180
// angles has; the reference evaluates it in numpy float32 (theta is np.float32), restated in hostmath.cpp, and the
// cos / sin of the resulting angle come from a host-built table (libm), so the rotation matrix is bit-identical to
// cv2.getRotationMatrix2D's.
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace {

constexpr int NANG = 180;

struct SkewJob {
    const uint8_t* src; int src_pitch, w, h;
    uint8_t* map;          // w*h dense: 0 weak candidate, 1 none, 2 strong
    int* label;            // w*h dense union-find parents (-1 = not a candidate)
    uint8_t* rootflag;     // w*h dense: component root has a strong pixel
    uint8_t* edges; int edges_pitch;      // may be null
    uint32_t* list;        // edge coordinates x | y << 16 (may be null)
    uint32_t* count;       // number of list entries
    int* accum;            // (NANG + 2) x (numrho + 2)
    int numrho;
    uint32_t* per_angle;   // NANG line counts
    uint2* cand; uint32_t* n_cand; int max_cand;     // optional (accumulator index, votes) of every line
};

// ---- hysteresis as connected components (atomic union-find) --------------------------------------------------------
__device__ __forceinline__ int uf_find(const int* L, int x) {
    int p = L[x];
    while (p != x) { x = p; p = L[x]; }
    return x;
}
// find with path halving for the merge phase: every visited node is re-pointed at its grandparent.  Racing writers only
// ever store an ancestor of the node, so the forest stays valid.
__device__ __forceinline__ int uf_find_halve(int* L, int x) {
    int p = L[x];
    while (p != x) {
        const int gp = L[p];
        if (gp != p) L[x] = gp;
        x = p; p = gp;
    }
    return x;
}
__device__ __forceinline__ void uf_union(int* L, int a, int b) {
    while (true) {
        a = uf_find_halve(L, a); b = uf_find_halve(L, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }      // the larger root is linked under the smaller one
        const int old = atomicMin(&L[a], b);
        if (old == a) return;
        a = old;
    }
}

// ---- Canny: gradient, non-maximum suppression, double threshold -------------------------------------------------
constexpr int CT_W = 64, CT_H = 16;

__global__ void __launch_bounds__(256) canny_nms_kernel(const SkewJob* __restrict__ jobs, int low, int high) {
    const SkewJob J = jobs[blockIdx.z];
    const int x0 = blockIdx.x * CT_W, y0 = blockIdx.y * CT_H;
    if (x0 >= J.w || y0 >= J.h) return;
    __shared__ uint8_t s_src[CT_H + 4][CT_W + 4];
    __shared__ short s_mag[CT_H + 2][CT_W + 2];
    const int tid = threadIdx.x;
    for (int i = tid; i < (CT_H + 4) * (CT_W + 4); i += 256) {
        const int ly = i / (CT_W + 4), lx = i - ly * (CT_W + 4);
        s_src[ly][lx] = J.src[(size_t)ds_clamp(y0 + ly - 2, 0, J.h - 1) * J.src_pitch + ds_clamp(x0 + lx - 2, 0, J.w - 1)];
    }
    __syncthreads();
    auto sobel = [&](int ly, int lx, int& gx, int& gy) {        // (ly, lx) in s_src coordinates of the centre pixel
        const int a = s_src[ly - 1][lx - 1], b = s_src[ly - 1][lx], c = s_src[ly - 1][lx + 1];
        const int d = s_src[ly][lx - 1], f = s_src[ly][lx + 1];
        const int g = s_src[ly + 1][lx - 1], hh = s_src[ly + 1][lx], k = s_src[ly + 1][lx + 1];
        gx = (c - a) + 2 * (f - d) + (k - g);
        gy = (g - a) + 2 * (hh - b) + (k - c);
    };
    for (int i = tid; i < (CT_H + 2) * (CT_W + 2); i += 256) {
        const int ly = i / (CT_W + 2), lx = i - ly * (CT_W + 2);
        const int gy_ = y0 + ly - 1, gx_ = x0 + lx - 1;
        int m = 0;
        if (gy_ >= 0 && gy_ < J.h && gx_ >= 0 && gx_ < J.w) {      // the magnitude plane has a zero border
            int gx, gy;
            sobel(ly + 1, lx + 1, gx, gy);
            m = abs(gx) + abs(gy);
        }
        s_mag[ly][lx] = (short)m;
    }
    __syncthreads();
    __shared__ int s_lab[CT_H * CT_W];               // tile-local union-find parents (-1 = no candidate)
    for (int i = tid; i < CT_H * CT_W; i += 256) {
        const int ly = i / CT_W, lx = i - ly * CT_W;
        const int y = y0 + ly, x = x0 + lx;
        s_lab[i] = -1;
        if (y >= J.h || x >= J.w) continue;
        int xs, ys;
        sobel(ly + 2, lx + 2, xs, ys);
        const int m = s_mag[ly + 1][lx + 1];
        bool cand = false;
        if (m > low) {
            const int ax = abs(xs);
            const long long ay = (long long)abs(ys) << 15, tg22x = (long long)ax * 13573;      // tan(22.5 deg) * 2^15
            if (ay < tg22x) cand = m > s_mag[ly + 1][lx] && m >= s_mag[ly + 1][lx + 2];
            else {
                const long long tg67x = tg22x + ((long long)ax << 16);
                if (ay > tg67x) cand = m > s_mag[ly][lx + 1] && m >= s_mag[ly + 2][lx + 1];
                else {
                    const int s = (xs ^ ys) < 0 ? -1 : 1;
                    cand = m > s_mag[ly][lx + 1 - s] && m > s_mag[ly + 2][lx + 1 + s];
                }
            }
        }
        const int p = y * J.w + x;
        J.map[p] = cand ? (m > high ? 2 : 0) : 1;
        if (cand) { s_lab[i] = i; J.rootflag[p] = 0; }       // only candidates are ever looked up
    }
    // Connected components inside the tile, in shared memory (the dependent loads and atomics of a union-find cost tens
    // of cycles here instead of hundreds in L2); ccl_merge_kernel then only stitches the tile borders together.
    __syncthreads();
    for (int i = tid; i < CT_H * CT_W; i += 256) {
        if (s_lab[i] < 0) continue;
        const int ly = i / CT_W, lx = i - ly * CT_W;
        if (lx > 0 && s_lab[i - 1] >= 0) uf_union(s_lab, i, i - 1);
        if (ly > 0) {
            const int q = i - CT_W;
            if (lx > 0 && s_lab[q - 1] >= 0) uf_union(s_lab, i, q - 1);
            if (s_lab[q] >= 0) uf_union(s_lab, i, q);
            if (lx + 1 < CT_W && s_lab[q + 1] >= 0) uf_union(s_lab, i, q + 1);
        }
    }
    __syncthreads();
    for (int i = tid; i < CT_H * CT_W; i += 256) {
        if (s_lab[i] < 0) continue;
        const int root = uf_find(s_lab, i);
        const int ly = i / CT_W, lx = i - ly * CT_W, ry = root / CT_W, rx = root - ry * CT_W;
        J.label[(y0 + ly) * J.w + x0 + lx] = (y0 + ry) * J.w + x0 + rx;
    }
}

__global__ void __launch_bounds__(256) ccl_merge_kernel(const SkewJob* __restrict__ jobs) {
    const SkewJob J = jobs[blockIdx.z];
    const int x = blockIdx.x * 64 + (threadIdx.x & 63), y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= J.w || y >= J.h) return;
    // canny_nms_kernel has already joined everything inside its CT_W x CT_H tiles: only neighbour pairs that straddle a
    // tile border are left
    const bool left = (x % CT_W) == 0, right = (x % CT_W) == CT_W - 1, top = (y % CT_H) == 0;
    if (!(left || right || top)) return;
    const int p = y * J.w + x;
    if (J.map[p] == 1) return;                       // most pixels are no candidates: decide on the byte plane
    if (left && x > 0 && J.map[p - 1] != 1) uf_union(J.label, p, p - 1);
    if (y > 0) {
        const int q = p - J.w;
        if ((left || top) && x > 0 && J.map[q - 1] != 1) uf_union(J.label, p, q - 1);
        if (top && J.map[q] != 1) uf_union(J.label, p, q);
        if ((right || top) && x + 1 < J.w && J.map[q + 1] != 1) uf_union(J.label, p, q + 1);
    }
}

__global__ void __launch_bounds__(256) ccl_flag_kernel(const SkewJob* __restrict__ jobs) {
    const SkewJob J = jobs[blockIdx.z];
    const int x = blockIdx.x * 64 + (threadIdx.x & 63), y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= J.w || y >= J.h) return;
    const int p = y * J.w + x;
    const int mv = J.map[p];
    if (mv == 1) return;
    const int root = uf_find(J.label, p);
    J.label[p] = root;                               // path compression (roots keep pointing at themselves)
    if (mv == 2) J.rootflag[root] = 1;
}

__global__ void __launch_bounds__(256) ccl_emit_kernel(const SkewJob* __restrict__ jobs) {
    const SkewJob J = jobs[blockIdx.z];
    const int x = blockIdx.x * 64 + (threadIdx.x & 63), y = blockIdx.y * 4 + (threadIdx.x >> 6);
    bool edge = false;
    if (x < J.w && y < J.h) {
        const int p = y * J.w + x;
        if (J.map[p] != 1) edge = J.rootflag[uf_find(J.label, p)] != 0;
        if (J.edges) J.edges[(size_t)y * J.edges_pitch + x] = edge ? 255 : 0;
    }
    if (J.list) {
        // one global atomic per CTA: warps publish their counts, the first warp reserves the block's range
        __shared__ uint32_t s_cnt[8], s_base;
        const uint32_t ballot = __ballot_sync(0xffffffffu, edge);
        const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
        if (lane == 0) s_cnt[wrp] = __popc(ballot);
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t total = 0;
            for (int k = 0; k < 8; k++) { const uint32_t c = s_cnt[k]; s_cnt[k] = total; total += c; }
            s_base = total ? atomicAdd(J.count, total) : 0u;
        }
        __syncthreads();
        if (edge) J.list[s_base + s_cnt[wrp] + __popc(ballot & ((1u << lane) - 1u))] = (uint32_t)x | ((uint32_t)y << 16);
    }
}

// edge list of an arbitrary edge image (cv2.HoughLines treats every non-zero pixel as an edge)
__global__ void __launch_bounds__(256) edge_list_kernel(const SkewJob* __restrict__ jobs) {
    const SkewJob J = jobs[blockIdx.z];
    const int x = blockIdx.x * 64 + (threadIdx.x & 63), y = blockIdx.y * 4 + (threadIdx.x >> 6);
    const bool edge = x < J.w && y < J.h && J.src[(size_t)y * J.src_pitch + x] != 0;
    const uint32_t ballot = __ballot_sync(0xffffffffu, edge);
    if (!ballot) return;
    const int lane = threadIdx.x & 31;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(J.count, __popc(ballot));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (edge) J.list[base + __popc(ballot & ((1u << lane) - 1u))] = (uint32_t)x | ((uint32_t)y << 16);
}

// ---- standard Hough transform: one CTA per (angle, page), votes in shared memory ----------------------------------
struct TrigTable { float c[NANG], s[NANG]; };

// VOTE_NA angles share one pass over the edge list (the list is streamed from L2 by every CTA of a page, so the number
// of passes is what the kernel costs); each angle has its own accumulator row in shared memory.
template <int VOTE_NA>
__global__ void __launch_bounds__(512) hough_vote_kernel(const SkewJob* __restrict__ jobs, const __grid_constant__ TrigTable T) {
    const SkewJob J = jobs[blockIdx.z];
    const int n0 = blockIdx.x * VOTE_NA;
    extern __shared__ int s_acc[];
    const int width = J.numrho + 2;
    for (int i = threadIdx.x; i < VOTE_NA * width; i += 512) s_acc[i] = 0;
    __syncthreads();
    float tc[VOTE_NA], ts[VOTE_NA];
#pragma unroll
    for (int a = 0; a < VOTE_NA; a++) { tc[a] = T.c[n0 + a]; ts[a] = T.s[n0 + a]; }
    const int half = (J.numrho - 1) / 2;
    const uint32_t cnt = *J.count;
    for (uint32_t e = threadIdx.x; e < cnt; e += 512) {
        const uint32_t v = J.list[e];
        const float fj = (float)(v & 0xffffu), fi = (float)(v >> 16);
#pragma unroll
        for (int a = 0; a < VOTE_NA; a++) {
            const int r = __float2int_rn(__fadd_rn(__fmul_rn(fj, tc[a]), __fmul_rn(fi, ts[a]))) + half;
            atomicAdd(&s_acc[a * width + r + 1], 1);
        }
    }
    __syncthreads();
    int* rows = J.accum + (size_t)(n0 + 1) * width;
    for (int i = threadIdx.x; i < VOTE_NA * width; i += 512) rows[i] = s_acc[i];
    if (n0 == 0) for (int i = threadIdx.x; i < width; i += 512) J.accum[i] = 0;                                   // border rows
    if (n0 == NANG - VOTE_NA) for (int i = threadIdx.x; i < width; i += 512) J.accum[(size_t)(NANG + 1) * width + i] = 0;
}

__global__ void __launch_bounds__(256) hough_peaks_kernel(const SkewJob* __restrict__ jobs, int threshold) {
    const SkewJob J = jobs[blockIdx.z];
    const int r = blockIdx.x * 256 + threadIdx.x, n = blockIdx.y;
    if (r >= J.numrho) return;
    const int width = J.numrho + 2;
    const int base = (n + 1) * width + r + 1;
    const int* a = J.accum;
    const int v = a[base];
    if (v > threshold && v > a[base - 1] && v >= a[base + 1] && v > a[base - width] && v >= a[base + width]) {
        atomicAdd(&J.per_angle[n], 1u);
        if (J.cand) {
            const uint32_t k = atomicAdd(J.n_cand, 1u);
            if (k < (uint32_t)J.max_cand) J.cand[k] = make_uint2((uint32_t)base, (uint32_t)v);
        }
    }
}

// ---- median angle + rotation matrix ------------------------------------------------------------------------------------
struct SkewTables {
    int order[NANG];          // angle indices in ascending order of their folded angle
    float folded[NANG];       // folded angle (degrees, numpy float32 arithmetic) per angle index
    const double2* trig;      // [NANG][NANG] (cos, sin) of the median of (folded[order[i]], folded[order[j]]) in radians
};

struct SkewOut {
    double* angle;            // per page
    WarpAJob* rot;            // per page: the inverse matrix of the deskew rotation is written into rot->m (may be null)
    int w, h;
};

__global__ void skew_finish_kernel(const SkewJob* __restrict__ jobs, const SkewOut* __restrict__ outs, int n_pages,
                                   const __grid_constant__ SkewTables T, double max_rotate) {
    const int pg = blockIdx.x * blockDim.x + threadIdx.x;
    if (pg >= n_pages) return;
    const uint32_t* cnt = jobs[pg].per_angle;
    uint32_t total = 0;
    for (int n = 0; n < NANG; n++) total += cnt[n];
    double angle = 0.0, ca = 1.0, sa = 0.0;
    if (total) {
        // np.median: the middle element, or the float32 mean of the two middle elements
        const uint32_t k_lo = (total - 1) / 2, k_hi = total / 2;
        int i_lo = -1, i_hi = -1;
        uint32_t run = 0;
        for (int i = 0; i < NANG; i++) {
            run += cnt[T.order[i]];
            if (i_lo < 0 && run > k_lo) i_lo = i;
            if (i_hi < 0 && run > k_hi) { i_hi = i; break; }
        }
        const float a_lo = T.folded[T.order[i_lo]], a_hi = T.folded[T.order[i_hi]];
        const float med = (total & 1u) ? a_lo : __fdiv_rn(__fadd_rn(a_lo, a_hi), 2.0f);
        if (!(fabs((double)med) > max_rotate)) {
            angle = (double)med;
            const double2 t = T.trig[i_lo * NANG + i_hi];
            ca = t.x; sa = t.y;
        }
    }
    const SkewOut O = outs[pg];
    O.angle[0] = angle;
    if (O.rot) {
        // cv2.getRotationMatrix2D((w/2, h/2), angle, 1.0) and the inverse cv::warpAffine derives from it (hostmath.cpp)
        const float fx = (float)(O.w / 2.0), fy = (float)(O.h / 2.0);
        double F[6];
        F[0] = ca; F[1] = sa; F[2] = __dsub_rn(__dmul_rn(__dsub_rn(1.0, ca), (double)fx), __dmul_rn(sa, (double)fy));
        F[3] = -sa; F[4] = ca; F[5] = __dadd_rn(__dmul_rn(sa, (double)fx), __dmul_rn(__dsub_rn(1.0, ca), (double)fy));
        double det = __dsub_rn(__dmul_rn(F[0], F[4]), __dmul_rn(F[1], F[3]));
        det = det != 0 ? __ddiv_rn(1.0, det) : 0;
        double* I = O.rot->m;
        I[0] = __dmul_rn(F[4], det);
        I[1] = __dmul_rn(F[1], -det);
        I[3] = __dmul_rn(F[3], -det);
        I[4] = __dmul_rn(F[0], det);
        I[2] = __dsub_rn(__dmul_rn(-I[0], F[2]), __dmul_rn(I[1], F[5]));
        I[5] = __dsub_rn(__dmul_rn(-I[3], F[2]), __dmul_rn(I[4], F[5]));
    }
}

template <int NA>
int launch_vote_t(docscan_ctx* ctx, const SkewJob* jd, int n, size_t smem, const TrigTable& T) {
    if (smem > 48 * 1024) DS_CUDA(ctx, cudaFuncSetAttribute(hough_vote_kernel<NA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    hough_vote_kernel<NA><<<dim3(NANG / NA, 1, n), 512, smem, ctx->stream>>>(jd, T);
    DS_CHECK_LAUNCH(ctx);
    return DOCSCAN_OK;
}

// as many angles per CTA (4, 2 or 1) as the accumulator rows of the largest page leave room for in shared memory
int launch_vote(docscan_ctx* ctx, const SkewJob* jd, int n, int max_rho) {
    TrigTable T;
    hm_hough_trig_table(T.c, T.s);
    const size_t row = sizeof(int) * (size_t)(max_rho + 2), cap = 100 * 1024;      // <= 100 KB: two CTAs per SM
    if (4 * row <= cap) return launch_vote_t<4>(ctx, jd, n, 4 * row, T);
    if (2 * row <= 2 * cap) return launch_vote_t<2>(ctx, jd, n, 2 * row, T);
    if (row <= 220 * 1024) return launch_vote_t<1>(ctx, jd, n, row, T);
    return ds_fail(ctx, DOCSCAN_ERR_UNSUPPORTED, "Hough transform: image too large for an accumulator row in shared memory");
}

int get_skew_tables(docscan_ctx* ctx, SkewTables* T) {
    float folded[NANG];
    hm_folded_angles(folded);
    int order[NANG];
    for (int i = 0; i < NANG; i++) order[i] = i;
    std::stable_sort(order, order + NANG, [&](int a, int b) { return folded[a] < folded[b]; });
    const uint64_t key = (uint64_t)11 << 32;
    auto it = ctx->tables.find(key);
    if (it == ctx->tables.end()) {
        std::vector<double2> trig((size_t)NANG * NANG);
        for (int i = 0; i < NANG; i++)
            for (int j = 0; j < NANG; j++) {
                const float a = folded[order[i]], b = folded[order[j]];
                const float med = i == j ? a : (a + b) / 2.0f;
                const double rad = (double)med * (3.1415926535897932384626433832795 / 180);
                trig[(size_t)i * NANG + j] = make_double2(std::cos(rad), std::sin(rad));
            }
        void* dev = nullptr;
        DS_CUDA(ctx, cudaMalloc(&dev, trig.size() * sizeof(double2)));
        DS_CUDA(ctx, cudaMemcpyAsync(dev, trig.data(), trig.size() * sizeof(double2), cudaMemcpyHostToDevice, ctx->stream));
        DS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        it = ctx->tables.emplace(key, dev).first;
    }
    for (int i = 0; i < NANG; i++) { T->order[i] = order[i]; T->folded[i] = folded[i]; }
    T->trig = (const double2*)it->second;
    return DOCSCAN_OK;
}

}  // namespace

size_t k_skew_scratch_bytes(int w, int h, bool want_list) {
    const size_t np = (size_t)w * h;
    const size_t numrho = 2 * ((size_t)w + h) + 1;
    return np + 4 * np + np + (want_list ? 4 * np : 0) + (NANG + 2) * (numrho + 2) * 4 + NANG * 4 + 4096;
}

// Canny (+ optional Hough + median angle) for a batch of gray planes.
//   edges_out  per page destination of the edge image, or null
//   angles_dev per page device double receiving the skew angle (null = stop after Canny)
//   rot_jobs   per page device WarpAJob whose inverse rotation matrix is to be filled in (entries may be null)
int k_skew_estimate(docscan_ctx* ctx, const DImg* gray, int n, double canny_low, double canny_high, int hough_threshold,
                    double max_rotate, const DImg* edges_out, double* const* angles_dev, WarpAJob* const* rot_jobs) {
    if (canny_low > canny_high) std::swap(canny_low, canny_high);
    const int low = (int)std::floor(canny_low), high = (int)std::floor(canny_high);
    const bool want_angle = angles_dev != nullptr;
    std::vector<SkewJob> jobs(n);
    std::vector<SkewOut> outs(n);
    int mw = 0, mh = 0, max_rho = 0;
    void* counters = nullptr;                                   // per page: list count + NANG per-angle counts
    DS_TRY(ds_arena_alloc(ctx, sizeof(uint32_t) * (size_t)n * (NANG + 4), &counters));
    DS_CUDA(ctx, cudaMemsetAsync(counters, 0, sizeof(uint32_t) * (size_t)n * (NANG + 4), ctx->stream));
    for (int i = 0; i < n; i++) {
        SkewJob& j = jobs[i];
        j = SkewJob{};
        const int w = gray[i].w, h = gray[i].h;
        if (w >= 65536 || h >= 65536) return ds_fail(ctx, DOCSCAN_ERR_UNSUPPORTED, "skew estimate: image larger than 65535 px");
        const size_t np = (size_t)w * h;
        j.src = gray[i].p; j.src_pitch = gray[i].pitch; j.w = w; j.h = h;
        void* p = nullptr;
        DS_TRY(ds_arena_alloc(ctx, np, &p)); j.map = (uint8_t*)p;
        DS_TRY(ds_arena_alloc(ctx, 4 * np, &p)); j.label = (int*)p;
        DS_TRY(ds_arena_alloc(ctx, np, &p)); j.rootflag = (uint8_t*)p;
        if (edges_out) { j.edges = edges_out[i].p; j.edges_pitch = edges_out[i].pitch; }
        uint32_t* c = (uint32_t*)counters + (size_t)i * (NANG + 4);
        j.count = c; j.per_angle = c + 4;
        if (want_angle) {
            DS_TRY(ds_arena_alloc(ctx, 4 * np, &p)); j.list = (uint32_t*)p;
            j.numrho = 2 * (w + h) + 1;
            DS_TRY(ds_arena_alloc(ctx, sizeof(int) * (size_t)(NANG + 2) * (j.numrho + 2), &p)); j.accum = (int*)p;
            outs[i].angle = angles_dev[i]; outs[i].rot = rot_jobs ? rot_jobs[i] : nullptr; outs[i].w = w; outs[i].h = h;
        }
        mw = std::max(mw, w); mh = std::max(mh, h); max_rho = std::max(max_rho, j.numrho);
    }
    void* dev = nullptr;
    DS_TRY(ds_upload(ctx, jobs.data(), sizeof(SkewJob) * n, &dev));
    const SkewJob* jd = (const SkewJob*)dev;
    double px = 0;
    for (int i = 0; i < n; i++) px += (double)gray[i].w * gray[i].h;
    {
        ProfScope prof(ctx, "canny_nms", 7.0 * px);
        canny_nms_kernel<<<dim3((mw + CT_W - 1) / CT_W, (mh + CT_H - 1) / CT_H, n), 256, 0, ctx->stream>>>(jd, low, high);
        DS_CHECK_LAUNCH(ctx);
    }
    const dim3 pgrid((mw + 63) / 64, (mh + 3) / 4, n);
    {
        ProfScope prof(ctx, "canny_hyst_merge", 0);
        ccl_merge_kernel<<<pgrid, 256, 0, ctx->stream>>>(jd);
        DS_CHECK_LAUNCH(ctx);
    }
    {
        ProfScope prof(ctx, "canny_hyst_flag", 0);
        ccl_flag_kernel<<<pgrid, 256, 0, ctx->stream>>>(jd);
        DS_CHECK_LAUNCH(ctx);
    }
    {
        ProfScope prof(ctx, "canny_hyst_emit", 0);
        ccl_emit_kernel<<<pgrid, 256, 0, ctx->stream>>>(jd);
        DS_CHECK_LAUNCH(ctx);
    }
    if (!want_angle) return DOCSCAN_OK;
    {
        ProfScope prof(ctx, "hough_vote", 0);
        DS_TRY(launch_vote(ctx, jd, n, max_rho));
    }
    {
        ProfScope prof(ctx, "hough_peaks", 0);
        hough_peaks_kernel<<<dim3((max_rho + 255) / 256, NANG, n), 256, 0, ctx->stream>>>(jd, hough_threshold);
        DS_CHECK_LAUNCH(ctx);
    }
    SkewTables ST;
    DS_TRY(get_skew_tables(ctx, &ST));
    DS_TRY(ds_upload(ctx, outs.data(), sizeof(SkewOut) * n, &dev));
    {
        ProfScope prof(ctx, "skew_finish", 0);
        skew_finish_kernel<<<(n + 63) / 64, 64, 0, ctx->stream>>>(jd, (const SkewOut*)dev, n, ST, max_rotate);
        DS_CHECK_LAUNCH(ctx);
    }
    return DOCSCAN_OK;
}

// cv2.HoughLines(edges, 1, pi/180, threshold) for one edge image: (accumulator index, votes) of every line and the
// accumulator width; the caller sorts them like OpenCV and converts to (rho, theta).
int k_hough_lines(docscan_ctx* ctx, const DImg& edges, int threshold, std::vector<uint2>* lines, int* numrho_out) {
    const int w = edges.w, h = edges.h;
    if (w >= 65536 || h >= 65536) return ds_fail(ctx, DOCSCAN_ERR_UNSUPPORTED, "hough_lines: image larger than 65535 px");
    SkewJob j{};
    j.src = edges.p; j.src_pitch = edges.pitch; j.w = w; j.h = h;
    j.numrho = 2 * (w + h) + 1;
    const int max_cand = NANG * j.numrho;                       // every accumulator cell could be a line
    void* p = nullptr;
    DS_TRY(ds_arena_alloc(ctx, sizeof(uint32_t) * (NANG + 8), &p));
    DS_CUDA(ctx, cudaMemsetAsync(p, 0, sizeof(uint32_t) * (NANG + 8), ctx->stream));
    j.count = (uint32_t*)p; j.n_cand = j.count + 1; j.per_angle = j.count + 4;
    DS_TRY(ds_arena_alloc(ctx, 4 * (size_t)w * h, &p)); j.list = (uint32_t*)p;
    DS_TRY(ds_arena_alloc(ctx, sizeof(int) * (size_t)(NANG + 2) * (j.numrho + 2), &p)); j.accum = (int*)p;
    DS_TRY(ds_arena_alloc(ctx, sizeof(uint2) * (size_t)max_cand, &p)); j.cand = (uint2*)p; j.max_cand = max_cand;
    void* dev = nullptr;
    DS_TRY(ds_upload(ctx, &j, sizeof(j), &dev));
    const SkewJob* jd = (const SkewJob*)dev;
    edge_list_kernel<<<dim3((w + 63) / 64, (h + 3) / 4, 1), 256, 0, ctx->stream>>>(jd);
    DS_CHECK_LAUNCH(ctx);
    DS_TRY(launch_vote(ctx, jd, 1, j.numrho));
    hough_peaks_kernel<<<dim3((j.numrho + 255) / 256, NANG, 1), 256, 0, ctx->stream>>>(jd, threshold);
    DS_CHECK_LAUNCH(ctx);
    uint32_t n_cand = 0;
    DS_CUDA(ctx, cudaMemcpyAsync(&n_cand, j.n_cand, sizeof(n_cand), cudaMemcpyDeviceToHost, ctx->stream));
    DS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    lines->resize(std::min<uint32_t>(n_cand, (uint32_t)max_cand));
    if (!lines->empty()) {
        DS_CUDA(ctx, cudaMemcpyAsync(lines->data(), j.cand, sizeof(uint2) * lines->size(), cudaMemcpyDeviceToHost, ctx->stream));
        DS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    *numrho_out = j.numrho;
    return DOCSCAN_OK;
}
