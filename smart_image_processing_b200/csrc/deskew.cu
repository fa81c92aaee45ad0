// The skew estimate of deskew() (DocScanner.py:218-231) on the device, so that the deskew rotation no longer needs a
// host round trip between the blend and the rotate kernels:
//
//   edges = cv2.Canny(gray, low, high)                      canny_bits_kernel + bit-parallel hysteresis (canny_hyst_kernel)
//   lines = cv2.HoughLines(edges, 1, pi/180, 150)           hough_vote_kernel + hough_peaks_kernel
//   angle = median of the folded line angles, 0 beyond max_rotate; getRotationMatrix2D   skew_finish_kernel
//
// Everything is exact: Canny is integer arithmetic (Sobel 3x3 with replicated borders, |dx|+|dy|, 15-bit fixed-point
// direction test); its hysteresis result is "the candidates 8-connected to a candidate above `high`", which does not
// depend on traversal order, so it is computed by bit-parallel relaxation sweeps over two bit planes instead of OpenCV's
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
    // bit planes, one 64-bit word per 64 pixels of a row (bit x & 63 of word x >> 6), `wpr` words per row:
    uint64_t* cand;        // pixels that survive non-maximum suppression with a magnitude above `low`
    uint64_t* act;         // starts as the survivors above `high`, grows into the edge map
    int wpr;
    uint8_t* map;          // union-find path: w*h dense: 0 weak candidate, 1 none, 2 strong
    int* label;            //   w*h dense union-find parents (-1 = not a candidate)
    uint8_t* rootflag;     //   w*h dense: component root has a strong pixel
    uint8_t* edges; int edges_pitch;      // may be null
    uint32_t* list;        // edge coordinates x | y << 16 (may be null)
    uint32_t* count;       // number of list entries
    // (NANG + 2) x (numrho + 2) votes.  16 bits are enough: an accumulator row must fit shared memory (w + h < 28 000), a cell
    // collects the pixels of at most two adjacent lattice lines, each no longer than min(w, h) < 14 000
    uint16_t* accum;
    int numrho;
    uint32_t* per_angle;   // NANG line counts
    uint2* lines; uint32_t* n_lines; int max_lines;     // optional (accumulator index, votes) of every line
};

// ---- Canny: gradient, non-maximum suppression, double threshold -> two bit planes ---------------------------------
// Tile = 64 x 32 pixels = one 64-bit word of 32 rows.  Source bytes are staged with replicated borders, the Sobel pair
// (dx, dy) of the tile and a one-pixel ring around it is kept in shared memory as two 16-bit lanes (zero outside the
// image: OpenCV's magnitude plane has a zero border), and the suppression test reads the two neighbours its direction
// selects.  A warp covers 32 consecutive pixels of a row and publishes its decisions with two ballots.
constexpr int CT_W = 64, CT_H = 32, CT_SW = CT_W + 8, CT_GW = CT_W + 4;
constexpr uint32_t G_BIAS = 0x04000400u;                              // both 16-bit lanes of a gradient word carry + 1024

__global__ void __launch_bounds__(256) canny_bits_kernel(const SkewJob* __restrict__ jobs, int low, int high) {
    const SkewJob J = jobs[blockIdx.z];
    const int x0 = blockIdx.x * CT_W, y0 = blockIdx.y * CT_H;
    if (blockIdx.x >= J.wpr || y0 >= J.h) return;
    __shared__ __align__(16) uint8_t s_src[CT_H + 4][CT_SW];          // origin (y0 - 2, x0 - 4)
    __shared__ __align__(16) uint32_t s_g[CT_H + 2][CT_GW];           // origin (y0 - 1, x0 - 1): (dx + 1024) | (dy + 1024) << 16
    const int tid = threadIdx.x;
    const bool al = ((reinterpret_cast<uintptr_t>(J.src) | (uintptr_t)J.src_pitch) & 3) == 0;
    for (int i = tid; i < (CT_H + 4) * (CT_SW / 4); i += 256) {
        const int ly = i / (CT_SW / 4), wi = i - ly * (CT_SW / 4);
        const uint8_t* rowp = J.src + (size_t)ds_clamp(y0 + ly - 2, 0, J.h - 1) * J.src_pitch;
        const int gx = x0 - 4 + 4 * wi;
        uint32_t word;
        if (al && gx >= 0 && gx + 3 < J.w) word = ds_ldg32(rowp + gx);
        else {
            word = 0;
            for (int b = 0; b < 4; b++) word |= (uint32_t)rowp[ds_clamp(gx + b, 0, J.w - 1)] << (8 * b);
        }
        *reinterpret_cast<uint32_t*>(&s_src[ly][4 * wi]) = word;
    }
    __syncthreads();
    // Sobel, four pixels per step on 16-bit lane pairs: with S = top + 2 mid + bottom and T = bottom - top per column,
    // dx(c) = S(c + 1) - S(c - 1) and dy(c) = T(c - 1) + 2 T(c) + T(c + 1); the columns c - 1 .. c + 4 of the four pixels are
    // bytes 2 .. 7 of two aligned words, and the differences of aligned pairs are the dx of the pixels in between
    for (int i = tid; i < (CT_H + 2) * (CT_GW / 4); i += 256) {
        const int ly = i / (CT_GW / 4), q = i - ly * (CT_GW / 4);       // pixels lx = 4 q .. 4 q + 3 of gradient row ly
        uint32_t S[3], Tb[3];                                          // column pairs (2,3) (4,5) (6,7) of the 8-byte window
        {
            const uint32_t* r0 = reinterpret_cast<const uint32_t*>(&s_src[ly][4 * q]);
            const uint32_t* r1 = reinterpret_cast<const uint32_t*>(&s_src[ly + 1][4 * q]);
            const uint32_t* r2 = reinterpret_cast<const uint32_t*>(&s_src[ly + 2][4 * q]);
            const uint32_t t0 = r0[0], t1 = r0[1], m0 = r1[0], m1 = r1[1], b0 = r2[0], b1 = r2[1];
            const uint32_t tp[3] = {__byte_perm(t0, 0, 0x4342), __byte_perm(t1, 0, 0x4140), __byte_perm(t1, 0, 0x4342)};
            const uint32_t mp[3] = {__byte_perm(m0, 0, 0x4342), __byte_perm(m1, 0, 0x4140), __byte_perm(m1, 0, 0x4342)};
            const uint32_t bp[3] = {__byte_perm(b0, 0, 0x4342), __byte_perm(b1, 0, 0x4140), __byte_perm(b1, 0, 0x4342)};
#pragma unroll
            for (int k = 0; k < 3; k++) { S[k] = tp[k] + bp[k] + 2 * mp[k]; Tb[k] = bp[k] + 0x01000100u - tp[k]; }
        }
        const uint32_t dx01 = S[1] + G_BIAS - S[0], dx23 = S[2] + G_BIAS - S[1];                     // lanes: pixels (0, 1), (2, 3)
        const uint32_t dy01 = Tb[0] + Tb[1] + 2 * __byte_perm(Tb[0], Tb[1], 0x5432);                 // 4 x 256 = the same bias
        const uint32_t dy23 = Tb[1] + Tb[2] + 2 * __byte_perm(Tb[1], Tb[2], 0x5432);
        uint32_t g[4] = {__byte_perm(dx01, dy01, 0x5410), __byte_perm(dx01, dy01, 0x7632),
                         __byte_perm(dx23, dy23, 0x5410), __byte_perm(dx23, dy23, 0x7632)};
        const int gy_ = y0 + ly - 1, gx_ = x0 + 4 * q - 1;
        const bool row_in = gy_ >= 0 && gy_ < J.h;
#pragma unroll
        for (int k = 0; k < 4; k++) if (!row_in || gx_ + k < 0 || gx_ + k >= J.w) g[k] = G_BIAS;      // zero border of the magnitude plane
        *reinterpret_cast<uint4*>(&s_g[ly][4 * q]) = make_uint4(g[0], g[1], g[2], g[3]);
    }
    __syncthreads();
    auto mag = [](uint32_t g) { return abs((int)(g & 0xffffu) - 1024) + abs((int)(g >> 16) - 1024); };
    const int lane = tid & 31, wrp = tid >> 5;
    uint32_t* cand32 = reinterpret_cast<uint32_t*>(J.cand);
    uint32_t* act32 = reinterpret_cast<uint32_t*>(J.act);
#pragma unroll
    for (int it = 0; it < 8; it++) {
        const int ly = wrp * 4 + (it >> 1), half = it & 1;
        const int y = y0 + ly;
        if (y >= J.h) break;                                           // warp-uniform
        const int lx = 32 * half + lane;
        const uint32_t* gp = &s_g[ly + 1][lx + 1];
        const uint32_t g = *gp;
        const int xs = (int)(g & 0xffffu) - 1024, ys = (int)(g >> 16) - 1024;
        const int m = abs(xs) + abs(ys);
        bool cand = false;
        if (__any_sync(0xffffffffu, m > low)) {                        // (pixels beyond the image carry m = 0)
            // one pair of neighbours, selected without branches: left / right below 22.5 degrees, up / down above 67.5, else the
            // diagonal the signs pick; the second comparison is >= for the axis directions and > for the diagonals
            const int ax = abs(xs);
            const int ay = abs(ys) << 15, tg22x = ax * 13573;          // tan(22.5 deg) * 2^15; |values| < 2^26
            const int tg67x = tg22x + (ax << 16);
            const bool horiz = ay < tg22x, vert = !horiz && ay > tg67x;
            const int s = (xs ^ ys) < 0 ? -1 : 1;
            const int o1 = horiz ? -1 : (vert ? -CT_GW : -CT_GW - s);
            const int ma = mag(gp[o1]), mb = mag(gp[-o1]);
            cand = m > low && m > ma && m + ((horiz || vert) ? 1 : 0) > mb;
        }
        const uint32_t cb = __ballot_sync(0xffffffffu, cand), sb = __ballot_sync(0xffffffffu, cand && m > high);
        if (lane == 0) {
            const size_t o = ((size_t)y * J.wpr + blockIdx.x) * 2 + half;
            cand32[o] = cb; act32[o] = sb;
        }
    }
}

// ---- hysteresis: grow the strong pixels through the candidates, 64 pixels per lane -----------------------------------
// The result ("candidates 8-connected to a strong one") does not depend on the order of traversal, so instead of
// OpenCV's stack flood fill the page is relaxed with bit-parallel sweeps.  One CTA per page, one warp per band of rows.
// A row step takes the row above (or below), widens it by one pixel either side, keeps what falls on candidates, and
// then fills along the candidate runs of the whole row in both directions at once: adding the seeds to the run mask
// lets the carry ripple to the end of every seeded run, and the carries between the lanes' 64-bit words (and between
// groups of 32 lanes) are resolved by one more integer addition on the ballots of "generates" / "propagates".
// A sweep therefore carries a label any distance horizontally and down (or up) in one pass; bands exchange their
// border rows through the plane itself and the CTA repeats until no band changed anything.
template <int G>
__device__ __forceinline__ bool hyst_row_step(const uint64_t (&c)[G], uint64_t (&a)[G], const uint64_t (&nb)[G], bool force, int lane) {
    constexpr uint32_t FULL = 0xffffffffu;
    uint64_t seeds[G];
    bool need = false;
#pragma unroll
    for (int g = 0; g < G; g++) {
        const uint64_t v = nb[g];
        uint64_t below = __shfl_up_sync(FULL, v, 1), above = __shfl_down_sync(FULL, v, 1);
        const uint64_t prev_last = __shfl_sync(FULL, g > 0 ? nb[g > 0 ? g - 1 : 0] : 0ull, 31);
        const uint64_t next_first = __shfl_sync(FULL, g + 1 < G ? nb[g + 1 < G ? g + 1 : g] : 0ull, 0);
        if (lane == 0) below = prev_last;
        if (lane == 31) above = next_first;
        const uint64_t d = v | (v << 1) | (v >> 1) | (below >> 63) | (above << 63);
        seeds[g] = a[g] | (d & c[g]);
        need = need || seeds[g] != a[g] || (force && seeds[g] != 0);
    }
    if (!__any_sync(FULL, need)) return false;
    uint64_t fill[G];
    uint32_t cin = 0;
#pragma unroll
    for (int g = 0; g < G; g++) {                                      // towards larger x
        const uint64_t t = c[g] + seeds[g];
        const uint32_t gm = __ballot_sync(FULL, t < c[g]), pm = __ballot_sync(FULL, t == ~0ull);
        const uint64_t u = (uint64_t)(gm | pm) + gm + cin;
        const uint64_t ci = (((uint32_t)u ^ pm) >> lane) & 1u;
        cin = (uint32_t)(u >> 32);
        fill[g] = ((c[g] ^ (t + ci)) & c[g]) | seeds[g];
    }
    cin = 0;
    bool changed = false;
#pragma unroll
    for (int g = G - 1; g >= 0; g--) {                                 // towards smaller x: the same on the reversed bits
        const uint64_t cr = __brevll(c[g]), t = cr + __brevll(seeds[g]);
        const uint32_t gm = __brev(__ballot_sync(FULL, t < cr)), pm = __brev(__ballot_sync(FULL, t == ~0ull));
        const uint64_t u = (uint64_t)(gm | pm) + gm + cin;
        const uint64_t ci = (((uint32_t)u ^ pm) >> (31 - lane)) & 1u;
        cin = (uint32_t)(u >> 32);
        const uint64_t na = fill[g] | __brevll((cr ^ (t + ci)) & cr);
        changed = changed || na != a[g];
        a[g] = na;
    }
    return changed;
}

template <int G>
__global__ void __launch_bounds__(1024) canny_hyst_kernel(const SkewJob* __restrict__ jobs) {
    constexpr uint32_t FULL = 0xffffffffu;
    const SkewJob J = jobs[blockIdx.x];
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int rpb = (J.h + nw - 1) / nw;
    const int yb0 = wrp * rpb, yb1 = min(J.h, yb0 + rpb);
    const int wpr = J.wpr;
    bool in[G];
#pragma unroll
    for (int g = 0; g < G; g++) in[g] = 32 * g + lane < wpr;
    auto load = [&](const uint64_t* plane, int y, uint64_t (&r)[G]) {
#pragma unroll
        for (int g = 0; g < G; g++) r[g] = in[g] ? plane[(size_t)y * wpr + 32 * g + lane] : 0ull;
    };
    auto zero = [&](uint64_t (&r)[G]) {
#pragma unroll
        for (int g = 0; g < G; g++) r[g] = 0ull;
    };
    constexpr int CH = G == 1 ? 4 : (G == 2 ? 2 : 1);                  // rows requested together
    bool first = true;
    while (true) {
        bool changed = false;
        if (yb0 < yb1) {
            uint64_t nb[G];
            // ---- downwards
            if (yb0 > 0) load(J.act, yb0 - 1, nb); else zero(nb);
            for (int y = yb0; y < yb1; y += CH) {
                uint64_t c[CH][G], a[CH][G];
#pragma unroll
                for (int k = 0; k < CH; k++)
                    if (y + k < yb1) { load(J.cand, y + k, c[k]); load(J.act, y + k, a[k]); }
#pragma unroll
                for (int k = 0; k < CH; k++) {
                    if (y + k >= yb1) break;
                    if (hyst_row_step<G>(c[k], a[k], nb, first, lane)) {
                        changed = true;
#pragma unroll
                        for (int g = 0; g < G; g++) if (in[g]) J.act[(size_t)(y + k) * wpr + 32 * g + lane] = a[k][g];
                    }
#pragma unroll
                    for (int g = 0; g < G; g++) nb[g] = a[k][g];
                }
            }
            // ---- upwards
            if (yb1 < J.h) load(J.act, yb1, nb); else zero(nb);
            for (int y = yb1 - 1; y >= yb0; y -= CH) {
                uint64_t c[CH][G], a[CH][G];
#pragma unroll
                for (int k = 0; k < CH; k++)
                    if (y - k >= yb0) { load(J.cand, y - k, c[k]); load(J.act, y - k, a[k]); }
#pragma unroll
                for (int k = 0; k < CH; k++) {
                    if (y - k < yb0) break;
                    if (hyst_row_step<G>(c[k], a[k], nb, false, lane)) {
                        changed = true;
#pragma unroll
                        for (int g = 0; g < G; g++) if (in[g]) J.act[(size_t)(y - k) * wpr + 32 * g + lane] = a[k][g];
                    }
#pragma unroll
                    for (int g = 0; g < G; g++) nb[g] = a[k][g];
                }
            }
        }
        first = false;
        // the barrier also orders the bands' border rows (global memory, same CTA) for the next round
        if (!__syncthreads_or(changed ? 1 : 0)) break;
    }
    // ---- the edge map is final: write the edge image and / or the coordinate list of this band's rows.
    if (!J.edges && !J.list) return;
    const bool e_al = J.edges && ((reinterpret_cast<uintptr_t>(J.edges) | (uintptr_t)J.edges_pitch) & 3) == 0;
    for (int ys = yb0; ys < yb1; ys += 16) {
#pragma unroll 1
        for (int g = 0; g < G; g++) {
            const int wi = 32 * g + lane, xw = 64 * wi;
            const bool inw = wi < wpr;
            uint64_t a[16];
#pragma unroll
            for (int i = 0; i < 16; i++) a[i] = (inw && ys + i < yb1) ? J.act[(size_t)(ys + i) * wpr + wi] : 0ull;
            if (J.edges && inw) {
#pragma unroll 1
                for (int i = 0; i < 16 && ys + i < yb1; i++) {
                    uint8_t* ep = J.edges + (size_t)(ys + i) * J.edges_pitch + xw;
                    uint64_t row = 0;
#pragma unroll
                    for (int k = 0; k < 16; k++) if (k == i) row = a[k];
#pragma unroll 4
                    for (int q = 0; q < 16; q++) {
                        const int x = xw + 4 * q;
                        if (x >= J.w) break;
                        const uint32_t nib = (uint32_t)(row >> (4 * q)) & 15u;
                        const uint32_t word = ((nib * 0x00204081u) & 0x01010101u) * 255u;      // bit b -> byte b
                        if (e_al && x + 3 < J.w) *reinterpret_cast<uint32_t*>(ep + 4 * q) = word;
                        else for (int b = 0; b < 4 && x + b < J.w; b++) ep[4 * q + b] = (uint8_t)(word >> (8 * b));
                    }
                }
            }
            if (J.list) {
                int n = 0;
#pragma unroll
                for (int i = 0; i < 16; i++) n += __popcll(a[i]);
                const int total = __reduce_add_sync(FULL, n);
                if (total == 0) continue;                              // warp-uniform
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(J.count, (uint32_t)total);
                base = __shfl_sync(FULL, base, 0);
                // entry k of the block comes from lane (k mod active lanes): 32 consecutive entries are pixels 64 columns apart,
                // whose accumulator cells differ for every angle but the vertical one
                const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    uint64_t bits = a[i];
                    while (true) {
                        const uint32_t mask = __ballot_sync(FULL, bits != 0);
                        if (!mask) break;
                        if (bits) {
                            const int b = __ffsll((long long)bits) - 1;
                            bits &= bits - 1;
                            J.list[base + __popc(mask & lt)] = (uint32_t)(xw + b) | ((uint32_t)(ys + i) << 16);
                        }
                        base += __popc(mask);
                    }
                }
            }
        }
    }
}

template <int G>
int launch_hyst_t(docscan_ctx* ctx, const SkewJob* jd, int n, int threads) {
    canny_hyst_kernel<G><<<n, threads, 0, ctx->stream>>>(jd);
    DS_CHECK_LAUNCH(ctx);
    return DOCSCAN_OK;
}

// ---- The parallel form for a few pages ---------------------------------------------------------------------------------
// The sweeps above take a number of row steps proportional to the vertical extent of the longest chain, whatever the number
// of bands: fine when hundreds of pages run side by side (one CTA each), slow for one big image (3840 x 2160: 1.3 ms).
// Single images and small batches therefore keep round 1's hysteresis: connected components by atomic union-find over the
// whole page (tile-local in shared memory, a stitch pass over tile borders, a flag pass, an emit pass), 0.19 ms for the same image.
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

// ---- Canny for a few pages: byte map + atomic union-find ------------------------------------------------------
constexpr int UT_W = 64, UT_H = 16;

__global__ void __launch_bounds__(256) canny_nms_kernel(const SkewJob* __restrict__ jobs, int low, int high) {
    const SkewJob J = jobs[blockIdx.z];
    const int x0 = blockIdx.x * UT_W, y0 = blockIdx.y * UT_H;
    if (x0 >= J.w || y0 >= J.h) return;
    __shared__ uint8_t s_src[UT_H + 4][UT_W + 4];
    __shared__ short s_mag[UT_H + 2][UT_W + 2];
    const int tid = threadIdx.x;
    for (int i = tid; i < (UT_H + 4) * (UT_W + 4); i += 256) {
        const int ly = i / (UT_W + 4), lx = i - ly * (UT_W + 4);
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
    for (int i = tid; i < (UT_H + 2) * (UT_W + 2); i += 256) {
        const int ly = i / (UT_W + 2), lx = i - ly * (UT_W + 2);
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
    __shared__ int s_lab[UT_H * UT_W];               // tile-local union-find parents (-1 = no candidate)
    for (int i = tid; i < UT_H * UT_W; i += 256) {
        const int ly = i / UT_W, lx = i - ly * UT_W;
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
    for (int i = tid; i < UT_H * UT_W; i += 256) {
        if (s_lab[i] < 0) continue;
        const int ly = i / UT_W, lx = i - ly * UT_W;
        if (lx > 0 && s_lab[i - 1] >= 0) uf_union(s_lab, i, i - 1);
        if (ly > 0) {
            const int q = i - UT_W;
            if (lx > 0 && s_lab[q - 1] >= 0) uf_union(s_lab, i, q - 1);
            if (s_lab[q] >= 0) uf_union(s_lab, i, q);
            if (lx + 1 < UT_W && s_lab[q + 1] >= 0) uf_union(s_lab, i, q + 1);
        }
    }
    __syncthreads();
    for (int i = tid; i < UT_H * UT_W; i += 256) {
        if (s_lab[i] < 0) continue;
        const int root = uf_find(s_lab, i);
        const int ly = i / UT_W, lx = i - ly * UT_W, ry = root / UT_W, rx = root - ry * UT_W;
        J.label[(y0 + ly) * J.w + x0 + lx] = (y0 + ry) * J.w + x0 + rx;
    }
}

__global__ void __launch_bounds__(256) ccl_merge_kernel(const SkewJob* __restrict__ jobs) {
    const SkewJob J = jobs[blockIdx.z];
    const int x = blockIdx.x * 64 + (threadIdx.x & 63), y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= J.w || y >= J.h) return;
    // canny_nms_kernel has already joined everything inside its UT_W x UT_H tiles: only neighbour pairs that straddle a
    // tile border are left
    const bool left = (x % UT_W) == 0, right = (x % UT_W) == UT_W - 1, top = (y % UT_H) == 0;
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
// A shared-memory increment costs two clocks per warp when its lanes hit 32 different banks, and that many times more when
// lanes share a bank or (short of all 32) an address (tests/tools/atoms_probe.cu: 16 increments per clock and SM at best,
// 10.6 for random cells, 7 when the lanes crowd into a window of 23 cells).  Two things keep the votes of a warp apart: the
// list interleaves pixels 64 columns apart (canny_hyst_kernel), and the cells of a row are stored XOR-swizzled
// (cell j at j ^ ((j >> 5) & 31)), so that the power-of-two strides such pixels produce at cos = 1/2, 1/4, .. spread
// over the banks as well.
template <int VOTE_NA>
__global__ void __launch_bounds__(512) hough_vote_kernel(const SkewJob* __restrict__ jobs, const __grid_constant__ TrigTable T) {
    const SkewJob J = jobs[blockIdx.z];
    const int n0 = blockIdx.x * VOTE_NA;
    extern __shared__ int s_acc[];
    const int width = J.numrho + 2, pitch = (width + 31) & ~31;
    for (int i = threadIdx.x; i < VOTE_NA * pitch; i += 512) s_acc[i] = 0;
    __syncthreads();
    float tc[VOTE_NA], ts[VOTE_NA];
#pragma unroll
    for (int a = 0; a < VOTE_NA; a++) { tc[a] = T.c[n0 + a]; ts[a] = T.s[n0 + a]; }
    const int off = (J.numrho - 1) / 2 + 1 - 0x4B400000;
    const uint32_t cnt = *J.count;
    auto cell = [&](int a, float rounded) {
        const int j = __float_as_int(rounded) + off;
        atomicAdd(&s_acc[a * pitch + (j ^ ((j >> 5) & 31))], 1);
    };
    auto vote = [&](uint32_t v) {
        // cvRound(x * cos + y * sin) without the conversion unit: |value| < 2^22, so adding 1.5 * 2^23 leaves the rounded integer
        // in the mantissa.  Two angles per packed multiply (every lane is an IEEE operation of its own); the sum of the two products is
        // taken with scalar adds, which ptxas never contracts into an fma (it does contract the packed pair: common.cuh).
        const float fj = (float)(v & 0xffffu), fi = (float)(v >> 16);
        const float2 fj2 = make_float2(fj, fj), fi2 = make_float2(fi, fi), magic = make_float2(12582912.0f, 12582912.0f);
#pragma unroll
        for (int a = 0; a + 1 < VOTE_NA; a += 2) {
            const float2 r = __fadd2_rn(ds_add2_unfused(ds_mul2_rn(fj2, make_float2(tc[a], tc[a + 1])), ds_mul2_rn(fi2, make_float2(ts[a], ts[a + 1]))), magic);
            cell(a, r.x); cell(a + 1, r.y);
        }
        if (VOTE_NA & 1) cell(VOTE_NA - 1, __fadd_rn(__fadd_rn(__fmul_rn(fj, tc[VOTE_NA - 1]), __fmul_rn(fi, ts[VOTE_NA - 1])), 12582912.0f));
    };
    uint32_t e = threadIdx.x;
    for (; e + 3 * 512 < cnt; e += 4 * 512) {
        const uint32_t v0 = J.list[e], v1 = J.list[e + 512], v2 = J.list[e + 1024], v3 = J.list[e + 1536];
        vote(v0); vote(v1); vote(v2); vote(v3);
    }
    for (; e < cnt; e += 512) vote(J.list[e]);
    __syncthreads();
    uint16_t* rows = J.accum + (size_t)(n0 + 1) * width;
    for (int i = threadIdx.x; i < VOTE_NA * width; i += 512) {
        const int a = i / width, j = i - a * width;
        rows[i] = (uint16_t)s_acc[a * pitch + (j ^ ((j >> 5) & 31))];
    }
    if (n0 == 0) for (int i = threadIdx.x; i < width; i += 512) J.accum[i] = 0;                                   // border rows
    if (n0 == NANG - VOTE_NA) for (int i = threadIdx.x; i < width; i += 512) J.accum[(size_t)(NANG + 1) * width + i] = 0;
}

__global__ void __launch_bounds__(512) hough_peaks_kernel(const SkewJob* __restrict__ jobs, int threshold) {
    const SkewJob J = jobs[blockIdx.z];
    const int n = blockIdx.x;                                   // one CTA per (angle, page): the row is streamed once
    const int width = J.numrho + 2;
    const uint16_t* a = J.accum;
    uint32_t found = 0;
    const int row = (n + 1) * width + 1;
    for (int r0 = threadIdx.x; r0 < J.numrho; r0 += 4 * 512) {
        int v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) v[u] = r0 + 512 * u < J.numrho ? (int)a[row + r0 + 512 * u] : 0;      // requested together
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int base = row + r0 + 512 * u;
            if (v[u] > threshold && v[u] > a[base - 1] && v[u] >= a[base + 1] && v[u] > a[base - width] && v[u] >= a[base + width]) {
                found++;
                if (J.lines) {
                    const uint32_t k = atomicAdd(J.n_lines, 1u);
                    if (k < (uint32_t)J.max_lines) J.lines[k] = make_uint2((uint32_t)base, (uint32_t)v[u]);
                }
            }
        }
    }
    if (found) atomicAdd(&J.per_angle[n], found);
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
    const size_t row = sizeof(int) * (size_t)((max_rho + 2 + 31) & ~31), cap = 100 * 1024;      // <= 100 KB: two CTAs per SM
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
    // bit planes of the batch path, or byte map + labels + root flags of the few-pages path (the larger of the two)
    const size_t planes = std::max(2 * ((size_t)(w + 63) / 64 * 8 * h + 256), 6 * np + 768);
    const size_t numrho = 2 * ((size_t)w + h) + 1;
    return planes + (want_list ? 4 * np + (NANG + 2) * (numrho + 2) * 2 + 512 : 0) + (NANG + 8) * 4 + 4096;
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
    // hysteresis: bit-parallel sweeps (one CTA per page) for batches, union-find over the whole page for a few pages
    bool sweeps = n >= 8;
    if (const char* e = getenv("DOCSCAN_CANNY_SWEEPS")) sweeps = atoi(e) != 0;
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
        j.wpr = (w + 63) / 64;
        if (sweeps) {
            DS_TRY(ds_arena_alloc(ctx, 8 * (size_t)j.wpr * h, &p)); j.cand = (uint64_t*)p;
            DS_TRY(ds_arena_alloc(ctx, 8 * (size_t)j.wpr * h, &p)); j.act = (uint64_t*)p;
        } else {
            DS_TRY(ds_arena_alloc(ctx, np, &p)); j.map = (uint8_t*)p;
            DS_TRY(ds_arena_alloc(ctx, 4 * np, &p)); j.label = (int*)p;
            DS_TRY(ds_arena_alloc(ctx, np, &p)); j.rootflag = (uint8_t*)p;
        }
        if (edges_out) { j.edges = edges_out[i].p; j.edges_pitch = edges_out[i].pitch; }
        uint32_t* c = (uint32_t*)counters + (size_t)i * (NANG + 4);
        j.count = c; j.per_angle = c + 4;
        if (want_angle) {
            DS_TRY(ds_arena_alloc(ctx, 4 * np, &p)); j.list = (uint32_t*)p;
            j.numrho = 2 * (w + h) + 1;
            DS_TRY(ds_arena_alloc(ctx, sizeof(uint16_t) * (size_t)(NANG + 2) * (j.numrho + 2), &p)); j.accum = (uint16_t*)p;
            outs[i].angle = angles_dev[i]; outs[i].rot = rot_jobs ? rot_jobs[i] : nullptr; outs[i].w = w; outs[i].h = h;
        }
        mw = std::max(mw, w); mh = std::max(mh, h); max_rho = std::max(max_rho, j.numrho);
    }
    void* dev = nullptr;
    DS_TRY(ds_upload(ctx, jobs.data(), sizeof(SkewJob) * n, &dev));
    const SkewJob* jd = (const SkewJob*)dev;
    double px = 0;
    for (int i = 0; i < n; i++) px += (double)gray[i].w * gray[i].h;
    if (sweeps) {
        {
            ProfScope prof(ctx, "canny_bits", 1.25 * px);
            canny_bits_kernel<<<dim3((mw + CT_W - 1) / CT_W, (mh + CT_H - 1) / CT_H, n), 256, 0, ctx->stream>>>(jd, low, high);
            DS_CHECK_LAUNCH(ctx);
        }
        // one CTA per page: 16 warps when the batch fills the device, 32 for fewer pages
        ProfScope prof(ctx, "canny_hyst", 0);
        const int threads = n >= ctx->sm_count ? 512 : 1024;
        const int groups = (mw + 2047) / 2048;
        if (groups <= 1) DS_TRY(launch_hyst_t<1>(ctx, jd, n, threads));
        else if (groups <= 2) DS_TRY(launch_hyst_t<2>(ctx, jd, n, threads));
        else if (groups <= 4) DS_TRY(launch_hyst_t<4>(ctx, jd, n, threads));
        else DS_TRY(launch_hyst_t<32>(ctx, jd, n, 512));
    } else {
        {
            ProfScope prof(ctx, "canny_nms", 7.0 * px);
            canny_nms_kernel<<<dim3((mw + UT_W - 1) / UT_W, (mh + UT_H - 1) / UT_H, n), 256, 0, ctx->stream>>>(jd, low, high);
            DS_CHECK_LAUNCH(ctx);
        }
        const dim3 pgrid((mw + 63) / 64, (mh + 3) / 4, n);
        ProfScope prof(ctx, "canny_hyst_uf", 0);
        ccl_merge_kernel<<<pgrid, 256, 0, ctx->stream>>>(jd);
        DS_CHECK_LAUNCH(ctx);
        ccl_flag_kernel<<<pgrid, 256, 0, ctx->stream>>>(jd);
        DS_CHECK_LAUNCH(ctx);
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
        hough_peaks_kernel<<<dim3(NANG, 1, n), 512, 0, ctx->stream>>>(jd, hough_threshold);
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
    const int max_lines = NANG * j.numrho;                       // every accumulator cell could be a line
    void* p = nullptr;
    DS_TRY(ds_arena_alloc(ctx, sizeof(uint32_t) * (NANG + 8), &p));
    DS_CUDA(ctx, cudaMemsetAsync(p, 0, sizeof(uint32_t) * (NANG + 8), ctx->stream));
    j.count = (uint32_t*)p; j.n_lines = j.count + 1; j.per_angle = j.count + 4;
    DS_TRY(ds_arena_alloc(ctx, 4 * (size_t)w * h, &p)); j.list = (uint32_t*)p;
    DS_TRY(ds_arena_alloc(ctx, sizeof(uint16_t) * (size_t)(NANG + 2) * (j.numrho + 2), &p)); j.accum = (uint16_t*)p;
    DS_TRY(ds_arena_alloc(ctx, sizeof(uint2) * (size_t)max_lines, &p)); j.lines = (uint2*)p; j.max_lines = max_lines;
    void* dev = nullptr;
    DS_TRY(ds_upload(ctx, &j, sizeof(j), &dev));
    const SkewJob* jd = (const SkewJob*)dev;
    edge_list_kernel<<<dim3((w + 63) / 64, (h + 3) / 4, 1), 256, 0, ctx->stream>>>(jd);
    DS_CHECK_LAUNCH(ctx);
    DS_TRY(launch_vote(ctx, jd, 1, j.numrho));
    hough_peaks_kernel<<<dim3(NANG, 1, 1), 512, 0, ctx->stream>>>(jd, threshold);
    DS_CHECK_LAUNCH(ctx);
    uint32_t n_cand = 0;
    DS_CUDA(ctx, cudaMemcpyAsync(&n_cand, j.n_lines, sizeof(n_cand), cudaMemcpyDeviceToHost, ctx->stream));
    DS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    lines->resize(std::min<uint32_t>(n_cand, (uint32_t)max_lines));
    if (!lines->empty()) {
        DS_CUDA(ctx, cudaMemcpyAsync(lines->data(), j.lines, sizeof(uint2) * lines->size(), cudaMemcpyDeviceToHost, ctx->stream));
        DS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    *numrho_out = j.numrho;
    return DOCSCAN_OK;
}
