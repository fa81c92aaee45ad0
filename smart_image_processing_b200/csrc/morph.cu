// cv2.erode / cv2.dilate with a MORPH_RECT kw x kh element (DocScanner.py:199-200,211-212,251-254;
// morph_seq grayscale_erosion / binary_closing).  A rectangle is separable, so one 2-D pass is a
// horizontal 1-D min/max followed by a vertical one.  Each 1-D pass is O(log k) per pixel: the tile is
// held in shared memory as packed bytes (4 pixels per word), window minima of length 1,2,4,..,P are
// built by doubling (A_2p[i] = op(A_p[i], A_p[i+p])) with ping-pong buffers, and the final window of
// length k is op(A_P[i], A_P[i+k-P]).  Pixels outside the image are ignored, exactly like OpenCV's
// default morphology border: they are loaded as the neutral element (255 for erode, 0 for dilate).
// The last pass can fuse the black-hat subtraction (close(src) - src) and a 256-bin histogram.
#include <algorithm>

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

// ---------------------------------------------------------------------------------------------------------
// Register-resident streaming variant for small rectangles (default anchor, one iteration): one warp owns a strip and
// marches down it, reading only the source row and writing only the result row (2 B/px), no block barrier.  KW, KH
// and the operation are template parameters, so every shift, ring index and window offset is a compile-time constant.
// Pixels are widened to 16-bit lanes (2 px per register) so that min/max is the native VIMNMX.U16x2.
// A lane owns 8 output pixels of a row (four u16x2 registers); its H halo comes from the raw words of the next
// lanes (shuffles of packed words), the H window doubling runs on a register array, the V doubling rings are
// registers (the top ring moves to shared memory only when kh - P > 4).  ~10 instructions per pixel.
template <bool DIL> __device__ __forceinline__ uint32_t opx(uint32_t a, uint32_t b) { return DIL ? __vmaxu2(a, b) : __vminu2(a, b); }
__device__ __forceinline__ uint32_t odd_shift(uint32_t lo, uint32_t hi) { return __byte_perm(lo, hi, 0x5432); }   // px (2i+1, 2i+2)

__host__ __device__ constexpr int ilog2_floor(int v) { int l = 0; while ((2 << l) <= v) l++; return l; }
__host__ __device__ constexpr int pow2_ceil(int v) { int p = 1; while (p < v) p *= 2; return p; }

struct FixedLaunch { int seg_rows; };

// slow path of a 4-pixel load: any alignment, neutral outside the row (kept out of line: it is rare and would
// otherwise be replicated in every unrolled row)
__device__ __noinline__ uint32_t load_word_slow(const uint8_t* rowp, int x, int w, uint32_t neutral_byte) {
    uint32_t wv = 0;
    for (int b = 0; b < 4; b++) wv |= ((x + b >= 0 && x + b < w) ? (uint32_t)rowp[x + b] : neutral_byte) << (8 * b);
    return wv;
}

// Histogram without atomics: every lane owns a private set of 256 16-bit counters, laid out [bin][lane] so that
// the 32 lanes of a warp always touch 32 different half-words (at most 2-way bank conflicts).  A plain
// load / add / store per pixel replaces one shared-memory atomic (2 LSU cycles per lane, the bottleneck of every
// per-pixel-atomic histogram); counters are folded into the page histogram once per CTA.
__device__ __forceinline__ void lane_hist_add(uint16_t* h, int lane, uint32_t val) {
    volatile uint16_t* p = h + val * 32 + lane;
    *p = (uint16_t)(*p + 1);
}

template <int KW, int KH, bool DIL>
__global__ void __launch_bounds__(128) morph_fixed_kernel(const MorphJob* __restrict__ jobs, const FixedLaunch L) {
    constexpr int AX = KW / 2, AY = KH / 2;
    constexpr int AXW = ((AX + 3) / 4) * 4;            // the warp's load window starts AXW px left of its first output
    constexpr int OFF = AXW - AX;                      // array index of the window start of the lane's first output
    constexpr int NPX = OFF + 8 + KW - 1;              // pixels a lane needs
    constexpr int NWORDS = (NPX + 3) / 4, NR = NWORDS * 2;
    constexpr int LANE_HALO = (NWORDS - 1) / 2;        // following lanes a lane borrows raw words from
    constexpr int STRIDE = 8 * (32 - LANE_HALO);       // valid output columns per warp
    constexpr int PH = 1 << ilog2_floor(KW);
    constexpr int NLEV = ilog2_floor(KH), P = 1 << NLEV, BACK = KH - P;
    constexpr int NREG = NLEV < 3 ? NLEV : 3;          // levels 0..2 (rings of 1, 2, 4 rows) live in registers
    constexpr int TR = BACK > 0 ? pow2_ceil(BACK) : 1;
    constexpr bool TOP_SMEM = BACK > 4;
    constexpr int U = 4;                               // row-loop unroll: multiple of every register ring size
    // shared-memory rings (uint4 per thread per slot): levels >= 3, then the top ring
    constexpr int SLOTS_L3 = NLEV > 3 ? 8 : 0, SLOTS_L4 = NLEV > 4 ? 16 : 0, SLOTS_TOP = TOP_SMEM ? TR : 0;
    constexpr int RING_SLOTS = SLOTS_L3 + SLOTS_L4 + SLOTS_TOP;
    constexpr uint32_t NEUTRAL = DIL ? 0u : 0x00ff00ffu, NEUTRAL_W = DIL ? 0u : 0xffffffffu;
    static_assert(NLEV <= 5, "kh up to 63");

    const MorphJob J = jobs[blockIdx.z];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int xo = (blockIdx.x * 4 + warp) * STRIDE + 8 * lane;     // lane's first output column
    const int y_begin = blockIdx.y * L.seg_rows;
    extern __shared__ __align__(16) uint32_t smem_u32[];
    uint4* s_ring = reinterpret_cast<uint4*>(smem_u32);               // [RING_SLOTS][128]
    uint16_t* s_hist = reinterpret_cast<uint16_t*>(smem_u32 + 4 * RING_SLOTS * 128) + warp * 256 * 32;   // [4][256][32]
    if (J.hist) {
        uint32_t* hz = smem_u32 + 4 * RING_SLOTS * 128;
        for (int i = threadIdx.x; i < 4 * 256 * 32 / 2; i += 128) hz[i] = 0;
        __syncthreads();
    }
    if ((blockIdx.x * 4 + warp) * STRIDE < J.w && y_begin < J.h) {
        const int y_end = min(J.h, y_begin + L.seg_rows);
        const int gx = xo - AXW;                                      // lane's first loaded column (multiple of 4)
        const bool al = ((reinterpret_cast<uintptr_t>(J.src) | (uintptr_t)J.src_pitch) & 3) == 0;
        const bool dst_al = ((reinterpret_cast<uintptr_t>(J.dst) | (uintptr_t)J.dst_pitch) & 3) == 0;
        const bool in0 = al && gx >= 0 && gx + 3 < J.w, in1 = al && gx + 4 >= 0 && gx + 7 < J.w;
        const bool store_lane = lane < 32 - LANE_HALO && xo < J.w;
        const int nvalid = min(8, J.w - xo);
        const int t_first = y_begin - AY;
        const int n_rows = (y_end - y_begin) + KH - 1;
        auto load_row = [&](int ti, uint32_t& w0, uint32_t& w1) {
            const int t = t_first + ti;
            w0 = NEUTRAL_W; w1 = NEUTRAL_W;
            if (t >= 0 && t < J.h && ti < n_rows) {
                const uint8_t* rowp = J.src + (size_t)t * J.src_pitch;
                w0 = in0 ? ds_ldg32(rowp + gx) : load_word_slow(rowp, gx, J.w, NEUTRAL_W & 255u);
                w1 = in1 ? ds_ldg32(rowp + gx + 4) : load_word_slow(rowp, gx + 4, J.w, NEUTRAL_W & 255u);
            }
        };
        uint32_t ring[NREG > 0 ? NREG : 1][4][4];                     // [level][slot][register]
#pragma unroll
        for (int l = 0; l < NREG; l++)
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int c = 0; c < 4; c++) ring[l][i][c] = NEUTRAL;
        uint32_t top[TOP_SMEM ? 1 : TR][4];
#pragma unroll
        for (int i = 0; i < (TOP_SMEM ? 1 : TR); i++)
#pragma unroll
            for (int c = 0; c < 4; c++) top[i][c] = NEUTRAL;
        for (int i = 0; i < RING_SLOTS; i++) s_ring[i * 128 + threadIdx.x] = make_uint4(NEUTRAL, NEUTRAL, NEUTRAL, NEUTRAL);
        uint32_t pw0[U], pw1[U];                                      // rows in flight (loaded U rows ahead)
#pragma unroll
        for (int u = 0; u < U; u++) load_row(u, pw0[u], pw1[u]);
        for (int t0 = 0; t0 < n_rows; t0 += U) {
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int ti = t0 + u;
                // ---- raw words: own two + halo from the following lanes
                uint32_t raw[NWORDS];
                raw[0] = pw0[u]; raw[1] = pw1[u];
#pragma unroll
                for (int j = 2; j < NWORDS; j++) raw[j] = __shfl_down_sync(0xffffffffu, (j & 1) ? pw1[u] : pw0[u], j >> 1);
                load_row(ti + U, pw0[u], pw1[u]);                     // refill the slot just consumed
                uint32_t r[NR];
#pragma unroll
                for (int j = 0; j < NWORDS; j++) { r[2 * j] = __byte_perm(raw[j], 0, 0x4140); r[2 * j + 1] = __byte_perm(raw[j], 0, 0x4342); }
                // ---- H pass: forward windows A_p[j] = op(px j..j+p-1), p = 1, 2, 4, .. PH, then KW
#pragma unroll
                for (int p = 1; p < PH; p *= 2) {
#pragma unroll
                    for (int i = 0; i < NR; i++) {
                        const int i2 = i + (p >> 1) < NR - 1 ? i + (p >> 1) : NR - 1;
                        const uint32_t other = p == 1 ? odd_shift(r[i], r[i + 1 < NR ? i + 1 : NR - 1]) : r[i2];
                        r[i] = opx<DIL>(r[i], other);
                    }
                }
                if (KW > PH) {
                    constexpr int O = KW - PH;
#pragma unroll
                    for (int i = 0; i < NR; i++) {
                        const int lo = i + O / 2 < NR - 1 ? i + O / 2 : NR - 1, hi = lo + 1 < NR ? lo + 1 : NR - 1;
                        const uint32_t other = (O & 1) ? odd_shift(r[lo], r[hi]) : r[lo];
                        r[i] = opx<DIL>(r[i], other);
                    }
                }
                uint32_t v[4];
#pragma unroll
                for (int c = 0; c < 4; c++) v[c] = (OFF & 1) ? odd_shift(r[OFF / 2 + c], r[OFF / 2 + c + 1]) : r[OFF / 2 + c];
                // ---- V pass: backward windows by doubling down the rows
#pragma unroll
                for (int l = 0; l < NREG; l++) {
                    const int slot = u & ((1 << l) - 1);                // static: u is an unrolled constant
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        const uint32_t old = ring[l][slot][c];
                        ring[l][slot][c] = v[c];
                        v[c] = opx<DIL>(v[c], old);
                    }
                }
                if (NLEV > 3) {
                    uint4* q = s_ring + (ti & 7) * 128 + threadIdx.x;
                    const uint4 old = *q;
                    *q = make_uint4(v[0], v[1], v[2], v[3]);
                    v[0] = opx<DIL>(v[0], old.x); v[1] = opx<DIL>(v[1], old.y); v[2] = opx<DIL>(v[2], old.z); v[3] = opx<DIL>(v[3], old.w);
                }
                if (NLEV > 4) {
                    uint4* q = s_ring + (SLOTS_L3 + (ti & 15)) * 128 + threadIdx.x;
                    const uint4 old = *q;
                    *q = make_uint4(v[0], v[1], v[2], v[3]);
                    v[0] = opx<DIL>(v[0], old.x); v[1] = opx<DIL>(v[1], old.y); v[2] = opx<DIL>(v[2], old.z); v[3] = opx<DIL>(v[3], old.w);
                }
                if (BACK > 0) {
                    if (TOP_SMEM) {
                        uint4* base = s_ring + (SLOTS_L3 + SLOTS_L4) * 128 + threadIdx.x;
                        const uint4 old = base[((ti - BACK) & (TR - 1)) * 128];
                        base[(ti & (TR - 1)) * 128] = make_uint4(v[0], v[1], v[2], v[3]);
                        v[0] = opx<DIL>(v[0], old.x); v[1] = opx<DIL>(v[1], old.y); v[2] = opx<DIL>(v[2], old.z); v[3] = opx<DIL>(v[3], old.w);
                    } else {
                        const int rd = (u - BACK) & (TR - 1), wr = u & (TR - 1);   // TR <= 4 divides U
#pragma unroll
                        for (int c = 0; c < 4; c++) {
                            const uint32_t old = top[TOP_SMEM ? 0 : rd][c];
                            top[TOP_SMEM ? 0 : wr][c] = v[c];
                            v[c] = opx<DIL>(v[c], old);
                        }
                    }
                }
                // ---- output row
                const int y = y_begin + ti - (KH - 1);
                if (ti >= KH - 1 && y < y_end && store_lane) {
                    uint32_t o0 = __byte_perm(v[0], v[1], 0x6420), o1 = __byte_perm(v[2], v[3], 0x6420);
                    if (J.ref) {                                      // black-hat: sat(close(src) - src)
                        const uint8_t* rp = J.ref + (size_t)y * J.ref_pitch + xo;
                        uint32_t r0, r1;
                        if (nvalid == 8 && (reinterpret_cast<uintptr_t>(rp) & 3) == 0) { r0 = ds_ldg32(rp); r1 = ds_ldg32(rp + 4); }
                        else { r0 = load_word_slow(rp - xo, xo, J.w, 0); r1 = load_word_slow(rp - xo, xo + 4, J.w, 0); }
                        o0 = __vsubus4(o0, r0); o1 = __vsubus4(o1, r1);
                    }
                    if (J.hist) {
#pragma unroll
                        for (int bb = 0; bb < 8; bb++)
                            if (bb < nvalid) lane_hist_add(s_hist, lane, ((bb < 4 ? o0 : o1) >> (8 * (bb & 3))) & 255u);
                    }
                    uint8_t* dp = J.dst + (size_t)y * J.dst_pitch + xo;
                    if (dst_al && nvalid == 8) { reinterpret_cast<uint32_t*>(dp)[0] = o0; reinterpret_cast<uint32_t*>(dp)[1] = o1; }
                    else for (int bb = 0; bb < nvalid; bb++) dp[bb] = (uint8_t)((bb < 4 ? o0 : o1) >> (8 * (bb & 3)));
                }
            }
        }
    }
    if (J.hist) {
        __syncwarp();
        // fold the 32 lane-private counters of each bin; lane j owns bins j, j+32, ...
        for (int bin = lane; bin < 256; bin += 32) {
            uint32_t sum = 0;
            for (int k = 0; k < 32; k++) sum += s_hist[bin * 32 + ((k + lane) & 31)];
            if (sum) atomicAdd(&J.hist[bin], sum);
        }
    }
}

template <int KW, int KH, bool DIL>
int launch_fixed_t(docscan_ctx* ctx, const MorphJob* jobs_host, int n, int max_w, int max_h) {
    constexpr int AX = KW / 2, AXW = ((AX + 3) / 4) * 4, OFF = AXW - AX, NPX = OFF + 8 + KW - 1, NWORDS = (NPX + 3) / 4;
    constexpr int STRIDE = 8 * (32 - (NWORDS - 1) / 2);
    constexpr int NLEV = ilog2_floor(KH), BACK = KH - (1 << NLEV), TR = BACK > 0 ? pow2_ceil(BACK) : 1;
    constexpr int RING_SLOTS = (NLEV > 3 ? 8 : 0) + (NLEV > 4 ? 16 : 0) + (BACK > 4 ? TR : 0);
    const int strips = (max_w + STRIDE - 1) / STRIDE, ctas_x = (strips + 3) / 4;
    int segs = (6 * ctx->sm_count + ctas_x * n - 1) / (ctas_x * n);
    if (segs < 1) segs = 1;
    int seg = std::max((max_h + segs - 1) / segs, std::max(32, 4 * KH));
    FixedLaunch L{seg};
    bool hist = false;
    double px = 0, refpx = 0;
    for (int i = 0; i < n; i++) {
        hist = hist || jobs_host[i].hist;
        px += (double)jobs_host[i].w * jobs_host[i].h;
        if (jobs_host[i].ref) refpx += (double)jobs_host[i].w * jobs_host[i].h;
    }
    const size_t smem = (size_t)RING_SLOTS * 128 * 16 + (hist ? (size_t)4 * 256 * 32 * 2 : 0);
    void* dev = nullptr;
    DS_TRY(ds_upload(ctx, jobs_host, sizeof(MorphJob) * n, &dev));
    dim3 grid(ctas_x, (max_h + seg - 1) / seg, n);
    ProfScope prof(ctx, "morph_fixed_" + std::to_string(KW) + "x" + std::to_string(KH), 2.0 * px + refpx);
    if (smem > 48 * 1024)
        DS_CUDA(ctx, cudaFuncSetAttribute(morph_fixed_kernel<KW, KH, DIL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    morph_fixed_kernel<KW, KH, DIL><<<grid, 128, smem, ctx->stream>>>((const MorphJob*)dev, L);
    DS_CHECK_LAUNCH(ctx);
    return DOCSCAN_OK;
}

// Rectangles with a static instantiation.  Measured on B200 (profiles/README.md): the register-resident kernel wins
// for small rectangles (3x3 close: 1.64 ms vs 2.73 ms per 256 pages); for 9x19 its one-warp-per-strip march has too
// few warps in flight and the shared-memory doubling kernels are faster, so larger shapes stay on those.
#define DS_FIXED_SHAPES(X) X(2, 2) X(3, 3) X(5, 5) X(7, 7)

bool launch_fixed(docscan_ctx* ctx, int is_dilate, int kw, int kh, int ax, int ay, const MorphJob* jobs_host, int n, int max_w,
                  int max_h, int* rc) {
    if (ax != kw / 2 || ay != kh / 2) return false;
#define DS_FIXED_CASE(W, H)                                                                                       \
    if (kw == W && kh == H) {                                                                                     \
        *rc = is_dilate ? launch_fixed_t<W, H, true>(ctx, jobs_host, n, max_w, max_h)                             \
                        : launch_fixed_t<W, H, false>(ctx, jobs_host, n, max_w, max_h);                           \
        return true;                                                                                              \
    }
    DS_FIXED_SHAPES(DS_FIXED_CASE)
#undef DS_FIXED_CASE
    return false;
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
    {
        int rc = DOCSCAN_OK;
        if (launch_fixed(ctx, is_dilate, kw, kh, ax, ay, jobs_host, n, max_w, max_h, &rc)) return rc;
    }
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
