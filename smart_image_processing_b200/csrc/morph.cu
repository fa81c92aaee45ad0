// cv2.erode / cv2.dilate with a MORPH_RECT kw x kh element (DocScanner.py:199-200,211-212,251-254;
// morph_seq grayscale_erosion / binary_closing).  A rectangle is separable, so one 2-D pass is a
// horizontal 1-D min/max followed by a vertical one.  Each 1-D pass is O(log k) per pixel: the tile is
// held in shared memory as packed bytes (4 pixels per word), window minima of length 1,2,4,..,P are
// built by doubling (A_2p[i] = op(A_p[i], A_p[i+p])) with ping-pong buffers, and the final window of
// length k is op(A_P[i], A_P[i+k-P]).  Pixels outside the image are ignored, exactly like OpenCV's
// default morphology border: they are loaded as the neutral element (255 for erode, 0 for dilate).
// The last pass can fuse the black-hat subtraction (close(src) - src) and a 256-bin histogram.
#include <algorithm>
#include <cstdlib>

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
// helpers for the register-array kernels: pixels widened to 16-bit lanes (2 px per register) so that min/max is the
// native VIMNMX.U16x2
template <bool DIL> __device__ __forceinline__ uint32_t opx(uint32_t a, uint32_t b) { return DIL ? __vmaxu2(a, b) : __vminu2(a, b); }
__device__ __forceinline__ uint32_t odd_shift(uint32_t lo, uint32_t hi) { return __byte_perm(lo, hi, 0x5432); }   // px (2i+1, 2i+2)

__host__ __device__ constexpr int ilog2_floor(int v) { int l = 0; while ((2 << l) <= v) l++; return l; }

// slow path of a 4-pixel load: any alignment, neutral outside the row (kept out of line: it is rare)
__device__ __noinline__ uint32_t load_word_slow(const uint8_t* rowp, int x, int w, uint32_t neutral_byte) {
    uint32_t wv = 0;
    for (int b = 0; b < 4; b++) wv |= ((x + b >= 0 && x + b < w) ? (uint32_t)rowp[x + b] : neutral_byte) << (8 * b);
    return wv;
}

// ---------------------------------------------------------------------------------------------------------
// Block-marching variant for the rectangles the reference uses (9x19 black-hat, 3x3 close, 2x2 erode) and the odd
// squares of BASELINE.json config 4, default anchor, one iteration: same layout as blur.cu — a CTA
// owns a 128-column strip and marches down it 32 rows at a time; rows are staged in shared memory (neutral outside
// the image), H-filtered into a ring of packed rows, then V-filtered out of the ring.  Both passes run the window
// doubling on register arrays of 16-bit lanes (native VIMNMX.U16x2) with every index a compile-time constant:
// H: 32 outputs per thread, V: one 4-pixel column x 8 rows per thread.  No vertical recompute inside a segment.
namespace march {
constexpr int TWm = 128, BRm = 32, NTm = 128, RPW = 36;   // strip width, rows per step, threads, ring pitch (words)

template <int N, int K, bool DIL>
__device__ __forceinline__ void windows_h(uint32_t (&r)[N]) {           // r[i] = px (2i, 2i+1)  ->  forward windows of K
    constexpr int PW = 1 << ilog2_floor(K);
    if (PW > 1) {
#pragma unroll
        for (int i = 0; i < N; i++) r[i] = opx<DIL>(r[i], odd_shift(r[i], r[i + 1 < N ? i + 1 : N - 1]));   // last: low lane only
    }
#pragma unroll
    for (int p = 2; p < PW; p *= 2) {
#pragma unroll
        for (int i = 0; i + p / 2 < N; i++) r[i] = opx<DIL>(r[i], r[i + p / 2]);
    }
    if (K > PW) {
        constexpr int O = K - PW;
#pragma unroll
        for (int i = 0; i + (O + 1) / 2 < N; i++) r[i] = opx<DIL>(r[i], (O & 1) ? odd_shift(r[i + O / 2], r[i + O / 2 + 1]) : r[i + O / 2]);
    }
}

struct MarchLaunch { int seg_rows, spw, ring_rows; };

template <int KW, int KH, bool DIL>
// small windows: capped at 6 CTAs' worth of registers (72; the prefetch registers would otherwise push it to 96)
__global__ void __launch_bounds__(NTm, (KH <= 19 ? 6 : 1)) morph_march_kernel(const MorphJob* __restrict__ jobs, const MarchLaunch L) {
    constexpr int AX = KW / 2, AY = KH / 2;
    constexpr int AXW = ((AX + 3) / 4) * 4, OFF = AXW - AX;
    constexpr int NWH = (OFF + 31 + KW + 3) / 4, NH = 2 * NWH;         // words / registers a thread needs per row (H)
    constexpr int NRW = 8 + KH - 1;                                     // ring rows a thread needs (V)
    constexpr int SPW = (24 + NWH + 1) | 1;                             // staged words per row == L.spw (launch_t)
    constexpr int PV = 1 << ilog2_floor(KH);
    constexpr int D = (KH - 1 + BRm - 1) / BRm;                         // V step lags the H step by D steps
    constexpr uint32_t NEUTRAL_W = DIL ? 0u : 0xffffffffu;
    const MorphJob J = jobs[blockIdx.z];
    const int x0 = blockIdx.x * TWm;
    const int y_begin = blockIdx.y * L.seg_rows;
    if (x0 >= J.w || y_begin >= J.h) return;
    const int y_end = min(J.h, y_begin + L.seg_rows);
    const int tid = threadIdx.x;
    extern __shared__ __align__(16) uint32_t smem_u32[];
    uint32_t* s_stage = smem_u32;                                        // BRm * spw
    // ring_rows * RPW, on the next 16-byte boundary.  The offset is rounded as an index: rounding the POINTER through an integer
    // cast makes it generic, and every ring access and histogram increment then compiles to LD / ST / ATOM instead of
    // LDS / STS / ATOMS (the dynamic shared array itself is 16-byte aligned).
    uint32_t* s_ring = s_stage + ((BRm * SPW + 3) & ~3);
    uint32_t* s_hist = s_ring + L.ring_rows * RPW;                       // 4 x 256 when J.hist
    if (J.hist) for (int i = tid; i < 4 * 256; i += NTm) s_hist[i] = 0;
    const bool src_al = ((reinterpret_cast<uintptr_t>(J.src) | (uintptr_t)J.src_pitch) & 3) == 0;
    const bool dst_al = ((reinterpret_cast<uintptr_t>(J.dst) | (uintptr_t)J.dst_pitch) & 3) == 0;
    const int n_vb = (y_end - y_begin + BRm - 1) / BRm;
    uint32_t zero_count = 0;
    // Staging is software-pipelined through registers in the strips whose staged rows lie inside the image and are
    // word-aligned (all but the first and last strip of a page): the words of step hb + 1 are requested right after step
    // hb's have been parked in shared memory, so their latency hides behind the H and V passes instead of stalling the STS.
    constexpr int NPF = (SPW + 3) / 4;
    const bool pipelined = src_al && x0 - AXW >= 0 && x0 - AXW + 4 * SPW <= J.w;
    uint32_t pf[NPF];
    auto fetch = [&](int hb) {   // virtual row v <-> source row y_begin - AY + v (neutral outside the image)
        const int gy = y_begin - AY + hb * BRm + (tid >> 2);
        const bool row_in = gy >= 0 && gy < J.h;
        const uint8_t* rowp = J.src + (size_t)(row_in ? gy : 0) * J.src_pitch + (x0 - AXW + 4 * (tid & 3));
#pragma unroll
        for (int j = 0; j < NPF; j++) {
            pf[j] = NEUTRAL_W;
            if (row_in && (tid & 3) + 4 * j < SPW) pf[j] = ds_ldg32(rowp + 16 * j);
        }
    };
    if (pipelined) fetch(0);
    for (int hb = 0; hb < n_vb + D; hb++) {
        if (pipelined) {   // ---- park the 32 staged source rows
            uint32_t* srow_w = s_stage + (tid >> 2) * SPW + (tid & 3);
#pragma unroll
            for (int j = 0; j < NPF; j++) if ((tid & 3) + 4 * j < SPW) srow_w[4 * j] = pf[j];
        } else {           // ---- stage 32 source rows, any alignment, neutral outside the image
            const int srow = tid >> 2;
            const int gy = y_begin - AY + hb * BRm + srow;
            const bool row_in = gy >= 0 && gy < J.h;
            const uint8_t* rowp = J.src + (size_t)(row_in ? gy : 0) * J.src_pitch;
            for (int wi = tid & 3; wi < SPW; wi += 4) {
                const int gx = x0 - AXW + 4 * wi;
                uint32_t word = NEUTRAL_W;
                if (row_in) word = (src_al && gx >= 0 && gx + 3 < J.w) ? ds_ldg32(rowp + gx) : load_word_slow(rowp, gx, J.w, NEUTRAL_W & 255u);
                s_stage[srow * SPW + wi] = word;
            }
        }
        __syncthreads();
        if (pipelined && hb + 1 < n_vb + D) fetch(hb + 1);
        {   // ---- H pass: thread = (row, 32 consecutive outputs)
            const int hr = tid >> 2, cg = tid & 3;
            const uint32_t* sp = s_stage + hr * SPW + cg * 8;
            uint32_t r[NH];
#pragma unroll
            for (int j = 0; j < NWH; j++) { const uint32_t w = sp[j]; r[2 * j] = __byte_perm(w, 0, 0x4140); r[2 * j + 1] = __byte_perm(w, 0, 0x4342); }
            windows_h<NH, KW, DIL>(r);
            uint32_t o[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint32_t a = (OFF & 1) ? odd_shift(r[OFF / 2 + 2 * j], r[OFF / 2 + 2 * j + 1]) : r[OFF / 2 + 2 * j];
                const uint32_t b = (OFF & 1) ? odd_shift(r[OFF / 2 + 2 * j + 1], r[OFF / 2 + 2 * j + 2]) : r[OFF / 2 + 2 * j + 1];
                o[j] = __byte_perm(a, b, 0x6420);
            }
            uint4* dst = reinterpret_cast<uint4*>(s_ring + ((hb * BRm + hr) % L.ring_rows) * RPW + cg * 8);
            dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
            dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
        }
        __syncthreads();
        if (hb < D) continue;
        // ---- V pass: thread = (4-pixel column, 8 rows)
        const int vb = hb - D;
        const int cw = tid & 31, rg = tid >> 5;
        int slot = (vb * BRm + rg * 8) % L.ring_rows;
        uint32_t va[NRW], vbq[NRW];
#pragma unroll
        for (int i = 0; i < NRW; i++) {
            const uint32_t w = s_ring[slot * RPW + cw];
            va[i] = __byte_perm(w, 0, 0x4140); vbq[i] = __byte_perm(w, 0, 0x4342);
            if (++slot == L.ring_rows) slot = 0;
        }
#pragma unroll
        for (int p = 1; p < PV; p *= 2) {
#pragma unroll
            for (int i = 0; i + p < NRW; i++) { va[i] = opx<DIL>(va[i], va[i + p]); vbq[i] = opx<DIL>(vbq[i], vbq[i + p]); }
        }
        if (KH > PV) {
#pragma unroll
            for (int i = 0; i < 8; i++) { va[i] = opx<DIL>(va[i], va[i + KH - PV]); vbq[i] = opx<DIL>(vbq[i], vbq[i + KH - PV]); }
        }
        const int x = x0 + 4 * cw;
        if (x < J.w) {
            const int nvalid = min(4, J.w - x);
            const int yb = y_begin + vb * BRm + rg * 8;
            uint32_t rw[8];                                           // black-hat: the 8 reference words, requested together
            if (J.ref) {
                const uint8_t* rp = J.ref + (size_t)yb * J.ref_pitch + x;
                const bool fast = nvalid == 4 && ((reinterpret_cast<uintptr_t>(rp) | (uintptr_t)J.ref_pitch) & 3) == 0;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    rw[i] = 0;
                    if (yb + i < y_end) rw[i] = fast ? ds_ldg32(rp) : load_word_slow(rp - x, x, J.w, 0);
                    rp += J.ref_pitch;
                }
            }
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int y = yb + i;
                if (y >= y_end) break;
                uint32_t res = __byte_perm(va[i], vbq[i], 0x6420);
                if (J.ref) res = __vsubus4(res, rw[i]);               // sat(close(src) - src)
                if (J.hist) {
                    // a black-hat page is mostly zeros: count those in a register instead of 32 lanes hammering bin 0
                    if (res == 0) zero_count += nvalid;
                    else if (nvalid == 4) {
                        // four predicated increments; the zero bytes are counted with one population count
                        uint32_t* hw = s_hist + (tid >> 5) * 256;
                        const uint32_t v0 = res & 255u, v1 = (res >> 8) & 255u, v2 = (res >> 16) & 255u, v3 = res >> 24;
                        if (v0) atomicAdd(hw + v0, 1u);
                        if (v1) atomicAdd(hw + v1, 1u);
                        if (v2) atomicAdd(hw + v2, 1u);
                        if (v3) atomicAdd(hw + v3, 1u);
                        zero_count += __popc(~(((res & 0x7f7f7f7fu) + 0x7f7f7f7fu) | res | 0x7f7f7f7fu));      // exact per byte: no carries between bytes
                    } else {
                        for (int b = 0; b < nvalid; b++) {
                            const uint32_t v = (res >> (8 * b)) & 255u;
                            if (v) atomicAdd(&s_hist[(tid >> 5) * 256 + v], 1u); else zero_count++;
                        }
                    }
                }
                uint8_t* dp = J.dst + (size_t)y * J.dst_pitch + x;
                if (dst_al && nvalid == 4) *reinterpret_cast<uint32_t*>(dp) = res;
                else for (int b = 0; b < nvalid; b++) dp[b] = (uint8_t)(res >> (8 * b));
            }
        }
    }
    if (J.hist) {
        for (int o = 16; o; o >>= 1) zero_count += __shfl_xor_sync(0xffffffffu, zero_count, o);
        if ((tid & 31) == 0 && zero_count) atomicAdd(&s_hist[(tid >> 5) * 256], zero_count);
        __syncthreads();
        for (int i = tid; i < 256; i += NTm) {
            const uint32_t sum = s_hist[i] + s_hist[256 + i] + s_hist[512 + i] + s_hist[768 + i];
            if (sum) atomicAdd(&J.hist[i], sum);
        }
    }
}

template <int KW, int KH, bool DIL>
int launch_t(docscan_ctx* ctx, const MorphJob* jobs_host, int n, int max_w, int max_h) {
    constexpr int AX = KW / 2, AXW = ((AX + 3) / 4) * 4, OFF = AXW - AX, NWH = (OFF + 31 + KW + 3) / 4;
    constexpr int D = (KH - 1 + BRm - 1) / BRm;
    MarchLaunch L{};
    L.spw = (24 + NWH + 1) | 1;
    L.ring_rows = (D + 1) * BRm;
    bool hist = false;
    double px = 0, refpx = 0;
    for (int i = 0; i < n; i++) {
        hist = hist || jobs_host[i].hist;
        px += (double)jobs_host[i].w * jobs_host[i].h;
        if (jobs_host[i].ref) refpx += (double)jobs_host[i].w * jobs_host[i].h;
    }
    const size_t smem = sizeof(uint32_t) * ((size_t)BRm * L.spw + 4 + (size_t)L.ring_rows * RPW + (hist ? 4 * 256 : 0));
    void* dev = nullptr;
    DS_TRY(ds_upload(ctx, jobs_host, sizeof(MorphJob) * n, &dev));
    if (smem > 48 * 1024)
        DS_CUDA(ctx, cudaFuncSetAttribute(morph_march_kernel<KW, KH, DIL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    DS_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, morph_march_kernel<KW, KH, DIL>, NTm, smem));
    L.seg_rows = ds_pick_seg_rows(per_sm * ctx->sm_count, n * ((max_w + TWm - 1) / TWm), max_h, std::max(64, 4 * KH), BRm);
    dim3 grid((max_w + TWm - 1) / TWm, (max_h + L.seg_rows - 1) / L.seg_rows, n);
    ProfScope prof(ctx, "morph_march_" + std::to_string(KW) + "x" + std::to_string(KH), 2.0 * px + refpx);
    morph_march_kernel<KW, KH, DIL><<<grid, NTm, smem, ctx->stream>>>((const MorphJob*)dev, L);
    DS_CHECK_LAUNCH(ctx);
    return DOCSCAN_OK;
}
}  // namespace march

#define DS_MARCH_SHAPES(X) X(2, 2) X(3, 3) X(5, 5) X(7, 7) X(9, 19) X(9, 9) X(11, 11) X(13, 13) X(15, 15) X(17, 17) X(19, 19) \
    X(21, 21) X(23, 23) X(25, 25) X(27, 27) X(29, 29) X(31, 31)

bool launch_march(docscan_ctx* ctx, int is_dilate, int kw, int kh, int ax, int ay, const MorphJob* jobs_host, int n, int max_w,
                  int max_h, int* rc) {
    if (ax != kw / 2 || ay != kh / 2) return false;
#define DS_MARCH_CASE(W, H)                                                                                       \
    if (kw == W && kh == H) {                                                                                     \
        *rc = is_dilate ? march::launch_t<W, H, true>(ctx, jobs_host, n, max_w, max_h)                            \
                        : march::launch_t<W, H, false>(ctx, jobs_host, n, max_w, max_h);                          \
        return true;                                                                                              \
    }
    DS_MARCH_SHAPES(DS_MARCH_CASE)
#undef DS_MARCH_CASE
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

// ---------------------------------------------------------------------------------------------------------
// morphologyEx(CLOSE | OPEN) with the 3x3 rectangle, one iteration (morph_cleanup's default, DocScanner.py:247-259), as
// ONE register-only pass for the library's own 16-byte-aligned planes.  A thread owns a 16-pixel column chunk (plus one
// word either side) and marches down its rows: the first operation's 3-row window and the second operation's 3-row
// window roll through registers, the horizontal windows come from funnel shifts on packed bytes.  Out-of-image pixels
// are the first operation's neutral element when it reads them and the second operation's neutral element when IT
// reads them (OpenCV ignores outside pixels in each pass separately).  HBM-shaped: the plane is read once, written once.
namespace fused3 {
constexpr int SEG_MAX = 64;      // output rows per thread: halved on the host until the launch fills the machine

// Pixels travel as 16-bit lanes, two per register, so that min / max is the native VIMNMX.U16x2 (the byte-wise
// __vminu4 / __vmaxu4 are emulated with ~20 logic instructions on this architecture).
template <bool DIL> __device__ __forceinline__ uint32_t op2(uint32_t a, uint32_t b) { return DIL ? __vmaxu2(a, b) : __vminu2(a, b); }

// 3-wide horizontal window over N pair registers: m[i] valid in every lane that has both neighbours in the array
template <bool DIL, int N>
__device__ __forceinline__ void hwin(const uint32_t (&r)[N], uint32_t (&m)[N]) {
    constexpr uint32_t NEUT = DIL ? 0u : 0x00ff00ffu;
#pragma unroll
    for (int i = 0; i < N; i++) {
        const uint32_t left = odd_shift(i ? r[i - 1] : NEUT, r[i]);           // (px-1, px0) of the pair (px0, px1)
        const uint32_t right = odd_shift(r[i], i + 1 < N ? r[i + 1] : NEUT);   // (px1, px2)
        m[i] = op2<DIL>(op2<DIL>(r[i], left), right);
    }
}

// FIRST_DIL = true: close (dilate then erode); false: open (erode then dilate)
template <bool FIRST_DIL>
__global__ void __launch_bounds__(128) close3_kernel(const MorphJob* __restrict__ jobs, int chunks, int segs, int SEG) {
    const MorphJob J = jobs[blockIdx.y];
    const int id = blockIdx.x * 128 + threadIdx.x;
    const int sg = id / chunks, xc = id - sg * chunks;
    const int x = xc * 16, y0 = sg * SEG;
    if (sg >= segs || x >= J.w || y0 >= J.h) return;
    constexpr uint32_t N1 = FIRST_DIL ? 0u : 0x00ff00ffu;       // neutral of the first pass
    constexpr uint32_t N2 = FIRST_DIL ? 0x00ff00ffu : 0u;       // neutral of the second pass
    // pair register i holds px (x - 2 + 2i, x - 1 + 2i), i = 0..9; vm: 0xff in the lanes whose column lies inside the image
    uint32_t vm[10];
#pragma unroll
    for (int i = 0; i < 10; i++) {
        const int p = x - 2 + 2 * i;
        vm[i] = ((p >= 0 && p < J.w) ? 0x000000ffu : 0u) | ((p + 1 >= 0 && p + 1 < J.w) ? 0x00ff0000u : 0u);
    }
    const bool has_left = x > 0, has_right = x + 16 < J.w;
    auto load_h = [&](int y, uint32_t (&m)[10]) {                // horizontal first-pass window of source row y
        uint32_t r[10];
        if (y < 0 || y >= J.h) {
#pragma unroll
            for (int i = 0; i < 10; i++) r[i] = N1;
        } else {
            const uint8_t* rowp = J.src + (size_t)y * J.src_pitch + x;
            const uint4 c = __ldg(reinterpret_cast<const uint4*>(rowp));
            const uint32_t wl = has_left ? ds_ldg32(rowp - 4) : 0u, wr = has_right ? ds_ldg32(rowp + 16) : 0u;
            r[0] = __byte_perm(wl, 0, 0x4342);
            r[1] = __byte_perm(c.x, 0, 0x4140); r[2] = __byte_perm(c.x, 0, 0x4342);
            r[3] = __byte_perm(c.y, 0, 0x4140); r[4] = __byte_perm(c.y, 0, 0x4342);
            r[5] = __byte_perm(c.z, 0, 0x4140); r[6] = __byte_perm(c.z, 0, 0x4342);
            r[7] = __byte_perm(c.w, 0, 0x4140); r[8] = __byte_perm(c.w, 0, 0x4342);
            r[9] = __byte_perm(wr, 0, 0x4140);
#pragma unroll
            for (int i = 0; i < 10; i++) r[i] = FIRST_DIL ? (r[i] & vm[i]) : (r[i] | (vm[i] ^ 0x00ff00ffu));
        }
        hwin<FIRST_DIL, 10>(r, m);
    };
    // second-pass horizontal window of first-pass row r: outside the image the first pass's result is replaced by the
    // second pass's neutral element
    auto second_h = [&](int r, const uint32_t (&a)[10], const uint32_t (&b)[10], const uint32_t (&c)[10], uint32_t (&out)[8]) {
        uint32_t d[10], m[10];
        const bool row_in = r >= 0 && r < J.h;
#pragma unroll
        for (int i = 0; i < 10; i++) {
            const uint32_t v = op2<FIRST_DIL>(op2<FIRST_DIL>(a[i], b[i]), c[i]);
            d[i] = !row_in ? N2 : (FIRST_DIL ? (v | (vm[i] ^ 0x00ff00ffu)) : (v & vm[i]));
        }
        hwin<!FIRST_DIL, 10>(d, m);
#pragma unroll
        for (int i = 0; i < 8; i++) out[i] = m[i + 1];
    };
    // rolling state, indexed modulo 3 with compile-time indices (the row loop is unrolled by 3): hs = first-pass
    // horizontal rows, he = second-pass horizontal rows
    uint32_t hs[3][10], he[3][8];
    const int y_end = min(y0 + SEG, J.h);
    load_h(y0 - 2, hs[0]);
    load_h(y0 - 1, hs[1]);
    load_h(y0, hs[2]);
    second_h(y0 - 1, hs[0], hs[1], hs[2], he[0]);               // second-pass row y0-1
    load_h(y0 + 1, hs[0]);
    second_h(y0, hs[1], hs[2], hs[0], he[1]);                   // row y0
    const bool full = x + 16 <= J.w;
    // invariant at the top of step u (output row y): hs[(u+1)%3] = row y, hs[(u+2)%3]... see the indices below
    for (int yb = y0; yb < y_end; yb += 3) {
#pragma unroll
        for (int u = 0; u < 3; u++) {
            const int y = yb + u;
            // rows y, y+1 live in hs[(u+2)%3], hs[u%3]; the new row y+2 replaces row y-1 in hs[(u+1)%3]
            load_h(y + 2, hs[(u + 1) % 3]);
            second_h(y + 1, hs[(u + 2) % 3], hs[u % 3], hs[(u + 1) % 3], he[(u + 2) % 3]);     // row y+1
            if (y < y_end) {
                uint32_t o[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t lo = op2<!FIRST_DIL>(op2<!FIRST_DIL>(he[0][2 * j], he[1][2 * j]), he[2][2 * j]);
                    const uint32_t hi = op2<!FIRST_DIL>(op2<!FIRST_DIL>(he[0][2 * j + 1], he[1][2 * j + 1]), he[2][2 * j + 1]);
                    o[j] = __byte_perm(lo, hi, 0x6420);
                }
                uint8_t* dp = J.dst + (size_t)y * J.dst_pitch + x;
                if (full) *reinterpret_cast<uint4*>(dp) = make_uint4(o[0], o[1], o[2], o[3]);
                else for (int b = 0; b < J.w - x; b++) dp[b] = (uint8_t)(o[b >> 2] >> (8 * (b & 3)));
            }
        }
    }
}
}  // namespace fused3

// true when the fused 3x3 close / open kernel took the batch
bool k_morph_close3(docscan_ctx* ctx, int open_not_close, const MorphJob* jobs_host, int n, int max_w, int max_h, int* rc) {
    for (int i = 0; i < n; i++) {
        const MorphJob& j = jobs_host[i];
        const int need = ((j.w + 15) >> 4) << 4;                 // whole 16-byte chunks are read
        if (((reinterpret_cast<uintptr_t>(j.src) | reinterpret_cast<uintptr_t>(j.dst) | (uintptr_t)j.src_pitch | (uintptr_t)j.dst_pitch) & 15) ||
            j.src_pitch < need || j.ref || j.hist)
            return false;
        const uint8_t* s0 = j.src; const uint8_t* s1 = j.src + (size_t)j.src_pitch * j.h;
        const uint8_t* d0 = j.dst; const uint8_t* d1 = j.dst + (size_t)j.dst_pitch * j.h;
        if (s0 < d1 && d0 < s1) return false;                   // in place: neighbours would be read after they were overwritten
    }
    void* dev = nullptr;
    *rc = ds_upload(ctx, jobs_host, sizeof(MorphJob) * n, &dev);
    if (*rc != DOCSCAN_OK) return true;
    const int chunks = (max_w + 15) >> 4;
    int seg = fused3::SEG_MAX;                                   // every segment re-reads 4 halo rows: keep them long ...
    while (seg > 8 && (long long)chunks * ((max_h + seg - 1) / seg) * n < 2LL * ctx->sm_count * 2048) seg >>= 1;   // ... unless SMs would idle
    const int segs = (max_h + seg - 1) / seg;
    double px = 0;
    for (int i = 0; i < n; i++) px += (double)jobs_host[i].w * jobs_host[i].h;
    ProfScope prof(ctx, open_not_close ? "morph_open3_fused" : "morph_close3_fused", 2.0 * px);
    const dim3 grid((chunks * segs + 127) / 128, n);
    if (open_not_close) fused3::close3_kernel<false><<<grid, 128, 0, ctx->stream>>>((const MorphJob*)dev, chunks, segs, seg);
    else fused3::close3_kernel<true><<<grid, 128, 0, ctx->stream>>>((const MorphJob*)dev, chunks, segs, seg);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    *rc = e == cudaSuccess ? DOCSCAN_OK : ds_fail(ctx, DOCSCAN_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
    return true;
}

// One 2-D erode/dilate pass for every job: H pass src -> tmp (arena), V pass tmp -> dst (+ epilogue).
// (kw, kh, ax, ay) already include the `iterations` enlargement.
int k_morph_jobs(docscan_ctx* ctx, int is_dilate, int kw, int kh, int ax, int ay, const MorphJob* jobs_host, int n,
                 int max_w, int max_h) {
    if (kw < 1 || kh < 1) return ds_fail(ctx, DOCSCAN_ERR_BAD_ARG, "bad structuring element %dx%d", kw, kh);
    {
        int rc = DOCSCAN_OK;
        if (launch_march(ctx, is_dilate, kw, kh, ax, ay, jobs_host, n, max_w, max_h, &rc)) return rc;
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
