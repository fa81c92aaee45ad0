// cv2.GaussianBlur(u8, 8.8 fixed point, BORDER_REFLECT_101) as two exact-integer banded-Toeplitz contractions on the
// 5th-generation tensor cores (tcgen05, kind::i8, s32 accumulators in tensor memory), with the same fused epilogues as
// blur.cu (subtract / divide / reverse subtract, min-max, histogram).  DocScanner.py:153,184.
//
// Every sum of the fixed-point blur is an exact integer, so regrouping it into matrix products is bit-exact:
//
//   pass 1 (vertical):   D1[y, x']  = sum_y'  Tv[y, y'] * S[y', x']        M = 128 output rows, N = 128 input columns, K = K1 rows
//   pass 2 (horizontal): D2[y, x ]  = sum_x'  D1[y, x'] * Th[x', x]        M = 128, N = NOUT output columns, K = 128
//   dst = (D2 + 32768) >> 16
//
// S  : the u8 source tile, K1 = 128 + 2R rows (rounded up to 32) x 128 columns, fetched by ONE TMA tensor-map request per
//      tile (128-byte swizzle, out-of-image elements zero-filled); image rows are contiguous along x, which makes the tile
//      an MN-major B operand as it lies.
// Tv : constant band matrix [128 x K1] (u8 coefficients of the quantised kernel), A operand, K-major, in shared memory.
//      The border rule is folded into the matrix: a top / bottom tile uses a variant in which the taps that REFLECT_101
//      maps back into the image are added onto the rows they land on — the kernel itself never sees a border.
// D1 : s32 in tensor memory, values <= 255 * 256.  The epilogue warps read it back (tcgen05.ld), split every value into
//      its low and high byte, pack four neighbours per 32-bit word and write the two byte planes back to tensor memory
//      (tcgen05.st) as the A operands of pass 2 — D1 never touches shared or global memory.
// Th : constant band matrix [NOUT x 128], B operand, K-major (left / right border variants like Tv).
// D2 : two accumulators, D2lo = A2lo * Th and D2hi = A2hi * Th;  V = D2lo + 256 * D2hi.
//
// One CTA = 8 warps, 256 TMEM columns, ~100 KB of shared memory -> two CTAs per SM: while one CTA drains its accumulators
// the other one's MMAs run.  Tiles are dealt round-robin to a persistent grid of 2 x SMs CTAs.
// Out-of-scope here (blur.cu keeps them): box sums, radii above 48, buffers that are not 16-byte aligned.
#include <cuda.h>      // CUtensorMap types; cuTensorMapEncodeTiled is fetched through cudaGetDriverEntryPoint (no -lcuda)

#include <array>

#include "common.cuh"
#include "tc05.cuh"

namespace {

constexpr int TM = 128;          // output rows per tile (MMA M, TMEM lanes)
constexpr int NIN = 128;         // input columns per tile (pass-1 N, pass-2 K)
constexpr int NT = 256;
constexpr int TMEM_COLS = 256;
constexpr int COL_D1 = 0, COL_D2LO = 0, COL_D2HI = 96, COL_A2LO = 192, COL_A2HI = 224;
constexpr int T_BYTES = 2 * TM * 128;      // two 128-byte K blocks
constexpr int TOE_BYTES = 96 * 128;
constexpr int MAX_R = 48;
constexpr int MAX_JOBS = 64, MAX_TABS = 256;      // page table / band-matrix pointer table kept in shared memory

struct TcJob {
    const uint8_t* src; uint8_t* dst;
    int src_pitch, dst_pitch, w, h;
    uint32_t* minmax; uint32_t* hist;
    int tile_base, ntx, nty;
    int t_off, toe_off;          // this page's first row / column variant pointer in `tabs`
};

struct TcLaunch {
    const TcJob* jobs;
    const CUtensorMap* maps;
    const uint8_t* const* tabs;
    uint32_t* dbg;               // debug dump of the first tile (DOCSCAN_TC_DEBUG), else null
    volatile uint32_t* status;   // pinned host words: [0] = which wait timed out, [1] = progress of CTA 0 (debug runs)
    int n_jobs, n_tabs, total_tiles;
    int crumbs;                  // debug: CTA 0 reports its progress to status[1] (slow: a system-scope fence per phase)
    int R, RL, K1, NOUT;         // RL: left margin of the source window (TMA needs its first byte 16-byte aligned)
    uint32_t idesc1, idesc2;
};

// pass 1 of one tile: NK MMAs of K = 32 source rows each (K-major band matrix in two 128-byte blocks, MN-major source tile)
template <int NK>
__device__ __forceinline__ void issue_pass1(uint32_t d_tmem, uint64_t dT, uint64_t dS, uint32_t idesc) {
#pragma unroll
    for (int s = 0; s < NK; s++)
        tc::mma_i8_ss(d_tmem, dT + (uint64_t)(((s >> 2) * (TM * 128) + (s & 3) * 32) >> 4), dS + (uint64_t)((s * 4096) >> 4), idesc, s > 0);
}

template <int EPI, bool STATS>
__global__ void __launch_bounds__(NT, 2) tc_blur_kernel(const __grid_constant__ TcLaunch L) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sT = base;
    uint8_t* sToe = sT + T_BYTES;
    uint8_t* sS[2] = {sToe + TOE_BYTES, sToe + TOE_BYTES + L.K1 * 128};
    uint32_t* s_hist = reinterpret_cast<uint32_t*>(sS[1] + L.K1 * 128);       // 8 x 256, only with STATS
    __shared__ uint64_t bar_s[2], bar_c, bar_d1, bar_d2;
    __shared__ uint32_t s_tmem;
    __shared__ TcJob s_jobs[MAX_JOBS];                   // the launch's page table and band-matrix pointers, read every tile
    __shared__ const uint8_t* s_tabs[MAX_TABS];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        tc::mbar_init(&bar_s[0], 1); tc::mbar_init(&bar_s[1], 1); tc::mbar_init(&bar_c, 1);
        tc::mbar_init(&bar_d1, 1); tc::mbar_init(&bar_d2, 1);
        tc::mbar_init_fence();
    }
    if (warp == 1) tc::tmem_alloc(&s_tmem, TMEM_COLS);
    for (int i = tid; i < L.n_jobs; i += NT) s_jobs[i] = L.jobs[i];
    for (int i = tid; i < L.n_tabs; i += NT) s_tabs[i] = L.tabs[i];
    if (STATS)
        for (int i = tid; i < 8 * 256; i += NT) s_hist[i] = 0;
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = s_tmem;
#define TC_CRUMB(v) do { if (L.crumbs && blockIdx.x == 0 && warp == 0) { L.status[1] = (v); __threadfence_system(); } } while (0)
#define TC_WAIT(bar, par, id) do { if (!tc::mbar_wait_bounded(bar, par)) { L.status[0] = (id); __threadfence_system(); __trap(); } } while (0)
#define TC_STAMP(k) do { if (L.dbg && blockIdx.x == 0 && warp == 0 && it < 64) { long long c_ = clock64(); L.dbg[40960 + it * 32 + 2 * (k)] = (uint32_t)c_; L.dbg[40960 + it * 32 + 2 * (k) + 1] = (uint32_t)(c_ >> 32); } } while (0)
    TC_CRUMB(1);
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;      // this warp's quarter of the TMEM lanes
    const int hf = warp >> 2;                                          // which half of the columns this warp drains
    const int row = (warp & 3) * 32 + lane;                            // tile row of this thread

    // tile index -> (job, tx, ty); tiles of a page are numbered row-major so that neighbours share their halo in L2
    auto locate = [&](int t, int& j, int& tx, int& ty) {
        while (j + 1 < L.n_jobs && t >= s_jobs[j + 1].tile_base) j++;
        const int idx = t - s_jobs[j].tile_base, ntx = s_jobs[j].ntx;
        ty = idx / ntx; tx = idx - ty * ntx;
    };

    // operand descriptors: constant but for the start address (low word)
    const uint64_t dT = tc::smem_desc_sw128(tc::smem_u32(sT), 16, 1024);
    const uint64_t dS[2] = {tc::smem_desc_sw128(tc::smem_u32(sS[0]), (uint32_t)L.K1 * 128, 1024),
                            tc::smem_desc_sw128(tc::smem_u32(sS[1]), (uint32_t)L.K1 * 128, 1024)};
    const uint64_t dToe = tc::smem_desc_sw128(tc::smem_u32(sToe), 16, 1024);

    // issuing thread's state
    const uint8_t* cur_t = nullptr; const uint8_t* cur_toe = nullptr;
    uint32_t ph_c = 0, ph_s[2] = {0, 0};
    int job = 0, tx = 0, ty = 0;
    if ((int)blockIdx.x < L.total_tiles) locate(blockIdx.x, job, tx, ty);
    if (warp == 0 && (int)blockIdx.x < L.total_tiles && tc::elect_one()) {
        tc::tmap_acquire(&L.maps[job]);
        tc::mbar_expect_tx(&bar_s[0], (uint32_t)L.K1 * 128);
        tc::tma_load_2d(sS[0], &L.maps[job], tx * L.NOUT - L.RL, ty * TM - L.R, &bar_s[0]);
        TC_CRUMB(12);
    }
    // running statistics of the current page (flushed when the page changes)
    uint32_t mn2 = 0x00FF00FFu, mx2 = 0;                 // min / max as two 16-bit lanes
    uint32_t* st_minmax = nullptr; uint32_t* st_hist = nullptr;

    auto flush_stats = [&]() {
        if (!STATS) return;
        if (st_minmax) {
            uint32_t lo = min(mn2 & 0xFFFFu, mn2 >> 16), hi = max(mx2 & 0xFFFFu, mx2 >> 16);
            for (int o = 16; o; o >>= 1) {
                lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
                hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
            }
            if (lane == 0 && lo <= hi) { atomicMin(&st_minmax[0], lo); atomicMax(&st_minmax[1], hi); }
        }
        if (st_hist) {
            __syncthreads();
            for (int i = tid; i < 256; i += NT) {
                uint32_t sum = 0;
#pragma unroll
                for (int wv = 0; wv < 8; wv++) { sum += s_hist[wv * 256 + i]; s_hist[wv * 256 + i] = 0; }
                if (sum) atomicAdd(&st_hist[i], sum);
            }
            __syncthreads();
        }
        mn2 = 0x00FF00FFu; mx2 = 0;
    };

    int it = 0;
    for (int t = blockIdx.x; t < L.total_tiles; t += gridDim.x, it++) {
        TC_STAMP(11);
        const TcJob& J = s_jobs[job];
        if (STATS && (st_minmax != J.minmax || st_hist != J.hist)) {   // statistics are per page
            if (it > 0) flush_stats();
            st_minmax = J.minmax; st_hist = J.hist;
        }
        const int x0 = tx * L.NOUT, y0 = ty * TM;
        const int stage = it & 1;
        // the tile after this one (every thread keeps the same view)
        const int tn = t + gridDim.x;
        int jobn = job, txn = 0, tyn = 0;
        TC_STAMP(0);

        if (warp == 0 && tc::elect_one()) {
            const uint8_t* want_t = s_tabs[J.t_off + ty];
            const uint8_t* want_toe = s_tabs[J.toe_off + tx];
            if (want_t != cur_t || want_toe != cur_toe) {
                // every MMA that read the old matrices has completed (bar_d2 of the previous tile was waited for)
                uint32_t bytes = 0;
                if (want_t != cur_t) bytes += T_BYTES;
                if (want_toe != cur_toe) bytes += (uint32_t)L.NOUT * 128;
                tc::mbar_expect_tx(&bar_c, bytes);
                if (want_t != cur_t) tc::bulk_load(sT, want_t, T_BYTES, &bar_c);
                if (want_toe != cur_toe) tc::bulk_load(sToe, want_toe, (uint32_t)L.NOUT * 128, &bar_c);
                cur_t = want_t; cur_toe = want_toe;
                TC_WAIT(&bar_c, ph_c, 1); ph_c ^= 1;
                TC_CRUMB(2);
            }
            TC_STAMP(1);
            TC_WAIT(&bar_s[stage], ph_s[stage], 2); ph_s[stage] ^= 1;
            TC_CRUMB(3);
            TC_STAMP(2);
            tc::fence_after_sync();
            // pass 1: D1[128 x 128] = Tv[128 x K1] * S[K1 x 128]
            switch (L.K1 >> 5) {
                case 5: issue_pass1<5>(tmem + COL_D1, dT, dS[stage], L.idesc1); break;
                case 6: issue_pass1<6>(tmem + COL_D1, dT, dS[stage], L.idesc1); break;
                case 7: issue_pass1<7>(tmem + COL_D1, dT, dS[stage], L.idesc1); break;
                default: issue_pass1<8>(tmem + COL_D1, dT, dS[stage], L.idesc1); break;
            }
            tc::mma_commit(&bar_d1);
            TC_CRUMB(4);
            TC_STAMP(3);
            // fetch the next tile's source window into the other stage (its last reader, MMA1 of the previous tile, is done)
            if (tn < L.total_tiles) {
                locate(tn, jobn, txn, tyn);
                if (jobn != job) tc::tmap_acquire(&L.maps[jobn]);
                tc::mbar_expect_tx(&bar_s[stage ^ 1], (uint32_t)L.K1 * 128);
                tc::tma_load_2d(sS[stage ^ 1], &L.maps[jobn], txn * L.NOUT - L.RL, tyn * TM - L.R, &bar_s[stage ^ 1]);
            }
        }
        __syncwarp();
        jobn = job;
        if (tn < L.total_tiles) locate(tn, jobn, txn, tyn);           // while pass 1 runs

        // ---- D1 -> byte planes (A operands of pass 2): this warp's 32 rows x 64 columns
        TC_WAIT(&bar_d1, it & 1, 3);
        tc::fence_after_sync();
        TC_CRUMB(5);
        TC_STAMP(4);
        {
            uint32_t lo[16], hi[16];
#pragma unroll
            for (int part = 0; part < 2; part++) {
                uint32_t v[32];
                tc::tmem_ld32(tmem + lane_base + COL_D1 + hf * 64 + part * 32, v);
                tc::tmem_wait_ld();
                if (L.dbg && t == 0)
                    for (int i = 0; i < 32; i++) L.dbg[row * 128 + hf * 64 + part * 32 + i] = v[i];
#pragma unroll
                for (int g = 0; g < 8; g++) {
                    const uint32_t t1 = __byte_perm(v[4 * g], v[4 * g + 1], 0x5140);        // a0 b0 a1 b1
                    const uint32_t t2 = __byte_perm(v[4 * g + 2], v[4 * g + 3], 0x5140);    // c0 d0 c1 d1
                    lo[part * 8 + g] = __byte_perm(t1, t2, 0x5410);                         // a0 b0 c0 d0
                    hi[part * 8 + g] = __byte_perm(t1, t2, 0x7632);                         // a1 b1 c1 d1
                }
            }
            if (hf) {
                // columns 126 and 127 carry no tap: they hold the rounding constant instead, 2 x (128 * 128) = 32768,
                // against the two 128s in the band matrix's last two slots
                lo[15] = (lo[15] & 0x0000FFFFu) | 0x80800000u;
                hi[15] &= 0x0000FFFFu;
            }
            TC_CRUMB(6);
            tc::tmem_st16(tmem + lane_base + COL_A2LO + hf * 16, lo);
            tc::tmem_st16(tmem + lane_base + COL_A2HI + hf * 16, hi);
            tc::tmem_wait_st();
            TC_CRUMB(7);
        }
        TC_STAMP(5);
        tc::fence_before_sync();
        __syncthreads();
        TC_STAMP(6);
        if (warp == 0 && tc::elect_one()) {
            tc::fence_after_sync();
            // pass 2: D2lo / D2hi [128 x NOUT] = A2lo / A2hi [128 x 128] (tensor memory) * Th[128 x NOUT]
#pragma unroll
            for (int s = 0; s < NIN / 32; s++) tc::mma_i8_ts(tmem + COL_D2LO, tmem + COL_A2LO + s * 8, dToe + (uint64_t)(s * 2), L.idesc2, s > 0);
#pragma unroll
            for (int s = 0; s < NIN / 32; s++) tc::mma_i8_ts(tmem + COL_D2HI, tmem + COL_A2HI + s * 8, dToe + (uint64_t)(s * 2), L.idesc2, s > 0);
            tc::mma_commit(&bar_d2);
            TC_CRUMB(8);
            TC_STAMP(7);
        }
        __syncwarp();

        // ---- epilogue: this warp's 32 rows x its half of the NOUT columns, 16 columns at a time
        const int H0 = ((L.NOUT >> 1) + 15) & ~15;
        const int c_begin = hf ? H0 : 0, c_end = hf ? L.NOUT : H0;
        const int y = y0 + row;
        const bool row_ok = y < J.h;
        uint8_t* drow = J.dst + (size_t)y * J.dst_pitch;
        // the centre pixels are in the source tile: row `row + R`, 16-byte chunk (c + RL) / 16, swizzled by the row number
        const int srow_i = row + L.R;
        const uint8_t* s_center = sS[stage] + srow_i * 128;
        const int chunk0 = L.RL >> 4, swz = srow_i & 7;
        uint32_t hc_ev = 0, hc_od = 0;                    // this tile's counts of the values 0..7 (8 bits each: even / odd bins)
        TC_WAIT(&bar_d2, it & 1, 4);
        tc::fence_after_sync();
        TC_CRUMB(9);
        TC_STAMP(8);
        for (int c = c_begin; c < c_end; c += 16) {
            uint32_t lo[16], hi[16];
            tc::tmem_ld16(tmem + lane_base + COL_D2LO + c, lo);
            tc::tmem_ld16(tmem + lane_base + COL_D2HI + c, hi);
            uint4 cw = make_uint4(0, 0, 0, 0);
            if (EPI != DS_EPI_BLUR) cw = *reinterpret_cast<const uint4*>(s_center + ((((c >> 4) + chunk0) ^ swz) << 4));
            tc::tmem_wait_ld();
            if (L.dbg && t == 0)
                for (int i = 0; i < 16; i++) {
                    L.dbg[16384 + row * 96 + c + i] = lo[i];
                    L.dbg[16384 + 12288 + row * 96 + c + i] = hi[i];
                }
            const int x = x0 + c;
            if (row_ok && x < J.w) {
                const uint32_t cws[4] = {cw.x, cw.y, cw.z, cw.w};
                const int nvalid = min(16, J.w - x);
                uint32_t out[4];
                uint32_t dl[8];                             // results as 16-bit lanes: dl[2g] = (px 4g, px 4g+2), dl[2g+1] = (px 4g+1, px 4g+3)
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    // blurred byte = bits 16..23 of D2lo + 256 * D2hi (the rounding constant is already in the sum; bits 24.. are 0)
                    const uint32_t e0 = lo[4 * g] + (hi[4 * g] << 8), e1 = lo[4 * g + 1] + (hi[4 * g + 1] << 8);
                    const uint32_t e2 = lo[4 * g + 2] + (hi[4 * g + 2] << 8), e3 = lo[4 * g + 3] + (hi[4 * g + 3] << 8);
                    const uint32_t b_ev = __byte_perm(e0, e2, 0x7632), b_od = __byte_perm(e1, e3, 0x7632);
                    uint32_t d_ev = b_ev, d_od = b_od;
                    if (EPI != DS_EPI_BLUR) {
                        const uint32_t s_ev = __byte_perm(cws[g], 0u, 0x4240), s_od = __byte_perm(cws[g], 0u, 0x4341);
                        if (EPI == DS_EPI_SUB) {            // sat(s - b) = max(s, b) - b, lane-wise without borrows
                            d_ev = __vmaxu2(s_ev, b_ev) - b_ev; d_od = __vmaxu2(s_od, b_od) - b_od;
                        } else if (EPI == DS_EPI_RSUB) {
                            d_ev = __vmaxu2(s_ev, b_ev) - s_ev; d_od = __vmaxu2(s_od, b_od) - s_od;
                        } else {                            // divide(s, b, 255) in fp32, per pixel
                            d_ev = (uint32_t)ds_div255((uint8_t)s_ev, (uint8_t)b_ev) | ((uint32_t)ds_div255((uint8_t)(s_ev >> 16), (uint8_t)(b_ev >> 16)) << 16);
                            d_od = (uint32_t)ds_div255((uint8_t)s_od, (uint8_t)b_od) | ((uint32_t)ds_div255((uint8_t)(s_od >> 16), (uint8_t)(b_od >> 16)) << 16);
                        }
                    }
                    dl[2 * g] = d_ev; dl[2 * g + 1] = d_od;
                    out[g] = __byte_perm(d_ev, d_od, 0x6240);
                }
                if (nvalid == 16) {
                    *reinterpret_cast<uint4*>(drow + x) = make_uint4(out[0], out[1], out[2], out[3]);
                } else {                                    // last columns of the page: never write past its width
                    for (int i = 0; i < nvalid; i++) drow[x + i] = (uint8_t)(out[i >> 2] >> (8 * (i & 3)));
                }
                if (STATS) {
                    if (nvalid < 16) {                      // rare: keep the columns past the width out of the statistics
                        for (int i = 0; i < nvalid; i++) {
                            const uint32_t v = (out[i >> 2] >> (8 * (i & 3))) & 0xFFu;
                            if (J.minmax) { mn2 = __vminu2(mn2, v | 0x00FF0000u); mx2 = __vmaxu2(mx2, v); }
                            if (J.hist) atomicAdd(&s_hist[warp * 256 + v], 1u);
                        }
                    } else {
                        if (J.minmax) {
#pragma unroll
                            for (int i = 0; i < 8; i++) { mn2 = __vminu2(mn2, dl[i]); mx2 = __vmaxu2(mx2, dl[i]); }
                        }
                        if (J.hist) {
                            // values 0..7: 4-bit counters in two registers (a 32-bit shift by 32 or more gives 0, so larger values
                            // add nothing here); at most 8 increments per field and chunk
                            uint32_t h0 = 0, h1 = 0;
#pragma unroll
                            for (int i = 0; i < 8; i++) {
                                const uint32_t d = dl[i];
                                h0 += tc::shl32(1u, (d << 2) & 0x3FCu);
                                h1 += tc::shl32(1u, (d >> 14) & 0x3FCu);
                                if (d & 0x00F800F8u) {      // a value of 8 or more: the shared-memory histogram
                                    const uint32_t v0 = d & 0xFFu, v1 = d >> 16;
                                    if (v0 >= 8) atomicAdd(&s_hist[warp * 256 + v0], 1u);
                                    if (v1 >= 8) atomicAdd(&s_hist[warp * 256 + v1], 1u);
                                }
                            }
                            hc_ev += (h0 & 0x0F0F0F0Fu) + (h1 & 0x0F0F0F0Fu);               // bins 0, 2, 4, 6 (8 bits each)
                            hc_od += ((h0 >> 4) & 0x0F0F0F0Fu) + ((h1 >> 4) & 0x0F0F0F0Fu);  // bins 1, 3, 5, 7
                        }
                    }
                }
            }
            __syncwarp();                                   // the tensor-memory loads of the next round are warp-wide
        }
        if (STATS && J.hist) {
            // the small values of this tile: 8-bit counters -> 16-bit lanes, summed over the warp, 8 atomics per warp
            uint32_t f0 = (hc_ev & 0xFFu) | ((hc_od & 0xFFu) << 16);                       // bins 0, 1
            uint32_t f1 = ((hc_ev >> 8) & 0xFFu) | (((hc_od >> 8) & 0xFFu) << 16);         // bins 2, 3
            uint32_t f2 = ((hc_ev >> 16) & 0xFFu) | (((hc_od >> 16) & 0xFFu) << 16);       // bins 4, 5
            uint32_t f3 = (hc_ev >> 24) | ((hc_od >> 24) << 16);                           // bins 6, 7
            for (int o = 16; o; o >>= 1) {
                f0 += __shfl_xor_sync(0xffffffffu, f0, o); f1 += __shfl_xor_sync(0xffffffffu, f1, o);
                f2 += __shfl_xor_sync(0xffffffffu, f2, o); f3 += __shfl_xor_sync(0xffffffffu, f3, o);
            }
            if (lane < 8) {
                const uint32_t f = (lane >> 1) == 0 ? f0 : (lane >> 1) == 1 ? f1 : (lane >> 1) == 2 ? f2 : f3;
                const uint32_t cnt = (f >> (16 * (lane & 1))) & 0xFFFFu;
                if (cnt) atomicAdd(&s_hist[warp * 256 + lane], cnt);
            }
        }
        TC_CRUMB(10);
        TC_STAMP(9);
        tc::fence_before_sync();
        __syncthreads();                                    // accumulators drained: the next tile may overwrite them
        tc::fence_after_sync();
        TC_STAMP(10);
        job = jobn; tx = txn; ty = tyn;
        TC_STAMP(12);
    }
    flush_stats();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem, TMEM_COLS);
    TC_CRUMB(11);
#undef TC_CRUMB
#undef TC_STAMP
#undef TC_WAIT
}

// ---- host ---------------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        cudaGetLastError();
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;
    return p;
}

// Byte (row, k) of a K-major operand with 128-byte swizzle: 128-byte rows, 16-byte chunks XORed with the row number mod 8;
// K beyond 128 continues in the next block of `rows` rows.
size_t sw128_offset(int rows, int row, int k) {
    const int blk = k >> 7, kk = k & 127;
    return (size_t)blk * rows * 128 + (size_t)row * 128 + (size_t)(((kk >> 4) ^ (row & 7)) << 4) + (kk & 15);
}

// Band matrix of one tile row (vertical, A operand [128 x K1]) or tile column (horizontal, B operand [NOUT x 128]):
// entry (o, slot) = sum of the taps of output o0 + o that the border rule maps onto source index o0 - margin + slot.
// Returns false when a folded coefficient does not fit 8 bits (images much smaller than the kernel).
bool build_band(const int32_t* q, int k_eff, int R, int margin, int o0, int len, int n_out, int n_slots, int rows_alloc, bool rounding_slots,
                uint8_t* img, size_t img_bytes) {
    std::vector<int> acc((size_t)n_out * n_slots, 0);
    for (int o = 0; o < n_out; o++) {
        const int p = o0 + o;
        if (p >= len) break;
        for (int t = 0; t < k_eff; t++) {
            const int sp = reflect101(p + t - R, len);
            const int slot = sp - (o0 - margin);
            if (slot < 0 || slot >= n_slots) return false;
            acc[(size_t)o * n_slots + slot] += q[t];
        }
    }
    if (rounding_slots)
        for (int o = 0; o < n_out; o++) {
            if (acc[(size_t)o * n_slots + n_slots - 2] || acc[(size_t)o * n_slots + n_slots - 1]) return false;
            acc[(size_t)o * n_slots + n_slots - 2] = acc[(size_t)o * n_slots + n_slots - 1] = 128;    // x 128 in the A operand, twice = 32768
        }
    memset(img, 0, img_bytes);
    for (int o = 0; o < n_out; o++)
        for (int s = 0; s < n_slots; s++) {
            const int v = acc[(size_t)o * n_slots + s];
            if (v > 255) return false;
            if (v) img[sw128_offset(rows_alloc, o, s)] = (uint8_t)v;
        }
    return true;
}

struct Variant { const uint8_t* dev; };

// device copy of one band matrix, cached per context: key = (axis, k, top/left distance or -1, bottom/right distance or -1)
int get_variant(docscan_ctx* ctx, int axis, int k, const int32_t* q, int k_eff, int R, int RL, int K1, int NOUT, int o0, int len, const uint8_t** out,
                bool* ok) {
    const int n_out = axis == 0 ? TM : NOUT, n_slots = axis == 0 ? K1 : NIN;
    const int a = (o0 - R < 0) ? o0 : -1;
    const int b = (o0 + n_out - 1 + R > len - 1) ? len - o0 : -1;
    const std::array<int, 6> key = {axis, k, a, b, NOUT, K1};
    auto it = ctx->tc_tables.find(key);
    if (it == ctx->tc_tables.end()) {
        const size_t bytes = axis == 0 ? (size_t)T_BYTES : (size_t)TOE_BYTES;
        std::vector<uint8_t> img(bytes);
        *ok = build_band(q, k_eff, R, axis == 0 ? R : RL, o0, len, n_out, n_slots, axis == 0 ? TM : NOUT, axis == 1, img.data(), bytes);
        if (!*ok) return DOCSCAN_OK;
        void* dev = nullptr;
        DS_CUDA(ctx, cudaMalloc(&dev, bytes));
        DS_CUDA(ctx, cudaMemcpyAsync(dev, img.data(), bytes, cudaMemcpyHostToDevice, ctx->stream));
        DS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        it = ctx->tc_tables.emplace(key, dev).first;
    }
    *ok = true;
    *out = reinterpret_cast<const uint8_t*>(it->second);
    return DOCSCAN_OK;
}

template <int EPI, bool STATS>
int launch_tc(docscan_ctx* ctx, const TcLaunch& L, size_t smem) {
    static bool attr_done = false;          // per template instance; the attribute is per function and device
    (void)attr_done;
    DS_CUDA(ctx, cudaFuncSetAttribute(tc_blur_kernel<EPI, STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = std::min(L.total_tiles, 2 * ctx->sm_count);
    if (const char* e = getenv("DOCSCAN_TC_GRID")) grid = std::max(1, std::min(grid, atoi(e)));
    tc_blur_kernel<EPI, STATS><<<grid, NT, smem, ctx->stream>>>(L);
    DS_CHECK_LAUNCH(ctx);
    return DOCSCAN_OK;
}

}  // namespace

// Returns false when the tensor-core path does not apply (the caller then runs blur.cu); otherwise *rc is the result.
bool k_tc_blur_jobs(docscan_ctx* ctx, int kind, int k, int epi, const BlurJob* jobs_host, int n, int* rc) {
    *rc = DOCSCAN_OK;
    if (kind != 0 || k < 3 || n <= 0 || n > MAX_JOBS) return false;
    if (epi != DS_EPI_BLUR && epi != DS_EPI_SUB && epi != DS_EPI_RSUB && epi != DS_EPI_DIV) return false;
    if (const char* e = getenv("DOCSCAN_TC")) if (atoi(e) == 0) return false;
    if (!encode_fn()) return false;
    std::vector<int32_t> q(k);
    if (docscan_gaussian_kernel_q8(k, q.data()) != DOCSCAN_OK) return false;
    int z = 0;
    while (z < k / 2 && q[z] == 0) z++;                 // zero tails of the quantised kernel
    const int k_eff = k - 2 * z, R = k_eff / 2;
    if (R < 1 || R > MAX_R) return false;
    for (int i = 0; i < k_eff; i++) if (q[z + i] > 255) return false;
    const int K1 = (TM + 2 * R + 31) / 32 * 32;
    const int RL = (R + 15) & ~15;                      // the window's first column must sit on a 16-byte boundary of its row
    // the last two source slots of a tile carry the rounding constant (see the kernel), so the taps must end before them
    const int NOUT = std::min(96, (NIN - 2 - RL - R) / 16 * 16);
    if (NOUT < 16 || K1 > 256) return false;
    bool stats = false;
    for (int i = 0; i < n; i++) {
        const BlurJob& j = jobs_host[i];
        const int w16 = (j.w + 15) & ~15;
        if (((uintptr_t)j.src | (uintptr_t)j.dst | (uintptr_t)j.src_pitch | (uintptr_t)j.dst_pitch) & 15) return false;
        if (j.src_pitch < w16 || j.dst_pitch < w16 || j.w < 1 || j.h < 1) return false;
        stats = stats || j.minmax || j.hist;
    }
    // per-page geometry: tile counts, border variants of the two band matrices, tensor map of the source plane
    std::vector<TcJob> jobs(n);
    std::vector<const uint8_t*> tabs;
    std::vector<CUtensorMap> maps(n);
    std::map<std::pair<int, int>, std::pair<int, int>> geom;          // (w, h) -> (t_off, toe_off)
    int total = 0;
    double px = 0;
    for (int i = 0; i < n; i++) {
        const BlurJob& b = jobs_host[i];
        TcJob& j = jobs[i];
        j.src = b.src; j.dst = b.dst; j.src_pitch = b.src_pitch; j.dst_pitch = b.dst_pitch; j.w = b.w; j.h = b.h;
        j.minmax = b.minmax; j.hist = b.hist;
        j.ntx = (b.w + NOUT - 1) / NOUT; j.nty = (b.h + TM - 1) / TM;
        j.tile_base = total;
        total += j.ntx * j.nty;
        px += (double)b.w * b.h;
        auto g = geom.find({b.w, b.h});
        if (g == geom.end()) {
            const int t_off = (int)tabs.size();
            for (int ty = 0; ty < j.nty; ty++) {
                const uint8_t* p = nullptr; bool ok = false;
                *rc = get_variant(ctx, 0, k, q.data() + z, k_eff, R, RL, K1, NOUT, ty * TM, b.h, &p, &ok);
                if (*rc != DOCSCAN_OK) return true;
                if (!ok) return false;
                tabs.push_back(p);
            }
            const int toe_off = (int)tabs.size();
            for (int tx = 0; tx < j.ntx; tx++) {
                const uint8_t* p = nullptr; bool ok = false;
                *rc = get_variant(ctx, 1, k, q.data() + z, k_eff, R, RL, K1, NOUT, tx * NOUT, b.w, &p, &ok);
                if (*rc != DOCSCAN_OK) return true;
                if (!ok) return false;
                tabs.push_back(p);
            }
            g = geom.emplace(std::make_pair(b.w, b.h), std::make_pair(t_off, toe_off)).first;
        }
        j.t_off = g->second.first; j.toe_off = g->second.second;
        const cuuint64_t dims[2] = {(cuuint64_t)b.w, (cuuint64_t)b.h};
        const cuuint64_t strides[1] = {(cuuint64_t)b.src_pitch};
        const cuuint32_t box[2] = {(cuuint32_t)NIN, (cuuint32_t)K1};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult cr = encode_fn()(&maps[i], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)b.src, dims, strides, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) return false;
    }
    if (tabs.size() > (size_t)MAX_TABS) return false;
    TcLaunch L{};
    void* dev = nullptr;
    // one upload: tensor maps (64-byte aligned) | jobs | variant pointers
    const size_t off_jobs = sizeof(CUtensorMap) * n, off_tabs = off_jobs + ((sizeof(TcJob) * n + 63) & ~(size_t)63);
    std::vector<uint8_t> blob(off_tabs + sizeof(void*) * tabs.size());
    memcpy(blob.data(), maps.data(), sizeof(CUtensorMap) * n);
    memcpy(blob.data() + off_jobs, jobs.data(), sizeof(TcJob) * n);
    memcpy(blob.data() + off_tabs, tabs.data(), sizeof(void*) * tabs.size());
    *rc = ds_upload(ctx, blob.data(), blob.size(), &dev);
    if (*rc != DOCSCAN_OK) return true;
    L.maps = reinterpret_cast<const CUtensorMap*>(dev);
    L.jobs = reinterpret_cast<const TcJob*>((uint8_t*)dev + off_jobs);
    L.tabs = reinterpret_cast<const uint8_t* const*>((uint8_t*)dev + off_tabs);
    L.n_jobs = n; L.n_tabs = (int)tabs.size(); L.total_tiles = total; L.R = R; L.RL = RL; L.K1 = K1; L.NOUT = NOUT;
    L.idesc1 = tc::idesc_i8(TM, NIN, 0, 0, 0, 1);
    L.idesc2 = tc::idesc_i8(TM, NOUT, 0, 0, 0, 0);
    if (!ctx->tc_status) {
        if (cudaHostAlloc((void**)&ctx->tc_status, 64, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return false; }
        memset(ctx->tc_status, 0, 64);
    }
    L.status = ctx->tc_status;
    const char* dbg_path = getenv("DOCSCAN_TC_DEBUG");
    const size_t dbg_words = 16384 + 2 * 12288 + 64 * 32;      // first tile's accumulators + phase clocks of CTA 0's first 64 tiles
    if (dbg_path) {
        void* d = nullptr;
        *rc = ds_arena_alloc(ctx, dbg_words * 4, &d);
        if (*rc != DOCSCAN_OK) return true;
        cudaMemsetAsync(d, 0xEE, dbg_words * 4, ctx->stream);
        L.dbg = (uint32_t*)d;
        L.crumbs = getenv("DOCSCAN_TC_CRUMBS") != nullptr;
    }
    const size_t smem = 1024 + T_BYTES + TOE_BYTES + 2 * (size_t)K1 * 128 + (stats ? 8 * 256 * 4 : 0);
    {
        ProfScope prof(ctx, std::string("tc_blur_k") + std::to_string(k), 2.0 * px);
#define DS_TC_CASE(E) case E: *rc = stats ? launch_tc<E, true>(ctx, L, smem) : launch_tc<E, false>(ctx, L, smem); break;
        switch (epi) {
            DS_TC_CASE(DS_EPI_BLUR)
            DS_TC_CASE(DS_EPI_SUB)
            DS_TC_CASE(DS_EPI_RSUB)
            DS_TC_CASE(DS_EPI_DIV)
        }
#undef DS_TC_CASE
    }
    if (dbg_path && *rc == DOCSCAN_OK) {
        std::vector<uint32_t> host(dbg_words);
        cudaMemcpyAsync(host.data(), L.dbg, dbg_words * 4, cudaMemcpyDeviceToHost, ctx->stream);
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (FILE* f = fopen(dbg_path, "wb")) {
            fprintf(stderr, "[tc debug] k=%d k_eff=%d R=%d RL=%d K1=%d NOUT=%d tiles=%d sync=%s timeout_id=%u progress=%u\n", k, k_eff, R, RL, K1, NOUT, total,
                    cudaGetErrorString(e), ctx->tc_status[0], ctx->tc_status[1]);
            fwrite(host.data(), 4, dbg_words, f);
            fclose(f);
        }
    }
    return true;
}
