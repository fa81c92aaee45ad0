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
constexpr int N_EPI_WARPS = 16;                  // drain / epilogue warps; warp 16 issues TMA and MMAs
constexpr int NT = (N_EPI_WARPS + 1) * 32;
constexpr int TMEM_COLS = 512;
// TMEM columns.  Blur: D1 | A2lo A2hi | D2lo D2hi.  Adaptive threshold (16-bit weights = two byte planes, 8.8 row means = two byte
// planes): D1h D1l | A2hi A2lo | D2a D2b D2c (products hi*hi, hi*lo + lo*hi, lo*lo).  D2 never overlaps D1: pass 1 of the next tile
// runs under the epilogue of this one.
constexpr int COL_D1 = 0, COL_A2LO = 128, COL_A2HI = 160, COL_D2LO = 192, COL_D2HI = 272, D2_STRIDE = 160;   // two pairs of D2 (NOUT <= 80)
constexpr int COLA_D1H = 0, COLA_D1L = 128, COLA_A2HI = 256, COLA_A2LO = 288, COLA_D2A = 320, COLA_D2B = 384, COLA_D2C = 448;
constexpr int NS = 3;                            // source-window stages: pass 1 of tile i+1 and the centre pixels of tile i are live together
constexpr int TOE_SLOTS = 3;                     // cached pass-2 band matrices (left / interior / right)
constexpr int T_BYTES = 2 * TM * 128;      // two 128-byte K blocks
constexpr int TOE_BYTES = 96 * 128;
constexpr int MAX_R = 48;                 // largest radius of the 128-column tile
constexpr int MAX_R_WIDE = 96;            // ... of the 256-column (wide) tile: 128 + 2 R source rows <= 320
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
    const CUtensorMap* maps;      // source planes (128-byte swizzle, box 128 x K1)
    const CUtensorMap* dmaps;     // destination planes (no swizzle, box NOUT x 128)
    const uint8_t* const* tabs;
    uint32_t* dbg;               // debug dump of the first tile (DOCSCAN_TC_DEBUG), else null
    volatile uint32_t* status;   // pinned host words: [0] = which wait timed out, [1] = progress of CTA 0 (debug runs)
    int n_jobs, n_tabs, total_tiles;
    // adaptive threshold (EPI == DS_EPI_AGAUSS)
    int c_param;                 // dst = src - mean > -c_param ? 255 : 0
    int band;                    // guard band in 1/65536 grey levels: inside it the pixel goes to the exact evaluation
    uint16_t* flag_list; uint32_t* flag_count; uint32_t flag_cap;   // per tile: count, then up to flag_cap entries (row << 6 | column)
    int flags;                   // debug: skip parts of the epilogue (DOCSCAN_TC_FLAGS), for timing experiments only
    int crumbs;                  // debug: CTA 0 reports its progress to status[1] (slow: a system-scope fence per phase)
    int t_slots;                 // cached pass-1 band matrices (top / interior / bottom): 3 when shared memory allows, else 2 or 1
    int toe_slots;               // cached pass-2 band matrices (left / interior / right): 3, or 2 for the wide tile
    int ns;                      // source-window stages: 3; wide tile (64..80 KB per window): 2 or 1
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

// One decoded tile: the issuing warp fills a ring of these 32 at a time (one lane per tile), so that nobody divides or walks
// the page table on the critical path.
struct TileRec { int job, tx, ty, pad; const uint8_t* t_mat; const uint8_t* toe_mat; };
constexpr int REC_RING = 64;

// One CTA per SM: 16 drain / epilogue warps (warp w: TMEM lane quarter w & 3, column group w >> 2) and one issuing warp that
// runs the TMA loads and both tensor-core passes one tile ahead, so that pass 1 of tile i+1 executes behind the epilogue of
// tile i.  No CTA-wide barrier inside the loop: the roles meet on mbarriers only.
//   bar_s[st]  source window of a tile has landed in stage st          (TMA transaction bytes)        issuer waits
//   bar_c      band matrices have landed                                (bulk-copy transaction bytes)  issuer waits
//   bar_d1     pass 1 of tile i complete: D1 readable                   (tcgen05.commit)               drain warps wait
//   bar_a2     all 16 warps have written their part of A2 for tile i (and so are done with D2 and the centre pixels of tile
//              i-1 and with D1 of tile i)                               (16 arrivals)                  issuer waits
//   bar_d2     pass 2 of tile i complete: D2lo / D2hi readable          (tcgen05.commit)               epilogue warps wait
// WIDE: the tile variant for radii 49..93 (8K scans, k = 101..217): 256 input columns per tile (two TMA boxes side by side = two
// MN blocks of the B operand of pass 1, two K blocks in pass 2), up to 320 source rows (three K blocks of the band matrix), 64
// output columns, one source stage, no lag between drain and epilogue (tensor memory: D1 256 | A2 128 | D2 128).
template <int EPI, bool STATS, bool DBG, bool WIDE>
__global__ void __launch_bounds__(NT, 1) tc_blur_kernel(const __grid_constant__ TcLaunch L) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte boundary (128-byte swizzle atoms), found as an OFFSET from the shared-space address: a pointer rounded through an
    // integer cast becomes generic, and the staging stores, centre-pixel loads and histogram increments then compile to
    // ST / LD / ATOM instead of STS / LDS / ATOMS
    uint8_t* base = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sT = base;                                            // L.t_slots band matrices of pass 1
    constexpr bool ADAPT = EPI == DS_EPI_AGAUSS;
    static_assert(!(ADAPT && WIDE), "no wide adaptive instance");
    constexpr bool DEFER = !ADAPT && !WIDE;                        // two pairs of pass-2 accumulators: the epilogue lags one tile
    constexpr int NINK = WIDE ? 256 : 128;                         // input columns per tile
    constexpr int DCOLS = NINK / 4;                                // D1 columns each of the four column groups drains
    // adaptive: two byte planes; wide: two or three 128-byte K blocks
    const uint32_t t_bytes = ADAPT ? 2 * T_BYTES : (WIDE ? (uint32_t)((L.K1 + 127) >> 7) * TM * 128 : (uint32_t)T_BYTES);
    constexpr int C_A2LO = ADAPT ? COLA_A2LO : (WIDE ? 256 : COL_A2LO), C_A2HI = ADAPT ? COLA_A2HI : (WIDE ? 320 : COL_A2HI);
    constexpr int C_D2LO = WIDE ? 384 : COL_D2LO, C_D2HI = WIDE ? 448 : COL_D2HI;
    uint8_t* sToe = sT + L.t_slots * t_bytes;                      // band matrices of pass 2
    const uint32_t toe_bytes = (uint32_t)L.NOUT * 128 * ((ADAPT || WIDE) ? 2 : 1);
    uint8_t* sS = sToe + L.toe_slots * toe_bytes;                  // source windows
    const uint32_t s_bytes_all = (uint32_t)L.K1 * NINK;
    uint8_t* s_out = sS + L.ns * s_bytes_all;                         // the tile's results, 128 dense rows of NOUT bytes: the source of the TMA store
    const int out_pitch = L.NOUT;
    uint32_t* s_hist = reinterpret_cast<uint32_t*>(s_out + TM * out_pitch);   // 8 x 256, only with STATS
    __shared__ uint64_t bar_s[NS], bar_c, bar_d1, bar_d2[2], bar_a2;
    __shared__ uint32_t s_tmem;
    __shared__ TcJob s_jobs[MAX_JOBS];                   // the launch's page table and band-matrix pointers, read every tile
    __shared__ const uint8_t* s_tabs[MAX_TABS];
    __shared__ TileRec s_rec[REC_RING];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < NS; i++) tc::mbar_init(&bar_s[i], 1);
        tc::mbar_init(&bar_c, 1); tc::mbar_init(&bar_d1, 1); tc::mbar_init(&bar_d2[0], 1); tc::mbar_init(&bar_d2[1], 1);
        tc::mbar_init(&bar_a2, N_EPI_WARPS);
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
#define TC_CRUMB(v) do { if (DBG && L.crumbs && blockIdx.x == 0) { L.status[1] = (v); __threadfence_system(); } } while (0)
#define TC_STAMP(who, i, k) do { if (DBG && L.dbg && blockIdx.x == 0 && (i) < 64) { long long c_ = clock64(); uint32_t* d_ = L.dbg + 40960 + (who) * 2048 + (i) * 32 + 2 * (k); d_[0] = (uint32_t)c_; d_[1] = (uint32_t)(c_ >> 32); } } while (0)
#define TC_WAIT(bar, par, id) do { if (!tc::mbar_wait_bounded(bar, par)) { L.status[0] = (id); __threadfence_system(); __trap(); } } while (0)

    // this CTA's tiles: a contiguous run of the launch's tile sequence (pages change once or twice per CTA, not every tile:
    // a page change costs a tensor-map acquire and a flush of the page statistics)
    const int per_cta = L.total_tiles / (int)gridDim.x, extra = L.total_tiles % (int)gridDim.x;
    const int n_mine = per_cta + ((int)blockIdx.x < extra ? 1 : 0);
    const int tile0 = (int)blockIdx.x * per_cta + min((int)blockIdx.x, extra);
    if (warp == N_EPI_WARPS) {
        // ================================================ issuing warp ================================================
        // tile i of this CTA is global tile tile0 + i; tiles of a page are numbered row-major, so consecutive tiles share
        // their halo columns through L2.  Lane l decodes tile i0 + l into the ring (32 tiles per refill).
        int jc = 0;                                                // this lane's page cursor (its tiles only move forward)
        auto refill = [&](int i0) {
            const int i = i0 + lane;
            if (i < n_mine) {
                const int t = tile0 + i;
                while (jc + 1 < L.n_jobs && t >= s_jobs[jc + 1].tile_base) jc++;
                const TcJob& J = s_jobs[jc];
                const int idx = t - J.tile_base;
                TileRec r;
                r.job = jc; r.ty = idx / J.ntx; r.tx = idx - r.ty * J.ntx; r.pad = 0;
                r.t_mat = s_tabs[J.t_off + r.ty]; r.toe_mat = s_tabs[J.toe_off + r.tx];
                s_rec[i % REC_RING] = r;
            }
            __threadfence_block();
            __syncwarp();
        };
        refill(0);
        refill(32);
        asm volatile("bar.arrive 2, %0;" ::"n"(NT) : "memory");       // the first records are there (the epilogue warps wait on this)

        // elected-lane state, all in registers
        const uint8_t* t_tag0 = nullptr; const uint8_t* t_tag1 = nullptr; const uint8_t* t_tag2 = nullptr;
        const uint8_t* e_tag0 = nullptr; const uint8_t* e_tag1 = nullptr; const uint8_t* e_tag2 = nullptr;
        int t_rr = 0, e_rr = 0;                                    // round-robin replacement
        int p1_issued = 0;                                         // pass 1 launched for tiles 0 .. p1_issued - 1
        uint32_t ph_c = 0, ph_s0 = 0, ph_s1 = 0, ph_s2 = 0;
        int last_map = -1;
        const uint32_t aT = tc::smem_u32(sT), aToe = tc::smem_u32(sToe), aS = tc::smem_u32(sS);
        const uint32_t s_bytes = s_bytes_all;                     // one source window; LBO of its descriptor = one 128-column block
        const uint64_t dT0 = tc::smem_desc_sw128(aT, 16, 1024), dToe0 = tc::smem_desc_sw128(aToe, 16, 1024);
        const uint64_t dS0 = tc::smem_desc_sw128(aS, (uint32_t)L.K1 * 128, 1024);

        // pass-1 band matrix `want` into a slot (returned); `keep` = slot an unfinished MMA may still be reading
        auto ensure_t = [&](const uint8_t* want, int keep) {
            if (want == t_tag0) return 0;
            if (want == t_tag1) return 1;
            if (L.t_slots > 2 && want == t_tag2) return 2;
            int v = t_rr;
            if (L.t_slots == 1) {
                // a single slot (adaptive threshold: 64 KB per matrix pair): the pass 1 that may still be reading it must finish first
                v = 0;
                if (p1_issued > 0) TC_WAIT(&bar_d1, (p1_issued - 1) & 1, 6);
            } else {
                if (v == keep) v = (v + 1 == L.t_slots) ? 0 : v + 1;
                t_rr = (v + 1 == L.t_slots) ? 0 : v + 1;
            }
            tc::mbar_expect_tx(&bar_c, t_bytes);
            tc::bulk_load(sT + (size_t)v * t_bytes, want, t_bytes, &bar_c);
            TC_WAIT(&bar_c, ph_c, 1); ph_c ^= 1;
            if (v == 0) t_tag0 = want; else if (v == 1) t_tag1 = want; else t_tag2 = want;
            return v;
        };
        auto ensure_toe = [&](const uint8_t* want, int keep) {
            if (want == e_tag0) return 0;
            if (want == e_tag1) return 1;
            if (L.toe_slots > 2 && want == e_tag2) return 2;
            int v = e_rr;
            if (v == keep) v = (v + 1 == L.toe_slots) ? 0 : v + 1;
            e_rr = (v + 1 == L.toe_slots) ? 0 : v + 1;
            tc::mbar_expect_tx(&bar_c, toe_bytes);
            tc::bulk_load(sToe + (size_t)v * toe_bytes, want, toe_bytes, &bar_c);
            TC_WAIT(&bar_c, ph_c, 1); ph_c ^= 1;
            if (v == 0) e_tag0 = want; else if (v == 1) e_tag1 = want; else e_tag2 = want;
            return v;
        };
        auto load_source = [&](int i) {
            const TileRec r = s_rec[i % REC_RING];
            if (r.job != last_map) { tc::tmap_acquire(&L.maps[r.job]); last_map = r.job; }
            const int st = i % L.ns;
            uint64_t* bar = &bar_s[st];
            tc::mbar_expect_tx(bar, s_bytes);
            if (!WIDE) {
                tc::tma_load_2d(sS + (size_t)st * s_bytes, &L.maps[r.job], r.tx * L.NOUT - L.RL, r.ty * TM - L.R, bar);
            } else {
                // two 128-column blocks side by side, each fetched as two boxes of K1 / 2 rows (a TMA box has at most 256 rows)
                const int half = L.K1 >> 1;
#pragma unroll
                for (int cb = 0; cb < 2; cb++)
#pragma unroll
                    for (int rb = 0; rb < 2; rb++)
                        tc::tma_load_2d(sS + (size_t)st * s_bytes + (size_t)cb * L.K1 * 128 + (size_t)rb * half * 128, &L.maps[r.job],
                                        r.tx * L.NOUT - L.RL + cb * 128, r.ty * TM - L.R + rb * half, bar);
            }
        };
        auto pass1 = [&](int i, int t_slot) {
            const int st = i % L.ns;
            if (st == 0) { TC_WAIT(&bar_s[0], ph_s0, 2); ph_s0 ^= 1; }
            else if (st == 1) { TC_WAIT(&bar_s[1], ph_s1, 2); ph_s1 ^= 1; }
            else { TC_WAIT(&bar_s[2], ph_s2, 2); ph_s2 ^= 1; }
            tc::fence_after_sync();
            const uint64_t dT = dT0 + (uint64_t)((t_slot * t_bytes) >> 4);
            const uint64_t dS = dS0 + (uint64_t)((st * s_bytes) >> 4);
            switch (L.K1 >> 5) {
                case 5: issue_pass1<5>(tmem + COL_D1, dT, dS, L.idesc1); break;
                case 6: issue_pass1<6>(tmem + COL_D1, dT, dS, L.idesc1); break;
                case 7: issue_pass1<7>(tmem + COL_D1, dT, dS, L.idesc1); break;
                case 8: issue_pass1<8>(tmem + COL_D1, dT, dS, L.idesc1); break;
                case 9: issue_pass1<9>(tmem + COL_D1, dT, dS, L.idesc1); break;
                default: issue_pass1<10>(tmem + COL_D1, dT, dS, L.idesc1); break;
            }
            if (ADAPT) {                                           // the low-byte plane of the weights into the second accumulator
                const uint64_t dTl = dT + (uint64_t)(T_BYTES >> 4);
                switch (L.K1 >> 5) {
                    case 5: issue_pass1<5>(tmem + COLA_D1L, dTl, dS, L.idesc1); break;
                    case 6: issue_pass1<6>(tmem + COLA_D1L, dTl, dS, L.idesc1); break;
                    case 7: issue_pass1<7>(tmem + COLA_D1L, dTl, dS, L.idesc1); break;
                    default: issue_pass1<8>(tmem + COLA_D1L, dTl, dS, L.idesc1); break;
                }
            }
            tc::mma_commit(&bar_d1);
            p1_issued = i + 1;
        };

        int t_slot = 0, toe_prev = -1;
        if (n_mine > 0 && tc::elect_one()) {
            load_source(0);
            if (n_mine > 1 && L.ns > 1) load_source(1);
            t_slot = ensure_t(s_rec[0].t_mat, -1);
            pass1(0, t_slot);
            TC_CRUMB(2);
        }
        __syncwarp();
        for (int i = 0; i < n_mine; i++) {
            if ((i & 31) == 0 && i > 0) refill(i + 32);          // whole warp: tiles i+32 .. i+63 (their slots were read 32 rounds ago)
            if (tc::elect_one()) {
                TC_STAMP(1, i, 0);
                // matrices this round needs: pass 2 of tile i, pass 1 of tile i+1 (fetched, if missing, while the warps drain D1)
                const int toe_slot = ensure_toe(s_rec[i % REC_RING].toe_mat, toe_prev);
                int t_next = t_slot;
                if (i + 1 < n_mine) t_next = ensure_t(s_rec[(i + 1) % REC_RING].t_mat, t_slot);
                TC_STAMP(1, i, 1);
                TC_WAIT(&bar_a2, i & 1, 5);
                tc::fence_after_sync();
                TC_STAMP(1, i, 2);
                // pass 2: D2lo / D2hi [128 x NOUT] = A2lo / A2hi [128 x 128] (tensor memory) * Th[128 x NOUT]
                const uint64_t dToe = dToe0 + (uint64_t)((toe_slot * toe_bytes) >> 4);
                const int b2 = DEFER ? (i & 1) : 0;                // the pair of accumulators tile i - 2 has left
                if (!ADAPT) {
                    const uint32_t d2 = tmem + b2 * D2_STRIDE;
                    const uint32_t kblk = ((uint32_t)L.NOUT * 128) >> 4;          // next 128-byte K block of the band matrix (wide tile)
#pragma unroll
                    for (int s2 = 0; s2 < NINK / 32; s2++)
                        tc::mma_i8_ts(d2 + C_D2LO, tmem + C_A2LO + s2 * 8, dToe + (uint64_t)((s2 >> 2) * kblk + (s2 & 3) * 2), L.idesc2, s2 > 0);
#pragma unroll
                    for (int s2 = 0; s2 < NINK / 32; s2++)
                        tc::mma_i8_ts(d2 + C_D2HI, tmem + C_A2HI + s2 * 8, dToe + (uint64_t)((s2 >> 2) * kblk + (s2 & 3) * 2), L.idesc2, s2 > 0);
                } else {
                    // 16-bit row means x 16-bit weights as byte planes: D2a = hi * Whi, D2b = hi * Wlo + lo * Whi, D2c = lo * Wlo
                    const uint64_t dWl = dToe + (uint64_t)(((uint32_t)L.NOUT * 128) >> 4);
#pragma unroll
                    for (int s2 = 0; s2 < NIN / 32; s2++) tc::mma_i8_ts(tmem + COLA_D2A, tmem + COLA_A2HI + s2 * 8, dToe + (uint64_t)(s2 * 2), L.idesc2, s2 > 0);
#pragma unroll
                    for (int s2 = 0; s2 < NIN / 32; s2++) tc::mma_i8_ts(tmem + COLA_D2B, tmem + COLA_A2HI + s2 * 8, dWl + (uint64_t)(s2 * 2), L.idesc2, s2 > 0);
#pragma unroll
                    for (int s2 = 0; s2 < NIN / 32; s2++) tc::mma_i8_ts(tmem + COLA_D2B, tmem + COLA_A2LO + s2 * 8, dToe + (uint64_t)(s2 * 2), L.idesc2, 1);
#pragma unroll
                    for (int s2 = 0; s2 < NIN / 32; s2++) tc::mma_i8_ts(tmem + COLA_D2C, tmem + COLA_A2LO + s2 * 8, dWl + (uint64_t)(s2 * 2), L.idesc2, s2 > 0);
                }
                tc::mma_commit(&bar_d2[b2]);
                TC_STAMP(1, i, 3);
                // a single source stage (wide tile): it is free now (pass 1 and the centre pixels of tile i are done: bar_a2)
                if (L.ns == 1 && i + 1 < n_mine) load_source(i + 1);
                if (i + 1 < n_mine) pass1(i + 1, t_next);       // runs behind pass 2 of tile i, under its epilogue
                TC_STAMP(1, i, 4);
                // three stages: stage (i + 2) % 3 held tile i - 1, whose epilogue is over (bar_a2 of tile i): refill it
                if (L.ns > 1 && i + 2 < n_mine) load_source(i + 2);
                TC_STAMP(1, i, 5);
                toe_prev = toe_slot; t_slot = t_next;
                TC_CRUMB(3);
            }
            __syncwarp();
        }
    } else {
        // ============================================= drain / epilogue warps =============================================
        const int q = warp & 3, cg = warp >> 2;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;          // this warp's quarter of the TMEM lanes
        const int row = q * 32 + lane;                                // tile row of this thread
        // columns of the epilogue in units of 8, dealt to the four column groups as evenly as possible
        const int units = L.NOUT >> 3;
        const int u_begin = (units * cg) >> 2, u_end = (units * (cg + 1)) >> 2;
        uint32_t mn2 = 0x00FF00FFu, mx2 = 0;                          // running min / max of the current page, two 16-bit lanes
        uint32_t* st_minmax = nullptr; uint32_t* st_hist = nullptr;
        uint32_t* my_hist = s_hist + (warp & 7) * 256;
        int last_dmap = -1;
        asm volatile("bar.sync 2, %0;" ::"n"(NT) : "memory");            // the issuing warp has decoded the first tiles

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
                asm volatile("bar.sync 1, %0;" ::"n"(N_EPI_WARPS * 32) : "memory");     // the 16 epilogue warps only
                for (int i = tid; i < 256; i += N_EPI_WARPS * 32) {
                    uint32_t sum = 0;
#pragma unroll
                    for (int wv = 0; wv < 8; wv++) { sum += s_hist[wv * 256 + i]; s_hist[wv * 256 + i] = 0; }
                    if (sum) atomicAdd(&st_hist[i], sum);
                }
                asm volatile("bar.sync 1, %0;" ::"n"(N_EPI_WARPS * 32) : "memory");
            }
            mn2 = 0x00FF00FFu; mx2 = 0;
        };

        // The epilogue of a tile runs one round after its drain (blur; the adaptive threshold has no room for a second set of
        // pass-2 accumulators): round i drains D1 of tile i, hands A2 to the issuing warp, and then — while pass 2 of tile i and
        // pass 1 of tile i+1 execute — finishes tile i-1 out of the other pair of accumulators.  Nobody waits for an MMA.
        constexpr int LAG = DEFER ? 1 : 0;
        uint2 cen_cur[3], cen_new[3];                              // centre pixels of this thread's units: tile being finished / drained
        int tr_job_cur = 0, tr_tx_cur = 0, tr_ty_cur = 0, tr_job_new = 0, tr_tx_new = 0, tr_ty_new = 0;   // likewise its decoded record
#pragma unroll
        for (int k = 0; k < 3; k++) { cen_cur[k] = make_uint2(0, 0); cen_new[k] = make_uint2(0, 0); }
        for (int i = 0; i < n_mine + LAG; i++) {
            if (i < n_mine) {
                // the tile's record now: by the time the tile is finished (next round) the issuing warp may have reused the ring slot
                tr_job_new = s_rec[i % REC_RING].job; tr_tx_new = s_rec[i % REC_RING].tx; tr_ty_new = s_rec[i % REC_RING].ty;
                if (tid == 0) TC_STAMP(0, i, 0);
                // ---- D1 -> byte planes (A operands of pass 2): this warp's 32 rows x 32 columns
                TC_WAIT(&bar_d1, i & 1, 3);
                tc::fence_after_sync();
                if (tid == 0) TC_STAMP(0, i, 1);
                {
                    // DCOLS = 32 (64 for the wide tile) columns of D1 per warp, 32 at a time
                    uint32_t lo[DCOLS / 4], hi[DCOLS / 4];
#pragma unroll
                    for (int part = 0; part < DCOLS / 32; part++) {
                        uint32_t v[32];
                        tc::tmem_ld32(tmem + lane_base + COL_D1 + cg * DCOLS + part * 32, v);
                        if (ADAPT) {
                            // row mean * 65536 = 256 * D1h + D1l (24 bits)  ->  8.8 fixed point, rounded
                            uint32_t vl[32];
                            tc::tmem_ld32(tmem + lane_base + COLA_D1L + cg * 32, vl);
                            tc::tmem_wait_ld();
#pragma unroll
                            for (int k = 0; k < 32; k++) v[k] = ((v[k] << 8) + vl[k] + 128u) >> 8;
                        }
                        tc::tmem_wait_ld();
                        if (DBG && !WIDE && L.dbg && blockIdx.x == 0 && i == 0)
                            for (int k = 0; k < 32; k++) L.dbg[row * 128 + cg * 32 + k] = v[k];
#pragma unroll
                        for (int g = 0; g < 8; g++) {
                            const uint32_t t1 = __byte_perm(v[4 * g], v[4 * g + 1], 0x5140);        // a0 b0 a1 b1
                            const uint32_t t2 = __byte_perm(v[4 * g + 2], v[4 * g + 3], 0x5140);    // c0 d0 c1 d1
                            lo[part * 8 + g] = __byte_perm(t1, t2, 0x5410);                         // a0 b0 c0 d0
                            hi[part * 8 + g] = __byte_perm(t1, t2, 0x7632);                         // a1 b1 c1 d1
                        }
                    }
                    if (!ADAPT && cg == 3) {
                        // the last two columns (126, 127; wide: 254, 255) carry no tap: they hold the rounding constant instead,
                        // 2 x (128 * 128) = 32768, against the two 128s in the band matrix's last two slots
                        lo[DCOLS / 4 - 1] = (lo[DCOLS / 4 - 1] & 0x0000FFFFu) | 0x80800000u;
                        hi[DCOLS / 4 - 1] &= 0x0000FFFFu;
                    }
                    if (DEFER && i > 0) {
                        // pass 2 of the previous tile reads A2: it must be over before A2 is rewritten (it usually finished long ago,
                        // under the epilogue this warp has just run)
                        TC_WAIT(&bar_d2[(i - 1) & 1], ((i - 1) >> 1) & 1, 7);
                        tc::fence_after_sync();
                    }
                    if (WIDE) {
                        tc::tmem_st16(tmem + lane_base + C_A2LO + cg * 16, lo);
                        tc::tmem_st16(tmem + lane_base + C_A2HI + cg * 16, hi);
                    } else {
                        tc::tmem_st8(tmem + lane_base + C_A2LO + cg * 8, lo);
                        tc::tmem_st8(tmem + lane_base + C_A2HI + cg * 8, hi);
                    }
                    tc::tmem_wait_st();
                }
                if (EPI != DS_EPI_BLUR && !(DBG && (L.flags & 2))) {
                    // this thread's centre pixels, out of the source window while it is still there: row `row + R`, byte RL + column,
                    // 16-byte chunks swizzled by the row number
                    const int srow_i = row + L.R, swz = srow_i & 7;
                    const uint8_t* s_center = sS + (size_t)(i % L.ns) * s_bytes_all + srow_i * 128;
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        const int cb = L.RL + (u_begin + k) * 8;           // byte in the window row; wide: its 128-column block first
                        const uint8_t* blk = s_center + (size_t)(cb >> 7) * L.K1 * 128;
                        const int cbb = cb & 127;
                        if (u_begin + k < u_end) cen_new[k] = *reinterpret_cast<const uint2*>(blk + ((((cbb >> 4) ^ swz) << 4) | (cbb & 8)));
                    }
                }
                tc::fence_before_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&bar_a2);
                if (tid == 0) TC_STAMP(0, i, 2);

            }
            if (LAG == 0) {
#pragma unroll
                for (int k = 0; k < 3; k++) cen_cur[k] = cen_new[k];
                tr_job_cur = tr_job_new; tr_tx_cur = tr_tx_new; tr_ty_cur = tr_ty_new;
            }
            if (i >= LAG) {
                const int e = i - LAG;                             // the tile this round finishes
                TileRec tr; tr.job = tr_job_cur; tr.tx = tr_tx_cur; tr.ty = tr_ty_cur;
                // the page's fields into registers: the shared-memory stores below would otherwise force re-reads of the table
                const int Jw = s_jobs[tr.job].w, Jh = s_jobs[tr.job].h, Jdp = s_jobs[tr.job].dst_pitch;
                uint8_t* const Jdst = s_jobs[tr.job].dst;
                uint32_t* const Jminmax = s_jobs[tr.job].minmax; uint32_t* const Jhist = s_jobs[tr.job].hist;
                if (STATS && (st_minmax != Jminmax || st_hist != Jhist)) {   // statistics are per page
                    if (e > 0) flush_stats();
                    st_minmax = Jminmax; st_hist = Jhist;
                }
                const int x0 = tr.tx * L.NOUT, y0 = tr.ty * TM;
                const bool first = DBG && L.dbg && blockIdx.x == 0 && e == 0;

                // ---- epilogue: 32 rows x this column group's units of 8 columns, the next unit's accumulators in flight
                const int y = y0 + row;
                const bool row_ok = y < Jh;
                uint32_t hc_ev = 0, hc_od = 0;                    // this tile's counts of the values 0..7 (8 bits each: even / odd bins)
                const int b2 = DEFER ? (e & 1) : 0;                 // which pair of accumulators
                TC_WAIT(&bar_d2[b2], (DEFER ? (e >> 1) : e) & 1, 4);
                tc::fence_after_sync();
                if (tid == 0) TC_STAMP(0, e, 3);
                // one unit of 8 columns: combine the two accumulators, apply the epilogue, store, update the statistics
                auto work = [&](int u, const uint32_t* lo, const uint32_t* hi, const uint2 cw) {
                    const int c = u * 8, x = x0 + c;
                    if (first)
                        for (int k = 0; k < 8; k++) {
                            L.dbg[16384 + row * 96 + c + k] = lo[k];
                            L.dbg[16384 + 12288 + row * 96 + c + k] = hi[k];
                        }
                    if (row_ok && x < Jw) {
                        const uint32_t cws[2] = {cw.x, cw.y};
                        const int nvalid = min(8, Jw - x);
                        uint32_t out[2];
                        uint32_t dl[4];                         // results as 16-bit lanes: dl[2g] = (px 4g, px 4g+2), dl[2g+1] = (px 4g+1, px 4g+3)
#pragma unroll
                        for (int g = 0; g < 2; g++) {
                            // blurred byte = bits 16..23 of D2lo + 256 * D2hi (the rounding constant is already in the sum; bits 24.. are 0)
                            const uint32_t e0 = lo[4 * g] + (hi[4 * g] << 8), e1 = lo[4 * g + 1] + (hi[4 * g + 1] << 8);
                            const uint32_t e2 = lo[4 * g + 2] + (hi[4 * g + 2] << 8), e3 = lo[4 * g + 3] + (hi[4 * g + 3] << 8);
                            const uint32_t b_ev = __byte_perm(e0, e2, 0x7632), b_od = __byte_perm(e1, e3, 0x7632);
                            uint32_t d_ev = b_ev, d_od = b_od;
                            if (EPI != DS_EPI_BLUR) {
                                const uint32_t s_ev = __byte_perm(cws[g], 0u, 0x4240), s_od = __byte_perm(cws[g], 0u, 0x4341);
                                if (EPI == DS_EPI_SUB) {        // sat(s - b) = max(s, b) - b, lane-wise without borrows
                                    d_ev = __vmaxu2(s_ev, b_ev) - b_ev; d_od = __vmaxu2(s_od, b_od) - b_od;
                                } else if (EPI == DS_EPI_RSUB) {
                                    d_ev = __vmaxu2(s_ev, b_ev) - s_ev; d_od = __vmaxu2(s_od, b_od) - s_od;
                                } else {                        // divide(s, b, 255) in fp32, per pixel
                                    d_ev = (uint32_t)ds_div255((uint8_t)s_ev, (uint8_t)b_ev) | ((uint32_t)ds_div255((uint8_t)(s_ev >> 16), (uint8_t)(b_ev >> 16)) << 16);
                                    d_od = (uint32_t)ds_div255((uint8_t)s_od, (uint8_t)b_od) | ((uint32_t)ds_div255((uint8_t)(s_od >> 16), (uint8_t)(b_od >> 16)) << 16);
                                }
                            }
                            dl[2 * g] = d_ev; dl[2 * g + 1] = d_od;
                            out[g] = __byte_perm(d_ev, d_od, 0x6240);
                        }
                        if (!(DBG && (L.flags & 8))) *reinterpret_cast<uint2*>(s_out + row * out_pitch + c) = make_uint2(out[0], out[1]);
                        if (STATS) {
                            if (nvalid < 8) {                   // rare: keep the columns past the width out of the statistics
                                for (int k = 0; k < nvalid; k++) {
                                    const uint32_t v = (out[k >> 2] >> (8 * (k & 3))) & 0xFFu;
                                    if (Jminmax) { mn2 = __vminu2(mn2, v | 0x00FF0000u); mx2 = __vmaxu2(mx2, v); }
                                    if (Jhist) atomicAdd(&my_hist[v], 1u);
                                }
                            } else {
                                if (Jminmax && !(DBG && (L.flags & 1))) {
#pragma unroll
                                    for (int k = 0; k < 4; k++) { mn2 = __vminu2(mn2, dl[k]); mx2 = __vmaxu2(mx2, dl[k]); }
                                }
                                if (Jhist && !(DBG && (L.flags & 4))) {
                                    // values 0..7: 4-bit counters in two registers (a 32-bit shift by 32 or more gives 0, so larger
                                    // values add nothing here); at most 4 increments per field and unit
                                    uint32_t h0 = 0, h1 = 0;
#pragma unroll
                                    for (int k = 0; k < 4; k++) {
                                        const uint32_t d = dl[k];
                                        h0 += tc::shl32(1u, (d << 2) & 0x3FCu);
                                        h1 += tc::shl32(1u, (d >> 14) & 0x3FCu);
                                        if (d & 0x00F800F8u) {  // a value of 8 or more: the shared-memory histogram
                                            const uint32_t v0 = d & 0xFFu, v1 = d >> 16;
                                            if (v0 >= 8) atomicAdd(&my_hist[v0], 1u);
                                            if (v1 >= 8) atomicAdd(&my_hist[v1], 1u);
                                        }
                                    }
                                    hc_ev += (h0 & 0x0F0F0F0Fu) + (h1 & 0x0F0F0F0Fu);               // bins 0, 2, 4, 6 (8 bits each)
                                    hc_od += ((h0 >> 4) & 0x0F0F0F0Fu) + ((h1 >> 4) & 0x0F0F0F0Fu);  // bins 1, 3, 5, 7
                                }
                            }
                        }
                    }
                };
                if (tid == 0) tc::tma_store_wait_read();
                asm volatile("bar.sync 3, %0;" ::"n"(N_EPI_WARPS * 32) : "memory");      // the previous tile has left the staging buffer
                if constexpr (ADAPT) {
                    // ---- adaptive threshold: mean * 65536 = 256 * D2a + D2b + D2c / 256 against (src + C - 0.5) * 65536
                    uint32_t A[16], B[16], Cc[16];
                    tc::tmem_ld16(tmem + lane_base + COLA_D2A + u_begin * 8, A);
                    tc::tmem_ld16(tmem + lane_base + COLA_D2B + u_begin * 8, B);
                    tc::tmem_ld16(tmem + lane_base + COLA_D2C + u_begin * 8, Cc);
                    tc::tmem_wait_ld();
#pragma unroll
                    for (int k2 = 0; k2 < 2; k2++) {
                        const int u = u_begin + k2;
                        if (u >= u_end) break;
                        const int c = u * 8, x = x0 + c;
                        const uint2 cw = cen_cur[k2];
                        if (first)
                            for (int k = 0; k < 8; k++) {
                                L.dbg[16384 + row * 96 + c + k] = A[8 * k2 + k];
                                L.dbg[16384 + 12288 + row * 96 + c + k] = B[8 * k2 + k];
                            }
                        if (row_ok && x < Jw) {
                            const uint32_t cws[2] = {cw.x, cw.y};
                            uint32_t out[2] = {0, 0};
                            uint32_t flagged = 0;
#pragma unroll
                            for (int p8 = 0; p8 < 8; p8++) {
                                const int sp = (int)((cws[p8 >> 2] >> (8 * (p8 & 3))) & 0xFFu);
                                const int V = (int)((A[8 * k2 + p8] << 8) + B[8 * k2 + p8] + (Cc[8 * k2 + p8] >> 8));
                                const int diff = V - (((sp + L.c_param) << 16) - 32768);      // < 0: mean below the boundary -> 255
                                if (diff < 0) out[p8 >> 2] |= 0xFFu << (8 * (p8 & 3));
                                if (abs(diff) <= L.band && x + p8 < Jw) flagged |= 1u << p8;
                            }
                            *reinterpret_cast<uint2*>(s_out + row * out_pitch + c) = make_uint2(out[0], out[1]);
                            while (flagged) {                   // rare: too close to the boundary for the fixed-point mean to decide
                                const int p8 = __ffs(flagged) - 1;
                                flagged &= flagged - 1;
                                const uint32_t slot = atomicAdd(&L.flag_count[tile0 + e], 1u);
                                if (slot < L.flag_cap) L.flag_list[(size_t)(tile0 + e) * L.flag_cap + slot] = (uint16_t)((row << 6) | (c + p8));
                            }
                        }
                    }
                    __syncwarp();
                } else {
                // both accumulators of this warp's columns in two wide loads (a TMEM load costs about the same whatever its width;
                // the columns past this group's share belong to a neighbour or to nobody and are ignored)
                uint32_t lo[24], hi[24];
                {
                    const uint32_t a_lo = tmem + lane_base + C_D2LO + b2 * D2_STRIDE + u_begin * 8;
                    const uint32_t a_hi = tmem + lane_base + C_D2HI + b2 * D2_STRIDE + u_begin * 8;
                    tc::tmem_ld16(a_lo, lo);
                    tc::tmem_ld16(a_hi, hi);
                    if (u_end - u_begin > 2) {                     // warp-uniform: this column group has a third unit
                        tc::tmem_ld8(a_lo + 16, lo + 16);
                        tc::tmem_ld8(a_hi + 16, hi + 16);
                    }
                }
                tc::tmem_wait_ld();
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    if (u_begin + k < u_end) work(u_begin + k, lo + 8 * k, hi + 8 * k, cen_cur[k]);
                }
                __syncwarp();
                }
                // ---- staging buffer -> global memory: one TMA store per tile (rows / columns outside the page are clipped by the
                // tensor map, so nothing is ever written past the page's width or height)
                tc::fence_async_smem();
                asm volatile("bar.sync 3, %0;" ::"n"(N_EPI_WARPS * 32) : "memory");
                if (tid == 0 && !(DBG && (L.flags & 16))) {
                    if (tr.job != last_dmap) { tc::tmap_acquire(&L.dmaps[tr.job]); last_dmap = tr.job; }
                    tc::tma_store_2d(&L.dmaps[tr.job], x0, y0, s_out);
                }
                if (tid == 0) TC_STAMP(0, e, 4);
                if (STATS && Jhist) {
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
                        if (cnt) atomicAdd(&my_hist[lane], cnt);
                    }
                }
                if (tid == 0) TC_STAMP(0, e, 5);
                tc::fence_before_sync();
            }
            if (LAG == 1) {
#pragma unroll
                for (int k = 0; k < 3; k++) cen_cur[k] = cen_new[k];
                tr_job_cur = tr_job_new; tr_tx_cur = tr_tx_new; tr_ty_cur = tr_ty_new;
            }
        }
        if (n_mine > 0) flush_stats();
        if (tid == 0) tc::tma_store_wait_all();
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem, TMEM_COLS);
    TC_CRUMB(11);
#undef TC_CRUMB
#undef TC_WAIT
#undef TC_STAMP
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

int border_index(int p, int len, int replicate) {
    if (replicate) return std::min(std::max(p, 0), len - 1);
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;      // BORDER_REFLECT_101
    return p;
}

// Byte (row, k) of a K-major operand with 128-byte swizzle: 128-byte rows, 16-byte chunks XORed with the row number mod 8;
// K beyond 128 continues in the next block of `rows` rows.
size_t sw128_offset(int rows, int row, int k) {
    const int blk = k >> 7, kk = k & 127;
    return (size_t)blk * rows * 128 + (size_t)row * 128 + (size_t)(((kk >> 4) ^ (row & 7)) << 4) + (kk & 15);
}

// What one launch computes: the taps (integers), the border rule, and how the two passes are split into byte planes.
struct TcSpec {
    int mode;                 // 0: 8.8 Gaussian blur (1 plane per pass, rounding constant in two spare slots); 1: adaptive threshold (2 planes)
    int k;                    // nominal kernel size (cache key)
    const int32_t* taps; int k_eff, R;
    int replicate;            // border rule: 0 BORDER_REFLECT_101, 1 BORDER_REPLICATE
    int epi, c_param, band;
};

// Band matrix of one tile row (vertical, A operand [128 x K1]) or tile column (horizontal, B operand [NOUT x 128]):
// entry (o, slot) = sum of the taps of output o0 + o that the border rule maps onto source index o0 - margin + slot.
// One byte plane (blur) or two (adaptive: high bytes, then low bytes, `plane_bytes` apart).  Returns false when a folded
// coefficient does not fit (images much smaller than the kernel).
bool build_band(const TcSpec& S, int margin, int o0, int len, int n_out, int n_slots, int rows_alloc, bool rounding_slots, uint8_t* img,
                size_t plane_bytes) {
    const int planes = S.mode == 1 ? 2 : 1;
    std::vector<int> acc((size_t)n_out * n_slots, 0);
    for (int o = 0; o < n_out; o++) {
        const int p = o0 + o;
        if (p >= len) break;
        for (int t = 0; t < S.k_eff; t++) {
            const int sp = border_index(p + t - S.R, len, S.replicate);
            const int slot = sp - (o0 - margin);
            if (slot < 0 || slot >= n_slots) return false;
            acc[(size_t)o * n_slots + slot] += S.taps[t];
        }
    }
    if (rounding_slots)
        for (int o = 0; o < n_out; o++) {
            if (acc[(size_t)o * n_slots + n_slots - 2] || acc[(size_t)o * n_slots + n_slots - 1]) return false;
            acc[(size_t)o * n_slots + n_slots - 2] = acc[(size_t)o * n_slots + n_slots - 1] = 128;    // x 128 in the A operand, twice = 32768
        }
    memset(img, 0, plane_bytes * planes);
    for (int o = 0; o < n_out; o++)
        for (int s = 0; s < n_slots; s++) {
            const int v = acc[(size_t)o * n_slots + s];
            if (v >= (planes == 2 ? 65536 : 256)) return false;
            if (!v) continue;
            const size_t off = sw128_offset(rows_alloc, o, s);
            if (planes == 1) img[off] = (uint8_t)v;
            else { img[off] = (uint8_t)(v >> 8); img[plane_bytes + off] = (uint8_t)(v & 255); }
        }
    return true;
}

// device copy of one band matrix, cached per context: key = (mode / axis, k, near-border distances or -1, tile geometry)
int get_variant(docscan_ctx* ctx, const TcSpec& S, int axis, bool wide, int RL, int K1, int NOUT, int o0, int len, const uint8_t** out, bool* ok) {
    const int n_out = axis == 0 ? TM : NOUT, n_slots = axis == 0 ? K1 : (wide ? 2 * NIN : NIN);
    const int a = (o0 - S.R < 0) ? o0 : -1;
    const int b = (o0 + n_out - 1 + S.R > len - 1) ? len - o0 : -1;
    const std::array<int, 6> key = {S.mode * 2 + axis + (wide ? 4 : 0), S.k, a, b, NOUT, K1};
    auto it = ctx->tc_tables.find(key);
    if (it == ctx->tc_tables.end()) {
        const size_t plane = axis == 0 ? (wide ? (size_t)((K1 + 127) / 128) * TM * 128 : (size_t)T_BYTES) : (size_t)NOUT * (wide ? 256 : 128);
        const size_t bytes = plane * (S.mode == 1 ? 2 : 1);
        std::vector<uint8_t> img(bytes);
        *ok = build_band(S, axis == 0 ? S.R : RL, o0, len, n_out, n_slots, axis == 0 ? TM : NOUT, axis == 1 && S.mode == 0, img.data(), plane);
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

template <int EPI, bool STATS, bool DBG = false, bool WIDE = false>
int launch_tc(docscan_ctx* ctx, const TcLaunch& L, size_t smem, bool wide = false) {
    if constexpr (!DBG) if (L.dbg) return launch_tc<EPI, STATS, true, WIDE>(ctx, L, smem, wide);      // DOCSCAN_TC_DEBUG: dumps, stamps, skip flags
    if constexpr (!WIDE && EPI != DS_EPI_AGAUSS) if (wide) return launch_tc<EPI, STATS, DBG, true>(ctx, L, smem, wide);
    DS_CUDA(ctx, cudaFuncSetAttribute(tc_blur_kernel<EPI, STATS, DBG, WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = std::min(L.total_tiles, ctx->sm_count);            // one CTA per SM (512 TMEM columns, ~210 KB of shared memory)
    if (const char* e = getenv("DOCSCAN_TC_GRID")) grid = std::max(1, std::min(grid, atoi(e)));
    tc_blur_kernel<EPI, STATS, DBG, WIDE><<<grid, NT, smem, ctx->stream>>>(L);
    DS_CHECK_LAUNCH(ctx);
    return DOCSCAN_OK;
}

// Tile geometry, band matrices, tensor maps, upload and launch.  false: not applicable (the caller runs the CUDA-core kernels).
bool tc_run(docscan_ctx* ctx, const TcSpec& S, const BlurJob* jobs_host, int n, const char* prof_name, TcFlagLists* fl, int* rc) {
    *rc = DOCSCAN_OK;
    if (n <= 0 || n > MAX_JOBS || !encode_fn()) return false;
    if (const char* e = getenv("DOCSCAN_TC")) if (atoi(e) == 0) return false;
    const int R = S.R;
    // radii up to 48: 128 input columns per tile; up to 96 (blur only): the wide tile, 256 input columns
    // (the narrow tile keeps only 32 output columns from radius 33 on; DOCSCAN_TC_WIDE_FROM moves the switch for measurements)
    int wide_from = 33;
    if (const char* e = getenv("DOCSCAN_TC_WIDE_FROM")) wide_from = std::max(1, atoi(e));
    const bool wide = S.mode == 0 && R >= std::min(wide_from, MAX_R + 1) && R <= MAX_R_WIDE;
    if (R < 1 || (R > MAX_R && !wide)) return false;
    const int K1 = (TM + 2 * R + 31) / 32 * 32;
    const int RL = (R + 15) & ~15;                      // the window's first column must sit on a 16-byte boundary of its row
    // blur: the last two source slots of a tile carry the rounding constant (see the kernel), so the taps must end before them;
    // adaptive: three 64-column accumulators
    const int NOUT = wide ? std::min(64, (2 * NIN - 2 - RL - R) / 16 * 16)
                          : S.mode == 0 ? std::min(80, (NIN - 2 - RL - R) / 16 * 16) : std::min(64, (NIN - RL - R) / 16 * 16);
    if (NOUT < 16 || K1 > (wide ? 320 : 256) || (S.mode == 1 && NOUT != 64)) return false;
    bool stats = false;
    for (int i = 0; i < n; i++) {
        const BlurJob& j = jobs_host[i];
        if (((uintptr_t)j.src | (uintptr_t)j.dst | (uintptr_t)j.src_pitch | (uintptr_t)j.dst_pitch) & 15) return false;
        if (j.src_pitch < j.w || j.dst_pitch < j.w || j.w < 1 || j.h < 1 || j.w > 65535 || j.h > 65535) return false;
        stats = stats || j.minmax || j.hist;
    }
    // per-page geometry: tile counts, border variants of the two band matrices, tensor maps of the source and destination planes
    std::vector<TcJob> jobs(n);
    std::vector<const uint8_t*> tabs;
    std::vector<CUtensorMap> maps(2 * (size_t)n);          // [0, n): sources, [n, 2n): destinations
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
                *rc = get_variant(ctx, S, 0, wide, RL, K1, NOUT, ty * TM, b.h, &p, &ok);
                if (*rc != DOCSCAN_OK) return true;
                if (!ok) return false;
                tabs.push_back(p);
            }
            const int toe_off = (int)tabs.size();
            for (int tx = 0; tx < j.ntx; tx++) {
                const uint8_t* p = nullptr; bool ok = false;
                *rc = get_variant(ctx, S, 1, wide, RL, K1, NOUT, tx * NOUT, b.w, &p, &ok);
                if (*rc != DOCSCAN_OK) return true;
                if (!ok) return false;
                tabs.push_back(p);
            }
            g = geom.emplace(std::make_pair(b.w, b.h), std::make_pair(t_off, toe_off)).first;
        }
        j.t_off = g->second.first; j.toe_off = g->second.second;
        const cuuint64_t dims[2] = {(cuuint64_t)b.w, (cuuint64_t)b.h};
        const cuuint64_t strides[1] = {(cuuint64_t)b.src_pitch};
        const cuuint32_t box[2] = {(cuuint32_t)NIN, (cuuint32_t)(wide ? K1 / 2 : K1)};      // wide: four boxes per window
        const cuuint32_t estr[2] = {1, 1};
        const CUresult cr = encode_fn()(&maps[i], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)b.src, dims, strides, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) return false;
        const cuuint64_t dstrides[1] = {(cuuint64_t)b.dst_pitch};
        const cuuint32_t dbox[2] = {(cuuint32_t)NOUT, (cuuint32_t)TM};
        const CUresult cr2 = encode_fn()(&maps[n + i], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)b.dst, dims, dstrides, dbox, estr,
                                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr2 != CUDA_SUCCESS) return false;
    }
    if (tabs.size() > (size_t)MAX_TABS) return false;
    TcLaunch L{};
    void* dev = nullptr;
    // one upload: tensor maps (64-byte aligned) | jobs | variant pointers
    const size_t off_jobs = sizeof(CUtensorMap) * 2 * n, off_tabs = off_jobs + ((sizeof(TcJob) * n + 63) & ~(size_t)63);
    std::vector<uint8_t> blob(off_tabs + sizeof(void*) * tabs.size());
    memcpy(blob.data(), maps.data(), sizeof(CUtensorMap) * 2 * n);
    memcpy(blob.data() + off_jobs, jobs.data(), sizeof(TcJob) * n);
    memcpy(blob.data() + off_tabs, tabs.data(), sizeof(void*) * tabs.size());
    *rc = ds_upload(ctx, blob.data(), blob.size(), &dev);
    if (*rc != DOCSCAN_OK) return true;
    L.maps = reinterpret_cast<const CUtensorMap*>(dev);
    L.dmaps = L.maps + n;
    L.jobs = reinterpret_cast<const TcJob*>((uint8_t*)dev + off_jobs);
    L.tabs = reinterpret_cast<const uint8_t* const*>((uint8_t*)dev + off_tabs);
    L.n_jobs = n; L.n_tabs = (int)tabs.size(); L.total_tiles = total; L.R = R; L.RL = RL; L.K1 = K1; L.NOUT = NOUT;
    L.idesc1 = tc::idesc_i8(TM, wide ? 2 * NIN : NIN, 0, 0, 0, 1);
    L.idesc2 = tc::idesc_i8(TM, NOUT, 0, 0, 0, 0);
    L.c_param = S.c_param; L.band = S.band;
    if (fl) {
        // one list per tile: the exact re-evaluation works tile by tile out of shared memory (adaptive.cu)
        void* cnt = nullptr; void* lst = nullptr;
        *rc = ds_arena_alloc(ctx, sizeof(uint32_t) * (size_t)total, &cnt);
        if (*rc == DOCSCAN_OK) *rc = ds_arena_alloc(ctx, sizeof(uint16_t) * (size_t)total * TC_TILE_FLAG_CAP, &lst);
        if (*rc != DOCSCAN_OK) return true;
        if (cudaMemsetAsync(cnt, 0, sizeof(uint32_t) * (size_t)total, ctx->stream) != cudaSuccess) { *rc = ds_fail(ctx, DOCSCAN_ERR_CUDA, "memset failed"); return true; }
        L.flag_count = (uint32_t*)cnt; L.flag_list = (uint16_t*)lst; L.flag_cap = TC_TILE_FLAG_CAP;
        fl->count = L.flag_count; fl->list = L.flag_list; fl->n_tiles = total; fl->RL = RL; fl->NOUT = NOUT;
        fl->tiles.resize(n);
        for (int i = 0; i < n; i++) { fl->tiles[i].tile_base = jobs[i].tile_base; fl->tiles[i].ntx = jobs[i].ntx; fl->tiles[i].nty = jobs[i].nty; }
    }
    if (!ctx->tc_status) {
        if (cudaHostAlloc((void**)&ctx->tc_status, 64, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return false; }
        memset(ctx->tc_status, 0, 64);
    }
    L.status = ctx->tc_status;
    const char* dbg_path = getenv("DOCSCAN_TC_DEBUG");
    const size_t dbg_words = 16384 + 2 * 12288 + 2 * 2048;     // first tile's accumulators + phase clocks of CTA 0 (epilogue warp 0, issuer)
    if (dbg_path) {
        void* d = nullptr;
        *rc = ds_arena_alloc(ctx, dbg_words * 4, &d);
        if (*rc != DOCSCAN_OK) return true;
        cudaMemsetAsync(d, 0xEE, dbg_words * 4, ctx->stream);
        L.dbg = (uint32_t*)d;
        L.crumbs = getenv("DOCSCAN_TC_CRUMBS") != nullptr;
        if (const char* f = getenv("DOCSCAN_TC_FLAGS")) L.flags = atoi(f);
    }
    const size_t planes = S.mode == 1 ? 2 : 1;
    size_t smem;
    if (!wide) {
        L.toe_slots = TOE_SLOTS; L.ns = NS;
        const size_t smem_fixed = 1024 + (size_t)TOE_SLOTS * NOUT * 128 * planes + (size_t)NS * K1 * 128 + (size_t)TM * NOUT + (stats ? 8 * 256 * 4 : 0);
        // + ~9 KB of static shared memory <= 227 KB; the adaptive threshold's matrix pair is 64 KB: one slot
        L.t_slots = S.mode == 1 ? 1 : (smem_fixed + 3 * (size_t)T_BYTES <= 218 * 1024 ? 3 : 2);
        smem = smem_fixed + (size_t)L.t_slots * T_BYTES * planes;
    } else {
        // wide tile: one window is 64..80 KB; two stages when they fit (K1 = 256), else the load of tile i+1 waits for pass 1 of tile i
        const size_t t_bytes = (size_t)((K1 + 127) / 128) * TM * 128, window = (size_t)K1 * 2 * NIN;
        L.toe_slots = 2; L.t_slots = 1;
        const size_t smem_fixed = 1024 + t_bytes + 2 * (size_t)NOUT * 256 + (size_t)TM * NOUT + (stats ? 8 * 256 * 4 : 0);
        L.ns = smem_fixed + 2 * window <= 218 * 1024 ? 2 : 1;
        smem = smem_fixed + (size_t)L.ns * window;
    }
    {
        ProfScope prof(ctx, prof_name, 2.0 * px);
#define DS_TC_CASE(E) case E: *rc = stats ? launch_tc<E, true>(ctx, L, smem, wide) : launch_tc<E, false>(ctx, L, smem, wide); break;
        switch (S.epi) {
            DS_TC_CASE(DS_EPI_BLUR)
            DS_TC_CASE(DS_EPI_SUB)
            DS_TC_CASE(DS_EPI_RSUB)
            DS_TC_CASE(DS_EPI_DIV)
            case DS_EPI_AGAUSS: *rc = launch_tc<DS_EPI_AGAUSS, false>(ctx, L, smem); break;
            default: return false;
        }
#undef DS_TC_CASE
    }
    if (dbg_path && *rc == DOCSCAN_OK) {
        std::vector<uint32_t> host(dbg_words);
        cudaMemcpyAsync(host.data(), L.dbg, dbg_words * 4, cudaMemcpyDeviceToHost, ctx->stream);
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (FILE* f = fopen(dbg_path, "wb")) {
            fprintf(stderr, "[tc debug] %s k=%d k_eff=%d R=%d RL=%d K1=%d NOUT=%d tiles=%d sync=%s timeout_id=%u progress=%u\n", prof_name, S.k, S.k_eff, R, RL,
                    K1, NOUT, total, cudaGetErrorString(e), ctx->tc_status[0], ctx->tc_status[1]);
            fwrite(host.data(), 4, dbg_words, f);
            fclose(f);
        }
    }
    return true;
}

}  // namespace

// cv2.GaussianBlur (8.8 fixed point) + epilogue.  Returns false when the tensor-core path does not apply (the caller then runs
// blur.cu); otherwise *rc is the result.
bool k_tc_blur_jobs(docscan_ctx* ctx, int kind, int k, int epi, const BlurJob* jobs_host, int n, int* rc) {
    *rc = DOCSCAN_OK;
    if (kind != 0 || k < 3) return false;
    if (epi != DS_EPI_BLUR && epi != DS_EPI_SUB && epi != DS_EPI_RSUB && epi != DS_EPI_DIV) return false;
    std::vector<int32_t> q(k);
    if (docscan_gaussian_kernel_q8(k, q.data()) != DOCSCAN_OK) return false;
    int z = 0;
    while (z < k / 2 && q[z] == 0) z++;                 // zero tails of the quantised kernel
    const int k_eff = k - 2 * z;
    for (int i = 0; i < k_eff; i++) if (q[z + i] > 255) return false;
    TcSpec S{};
    S.mode = 0; S.k = k; S.taps = q.data() + z; S.k_eff = k_eff; S.R = k_eff / 2; S.replicate = 0; S.epi = epi;
    const std::string name = std::string("tc_blur_k") + std::to_string(k);
    return tc_run(ctx, S, jobs_host, n, name.c_str(), nullptr, rc);
}

// cv2.adaptiveThreshold(GAUSSIAN_C): the local mean in fixed point on the tensor cores, decided wherever it is further from the
// rounding boundary than its own error bound; the pixels inside the guard band are listed for the exact evaluation
// (adaptive.cu: k_adaptive_fix_flagged).  `w16` = the fp32 Gaussian taps * 65536, rounded; `band` in 1/65536 grey levels.
bool k_tc_adaptive_jobs(docscan_ctx* ctx, int k, int c_param, const int32_t* w16, int band, const AdaptJob* jobs_host, int n, TcFlagLists* fl, int* rc) {
    *rc = DOCSCAN_OK;
    if (k < 3 || (k & 1) == 0) return false;
    std::vector<BlurJob> bj(n);
    for (int i = 0; i < n; i++) {
        bj[i] = BlurJob{};
        bj[i].src = jobs_host[i].src; bj[i].dst = jobs_host[i].dst; bj[i].src_pitch = jobs_host[i].src_pitch; bj[i].dst_pitch = jobs_host[i].dst_pitch;
        bj[i].w = jobs_host[i].w; bj[i].h = jobs_host[i].h;
        if (jobs_host[i].w == 1 || jobs_host[i].h == 1) return false;      // cv2 shrinks the kernel on 1-pixel axes
    }
    TcSpec S{};
    S.mode = 1; S.k = k; S.taps = w16; S.k_eff = k; S.R = k / 2; S.replicate = 1; S.epi = DS_EPI_AGAUSS; S.c_param = c_param; S.band = band;
    const std::string name = std::string("tc_adaptive_k") + std::to_string(k);
    return tc_run(ctx, S, bj.data(), n, name.c_str(), fl, rc);
}
