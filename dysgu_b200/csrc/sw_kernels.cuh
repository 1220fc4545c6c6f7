// Smith-Waterman device kernels for sm_100a.
//
// What they compute is the observable behaviour of dysgu's vendored striped Smith-Waterman
// (reference: dysgu/scikitbio/ssw.c:124-346, 372-548 scoring passes; 550-728 banded traceback;
// 772-857 driver) - not its SIMD layout.  The design is B200-first:
//
//   * The target ("ref", DP columns) is DISTRIBUTED over virtual lanes, the query ("read", DP rows)
//     is STREAMED.  Each thread carries two virtual lanes in the two int16 halves of its 32-bit
//     registers (s16x2), skewed by one row, so every DPX instruction (VIADDMNMX.S16x2[.RELU],
//     VIMNMX.S16x2) updates two independent cells.  A group of G threads (G = 1..32) forms a
//     wavefront of 2G virtual lanes x K columns; 32/G pairs share a warp.
//   * The per-pair substitution profile lives in shared memory, laid out so that every LDS.128
//     of a quarter warp hits 32 distinct banks; sequences are read as packed 4-bit codes with
//     128-bit / 32-bit loads.
//   * Targets longer than 2*G*K columns are processed in column tiles ("passes"); the right-edge
//     column of a tile travels to the next pass through a per-warp boundary array.
//   * Tie-breaking of the reference (first column holding the maximum, then the smallest row in
//     that column) is reproduced from a per-lane running maximum plus per-column maxima; the rare
//     pair with several maximal columns inside one virtual lane is re-run by the WATCH variant.
#pragma once
#include "engine.cuh"

namespace db200 {

constexpr int SW_PAD_CODE = 5;       // nibble used for padding; scores SW_NEG against everything
constexpr int SW_NEG = -16384;
constexpr int SW_NCFG = 60;          // kernel configurations (K, G): 11 classic (K = 16), 48 fine, the column-tiled one (sw_launch.cuh)
constexpr int SW_LBINS = 512;        // length bins per (config, colmax) segment, 32 rows each
constexpr int SW_NSEG = SW_NCFG * 2; // (config, needs column maxima)
constexpr int SW_MULTI_MAX_WARPS_PER_SM = 8;
#ifndef SW_PASS_THREADS
#define SW_PASS_THREADS 128       // four warps per block (two-warp blocks with nine blocks per SM measured the same)
#endif
#ifndef SW_PASS_MIN_BLOCKS
#define SW_PASS_MIN_BLOCKS 4
#endif // bounds the per-warp boundary scratch of the column-tiled kernel
#ifndef SW_FINE_KMAX
#define SW_FINE_KMAX 20           // largest K that still runs four blocks per SM (registers and 768*K bytes of profile per warp)
#endif

struct PairMeta {      // one alignment problem on packed sequences
    uint32_t q_unit;   // offset of the streamed sequence in the nibble pool, in 16-byte units (32 bases)
    uint32_t t_unit;   // offset of the distributed sequence
    int32_t m;         // streamed length (rows)
    int32_t n;         // distributed length (columns)
};

struct PassOut {       // result of one scoring pass
    int32_t score;
    int32_t end_t;     // column of the best cell (first column holding the maximum)
    int32_t end_q;     // smallest row holding the maximum in that column
    int32_t amb;       // several maximal columns inside one virtual lane: end_q must be recomputed
};

struct SwScoring {
    int16_t mat[36];   // 6x6: ACGTN + pad; row = distributed code, column = streamed code
    int32_t gap_open, gap_extend;
};

struct PassArgs {
    const uint4 *pool;
    const PairMeta *meta;
    const int32_t *tasks;      // pair ids ordered by (segment, length bin)
    const int32_t *seg_begin;  // [SW_NSEG] first task of each segment
    const int32_t *seg_count;  // [SW_NSEG]
    int32_t seg;
    int32_t *counter;          // work counter of this launch
    PassOut *out;
    int16_t *colmax;           // COLMAX: per-column maxima then last-row values, at colmax_off[pair]
    const int64_t *colmax_off;
    uint32_t *boundary;        // MULTI: per-warp right-edge column (fallback when the per-pair pool is exhausted)
    int32_t boundary_stride;
    uint32_t *bnd_pool;        // MULTI: per-pair, per-pass right-edge columns, kept so that an ambiguous pair only
    unsigned long long *bnd_cursor; //        has to re-run the winning pass
    unsigned long long bnd_cap;
    long long *bnd_off;        // per pair: offset into bnd_pool (words), -1 if not kept
    int32_t *amb_list;         // pairs whose end_q is ambiguous (appended at seg_begin[amb_base_seg])
    int32_t *amb_count;
    int32_t amb_base_seg;
    uint32_t k65536;           // 65536 passed at run time so the score merge stays an IMAD
    // PIPE (column-tiled pairs, one warp per pass): work units are (pair, pass) entries of an expanded list
    const int32_t *vt_pair;    // unit -> pair
    const int32_t *vt_pass;    // unit -> pass; -1: run every pass of the pair in this warp (boundary pool exhausted), -2: nothing to do
    const int64_t *vt_total;   // number of units
    int32_t *vt_prog;          // unit -> rows of its right-edge column that are complete (read by the next pass)
    PassOut *vt_res;           // unit -> candidate of its pass
    int32_t *pair_done;        // pair -> passes finished; the last one merges the candidates
    int32_t *abort_flag;       // set when a wait for the producer pass timed out (never expected; avoids a hang)
    SwScoring sc;
};

__device__ __forceinline__ uint32_t pack16(int v) { return (uint32_t)(v & 0xFFFF) * 0x10001u; }
__device__ __forceinline__ int lo16(uint32_t x) { return (int)(int16_t)(x & 0xFFFF); }
__device__ __forceinline__ int hi16(uint32_t x) { return (int)(int16_t)(x >> 16); }

// PTX prmt in its generic form: selector nibbles 0..7 pick a byte of {a (0..3), b (4..7)}, bit 3 of a nibble replicates
// that byte's sign bit instead (0x00 for a non-negative value) - __byte_perm documents three selector bits only
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

__device__ __forceinline__ uint32_t unit_word(const uint4 &u, int w)
{
    return w == 0 ? u.x : (w == 1 ? u.y : (w == 2 ? u.z : u.w));
}

// 4-bit code at position idx (>= 0) of a packed sequence starting at unit `unit`
__device__ __forceinline__ int pool_code(const uint4 *pool, uint32_t unit, int idx)
{
    const uint32_t *w = reinterpret_cast<const uint32_t *>(pool + unit);
    return (int)((w[idx >> 3] >> ((idx & 7) * 4)) & 15u);
}

// TAG: scores are pre-scaled by 2^TAGB (host side) and the column index inside the virtual lane rides in the low
// bits of the running maximum, so the best cell's column and row are exact without per-column maxima.
// PIPE: the passes (2048-column tiles) of one long pair are work units of their own, taken by different warps.  Pass p
// consumes the right-edge column of pass p-1 in 32-row chunks as soon as they are complete (a progress word per unit),
// so the tiles of a 10 kb x 10 kb pair run as a pipeline of five warps instead of one warp after the other.  Units are
// dispensed in list order and a pass only ever waits for the unit right before it, which a resident warp already holds.
template <int K, int G, bool COLMAX, bool MULTI, bool WATCH, bool TAG, bool PIPE = false>
__global__ void __launch_bounds__(SW_PASS_THREADS, (K <= SW_FINE_KMAX && !MULTI) ? SW_PASS_MIN_BLOCKS : 1) sw_pass_kernel(PassArgs a)
{
    static_assert(!PIPE || (MULTI && !WATCH), "pipelined passes are the column-tiled scoring kernel");
    constexpr int V = 2 * G;          // virtual lanes per group
    constexpr int C = V * K;          // columns per pass
    constexpr int GROUPS = 32 / G;    // pairs per warp (G need not divide 32: the remaining lanes idle)
    // profile row of one streamed code, per warp: 32-bit word [k][lane] = (score of the low-half column k) |
    // (score of the high-half column k) << 16.  Every lane owns one bank, so the 16-bit loads below never conflict
    // whatever mix of rows the lanes read.
    constexpr int ROW16 = 2 * K * 32; // uint16 entries per row
    constexpr int TAGB = (K <= 16) ? 4 : 5;
    constexpr bool VMAX = !TAG || COLMAX;
    constexpr bool LAZY = (G > 1);    // strict-improvement test once per SW_UNR rows instead of every row
    constexpr bool ALIGNED = (K % 16 == 0);   // a lane's 2K columns are whole 16-byte units of the pool
    constexpr int UNITS = ALIGNED ? (2 * K) / 32 : 1;
    constexpr int TW = (2 * K + 7) / 8 + 1;   // !ALIGNED: 32-bit words that can hold a lane's 2K codes at any nibble offset
    static_assert(K >= 2 && K <= 32, "columns per virtual lane");
    static_assert(!MULTI || G == 32, "column tiling runs one pair per warp");
    static_assert(!(TAG && (WATCH || MULTI)), "the tagged variant is a single-pass kernel");

    extern __shared__ uint4 smem[];
    __shared__ int16_t mat_s[36];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lig = lane % G, grp = lane / G;
    uint16_t *prof = reinterpret_cast<uint16_t *>(smem) + warp * (6 * ROW16);
    if (threadIdx.x < 36) mat_s[threadIdx.x] = a.sc.mat[threadIdx.x];
    __syncthreads();

    const uint32_t nGE2 = pack16(-a.sc.gap_extend), nGO2 = pack16(-a.sc.gap_open);
    // lane hand-over: low half <- previous lane's high half, high half <- own low half, as one byte permute.  The first
    // lane of a group takes "sign of the high byte of its own low half" instead of the neighbour's bytes: 0 for H (never
    // negative), 0 or -1 for the horizontal gap state - any non-positive value entering column 0 gives the same H
    // everywhere (a gap state only matters once it is positive, and then it was re-opened from H).  No select per row.
    const uint32_t selIn = (lig == 0) ? 0x54DDu : 0x5432u;
    const int seg_begin = a.seg_begin[a.seg], seg_count = a.seg_count[a.seg];
    const unsigned FULL = 0xFFFFFFFFu;
    uint32_t *bnd = nullptr;
    if (MULTI) bnd = a.boundary + (size_t)(blockIdx.x * (blockDim.x >> 5) + warp) * (size_t)a.boundary_stride;

    while (true) {
        // The warp must run every unit converged: ncu's source counters showed the first unit of each warp executing as
        // two 16-thread fragments (every instruction issued twice) when the result store at the bottom of the loop
        // was the last statement of the body; the explicit warp barriers here and there pin the reconvergence points.
        __syncwarp();
        int base = 0;
        if (lane == 0) base = atomicAdd(a.counter, GROUPS);
        __syncwarp();
        base = __shfl_sync(FULL, base, 0);
        int my_pass = -1;                  // PIPE: the pass this unit stands for (-1: all passes, as without PIPE)
        if (PIPE) {
            if ((int64_t)base >= *a.vt_total) break;
            my_pass = a.vt_pass[base];
            if (my_pass == -2) continue;
        } else {
            if (base >= seg_count) break;
        }
        const int ti = base + grp;
        const bool active = PIPE ? true : (grp < GROUPS && ti < seg_count);
        const int pair = PIPE ? a.vt_pair[base] : (active ? a.tasks[seg_begin + ti] : -1);
        PairMeta pm;
        pm.q_unit = 0; pm.t_unit = 0; pm.m = 0; pm.n = 0;
        if (active) pm = a.meta[pair];
        const int m = pm.m, n = pm.n;
        const uint32_t *qw = reinterpret_cast<const uint32_t *>(a.pool + pm.q_unit);
        const int q_words = ((m + 31) >> 5) << 2;

        int watch_score = -1, watch_slot = -1;
        if (WATCH && active) { watch_score = a.out[pair].score; watch_slot = a.out[pair].end_t; }

        int steps = m + V - 1;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) steps = max(steps, __shfl_xor_sync(FULL, steps, o));
        const int npass = MULTI ? max(1, (n + C - 1) / C) : 1;

        int run_score = -1, run_slot = 0, run_row = 0, run_amb = 0, run_wrow = -1;

        // column-tiled pairs keep the right-edge column of every pass (rows padded to 32 per pass)
        int ps_first = 0, ps_last = npass;
        uint32_t *pbnd = nullptr;          // per-pair boundary storage, pass p at pbnd + p * bstride
        int bstride = 0;
        if (MULTI && active && npass > 1) {
            bstride = (m + 31) & ~31;
            long long off = -1;
            if (PIPE) {
                off = a.bnd_off[pair];     // assigned by sw_expand_kernel
                if (my_pass >= 0) { ps_first = my_pass; ps_last = my_pass + 1; }
            } else if (!WATCH) {
                if (lane == 0) {
                    const unsigned long long need = (unsigned long long)bstride * (unsigned long long)(npass - 1);
                    const unsigned long long o = atomicAdd(a.bnd_cursor, need);
                    off = (o + need <= a.bnd_cap) ? (long long)o : -1;
                    a.bnd_off[pair] = off;
                }
                off = __shfl_sync(FULL, off, 0);
            } else {
                off = a.bnd_off[pair];
                if (off >= 0) { ps_first = watch_slot / C; ps_last = ps_first + 1; } // only the winning pass is re-run
            }
            if (off >= 0) pbnd = a.bnd_pool + off;
        }

        for (int ps = ps_first; ps < ps_last; ++ps) {
            // ---- profile of this pass's columns -> shared memory -------------------------------
            __syncwarp();
            {
                const int j0 = ps * C + lig * 2 * K;
                if constexpr (ALIGNED) {
                    uint4 tu[UNITS];
#pragma unroll
                    for (int u = 0; u < UNITS; ++u) {
                        const int jb = j0 + u * 32;
                        if (jb < n) tu[u] = a.pool[pm.t_unit + (jb >> 5)];
                        else tu[u] = make_uint4(0x55555555u, 0x55555555u, 0x55555555u, 0x55555555u);
                    }
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
#pragma unroll
                        for (int k = 0; k < K; ++k) {
                            const int jj = h * K + k; // position inside this lane's 2K columns
                            int tc = (int)((unit_word(tu[jj >> 5], (jj >> 3) & 3) >> ((jj & 7) * 4)) & 15u);
                            tc = min(tc, SW_PAD_CODE) * 6;
#pragma unroll
                            for (int c = 0; c < 6; ++c)
                                prof[c * ROW16 + (k * 32 + lane) * 2 + h] = (uint16_t)mat_s[tc + c];
                        }
                    }
                } else {
                    // any K: the lane's codes start at nibble j0 of the target; words past the sequence's last unit read as padding
                    const uint32_t *tw = reinterpret_cast<const uint32_t *>(a.pool + pm.t_unit);
                    const int t_words = ((n + 31) >> 5) << 2;
                    const int w0 = j0 >> 3;
                    const int sh = (j0 & 7) * 4;
                    uint32_t raw[TW];
#pragma unroll
                    for (int u = 0; u < TW; ++u) raw[u] = (w0 + u < t_words) ? tw[w0 + u] : 0x55555555u;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
#pragma unroll
                        for (int k = 0; k < K; ++k) {
                            const int jj = h * K + k;
                            const uint32_t w = __funnelshift_r(raw[jj >> 3], raw[(jj >> 3) + 1], sh);
                            int tc = (int)((w >> ((jj & 7) * 4)) & 15u);
                            tc = min(tc, SW_PAD_CODE) * 6;
#pragma unroll
                            for (int c = 0; c < 6; ++c)
                                prof[c * ROW16 + (k * 32 + lane) * 2 + h] = (uint16_t)mat_s[tc + c];
                        }
                    }
                }
            }
            __syncwarp();

            // ---- state ---------------------------------------------------------------------------
            uint32_t H[K], U[K], HmO[K], vmax[VMAX ? K : 1];
#pragma unroll
            for (int k = 0; k < K; ++k) { H[k] = 0; U[k] = 0; HmO[k] = 0; }
#pragma unroll
            for (int k = 0; k < (VMAX ? K : 1); ++k) vmax[k] = 0;
            uint32_t best2 = 0;
            int bstep_lo = 0, bstep_hi = 0;
            uint32_t OH = 0, OV = 0, prevIn = 0;
            int wstep_lo = -1, wstep_hi = -1;
            int kstar_lo = -1, kstar_hi = -1;
            if (WATCH) {
                const int b0 = ps * C + (2 * lig) * K;
                if (watch_slot >= b0 && watch_slot < b0 + K) kstar_lo = watch_slot - b0;
                if (watch_slot >= b0 + K && watch_slot < b0 + 2 * K) kstar_hi = watch_slot - b0 - K;
            }

            // streamed codes: this lane's low half sits at row t - 2*lig
            const int phi4 = ((-2 * lig) & 7) * 4;
            int widx = (-2 * lig) >> 3; // floor
            uint32_t prevw = (widx >= 0 && widx < q_words) ? qw[widx] : 0x55555555u;
            uint32_t cur = 0;
            const uint16_t *row_hi = prof + SW_PAD_CODE * ROW16 + lane * 2 + 1; // row -1: padding
            uint32_t bchunk = 0, wchunk = 0;

            // eight rows per trip (one 32-bit word of query codes).  Rows past the end of the query read padding
            // codes and cannot change any result.  Unrolling four rows lets the compiler rename the rotating H / diag
            // registers instead of moving them and drops loop counters from the ALU pipe (measured on B200, forward
            // stage of the 524k-pair step: 18.1 ms rolled, 17.9 / 17.4 / 17.6 ms unrolled by 2 / 4 / 8).
#ifndef SW_UNR
#define SW_UNR 4
#endif
            constexpr int UNR = SW_UNR;
            static_assert(UNR == 4, "the query decode below takes the codes of four rows from two 8-bit windows");
            const int trips = (steps + 7) >> 3;
            for (int t8 = 0; t8 < trips; ++t8) {
              {
                ++widx;
                const uint32_t neww = (widx >= 0 && widx < q_words) ? qw[widx] : 0x55555555u;
                cur = __funnelshift_r(prevw, neww, phi4);
                prevw = neww;
              }
              if (MULTI && (t8 & 3) == 0) {
                // every 32 rows, outside the unrolled row loop (which stays the same code with and without PIPE):
                // publish what this pass has completed, then fetch the next 32 rows of the left neighbour's right edge
                const int t = t8 * 8;
                if (PIPE && my_pass >= 0 && ps + 1 < npass && t >= 96) {
                    // the chunk stored at step t - 2 ends at row t - 65: rows below ((t - 64) & ~31) are complete
                    __threadfence();
                    __syncwarp();
                    if (lane == 0) *(volatile int32_t *)(a.vt_prog + base) = min((t - 64) & ~31, m);
                }
                const uint32_t *src = pbnd ? pbnd + (size_t)(ps - 1) * bstride : bnd;
                if (PIPE && my_pass > 0 && t < m) {
                    const int need = min(t + 32, m);
                    if (lane == 0) {
                        const volatile int32_t *pg = a.vt_prog + (base - 1);
                        unsigned spins = 0;
#pragma unroll 1
                        while (*pg < need && *(volatile int32_t *)a.abort_flag == 0) {
                            __nanosleep(100);
                            if (++spins > (1u << 24)) { atomicExch(a.abort_flag, 1); break; }
                        }
                    }
                    __syncwarp();
                    __threadfence();
                    bchunk = (t + lane < m) ? __ldcg(src + t + lane) : 0u;
                } else {
                    bchunk = (ps > 0 && t + lane < m) ? src[t + lane] : 0u;
                }
              }
#pragma unroll 1
              for (int hb = 0; hb < 8 / UNR; ++hb) {
              uint32_t m2r[UNR];
              // The running maximum of a virtual lane is one chain through all its cells: a row group starts from the best
              // value so far (no zero per row, no fold at the end) and m2r[] keeps what it was after each row.
              uint32_t m2 = best2;
              // codes of this group's rows: one variable shift per group, then a constant mask and an IMAD per row (the
              // mask keeps the code in place, the multiply scales it to the profile row: (x & 0xF0) * (row / 16) ...)
              const uint32_t curh = cur >> (4 * UNR * hb);
              const uint32_t curh2 = curh >> 8;
#pragma unroll
              for (int bb = 0; bb < UNR; ++bb) {
                const int b = hb * UNR + bb;
                const int t = t8 * 8 + b;
                // the pool only holds codes 0..5
                const uint32_t cw = (bb & 2) ? curh2 : curh;
                const uint16_t *row_lo = (bb & 1) ? prof + ((cw & 0xF0u) * (uint32_t)(ROW16 / 16)) + lane * 2
                                                  : prof + ((cw & 0x0Fu) * (uint32_t)ROW16) + lane * 2;

                // ---- inputs from the preceding virtual lane ---------------------------------------
                uint32_t sh = OH, sv = OV;
                if (G > 1) {
                    sh = __shfl_sync(FULL, OH, lane - 1);
                    sv = __shfl_sync(FULL, OV, lane - 1);
                }
                uint32_t inH, inV;
                if (MULTI) {
                    const uint32_t y = __shfl_sync(FULL, bchunk, t & 31);
                    if (lig == 0) { sh = y << 16; sv = y & 0xFFFF0000u; }
                    inH = __funnelshift_l(sh, OH, 16); // lo <- previous lane's hi, hi <- own lo
                    inV = __funnelshift_l(sv, OV, 16);
                } else if (G == 1) {
                    inH = OH << 16;   // no neighbour: two IMAD.SHL on the FMA pipe
                    inV = OV << 16;
                } else {
                    inH = prmt(sh, OH, selIn);
                    inV = prmt(sv, OV, selIn);
                }

                uint32_t Vv = inV;
                uint32_t diag = prevIn;
                prevIn = inH;
                if (!LAZY) m2 = best2;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    {
                        // both halves' scores: two zero-extending 16-bit shared loads merged by one IMAD (FMA pipe),
                        // keeping the ALU pipe for the DPX recurrence
                        const uint32_t wl = row_lo[k * 64], wh = row_hi[k * 64];
                        const uint32_t s = wh * a.k65536 + wl;
                        U[k] = __viaddmax_s16x2_relu(U[k], nGE2, HmO[k]);
                        const uint32_t Hp = __viaddmax_s16x2_relu(diag, s, U[k]);
                        diag = H[k];
                        HmO[k] = __vadd2(Hp, nGO2);
                        H[k] = __vmaxs2(Hp, Vv);
                        Vv = __viaddmax_s16x2(Vv, nGE2, HmO[k]);
                        if (VMAX) vmax[k] = __vmaxs2(vmax[k], H[k]);
                        if (TAG) m2 = __viaddmax_s16x2(H[k], pack16(K - 1 - k), m2);
                        else m2 = __vmaxs2(m2, H[k]);
                    }
                }
                OH = H[K - 1];
                OV = Vv;
                row_hi = row_lo + 1;
                if (LAZY) m2r[bb] = m2;   // running maximum of each virtual lane: examined once per UNR rows, below
                else {
                    const uint32_t chg = m2 ^ best2;
                    if (chg & 0x0000FFFFu) bstep_lo = t;
                    if (chg & 0xFFFF0000u) bstep_hi = t;
                    best2 = m2;
                }

                if (WATCH) {
                    uint32_t hs_lo = 0xFFFFFFFFu, hs_hi = 0xFFFFFFFFu;
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        if (k == kstar_lo) hs_lo = H[k] & 0xFFFFu;
                        if (k == kstar_hi) hs_hi = H[k] >> 16;
                    }
                    if (wstep_lo < 0 && hs_lo == (uint32_t)watch_score) wstep_lo = t;
                    if (wstep_hi < 0 && hs_hi == (uint32_t)watch_score) wstep_hi = t;
                }

                if (COLMAX) {
                    // last-row values H(m-1, j): lo half is at row t-2*lig, hi half one row behind
                    const unsigned d = (unsigned)(t - 2 * lig - (m - 1));
                    if (d <= 1u && active) {
                        int16_t *lr = a.colmax + a.colmax_off[pair] + n;
                        const int jb = ps * C + (2 * lig + (int)d) * K;
#pragma unroll
                        for (int k = 0; k < K; ++k)
                            if (jb + k < n) lr[jb + k] = (int16_t)((d ? (H[k] >> 16) : (H[k] & 0xFFFFu)) >> (TAG ? TAGB : 0));
                    }
                }

                if (MULTI) {
                    // right-edge column of this pass (virtual lane V-1 = hi half of lane 31), row t-(V-1)
                    const int r = t - (V - 1);
                    const uint32_t x = __byte_perm(OH, OV, 0x7632);
                    const uint32_t val = __shfl_sync(FULL, x, 31);
                    if (r >= 0) {
                        if (lane == (r & 31)) wchunk = val;
                        if ((r & 31) == 31 || t == steps - 1) {
                            const int r0 = r & ~31;
                            uint32_t *dst = pbnd ? pbnd + (size_t)ps * bstride : bnd;
                            if (lane <= (r & 31) && r0 + lane < m && ps + 1 < npass && !(WATCH && pbnd)) dst[r0 + lane] = wchunk;
                        }
                    }
                }
              }
              // ---- running maximum of each virtual lane (strict improvement -> step) ------------
              // The chain's value after the group tells whether any of the UNR rows improved a lane's maximum; only then
              // (a warp-uniform branch) the rows' snapshots are examined one by one for the first step of strict improvement.
              // With 32 pairs in a warp (G = 1) some pair improves in most row groups and the per-row test is cheaper
              // (B200, 32-column targets: 110 issue slots per row against 121).
              // (__vibmax_s16x2's predicate outputs would do it per row in one instruction, but the variant built with
              //  it produced wrong rows on B200 with CUDA 12.9.)
              if (LAZY) {
                if (__any_sync(FULL, m2 != best2)) {
#pragma unroll
                    for (int bb = 0; bb < UNR; ++bb) {
                        const uint32_t chg = m2r[bb] ^ best2;   // m2r[] never decreases: a changed half is an improved half
                        if (chg & 0x0000FFFFu) bstep_lo = t8 * 8 + hb * UNR + bb;
                        if (chg & 0xFFFF0000u) bstep_hi = t8 * 8 + hb * UNR + bb;
                        best2 = m2r[bb];
                    }
                }
              }
              }
            }

            if (PIPE && my_pass >= 0 && ps + 1 < npass) {   // the whole right edge of this pass is in place
                __threadfence();
                __syncwarp();
                if (lane == 0) *(volatile int32_t *)(a.vt_prog + base) = m;
            }
            // ---- end of pass: column maxima, lane winners -----------------------------------------
            if (COLMAX && active) {
                int16_t *cm = a.colmax + a.colmax_off[pair];
                const int jb = ps * C + 2 * lig * K;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    if (jb + k < n) cm[jb + k] = (int16_t)((vmax[k] & 0xFFFFu) >> (TAG ? TAGB : 0));
                    if (jb + K + k < n) cm[jb + K + k] = (int16_t)((vmax[k] >> 16) >> (TAG ? TAGB : 0));
                }
            }
            int lb_lo = (int)(best2 & 0xFFFFu), lb_hi = (int)(best2 >> 16);
            int k1_lo = 0, cnt_lo = 0, k1_hi = 0, cnt_hi = 0;
            if (TAG) {
                k1_lo = K - 1 - (lb_lo & ((1 << TAGB) - 1)); k1_hi = K - 1 - (lb_hi & ((1 << TAGB) - 1));
                lb_lo >>= TAGB; lb_hi >>= TAGB;
                cnt_lo = cnt_hi = 1;
            } else {
#pragma unroll
                for (int k = K - 1; k >= 0; --k) {
                    if ((int)(vmax[k] & 0xFFFFu) == lb_lo) { k1_lo = k; ++cnt_lo; }
                    if ((int)(vmax[k] >> 16) == lb_hi) { k1_hi = k; ++cnt_hi; }
                }
            }
            const int comp_lo = (lb_lo << 8) | (255 - 2 * lig);
            const int comp_hi = (lb_hi << 8) | (255 - (2 * lig + 1));
            int gmax = max(comp_lo, comp_hi);
            {
                const int mine = gmax;
#pragma unroll
                for (int o = 1; o < G; ++o) {
                    const int other = __shfl_sync(FULL, mine, grp * G + (lig + o) % G);
                    gmax = max(gmax, other);
                }
            }
            const int vw = 255 - (gmax & 255);
            const int src = (grp * G + (vw >> 1)) & 31;
            const int hsel = vw & 1;
            int d_k1 = hsel ? k1_hi : k1_lo, d_cnt = hsel ? cnt_hi : cnt_lo, d_step = hsel ? bstep_hi : bstep_lo;
            d_k1 = __shfl_sync(FULL, d_k1, src);
            d_cnt = __shfl_sync(FULL, d_cnt, src);
            d_step = __shfl_sync(FULL, d_step, src);
            const int p_score = gmax >> 8;
            if (p_score > run_score) {
                run_score = p_score;
                run_slot = ps * C + vw * K + d_k1;
                run_row = d_step - vw;
                run_amb = d_cnt > 1;
            }
            if (WATCH) {
                int w = -1;
                if (wstep_lo >= 0) w = wstep_lo - 2 * lig;
                if (wstep_hi >= 0) w = wstep_hi - (2 * lig + 1);
                {
                    const int mine = w;
#pragma unroll
                    for (int o = 1; o < G; ++o) w = max(w, __shfl_sync(FULL, mine, grp * G + (lig + o) % G));
                }
                if (w >= 0) run_wrow = w;
            }
        }

        if (PIPE && my_pass >= 0 && npass > 1) {
            // one pass of a pipelined pair: leave the candidate, the pass that finishes last merges them in pass order
            // (first strictly greater score wins, as the sequential loop above does)
            if (lane == 0) {
                PassOut c;
                c.score = run_score; c.end_t = run_slot; c.end_q = run_row; c.amb = run_amb;
                a.vt_res[base] = c;
                __threadfence();
                const int done = atomicAdd(a.pair_done + pair, 1);
                if (done == npass - 1) {
                    __threadfence();
                    const volatile PassOut *rs = a.vt_res + (base - my_pass);
                    PassOut o;
                    o.score = -1; o.end_t = 0; o.end_q = 0; o.amb = 0;
                    for (int q = 0; q < npass; ++q) {
                        const int sc_q = rs[q].score;
                        if (sc_q > o.score) { o.score = sc_q; o.end_t = rs[q].end_t; o.end_q = rs[q].end_q; o.amb = rs[q].amb; }
                    }
                    o.amb = (o.amb && o.score > 0) ? 1 : 0;
                    a.out[pair] = o;
                    if (o.amb) {
                        const int pos = atomicAdd(a.amb_count, 1);
                        a.amb_list[a.seg_begin[a.amb_base_seg] + pos] = pair;
                    }
                }
            }
        } else if (active && lig == 0) {
            if (WATCH) {
                a.out[pair].end_q = run_wrow;
                a.out[pair].amb = run_wrow < 0 ? 2 : 0;
            } else {
                PassOut o;
                o.score = run_score;
                o.end_t = run_slot;
                o.end_q = run_row;
                o.amb = (run_amb && run_score > 0) ? 1 : 0;
                a.out[pair] = o;
                if (o.amb) {
                    const int pos = atomicAdd(a.amb_count, 1);
                    a.amb_list[a.seg_begin[a.amb_base_seg] + pos] = pair;
                }
            }
        }
        __syncwarp();
    }
}

}  // namespace db200
