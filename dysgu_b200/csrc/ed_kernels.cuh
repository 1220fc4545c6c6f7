// Bit-parallel edit distance kernels for sm_100a (Myers 1999 / Hyyro block formulation).
//
// They compute the observable results of dysgu's vendored edlib (reference:
// dysgu/edlib/src/edlib.cpp:141-296 driver, 411-443 word step, 546-709 HW/SHW, 735-944 NW):
// edit distance, every end location, and for HW the smallest start of each end location.
// The reference's Ukkonen band only skips cells that are provably > k, so the full-column
// evaluation done here yields identical results.
//
// Layout: the query's vertical delta vectors live in registers, WL 64-bit words per thread; a group
// of G threads (G = 1..32) covers 64*G*WL query symbols and runs a wavefront over the target
// columns (thread l works on column t-l at step t), the horizontal delta of a word boundary
// travels to the next thread with __shfl_up_sync.  32/G pairs share a warp.  The match vectors
// (Peq, one row per distinct target symbol) are built once per pair by ed_build_kernel.
#pragma once
#include "engine.cuh"

namespace db200 {

constexpr int ED_NCFG = 18;
constexpr int ED_LBINS = 256;   // target-length bins (64 columns each) per config, longest first

struct EdPair {
    int64_t q_off, t_off;   // offsets into the raw query / target byte buffers
    int64_t peq_off;        // u64 words: Peq rows [At][W]; reversed-query rows follow when needed
    int64_t hlog_off;       // u32 words: 2-bit horizontal deltas of the last row, 16 columns per word
    int32_t m, n;
    int32_t W;              // 64-bit words of the query
    int32_t At;             // distinct target symbols = Peq rows
    int32_t alphabet;       // distinct symbols of query and target (alphabetLength)
    int32_t k;              // -1: unlimited
    int32_t best;           // edit distance (or -1)
    int32_t nloc;
    int64_t loc_off;
    int32_t special;        // 1: handled without DP (an empty sequence)
    int32_t cfg;
    int32_t band_round;     // NW: first round of the banded kernel this pair enters (-1: full-column kernel only)
    int32_t pad_;
};

struct EdArgs {
    const EdPair *pairs;
    const uint64_t *peq;
    const uint8_t *tcodes;       // target re-coded as Peq row indices (same layout as the raw target buffer)
    uint32_t *hlog;
    const int32_t *tasks;        // forward: pair ids; reverse: location indices
    const int32_t *seg_begin;
    const int32_t *seg_count;
    int32_t seg;
    int32_t *counter;
    int32_t mode;                // DB200_ED_MODE_*
    int32_t *nw_score;           // forward, NW: final score per pair
    // reverse (start locations)
    const int32_t *loc_pair;     // location index -> pair
    int32_t *loc_buf;            // (start, end) pairs
    unsigned long long *wsteps;  // word steps evaluated (statistics)
};

// One Myers column step on a 64-bit word.  hin/hout are (plus, minus) bit pairs.
__device__ __forceinline__ void myers_word(uint64_t &Pv, uint64_t &Mv, uint64_t Eq, uint32_t hin_p, uint32_t hin_m, uint32_t &hout_p,
                                           uint32_t &hout_m, uint64_t &Ph_raw, uint64_t &Mh_raw)
{
    const uint64_t Xv = Eq | Mv;
    Eq |= (uint64_t)hin_m;
    const uint64_t Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
    uint64_t Ph = Mv | ~(Xh | Pv);
    uint64_t Mh = Pv & Xh;
    Ph_raw = Ph;
    Mh_raw = Mh;
    hout_p = (uint32_t)(Ph >> 63);
    hout_m = (uint32_t)(Mh >> 63);
    Ph = (Ph << 1) | (uint64_t)hin_p;
    Mh = (Mh << 1) | (uint64_t)hin_m;
    Pv = Mh | ~(Xv | Ph);
    Mv = Ph & Xv;
}

constexpr int ED_SMEM_SYMS = 8;   // match vectors of up to 8 distinct target symbols are staged in shared memory

template <int WL, int G, bool REV>
__global__ void __launch_bounds__(128) ed_myers_kernel(EdArgs a)
{
    constexpr int GROUPS = 32 / G;
    constexpr bool SMEM_PEQ = (WL <= 8);
    // per warp: [symbol][word][lane] 64-bit entries -> conflict-free LDS.64 whatever symbols the lanes read
    extern __shared__ unsigned long long ed_sm[];
    unsigned long long *peq_s = ed_sm + (size_t)(threadIdx.x >> 5) * (SMEM_PEQ ? ED_SMEM_SYMS * WL * 32 : 0);
    const int lane = threadIdx.x & 31;
    const int lig = lane % G, grp = lane / G;
    const unsigned FULL = 0xFFFFFFFFu;
    const int seg_begin = a.seg_begin[a.seg], seg_count = a.seg_count[a.seg];

    while (true) {
        // explicit warp barriers at the loop edges keep every unit running converged (see sw_pass_kernel)
        __syncwarp();
        int base = 0;
        if (lane == 0) base = atomicAdd(a.counter, GROUPS);
        __syncwarp();
        base = __shfl_sync(FULL, base, 0);
        if (base >= seg_count) break;
        const int ti = base + grp;
        const bool active = grp < GROUPS && ti < seg_count;
        int task = active ? a.tasks[seg_begin + ti] : -1;
        int pair = -1, end = 0, ncols = 0;
        EdPair pr;
        memset(&pr, 0, sizeof(pr));
        if (active) {
            pair = REV ? a.loc_pair[task] : task;
            pr = a.pairs[pair];
            if (REV) {
                end = a.loc_buf[2 * (int64_t)task + 1];
                ncols = min(end + 1, pr.m + pr.best);   // longer windows cannot reach the distance again
            } else {
                ncols = pr.n;
            }
        }
        const int W = pr.W, m = pr.m;
        const int w0 = lig * WL;
        const int nw = max(0, min(WL, W - w0));
        const bool has_last = active && nw > 0 && (w0 + nw == W);
        const int last_w = nw - 1;
        const int last_bit = (m - 1) & 63;
        const uint64_t *peq = a.peq + pr.peq_off + (REV ? (int64_t)pr.At * W : 0) + w0;
        const uint8_t *tc = a.tcodes + pr.t_off;
        // stage this lane's words of every symbol row (warp-uniform decision so that the shuffles below stay converged)
        bool use_smem = false;
        if (SMEM_PEQ) {
            const int at_max = __reduce_max_sync(FULL, pr.At);
            use_smem = at_max <= ED_SMEM_SYMS;
            if (use_smem) {
                __syncwarp();
#pragma unroll
                for (int w = 0; w < WL; ++w)
                    for (int c = 0; c < pr.At; ++c)
                        peq_s[(c * WL + w) * 32 + lane] = (w < nw) ? peq[(int64_t)c * W + w] : 0ull;
                __syncwarp();
            }
        }
        // forward streams read the re-coded target as 32-bit words (4 columns), shifted by the lane's column lag and the
        // byte misalignment of the pair's buffer; reverse streams (short windows) keep byte loads
        const int tphase = (int)(((uintptr_t)tc - (uintptr_t)lig) & 3u);
        const uint32_t *tcw = reinterpret_cast<const uint32_t *>((uintptr_t)tc & ~(uintptr_t)3);
        const int tmis = (int)((uintptr_t)tc & 3u);
        int twidx = (tmis - lig) >> 2;   // floor((tmis + (0 - lig)) / 4): word holding column -lig
        const int twords = (tmis + ncols + 3) >> 2;
        uint32_t tprev = (!REV && twidx >= 0 && twidx < twords) ? tcw[twidx] : 0u;
        uint32_t tcur = 0;
        const uint32_t hin0 = (REV || a.mode != DB200_ED_MODE_HW) ? 1u : 0u; // SHW / NW: gap before the query costs

        uint64_t Pv[WL], Mv[WL];
#pragma unroll
        for (int w = 0; w < WL; ++w) { Pv[w] = ~0ull; Mv[w] = 0ull; }
        int score = m;
        int rmin = (m & 63) ? m : 0x7FFFFFFF; // column -1 is only a candidate when the last word is padded (edlib.cpp:663-676)
        int rpos = -1;
        uint32_t logw = 0;
        uint32_t *hlog = a.hlog + pr.hlog_off;

        int steps = ncols + G - 1;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) steps = max(steps, __shfl_xor_sync(FULL, steps, o));
        uint32_t carry = 0; // bit0 = plus, bit1 = minus: horizontal delta leaving this thread's words

        for (int t = 0; t < steps; ++t) {
            const int j = t - lig;
            const bool valid = active && nw > 0 && j >= 0 && j < ncols;
            if (!REV && (t & 3) == 0) {
                ++twidx;
                const uint32_t tnew = (twidx >= 0 && twidx < twords) ? tcw[twidx] : 0u;
                tcur = __funnelshift_r(tprev, tnew, tphase * 8);
                tprev = tnew;
            }
            const int c_fwd = (int)((tcur >> ((t & 3) * 8)) & 0xFFu);
            uint32_t cin = carry;
            if (G > 1) cin = __shfl_up_sync(FULL, carry, 1);   // G need not divide 32: lanes past the last whole group idle
            if (lig == 0) cin = hin0;
            if (valid) {
                const int c = REV ? tc[end - j] : c_fwd;
                const uint64_t *row = peq + (int64_t)c * W;
                const unsigned long long *srow = peq_s + (c * WL) * 32 + lane;
                uint32_t hp = cin & 1u, hm = (cin >> 1) & 1u;
                // Every thread that owns at least one word steps all WL of them: the words past the end of the query hold
                // zero match vectors (or whatever follows in the pool) and feed nothing - only the last thread of a pair
                // has them and its carry goes nowhere.  No per-word predicate, no per-word select of the last row's word.
                uint64_t Phr[WL], Mhr[WL];
                if (SMEM_PEQ && use_smem) {
#pragma unroll
                    for (int w = 0; w < WL; ++w) {
                        uint32_t op, om;
                        myers_word(Pv[w], Mv[w], srow[w * 32], hp, hm, op, om, Phr[w], Mhr[w]);
                        hp = op; hm = om;
                    }
                } else {
#pragma unroll
                    for (int w = 0; w < WL; ++w) {
                        uint32_t op, om;
                        myers_word(Pv[w], Mv[w], __ldg(row + w), hp, hm, op, om, Phr[w], Mhr[w]);
                        hp = op; hm = om;
                    }
                }
                carry = hp | (hm << 1);
                if (has_last) {   // one lane per pair: the horizontal delta of the query's last row
                    uint64_t Ph_l = Phr[0], Mh_l = Mhr[0];
#pragma unroll
                    for (int w = 1; w < WL; ++w)
                        if (w == last_w) { Ph_l = Phr[w]; Mh_l = Mhr[w]; }
                    const int dp = (int)((Ph_l >> last_bit) & 1ull), dm = (int)((Mh_l >> last_bit) & 1ull);
                    score += dp - dm;
                    if (REV) {
                        if (score <= rmin) { rmin = score; rpos = j; }
                    } else if (a.mode != DB200_ED_MODE_NW) {
                        logw |= (uint32_t)(dp | (dm << 1)) << ((j & 15) * 2);
                        if ((j & 15) == 15 || j == ncols - 1) { hlog[j >> 4] = logw; logw = 0; }
                    }
                }
            }
        }
        if (has_last) {
            if (REV) a.loc_buf[2 * (int64_t)task] = end - rpos;
            else if (a.mode == DB200_ED_MODE_NW) a.nw_score[pair] = score;
            if (!REV) atomicAdd(a.wsteps, (unsigned long long)ncols * (unsigned long long)W);   // forward passes only
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// Banded global (NW) edit distance: one thread per pair, a window of WB 64-row words in registers that slides
// down the diagonal.  Same results as the reference's Ukkonen-banded block algorithm (edlib.cpp:735-944), by a
// different argument: cells outside the window are replaced by upper bounds (a new bottom word starts with all
// vertical deltas +1, the row above the window advances by +1 per column), so every computed value is >= the
// true one and exact on every alignment path that stays inside the window.  A path of cost <= c from corner to
// corner visits only diagonals j - i in [-(c - D) / 2, (c + D) / 2] (D = n - m); the window placed at
// 64 * floor((j - (c + D) / 2) / 64) covers them for c = 64 * (WB - 1).  Hence: banded distance <= c  =>  it is
// the edit distance; banded distance > c  =>  the edit distance is > c and the pair moves to the next, wider
// round (2, 3, 5, 9, 17 words, then the full-column kernel) - the same doubling edlib applies to k (edlib.cpp:199-214).
// ------------------------------------------------------------------------------------------------
constexpr int ED_BAND_ROUNDS = 5;
__host__ __device__ constexpr int ed_band_words(int r) { return r == 0 ? 2 : (r == 1 ? 3 : (r == 2 ? 5 : (r == 3 ? 9 : 17))); }

struct EdBandArgs {
    const EdPair *pairs;
    const uint64_t *peq;
    const uint8_t *tcodes;
    const int32_t *tasks;        // pairs entering at this round, longest targets first (segment `round` of the bucket list)
    const int32_t *seg_begin;
    const int32_t *seg_count;
    const int32_t *carry;        // pairs handed on by the previous round
    const int32_t *carry_count;
    int32_t *next;               // where this round hands its failures: the next round's carry list ...
    int32_t *next_count;
    int32_t *fb_tasks;           // ... or, after the last round, the full-column kernel's task list of the pair's configuration
    int32_t *fb_count;           // [ED_NCFG]
    int64_t fb_stride;
    int32_t round;
    int32_t *nw_score;
    unsigned long long *wsteps;  // word steps evaluated (statistics)
};

// Columns whose target symbols and match-vector words are fetched together before any of them is stepped: the loads of
// a batch are independent (the window position is a function of the column index alone), so their latencies overlap
// instead of adding up once per column - ncu of the one-column version showed 12.8 of 16.7 cycles per issued
// instruction waiting on the long scoreboard (profiles/r02_ed_banded_nw_before.txt).
__host__ __device__ constexpr int ed_band_cols(int wb) { return wb <= 3 ? 8 : (wb <= 5 ? 4 : 2); }

template <int WB, bool LAST>
__global__ void __launch_bounds__(128) ed_banded_nw_kernel(EdBandArgs a)
{
    constexpr int CB = ed_band_cols(WB);
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_first = a.seg_count[a.round];
    const int n_carry = a.carry_count ? *a.carry_count : 0;
    if (idx >= n_first + n_carry) return;
    const int pair = idx < n_first ? a.tasks[a.seg_begin[a.round] + idx] : a.carry[idx - n_first];
    const EdPair pr = a.pairs[pair];
    const int m = pr.m, n = pr.n, W = pr.W;
    const int nwin = min(WB, W);
    const bool full = W <= WB;                    // the window is the whole column: exact whatever the distance
    const int c = 64 * (WB - 1);
    const int H = (c + (n - m)) >> 1;             // rows above the diagonal a cost-c path can reach (routing guarantees c >= |n - m|)
    const uint64_t *peq = a.peq + pr.peq_off;
    const uint8_t *tc = a.tcodes + pr.t_off;

    uint64_t Pv[WB], Mv[WB];
#pragma unroll
    for (int w = 0; w < WB; ++w) { Pv[w] = ~0ull; Mv[w] = 0ull; }
    int wb0 = 0;
    int score = min(64 * nwin, m);                // value of the window's bottom row in column -1
    const int wb_max = W - nwin;
    bool failed = false;
    int jdone = n;
    for (int j0 = 0; j0 < n && !failed; j0 += CB) {
        // ---- fetch the batch: symbols, then the window's match-vector words of every column ----------------------
        uint64_t eq[CB][WB];
        int wbs[CB];
#pragma unroll
        for (int b = 0; b < CB; ++b) {
            const int j = min(j0 + b, n - 1);
            wbs[b] = full ? 0 : min(max(j - H, 0) >> 6, wb_max);
            const uint64_t *row = peq + (int64_t)tc[j] * W + wbs[b];
#pragma unroll
            for (int w = 0; w < WB; ++w) eq[b][w] = (w < nwin) ? __ldg(row + w) : 0ull;
        }
        // ---- step the columns --------------------------------------------------------------------------------------
#pragma unroll
        for (int b = 0; b < CB; ++b) {
            const int j = j0 + b;
            if (j >= n || failed) break;
            if (wbs[b] > wb0) {                   // the window moves down by one word
                const int old_b = 64 * (wb0 + nwin) - 1;
#pragma unroll
                for (int w = 0; w + 1 < WB; ++w) { Pv[w] = Pv[w + 1]; Mv[w] = Mv[w + 1]; }
                Pv[WB - 1] = ~0ull; Mv[WB - 1] = 0ull;
                ++wb0;
                score += min(64 * (wb0 + nwin), m) - 1 - old_b;
            }
            uint32_t hp = 1u, hm = 0u;
            uint64_t Ph_l = 0, Mh_l = 0;
#pragma unroll
            for (int w = 0; w < WB; ++w) {
                if (w < nwin) {
                    uint32_t op, om;
                    myers_word(Pv[w], Mv[w], eq[b][w], hp, hm, op, om, Ph_l, Mh_l);
                    hp = op; hm = om;
                }
            }
            const int bit = (wb0 + nwin == W) ? ((m - 1) & 63) : 63;
            score += (int)((Ph_l >> bit) & 1ull) - (int)((Mh_l >> bit) & 1ull);
            // Ukkonen cut-off, once per 64 columns: every path through column j costs at least the value on the corner's
            // diagonal (row j - (n - m)): vertical neighbours differ by at most 1 and a detour of r rows costs r again on
            // the way back.  Once that value exceeds c the round cannot succeed.
            if (!full && (j & 63) == 63) {
                const int i0 = j - (n - m);
                if (i0 >= 0) {
                    int v = score;
#pragma unroll
                    for (int w = 0; w < WB; ++w) {
                        const int lo = i0 + 1 - 64 * (wb0 + w);          // rows of this word above the diagonal row are excluded
                        uint64_t mask = lo <= 0 ? ~0ull : (lo >= 64 ? 0ull : (~0ull << lo));
                        if (w == WB - 1 && bit != 63) mask &= (2ull << bit) - 1ull;
                        v -= __popcll(Pv[w] & mask) - __popcll(Mv[w] & mask);
                    }
                    if (v > c) { score = v; failed = true; jdone = j + 1; }
                }
            }
        }
    }
    {   // statistics: one atomic per converged group of threads instead of one per pair
        const unsigned mask = __activemask();
        const unsigned total = __reduce_add_sync(mask, (unsigned)jdone * (unsigned)nwin);
        if ((threadIdx.x & 31) == __ffs(mask) - 1) atomicAdd(a.wsteps, (unsigned long long)total);
    }
    if (full || (!failed && score <= c) || (pr.k >= 0 && c >= pr.k)) {
        a.nw_score[pair] = score;                 // exact, or known to exceed k (ed_count_kernel turns that into -1)
    } else if (!LAST) {
        a.next[atomicAdd(a.next_count, 1)] = pair;
    } else {
        a.fb_tasks[(int64_t)pr.cfg * a.fb_stride + atomicAdd(&a.fb_count[pr.cfg], 1)] = pair;
    }
}

}  // namespace db200
