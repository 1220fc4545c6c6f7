// Batched Smith-Waterman pipeline on the device: encode/pack -> bucket -> scoring pass ->
// (re-run of ambiguous pairs) -> reversed extraction -> reverse scoring pass -> second-best /
// finalize -> banded traceback -> cigar compaction.  Host side only enqueues kernels.
//
// Reference behaviour being reproduced (dysgu v1.9.0):
//   _ssw_wrapper.pyx:60-68 (encoding), 583-588 (mask length), ssw.c:772-857 (driver),
//   ssw.c:124-346 / 372-548 (scoring passes), ssw.c:550-728 (banded traceback).
#include "sw_launch.cuh"

namespace db200 {

// per-call view of the configuration table (sw_launch.cuh) for the bucketing kernels
struct CfgSel {
    uint16_t cap[SW_NCFG];    // columns per pass
    uint16_t cost[SW_NCFG];   // relative issue cost of one row of one pair; 0 = not usable in this call
    uint8_t skew[SW_NCFG];    // 2G - 1: wavefront steps beyond the rows
};
constexpr int SW_SINGLE_PASS_MAX = 1024;   // widest single-pass configuration (K = 16, G = 32); longer targets are column-tiled

struct PairState {          // per-pair bookkeeping of the pipeline
    int32_t mask_len;
    int32_t need_colmax;
    int32_t use_word;
    int32_t status;
    int32_t want_rev;       // reverse pass required
    int32_t want_cigar;     // traceback required
    int32_t cigar_len;
    int32_t trace_bw;       // next band half-width the traceback has to try (travels with a pair between levels)
    int64_t cigar_pos;      // position in the cigar pool
};

// ------------------------------------------------------------------------------------------------
// 1. sequence units + packing (ASCII -> 4-bit codes, _ssw_wrapper.pyx:60-68)
// ------------------------------------------------------------------------------------------------
// A side of the batch is either CSR (len == nullptr: sequence p = buf[off[p], off[p + 1])) or slices of a shared source
// buffer (sequence p = buf[off[p], off[p] + len[p]): overlapping windows of one reference region, re_map.py:429-434).
__global__ void sw_units_kernel(const int64_t *qoff, const int64_t *toff, const int32_t *qlen, const int32_t *tlen, int64_t npairs, int32_t *units)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= 2 * npairs) return;
    const int64_t *off = i < npairs ? qoff : toff;
    const int32_t *lens = i < npairs ? qlen : tlen;
    int64_t p = i < npairs ? i : i - npairs;
    int64_t len = lens ? (int64_t)lens[p] : off[p + 1] - off[p];
    units[i] = (int32_t)((len + 31) >> 5);
}

__device__ __forceinline__ uint32_t nt_code(uint8_t ch)
{
    switch (ch) {
        case 'A': case 'a': case 'U': case 'u': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return 4;
    }
}

// Four ASCII bytes -> four code bytes (0..4), branch-free.  Folding bit 5 maps lower to upper case; the low three bits
// of the five letters are distinct (A=1 C=3 G=7 T=4 U=5), so two byte permutes look up the letter each index stands
// for and its code; a byte whose folded value differs from that letter is not a nucleotide and becomes 4 (N).
__device__ __forceinline__ uint32_t nibbles_of_bytes(uint32_t v)   // four bytes (each < 16) -> 16 bits of nibbles
{
    v = (v | (v >> 4)) & 0x00FF00FFu;
    return (v | (v >> 8)) & 0xFFFFu;
}

__device__ __forceinline__ uint32_t nt_codes4(uint32_t x)
{
    const uint32_t u = x & 0xDFDFDFDFu;
    const uint32_t sel = nibbles_of_bytes(u & 0x07070707u);
    const uint32_t expect = __byte_perm(0x43FF41FFu, 0x47FF5554u, sel);   // idx: -, A, -, C | T, U, -, G  (0xFF never matches: bit 5 set)
    const uint32_t code = __byte_perm(0x01040004u, 0x02040003u, sel);     // idx: 4, 0, 4, 1 | 3, 0, 4, 2
    const uint32_t y = expect ^ u;
    const uint32_t nz = (((y & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | y) & 0x80808080u;   // high bit where the byte differs
    const uint32_t bad = nz >> 7;                                         // 0x01 per non-nucleotide byte
    return (code & ~(bad * 0xFFu)) | (bad << 2);
}

// one warp per sequence; each lane turns 8 bases into one 32-bit word per iteration.  The bytes arrive through two
// aligned 64-bit loads and a funnel shift (sequences start at arbitrary byte offsets of the CSR buffer).
__global__ void __launch_bounds__(256) sw_pack_kernel(const uint8_t *qbuf, const int64_t *qoff, const uint8_t *tbuf, const int64_t *toff,
                                                      const int32_t *qlen, const int32_t *tlen,
                                                      int64_t npairs, const int64_t *unit_off, uint32_t *pool, PairMeta *meta)
{
    const int64_t wid = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wid >= 2 * npairs) return;
    const bool isq = wid < npairs;
    const int64_t p = isq ? wid : wid - npairs;
    const int64_t *off = isq ? qoff : toff;
    const int32_t *lens = isq ? qlen : tlen;
    const uint8_t *src = (isq ? qbuf : tbuf) + off[p];
    const int64_t len = lens ? (int64_t)lens[p] : off[p + 1] - off[p];
    const int64_t u0 = unit_off[wid];
    const int64_t nwords = ((len + 31) >> 5) << 2;
    const int64_t fullw = (len + 7) >> 3;          // words holding at least one base
    uint32_t *dst = pool + u0 * 4;
    const unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(src) & 7u);
    const uint8_t *abase = src - mis;
    const uint8_t *aend = src + len;               // first byte that must not be the start of an aligned load
    const unsigned sh = mis * 8u;
    for (int64_t w = lane; w < nwords; w += 32) {
        uint32_t x = 0x55555555u;                  // padding code in every nibble
        if (w < fullw) {
            const uint8_t *a = abase + w * 8;
            const uint2 lo = *reinterpret_cast<const uint2 *>(a);
            uint2 hi = make_uint2(0u, 0u);
            if (mis != 0 && a + 8 < aend) hi = *reinterpret_cast<const uint2 *>(a + 8);
            uint32_t b0, b1;
            if (sh >= 32u) { b0 = __funnelshift_r(lo.y, hi.x, sh - 32u); b1 = __funnelshift_r(hi.x, hi.y, sh - 32u); }
            else { b0 = __funnelshift_r(lo.x, lo.y, sh); b1 = __funnelshift_r(lo.y, hi.x, sh); }
            x = nibbles_of_bytes(nt_codes4(b0)) | (nibbles_of_bytes(nt_codes4(b1)) << 16);
            const int64_t rem = len - w * 8;       // bases in this word
            if (rem < 8) {
                const uint32_t keep = (1u << (4 * (int)rem)) - 1u;
                x = (x & keep) | (0x55555555u & ~keep);
            }
        }
        dst[w] = x;
    }
    if (lane == 0) {
        if (isq) { meta[p].q_unit = (uint32_t)u0; meta[p].m = (int32_t)len; }
        else { meta[p].t_unit = (uint32_t)u0; meta[p].n = (int32_t)len; }
    }
}

// pre-packed input (db200_sw_batch_*_packed): the pool content arrived as is, only the pair records are built here
__global__ void sw_meta_kernel(const int32_t *qlen, const int32_t *tlen, int64_t npairs, const int64_t *unit_off, PairMeta *meta)
{
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= npairs) return;
    PairMeta pm;
    pm.q_unit = (uint32_t)unit_off[p]; pm.m = qlen[p];
    pm.t_unit = (uint32_t)unit_off[npairs + p]; pm.n = tlen[p];
    meta[p] = pm;
}

// ------------------------------------------------------------------------------------------------
// 2. bucketing: counting sort of pairs by (config, needs column maxima, descending length bin)
// ------------------------------------------------------------------------------------------------
// cheapest usable configuration whose pass holds n columns: cost per row x (rows + wavefront skew).  With the fine
// configurations switched off the costs grow with the capacity, so this is "smallest capacity that fits" as before.
__device__ __forceinline__ int sw_cfg_for(const CfgSel &s, int m, int n)
{
    int best = SW_CFG_MULTI;
    unsigned long long bw = ~0ull;
    for (int c = 0; c < SW_CFG_MULTI; ++c) {
        if (s.cost[c] == 0 || (int)s.cap[c] < n) continue;
        const unsigned long long w = (unsigned long long)s.cost[c] * (unsigned long long)(m + (int)s.skew[c]);
        if (w < bw) { bw = w; best = c; }
    }
    return best;
}

__device__ __forceinline__ int sw_bin_key(const CfgSel &sel, int m, int n, int need_colmax)
{
    const int c = sw_cfg_for(sel, m, n);
    const int lb = min(m >> 5, SW_LBINS - 1);
    return ((c * 2 + need_colmax) * SW_LBINS) + (SW_LBINS - 1 - lb);
}

struct InitArgs {
    const PairMeta *meta;
    const int32_t *mask_len;   // may be null -> auto
    PairState *state;
    int64_t npairs;
    int32_t max_match;         // largest matrix entry
    int32_t flag;
    CfgSel sel;
};

// per pair: status, mask length, colmax need, histogram of forward bins
__global__ void sw_init_kernel(InitArgs a, int32_t *hist)
{
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= a.npairs) return;
    const PairMeta pm = a.meta[p];
    PairState s;
    memset(&s, 0, sizeof(s));
    s.mask_len = a.mask_len ? a.mask_len[p] : max(pm.m / 2, 15);
    s.status = DB200_PAIR_OK;
    if (pm.m <= 0 || pm.n <= 0) s.status = DB200_PAIR_EMPTY;
    else if ((int64_t)a.max_match * min(pm.m, pm.n) > 30000) s.status = DB200_PAIR_SCORE_RANGE;
    // column maxima are only needed when a column can fall outside the mask window (ssw.c:326-341)
    s.need_colmax = (s.mask_len >= 15 && pm.n > s.mask_len - 1) ? 1 : 0;
    a.state[p] = s;
    if (s.status == DB200_PAIR_OK) atomicAdd(&hist[sw_bin_key(a.sel, pm.m, pm.n, s.need_colmax)], 1);
}

// single block: exclusive scan of the bin histogram, segment ranges
// a warp per segment at a time: totals by warp reduction, bin starts by a warp scan over 32 bins at a time
constexpr int SW_BINSCAN_THREADS = 1024;
__global__ void __launch_bounds__(SW_BINSCAN_THREADS) sw_binscan_kernel(const int32_t *hist, int32_t *bin_start, int32_t *seg_begin, int32_t *seg_count)
{
    __shared__ int32_t seg_tot[SW_NSEG];
    __shared__ int32_t seg_off[SW_NSEG + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int NW = SW_BINSCAN_THREADS / 32;
    for (int seg = warp; seg < SW_NSEG; seg += NW) {
        const int32_t *h = hist + seg * SW_LBINS;
        int s = 0;
        for (int b = lane; b < SW_LBINS; b += 32) s += h[b];
        s = __reduce_add_sync(0xFFFFFFFFu, s);
        if (lane == 0) seg_tot[seg] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int k = 0; k < SW_NSEG; ++k) { seg_off[k] = acc; acc += seg_tot[k]; }
        seg_off[SW_NSEG] = acc;
    }
    __syncthreads();
    for (int seg = warp; seg < SW_NSEG; seg += NW) {
        if (lane == 0) { seg_begin[seg] = seg_off[seg]; seg_count[seg] = seg_tot[seg]; }
        if (seg_tot[seg] == 0) continue;   // nothing is scattered into an empty segment
        const int32_t *h = hist + seg * SW_LBINS;
        int acc = seg_off[seg];
        for (int b0 = 0; b0 < SW_LBINS; b0 += 32) {
            const int v = h[b0 + lane];
            int inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += t; }
            bin_start[seg * SW_LBINS + b0 + lane] = acc + inc - v;
            acc += __shfl_sync(0xFFFFFFFFu, inc, 31);
        }
    }
}

__global__ void sw_scatter_kernel(const PairMeta *meta, const PairState *state, int64_t npairs, int which, const CfgSel sel,
                                  const int32_t *bin_start, int32_t *bin_fill, int32_t *tasks)
{
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= npairs) return;
    const PairState &s = state[p];
    if (s.status != DB200_PAIR_OK) return;
    if (which == 1 && !s.want_rev) return;
    const PairMeta pm = meta[p];
    const int key = sw_bin_key(sel, pm.m, pm.n, which == 0 ? s.need_colmax : 0);
    const int pos = bin_start[key] + atomicAdd(&bin_fill[key], 1);
    tasks[pos] = (int32_t)p;
}

// ------------------------------------------------------------------------------------------------
// 3. after the forward pass: decide what else a pair needs, build the reversed problem
//    (ssw.c:825-841: flag handling; ssw.c:827-835: reversed read prefix, columns ref_end1..0)
// ------------------------------------------------------------------------------------------------
struct MidArgs {
    const PassOut *fwd;
    const PairMeta *meta;
    PairMeta *rmeta;
    PairState *state;
    int64_t npairs;
    int32_t bias, score_size, flag, filters;
    int32_t max_match, gap_open, gap_extend;
    uint32_t *unit_cursor;   // bump allocator for reversed sequences (units)
    uint32_t unit_base;
    CfgSel sel;
    const uint32_t *pool;    // packed codes (perfect-diagonal shortcut)
    PassOut *rev;
    int32_t uniform_match;   // the matrix is match / mismatch with zeros for N (no custom matrix): shortcut allowed
};

// eight packed codes ending at position p >= 7 of the sequence starting at word array w (position p in the top nibble)
__device__ __forceinline__ uint32_t codes_window(const uint32_t *w, int p)
{
    const int a = p - 7;
    return __funnelshift_r(w[a >> 3], w[p >> 3], (a & 7) * 4);
}

// Does the alignment ending at (end_q, end_t) consist of L identical A/C/G/T pairs?  Then its score is L x match, no
// alignment over fewer target columns or fewer query rows can reach that score, and the reverse pass of ssw.c:827-835
// (first reversed column whose maximum reaches the score, smallest row in it) must stop at (L - 1, L - 1).
__device__ __forceinline__ bool perfect_diagonal(const uint32_t *qw, const uint32_t *tw, int end_q, int end_t, int L)
{
    int eq = end_q, et = end_t, left = L;
    while (left >= 8) {
        const uint32_t a = codes_window(qw, eq), b = codes_window(tw, et);
        if ((a ^ b) != 0u || (a & 0x44444444u) != 0u) return false;
        eq -= 8; et -= 8; left -= 8;
    }
    for (; left > 0; --left, --eq, --et) {
        const uint32_t a = (qw[eq >> 3] >> ((eq & 7) * 4)) & 15u, b = (tw[et >> 3] >> ((et & 7) * 4)) & 15u;
        if (a != b || a >= 4u) return false;
    }
    return true;
}

// what a pair still needs after the forward pass; returns the 16-byte units its reversed problem takes (0: none)
__device__ __forceinline__ uint32_t sw_mid_pair(const MidArgs &a, int64_t p, PairMeta &r, uint32_t &uq)
{
    PairState &s = a.state[p];
    if (s.status != DB200_PAIR_OK) return 0;
    const PairMeta pm = a.meta[p];
    const PassOut fw = a.fwd[p];
    const int score = fw.score;
    const bool overflow8 = score + a.bias >= 255;
    s.use_word = (a.score_size == 1) || (a.score_size == 2 && overflow8);
    if (a.score_size == 0 && overflow8) { s.status = DB200_PAIR_OVERFLOW; return 0; }
    const bool want_begin = !(a.flag == 0 || (a.flag == 2 && score < a.filters));
    s.want_rev = 0;
    s.want_cigar = 0;
    if (!want_begin) return 0;
    if (score == 0) return 0; // degenerate: handled in finalize (no cells to walk)
    if (a.uniform_match && score % a.max_match == 0) {
        const int L = score / a.max_match;
        if (L <= fw.end_q + 1 && L <= fw.end_t + 1 &&
            perfect_diagonal(a.pool + (size_t)pm.q_unit * 4, a.pool + (size_t)pm.t_unit * 4, fw.end_q, fw.end_t, L)) {
            PassOut o;
            o.score = score; o.end_t = L - 1; o.end_q = L - 1; o.amb = 0;
            a.rev[p] = o;      // what the reverse pass would report; no reversed problem is built
            return 0;
        }
    }
    s.want_rev = 1;
    // rows the reverse pass can possibly need: the alignment spans at most tspan columns and
    // g query-only gap bases with  score <= max_match*tspan - (gap_open + (g-1)*gap_extend)
    const int tspan = fw.end_t + 1;
    int g = (a.max_match * tspan - score - a.gap_open) / a.gap_extend + 1;
    if (a.max_match * tspan - score - a.gap_open < 0) g = 0;
    const int rows = min(fw.end_q + 1, tspan + max(g, 0));
    uq = (uint32_t)((rows + 31) >> 5);
    const uint32_t ut = (uint32_t)((tspan + 31) >> 5);
    r.m = rows; r.n = tspan;
    return uq + ut;
}

// The units of the reversed sequences come from one bump cursor: a warp adds up what its pairs need and takes it with one
// atomic (one returning atomic per pair on a single address took 0.2 ms per 524,288 pairs - as long as the rest of the stage).
__global__ void sw_mid_kernel(MidArgs a, int32_t *hist)
{
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    PairMeta r;
    r.q_unit = 0; r.t_unit = 0; r.m = 0; r.n = 0;
    uint32_t uq = 0;
    const uint32_t need = p < a.npairs ? sw_mid_pair(a, p, r, uq) : 0u;
    uint32_t incl = need;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
    const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    uint32_t base = 0;
    if (lane == 31 && total) base = atomicAdd(a.unit_cursor, total);
    base = __shfl_sync(0xFFFFFFFFu, base, 31);
    if (need) {
        const uint32_t u0 = a.unit_base + base + (incl - need);
        r.q_unit = u0; r.t_unit = u0 + uq;
        a.rmeta[p] = r;
        atomicAdd(&hist[sw_bin_key(a.sel, r.m, r.n, 0)], 1);
    }
}

// the eight codes of x in reversed order (byte reversal, then the two codes of every byte swapped)
__device__ __forceinline__ uint32_t codes_reversed(uint32_t x)
{
    x = __byte_perm(x, 0u, 0x0123);
    return ((x & 0x0F0F0F0Fu) << 4) | ((x >> 4) & 0x0F0F0F0Fu);
}

// words of R[i] = S[end - i], i < len (positions past len hold the padding code), written to dst
__device__ __forceinline__ void reverse_prefix_words(const uint32_t *src, int end, int len, uint32_t *dst, int lane)
{
    const int nwords = ((len + 31) >> 5) << 2;
    for (int w = lane; w < nwords; w += 32) {
        const int p = end - 8 * w;          // source position of the word's first code
        const int cnt = len - 8 * w;        // codes of this word that exist
        uint32_t x;
        if (cnt <= 0) x = 0x55555555u;
        else if (p >= 7) {
            x = codes_reversed(codes_window(src, p));
            if (cnt < 8) { const uint32_t keep = (1u << (4 * cnt)) - 1u; x = (x & keep) | (0x55555555u & ~keep); }
        } else {
            x = 0;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int i = p - e;
                const uint32_t c = (e < cnt && i >= 0) ? ((src[i >> 3] >> ((i & 7) * 4)) & 15u) : (uint32_t)SW_PAD_CODE;
                x |= c << (4 * e);
            }
        }
        dst[w] = x;
    }
}

// one warp per pair: write reversed prefixes as packed codes, a 32-bit word (eight codes) at a time
__global__ void sw_reverse_extract_kernel(const PairMeta *meta, const PairMeta *rmeta, const PairState *state,
                                          const PassOut *fwd, int64_t npairs, const uint32_t *src_pool, uint32_t *pool)
{
    const int64_t p = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (p >= npairs) return;
    const PairState &s = state[p];
    if (s.status != DB200_PAIR_OK || !s.want_rev) return;
    const PairMeta pm = meta[p], r = rmeta[p];
    const PassOut fw = fwd[p];
    // the forward sequences may be the caller's packed batch; reversed query prefix R[i] = Q[end_q - i], i < r.m, and
    // reversed target prefix R[j] = T[end_t - j], j < r.n
    reverse_prefix_words(src_pool + (size_t)pm.q_unit * 4, fw.end_q, r.m, pool + (size_t)r.q_unit * 4, lane);
    reverse_prefix_words(src_pool + (size_t)pm.t_unit * 4, fw.end_t, r.n, pool + (size_t)r.t_unit * 4, lane);
}

// ------------------------------------------------------------------------------------------------
// 4. finalize: coordinates, second best (ssw.c:326-341 / 530-543 incl. the padded profile rows,
//    ssw.c:109 / 364), flags -> traceback request
// ------------------------------------------------------------------------------------------------
struct FinalArgs {
    const PassOut *fwd;
    const PassOut *rev;
    const PairMeta *meta;
    PairState *state;
    const int16_t *colmax;
    const int64_t *colmax_off;
    db200_sw_result *results;
    int32_t *status_out;
    int64_t npairs;
    int32_t flag, filters, filterd;
    const int32_t *pipe_abort;   // set if a pipelined pass gave up waiting for its predecessor (never expected)
    int32_t tiled_from;          // targets longer than this run in the column-tiled kernel
    int32_t defer_second_from;   // targets longer than this leave the second-best scan to sw_second_best_kernel (0: never)
};

// second best over the columns outside the mask window (ssw.c:326-341 / 530-543): the per-column maxima of the query
// padded with P phantom rows (ssw.c:109 / 364).  Columns lo <= j < hi, stride `step`; first column holding the maximum.
__device__ __forceinline__ void second_best_scan(const int16_t *cm, const int16_t *lr, int P, int lo, int hi, int step, int &best, int &bref)
{
    for (int j = lo; j < hi; j += step) {
        int x = cm[j];
        for (int k = 0; k < P && j - 1 - k >= 0; ++k) x = max(x, (int)lr[j - 1 - k]);
        if (x > best) { best = x; bref = j; }
    }
}

__global__ void sw_finalize_kernel(FinalArgs a)
{
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= a.npairs) return;
    PairState &s = a.state[p];
    db200_sw_result r;
    r.score1 = 0; r.score2 = 0; r.ref_begin1 = -1; r.ref_end1 = -1; r.read_begin1 = -1; r.read_end1 = -1;
    r.ref_end2 = 0; r.cigar_len = 0;
    if (s.status != DB200_PAIR_OK) {
        a.results[p] = r;
        if (a.status_out) a.status_out[p] = s.status;
        return;
    }
    const PairMeta pm = a.meta[p];
    if (a.pipe_abort && *a.pipe_abort && pm.n > a.tiled_from) {
        // not a result: reported like scratch exhaustion so that the host entry point re-runs the pair in a small group
        s.status = DB200_PAIR_NOSCRATCH;
        a.results[p] = r;
        if (a.status_out) a.status_out[p] = s.status;
        return;
    }
    const PassOut fw = a.fwd[p], rv = a.rev[p];
    const int score = fw.score;
    r.score1 = score;
    if (score == 0) {
        // no positive cell: ssw.c:146 (8-bit: end_ref stays -1) vs ssw.c:389 (16-bit: 0); read end collapses to 0
        r.ref_end1 = s.use_word ? 0 : -1;
        r.read_end1 = 0;
    } else {
        r.ref_end1 = fw.end_t;
        r.read_end1 = fw.end_q;
    }
    // ---- second best ----------------------------------------------------------------------------
    if (s.mask_len < 15) { r.score2 = 0; r.ref_end2 = -1; }
    else if (s.need_colmax && score > 0) {
        const int16_t *cm = a.colmax + a.colmax_off[p];
        const int16_t *lr = cm + pm.n;
        const int L = s.use_word ? 8 : 16;
        const int P = ((pm.m + L - 1) / L) * L - pm.m; // padded profile rows
        if (a.defer_second_from > 0 && pm.n > a.defer_second_from) {
            s.need_colmax = 2;   // long target: a warp scans the columns (sw_second_best_kernel)
        } else {
            int best = 0, bref = 0;
            second_best_scan(cm, lr, P, 0, max(r.ref_end1 - s.mask_len, 0), 1, best, bref);
            second_best_scan(cm, lr, P, min(r.ref_end1 + s.mask_len, pm.n) + (s.use_word ? 0 : 1), pm.n, 1, best, bref);
            r.score2 = best;
            r.ref_end2 = bref;
        }
    }
    // ---- begin coordinates ------------------------------------------------------------------------
    const bool want_begin = !(a.flag == 0 || (a.flag == 2 && score < a.filters));
    s.want_cigar = 0;
    if (want_begin) {
        if (score == 0) {
            // reverse pass over zero (8-bit) or one (16-bit) column: ssw.c:828-839
            r.ref_begin1 = s.use_word ? 0 : -1;
            r.read_begin1 = 0;
        } else {
            r.ref_begin1 = fw.end_t - rv.end_t;
            r.read_begin1 = fw.end_q - rv.end_q;
        }
        const bool stop = ((7 & a.flag) == 0) || ((2 & a.flag) != 0 && score < a.filters) ||
                          ((4 & a.flag) != 0 && (r.ref_end1 - r.ref_begin1 > a.filterd || r.read_end1 - r.read_begin1 > a.filterd));
        if (!stop) s.want_cigar = 1;
    }
    a.results[p] = r;
    if (a.status_out) a.status_out[p] = s.status;
}

// one warp per pair marked by sw_finalize_kernel: lanes stride over the columns, then (largest value, smallest column)
__global__ void sw_second_best_kernel(FinalArgs a)
{
    const int64_t p = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (p >= a.npairs) return;
    const PairState &s = a.state[p];
    if (s.status != DB200_PAIR_OK || s.need_colmax != 2) return;
    const PairMeta pm = a.meta[p];
    const int ref_end1 = a.results[p].ref_end1;
    const int16_t *cm = a.colmax + a.colmax_off[p];
    const int16_t *lr = cm + pm.n;
    const int L = s.use_word ? 8 : 16;
    const int P = ((pm.m + L - 1) / L) * L - pm.m;
    int best = 0, bref = 0;
    second_best_scan(cm, lr, P, lane, max(ref_end1 - s.mask_len, 0), 32, best, bref);
    second_best_scan(cm, lr, P, min(ref_end1 + s.mask_len, pm.n) + (s.use_word ? 0 : 1) + lane, pm.n, 32, best, bref);
    // value in the high bits, complemented column in the low bits: the maximum is the first column holding the best value
    unsigned long long key = ((unsigned long long)(unsigned)best << 32) | (unsigned)(0x7FFFFFFF - bref);
    for (int o = 16; o >= 1; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xFFFFFFFFu, key, o);
        key = other > key ? other : key;
    }
    if (lane == 0) {
        const int b = (int)(key >> 32);
        a.results[p].score2 = b;
        a.results[p].ref_end2 = b > 0 ? 0x7FFFFFFF - (int)(key & 0xFFFFFFFFu) : 0;
    }
}

}  // namespace db200
#include "sw_trace.cuh"
#include "sw_small.cuh"
namespace db200 {

// ------------------------------------------------------------------------------------------------
// latency path: up to SW_SMALL_MAX_PAIRS short pairs in one launch (sw_small.cuh)
// ------------------------------------------------------------------------------------------------
// Returns DB200_OK with *handled = 1 when the batch was taken (pairs flagged SW_SMALL_FALLBACK in status[] are left to the
// caller), *handled = 0 when the batch is not eligible.
int sw_small_host(Engine *e, const db200_sw_params *prm, const uint8_t *qbuf, const int64_t *qoff, const uint8_t *tbuf,
                  const int64_t *toff, int64_t npairs, const int32_t *mask_len, db200_sw_result *results, uint32_t *cigar_buf,
                  int64_t cigar_cap, int64_t *cigar_off, int32_t *status, int *handled)
{
    *handled = 0;
    if (npairs <= 0 || npairs > SW_SMALL_MAX_PAIRS || getenv("DB200_NO_SMALL_PATH")) return DB200_OK;
    if (!(prm->flag == 0 || prm->flag == 1 || prm->flag == 8 || prm->flag == 9)) return DB200_OK;
    if (prm->gap_open <= 0 || prm->gap_extend <= 0 || prm->gap_open > 255 || prm->gap_extend > 255 || (prm->mat && prm->n != 5) ||
        prm->score_size < 0 || prm->score_size > 2)
        return DB200_OK;   // the batch pipeline reports the argument error
    int64_t qbytes = 0, tbytes = 0;
    int maxq = 0, eligible = 0;
    for (int64_t i = 0; i < npairs; ++i) {
        const int64_t ql = qoff[i + 1] - qoff[i], tl = toff[i + 1] - toff[i];
        if (ql < 0 || tl < 0) return DB200_OK;
        if (ql <= SW_SMALL_MAX_Q && tl <= SW_SMALL_MAX_T) { ++eligible; maxq = std::max<int>(maxq, (int)ql); qbytes += ql; tbytes += tl; }
        else { qbytes += 0; }
    }
    if (eligible * 2 < npairs) return DB200_OK;   // mostly long pairs: one trip through the batch pipeline
    // staging area in page-locked memory mapped into the device address space
    const size_t in_bytes = (size_t)(qbytes + tbytes) + sizeof(int32_t) * (size_t)(3 * npairs + 2) + 64;
    const size_t out_bytes = (size_t)npairs * (sizeof(db200_sw_result) + 2 * sizeof(int32_t) + sizeof(uint32_t) * SW_SMALL_CIGAR_STRIDE) + 64;
    const size_t need = ((in_bytes + 255) & ~(size_t)255) + out_bytes;
    if (e->small_cap < need) {
        if (e->small_host) cudaFreeHost(e->small_host);
        e->small_host = nullptr; e->small_cap = 0;
        const size_t cap = std::max<size_t>(need * 2, (size_t)1 << 20);
        if (cudaHostAlloc(&e->small_host, cap, cudaHostAllocMapped) != cudaSuccess) { cudaGetLastError(); return DB200_OK; }
        if (cudaHostGetDevicePointer(&e->small_dev, e->small_host, 0) != cudaSuccess) { cudaGetLastError(); cudaFreeHost(e->small_host); e->small_host = nullptr; return DB200_OK; }
        e->small_cap = cap;
    }
    uint8_t *h = reinterpret_cast<uint8_t *>(e->small_host), *d = reinterpret_cast<uint8_t *>(e->small_dev);
    int32_t *h_qoff = reinterpret_cast<int32_t *>(h), *h_toff = h_qoff + npairs + 1, *h_mask = h_toff + npairs + 1;
    uint8_t *h_seq = reinterpret_cast<uint8_t *>(h_mask + npairs);
    int32_t pos = 0;
    for (int64_t i = 0; i < npairs; ++i) {     // ineligible pairs travel as empty-length placeholders and are re-flagged below
        const int64_t ql = qoff[i + 1] - qoff[i], tl = toff[i + 1] - toff[i];
        const bool ok = ql <= SW_SMALL_MAX_Q && tl <= SW_SMALL_MAX_T;
        h_qoff[i] = pos;
        if (ok) { memcpy(h_seq + pos, qbuf + qoff[i], (size_t)ql); pos += (int32_t)ql; }
        h_mask[i] = mask_len ? mask_len[i] : std::max<int32_t>((int32_t)(ql / 2), 15);
    }
    h_qoff[npairs] = pos;
    for (int64_t i = 0; i < npairs; ++i) {
        const int64_t ql = qoff[i + 1] - qoff[i], tl = toff[i + 1] - toff[i];
        const bool ok = ql <= SW_SMALL_MAX_Q && tl <= SW_SMALL_MAX_T;
        h_toff[i] = pos;
        if (ok) { memcpy(h_seq + pos, tbuf + toff[i], (size_t)tl); pos += (int32_t)tl; }
    }
    h_toff[npairs] = pos;
    // the query block ends where the target block starts: qoff[n] == toff[0]
    const size_t out0 = (in_bytes + 255) & ~(size_t)255;
    db200_sw_result *h_res = reinterpret_cast<db200_sw_result *>(h + out0);
    int32_t *h_status = reinterpret_cast<int32_t *>(h_res + npairs);
    volatile int32_t *h_done = reinterpret_cast<volatile int32_t *>(h_status + npairs);
    uint32_t *h_cig = reinterpret_cast<uint32_t *>(h_status + 2 * npairs);
    SmallArgs a;
    a.seq = d + (h_seq - h);
    a.qoff = reinterpret_cast<const int32_t *>(d); a.toff = a.qoff + npairs + 1; a.mask_len = a.toff + npairs + 1;
    a.results = reinterpret_cast<db200_sw_result *>(d + out0);
    a.status = reinterpret_cast<int32_t *>(a.results + npairs);
    a.done = reinterpret_cast<volatile int32_t *>(a.status + npairs);
    a.cigar = reinterpret_cast<uint32_t *>(a.status + 2 * npairs);
    a.epoch = ++e->small_epoch;
    if (a.epoch == 0) a.epoch = ++e->small_epoch;
    for (int64_t i = 0; i < npairs; ++i) h_done[i] = 0;
    a.maxq = (maxq + 15) & ~15;
    int bias = 0;
    for (int x = 0; x < 6; ++x)
        for (int y = 0; y < 6; ++y) {
            int v = SW_NEG;
            if (x < 5 && y < 5) v = prm->mat ? prm->mat[x * 5 + y] : ((x == 4 || y == 4) ? 0 : (x == y ? prm->match : prm->mismatch));
            a.sc.mat[x * 6 + y] = (int16_t)v;
            if (x < 5 && y < 5 && v < bias) bias = v;
        }
    a.sc.gap_open = prm->gap_open; a.sc.gap_extend = prm->gap_extend;
    a.bias = -bias; a.score_size = prm->score_size; a.flag = prm->flag;
    const size_t smem = (size_t)a.maxq + SW_SMALL_MAX_T + 2 * SW_SMALL_MAX_T * sizeof(int16_t) + 3 * (SW_SMALL_XW + 2) * sizeof(int16_t) +
                        SW_SMALL_RUNS * sizeof(uint32_t) + SW_SMALL_DIR_BYTES;
    static bool attr_done[16] = {false};
    if (!attr_done[e->device & 15]) {
        DB200_CUDA_CHECK(cudaFuncSetAttribute(sw_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)(SW_SMALL_MAX_Q + SW_SMALL_MAX_T + 2 * SW_SMALL_MAX_T * sizeof(int16_t) +
                                                    3 * (SW_SMALL_XW + 2) * sizeof(int16_t) + SW_SMALL_RUNS * sizeof(uint32_t) + SW_SMALL_DIR_BYTES)));
        attr_done[e->device & 15] = true;
    }
    sw_small_kernel<<<(unsigned)npairs, 32, smem, e->stream>>>(a);
    count_launch(e);
    DB200_CUDA_CHECK(cudaGetLastError());
    {
        // wait on the per-pair flags the kernel writes into mapped memory after its results (cheaper than a stream
        // synchronisation for a call of a few tens of microseconds); the driver call remains the fallback and the error path
        timespec t0; clock_gettime(CLOCK_MONOTONIC, &t0);
        int64_t i = 0;
        unsigned spins = 0;
        while (i < npairs) {
            if (h_done[i] == a.epoch) { ++i; continue; }
            if ((++spins & 1023u) == 0) {
                timespec t1; clock_gettime(CLOCK_MONOTONIC, &t1);
                if ((t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6 > 2.0) break;   // long pairs: let the driver wait
            }
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
        if (i < npairs) DB200_CUDA_CHECK(cudaStreamSynchronize(e->stream));
    }
    int64_t cpos = 0;
    for (int64_t i = 0; i < npairs; ++i) {
        const int64_t ql = qoff[i + 1] - qoff[i], tl = toff[i + 1] - toff[i];
        int st = h_status[i];
        if (!(ql <= SW_SMALL_MAX_Q && tl <= SW_SMALL_MAX_T)) st = SW_SMALL_FALLBACK;
        results[i] = h_res[i];
        status[i] = st;
        if (cigar_off) cigar_off[i] = cpos;
        if (st == DB200_PAIR_OK && h_res[i].cigar_len > 0) {
            if (cigar_buf) {
                if (cpos + h_res[i].cigar_len > cigar_cap) { set_error("sw_batch_host: cigar buffer too small"); return DB200_ERR_CAPACITY; }
                memcpy(cigar_buf + cpos, h_cig + (size_t)i * SW_SMALL_CIGAR_STRIDE, sizeof(uint32_t) * (size_t)h_res[i].cigar_len);
            }
            cpos += h_res[i].cigar_len;
        }
    }
    if (cigar_off) cigar_off[npairs] = cpos;
    *handled = 1;
    return DB200_OK;
}

__global__ void sw_colmax_len_kernel(const PairMeta *meta, const PairState *state, int64_t npairs, int32_t *lens)
{
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= npairs) return;
    lens[p] = (state[p].status == DB200_PAIR_OK && state[p].need_colmax) ? 2 * meta[p].n : 0;
}

__global__ void sw_cigar_len_kernel(const PairState *state, int64_t npairs, int32_t *lens)
{
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= npairs) return;
    lens[p] = state[p].cigar_len;
}

// `off` are the exclusive cigar offsets of this (sub-)batch.  A sub-batch of a larger device call is linked to its
// predecessor: *base is the first cigar slot it may use (the predecessor's end), the rebased offsets go to out_off
// and the sub-batch's own end to *end_word.
__global__ void sw_cigar_gather_kernel(const PairState *state, int64_t npairs, const int64_t *off, const uint32_t *pool,
                                       uint32_t *out, int64_t out_cap, int32_t *status_out, int32_t *overflow_flag,
                                       const int64_t *base, int64_t *out_off, int64_t *end_word)
{
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= npairs) return;
    const int64_t b = base ? *base : 0;
    if (out_off) {
        out_off[p] = b + off[p];
        if (p == npairs - 1) out_off[npairs] = b + off[npairs];
    }
    if (end_word && p == npairs - 1) *end_word = b + off[npairs];
    if (status_out) status_out[p] = state[p].status;
    const int len = state[p].cigar_len;
    if (len <= 0 || !out) return;
    if (b + off[p] + len > out_cap) { atomicExch(overflow_flag, 3); return; }
    const uint32_t *src = pool + state[p].cigar_pos;
    uint32_t *dst = out + b + off[p];
    for (int i = 0; i < len; ++i) dst[i] = src[i];
}

// ------------------------------------------------------------------------------------------------
// pipelined column tiles: the tasks of a column-tiled segment are expanded into (pair, pass) units
// ------------------------------------------------------------------------------------------------

__global__ void sw_npass_kernel(const int32_t *tasks, const int32_t *seg_begin, const int32_t *seg_count, int seg, const PairMeta *meta,
                                int64_t npairs, int32_t *lens)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= npairs) return;
    int v = 0;
    if (i < seg_count[seg]) v = max(1, (meta[tasks[seg_begin[seg] + i]].n + SW_MULTI_COLS - 1) / SW_MULTI_COLS);
    lens[i] = v;
}

struct ExpandArgs {
    const int32_t *tasks, *seg_begin, *seg_count;
    int32_t seg;
    const PairMeta *meta;
    const int64_t *vt_off;       // exclusive scan of the pass counts, in task order (longest queries first)
    int32_t *vt_pair, *vt_pass;
    long long *bnd_off;
    unsigned long long *bnd_cursor;
    unsigned long long bnd_cap;
    int32_t pipe_max_tasks;      // segments with more tasks than this keep one warp per pair
};

__global__ void sw_expand_kernel(ExpandArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.seg_count[a.seg]) return;
    const int pair = a.tasks[a.seg_begin[a.seg] + i];
    const PairMeta pm = a.meta[pair];
    const int npass = max(1, (pm.n + SW_MULTI_COLS - 1) / SW_MULTI_COLS);
    const int64_t base = a.vt_off[i];
    long long off = -1;
    if (npass > 1) {
        // right-edge columns of every pass but the last, rows padded to 32 (what sw_pass_kernel indexes)
        const unsigned long long need = (unsigned long long)((pm.m + 31) & ~31) * (unsigned long long)(npass - 1);
        const unsigned long long o = atomicAdd(a.bnd_cursor, need);
        off = (o + need <= a.bnd_cap) ? (long long)o : -1;
        a.bnd_off[pair] = off;
    }
    // Without room for its boundary columns a pair cannot be pipelined: its first unit runs all passes in one warp.
    // The same holds for every pair of a large segment: with more pairs than the GPU has warps for, one warp per pair
    // already fills the machine and the hand-over between warps only costs (B200, long-read contigs 1-10 kb, forward
    // stage: 1,000 pairs 25.5 -> 7.1 ms, 4,000 pairs 27.2 -> 17.0 ms, but 16,000 pairs 52.3 -> 66.7 ms when pipelined).
    const bool seq = npass > 1 && (off < 0 || a.seg_count[a.seg] > a.pipe_max_tasks);
    for (int ps = 0; ps < npass; ++ps) {
        a.vt_pair[base + ps] = pair;
        a.vt_pass[base + ps] = seq ? (ps == 0 ? -1 : -2) : ps;
    }
}

// ------------------------------------------------------------------------------------------------
// host orchestration
// ------------------------------------------------------------------------------------------------
template <bool COLMAX, bool WATCH, bool TAG>
static int launch_cfg(Engine *e, int cfg, const PassArgs &a, cudaStream_t st)
{
    switch (cfg) {
        case 0: return launch_pass<16, 1, COLMAX, false, WATCH, TAG>(e, a, st);
        case 1: return launch_pass<16, 2, COLMAX, false, WATCH, TAG>(e, a, st);
        case 2: return launch_pass<16, 3, COLMAX, false, WATCH, TAG>(e, a, st);
        case 3: return launch_pass<16, 4, COLMAX, false, WATCH, TAG>(e, a, st);
        case 4: return launch_pass<16, 5, COLMAX, false, WATCH, TAG>(e, a, st);
        case 5: return launch_pass<16, 6, COLMAX, false, WATCH, TAG>(e, a, st);
        case 6: return launch_pass<16, 7, COLMAX, false, WATCH, TAG>(e, a, st);
        case 7: return launch_pass<16, 8, COLMAX, false, WATCH, TAG>(e, a, st);
        case 8: return launch_pass<16, 10, COLMAX, false, WATCH, TAG>(e, a, st);
        case 9: return launch_pass<16, 16, COLMAX, false, WATCH, TAG>(e, a, st);
        case 10: return launch_pass<16, 32, COLMAX, false, WATCH, TAG>(e, a, st);
        case SW_CFG_MULTI:
            if constexpr (TAG) return DB200_ERR_INTERNAL;
            else return launch_pass<32, 32, COLMAX, true, WATCH, false>(e, a, st);
        default: break;
    }
    // fine configurations: tagged kernels only, one translation unit per group size
    if constexpr (TAG && !WATCH) {
        const SwCfg f = sw_cfg_table()[cfg];
        switch (f.G) {
            case 1: return sw_launch_fine_g1(e, f.K, COLMAX, a, st);
            case 2: return sw_launch_fine_g2(e, f.K, COLMAX, a, st);
            case 4: return sw_launch_fine_g4(e, f.K, COLMAX, a, st);
            case 8: return sw_launch_fine_g8(e, f.K, COLMAX, a, st);
            case 16: return sw_launch_fine_g16(e, f.K, COLMAX, a, st);
            case 32: return sw_launch_fine_g32(e, f.K, COLMAX, a, st);
            default: break;
        }
    }
    set_error("sw_batch: configuration %d has no kernel for this variant", cfg);
    return DB200_ERR_INTERNAL;
}

int sw_batch_device_impl(Engine *e, Arena &ws, const db200_sw_params *prm, const uint8_t *d_qbuf, const int64_t *d_qoff,
                         int64_t q_bytes, const uint8_t *d_tbuf, const int64_t *d_toff, int64_t t_bytes, int64_t npairs,
                         int32_t max_query_len, int32_t max_target_len, const int32_t *d_mask_len, db200_sw_result *d_results,
                         uint32_t *d_cigar_buf,
                         int64_t cigar_cap, int64_t *d_cigar_off, int32_t *d_status, int64_t *cigar_total, cudaStream_t st,
                         const SwLink *link, const int32_t *d_qlen, const int32_t *d_tlen, const uint32_t *d_packed, int64_t packed_units)
{
    if (npairs < 0 || !prm || !d_results) { set_error("sw_batch: bad arguments"); return DB200_ERR_ARG; }
    if (prm->gap_open <= 0 || prm->gap_extend <= 0 || prm->gap_open > 255 || prm->gap_extend > 255) {
        set_error("sw_batch: gap penalties must be in 1..255"); return DB200_ERR_ARG;
    }
    if (prm->mat && prm->n != 5) { set_error("sw_batch: only 5x5 nucleotide matrices are supported"); return DB200_ERR_ARG; }
    if (prm->score_size < 0 || prm->score_size > 2) { set_error("sw_batch: score_size must be 0, 1 or 2"); return DB200_ERR_ARG; }
    if (npairs == 0) {
        if (d_cigar_off) DB200_CUDA_CHECK(cudaMemsetAsync(d_cigar_off, 0, sizeof(int64_t), st));
        if (cigar_total) *cigar_total = 0;
        return DB200_OK;
    }
    if (npairs > (int64_t)1 << 30) { set_error("sw_batch: at most 2^30 pairs per call"); return DB200_ERR_ARG; }
    // ---- scoring ---------------------------------------------------------------------------------
    SwScoring sc;
    int8_t m5[25];
    for (int a = 0; a < 5; ++a)
        for (int b = 0; b < 5; ++b)
            m5[a * 5 + b] = prm->mat ? prm->mat[a * 5 + b] : (int8_t)((a == 4 || b == 4) ? 0 : (a == b ? prm->match : prm->mismatch));
    int bias = 0, max_match = 0;
    for (int i = 0; i < 25; ++i) { if (m5[i] < bias) bias = m5[i]; if (m5[i] > max_match) max_match = m5[i]; }
    bias = -bias;
    for (int a = 0; a < 6; ++a)
        for (int b = 0; b < 6; ++b)
            sc.mat[a * 6 + b] = (a < 5 && b < 5) ? (int16_t)m5[a * 5 + b] : (int16_t)SW_NEG;
    sc.gap_open = prm->gap_open;
    sc.gap_extend = prm->gap_extend;
    // tagged kernels run on scores scaled by 2^TAGB: 16 for K <= 16 columns per virtual lane (4 tag bits), 32 above
    SwScoring sc16 = sc, sc32 = sc;
    for (int i = 0; i < 36; ++i) if (sc.mat[i] != SW_NEG) { sc16.mat[i] = (int16_t)(sc.mat[i] * 16); sc32.mat[i] = (int16_t)(sc.mat[i] * 32); }
    sc16.gap_open = sc.gap_open * 16; sc16.gap_extend = sc.gap_extend * 16;
    sc32.gap_open = sc.gap_open * 32; sc32.gap_extend = sc.gap_extend * 32;
    // ---- configurations usable in this call ------------------------------------------------------------
    // Fine configurations (sw_launch.cuh) pay for themselves when the buckets are large: ~60 usable configurations split a
    // batch (or a chunk of the host pipeline) into as many persistent launches.  The ASCII host call, whose 32k-147k-pair
    // chunks arrive in submission order, went from 27.4 to 34 ms per 524,288 pairs with them (500-2,500 pairs per launch);
    // whole batches of that size gain 6 %.  Smaller calls keep the eleven classic configurations.
    const SwCfg *cfgs = sw_cfg_table();
    int64_t fine_min_pairs = 196608;
    if (const char *fm = getenv("DB200_SW_FINE_MIN_PAIRS")) fine_min_pairs = atoll(fm);
    const bool fine = npairs >= fine_min_pairs && !getenv("DB200_SW_NO_FINE");
    int ovh = 8;    // ALU-pipe instructions per row outside the recurrence (hand-over, query decode; measured optimum on the headline step)
    if (const char *ov = getenv("DB200_SW_ROW_OVERHEAD")) ovh = atoi(ov);
    int force_k = 0, force_g = 0;   // measurement aid: DB200_SW_FORCE_CFG=K,G leaves that configuration alone in the table
    if (const char *fc = getenv("DB200_SW_FORCE_CFG")) sscanf(fc, "%d,%d", &force_k, &force_g);
    bool tag_ok[SW_NCFG];
    CfgSel sel;
    memset(&sel, 0, sizeof(sel));
    for (int c = 0; c < SW_NCFG; ++c) {
        const int K = cfgs[c].K, G = cfgs[c].G, cap = 2 * K * G;
        const bool is_fine = c >= SW_NCLASSIC && c < SW_CFG_MULTI;
        tag_ok[c] = c < SW_CFG_MULTI && max_match > 0 && (int64_t)max_match * cap <= (K <= 16 ? 2047 : 1023);
        sel.cap[c] = (uint16_t)cap;
        sel.skew[c] = (uint8_t)(2 * G - 1);
        // issue slots of one row of one pair: (6 K + overhead) per warp row, 32 / G pairs per warp
        const bool usable = c < SW_CFG_MULTI && (!is_fine || (fine && tag_ok[c])) && (force_k == 0 || (K == force_k && G == force_g));
        sel.cost[c] = usable ? (uint16_t)((64 * (6 * K + ovh) + (32 / G) / 2) / (32 / G)) : 0;
    }
    // a configuration is never chosen when another one that holds every target of this batch is cheaper in both terms
    bool launchable[SW_NCFG];
    for (int c = 0; c < SW_NCFG; ++c) {
        launchable[c] = c == SW_CFG_MULTI || sel.cost[c] != 0;
        if (c == SW_CFG_MULTI || !launchable[c] || max_target_len <= 0) continue;
        for (int d = 0; d < SW_CFG_MULTI && launchable[c]; ++d) {
            if (d == c || sel.cost[d] == 0 || (int)sel.cap[d] < max_target_len || sel.skew[d] > sel.skew[c]) continue;
            if (sel.cost[d] < sel.cost[c] || (sel.cost[d] == sel.cost[c] && d < c)) launchable[c] = false;
        }
    }
    if (max_target_len > 0 && max_target_len <= SW_SINGLE_PASS_MAX) launchable[SW_CFG_MULTI] = false;
    // launch order: widest passes first, their drain overlaps the narrower buckets
    int order[SW_NCFG];
    for (int c = 0; c < SW_NCFG; ++c) order[c] = c;
    std::stable_sort(order, order + SW_NCFG, [&](int x, int y) { return 2 * cfgs[x].K * cfgs[x].G * (x == SW_CFG_MULTI ? 64 : 1) > 2 * cfgs[y].K * cfgs[y].G * (y == SW_CFG_MULTI ? 64 : 1); });

    // ---- workspace ---------------------------------------------------------------------------------
    const int64_t nseq = 2 * npairs;
    const size_t seq_units = (size_t)((q_bytes + t_bytes) / 32 + nseq + 64);
    if (seq_units * 2 > 0xFFFFFFF0ull) { set_error("sw_batch: batch too large (packed pool exceeds 2^32 units)"); return DB200_ERR_ARG; }
    const size_t nb = (size_t)SW_NSEG * SW_LBINS;
    const int bstride = ((max_query_len > 0 ? max_query_len : 0) + 63) & ~31;
    const size_t bwarps = (size_t)e->sm_count * SW_MULTI_MAX_WARPS_PER_SM;
    const int64_t cig_pool_cap = cigar_cap > 0 ? cigar_cap : (q_bytes + t_bytes + 2 * npairs);

    int32_t *units = ws.alloc<int32_t>(nseq + 16);
    int64_t *unit_off = ws.alloc<int64_t>(nseq + 16);
    int64_t *scan_tmp = ws.alloc<int64_t>(scan_tmp_count(nseq) * 2);
    uint32_t *pool = ws.alloc<uint32_t>(seq_units * 4 * 2);
    PairMeta *meta = ws.alloc<PairMeta>(npairs);
    PairMeta *rmeta = ws.alloc<PairMeta>(npairs);
    PairState *state = ws.alloc<PairState>(npairs);
    PassOut *fwd = ws.alloc<PassOut>(npairs);
    PassOut *rev = ws.alloc<PassOut>(npairs);
    int32_t *tasks_f = ws.alloc<int32_t>(npairs);
    int32_t *tasks_r = ws.alloc<int32_t>(npairs);
    int32_t *amb_f = ws.alloc<int32_t>(npairs);
    int32_t *amb_r = ws.alloc<int32_t>(npairs);
    int64_t *colmax_off = ws.alloc<int64_t>(npairs + 1);
    int64_t *cig_off_tmp = ws.alloc<int64_t>(npairs + 1);
    int32_t *lens = ws.alloc<int32_t>(npairs);
    int16_t *colmax = ws.alloc<int16_t>((size_t)(t_bytes + npairs) * 2);
    int32_t *bins = ws.alloc<int32_t>(nb * 6 + 1024);
    uint32_t *cigar_pool = ws.alloc<uint32_t>((size_t)cig_pool_cap + 16);
    int32_t *counters = ws.alloc<int32_t>(4096);
    uint32_t *boundary = ws.alloc<uint32_t>(bwarps * (size_t)bstride + 64);
    // per-pair boundary columns of column-tiled pairs: rows x (passes - 1) words each; bounded by the query bytes
    // times the largest pass count any target of this batch can need
    const int64_t max_t = max_target_len > 0 ? max_target_len : 0;
    const int64_t max_extra_passes = max_t > 2048 ? (max_t + 2047) / 2048 - 1 : 0;
    const size_t bnd_cap = (size_t)std::min<int64_t>(2 * (q_bytes + 32 * npairs) * max_extra_passes, (int64_t)1 << 30); // forward + reverse
    uint32_t *bnd_pool = ws.alloc<uint32_t>(bnd_cap + 64);
    long long *bnd_off = ws.alloc<long long>(2 * (size_t)npairs);
    // Pipelined column tiles (sw_pass_kernel PIPE): only when a target can exceed the widest single-pass configuration and
    // the batch is small enough for the hand-over between warps to pay.  The pair count of the batch bounds the number of
    // column-tiled pairs from above; beyond pipe_max_tasks the plain kernel runs (its instantiation is also ~10 % faster
    // than the PIPE instantiation running one warp per pair: 16,000 long-read contigs 52.5 vs 58.3 ms forward).
    int pipe_max_tasks = 6144;
    if (const char *pt = getenv("DB200_SW_PIPE_MAX_TASKS")) pipe_max_tasks = atoi(pt);
    const bool pipe = max_query_len > 0 && max_target_len > SW_SINGLE_PASS_MAX && npairs <= pipe_max_tasks && !getenv("DB200_SW_NO_PIPE");
    const size_t vt_cap = pipe ? (size_t)(t_bytes / SW_MULTI_COLS + npairs + 16) : 0;
    constexpr int NPIPE = 3;   // forward without / with column maxima, reverse
    int32_t *vt_len = nullptr, *vt_pair = nullptr, *vt_pass = nullptr, *vt_prog = nullptr, *pair_done = nullptr;
    int64_t *vt_off = nullptr, *vt_scan_tmp = nullptr;
    PassOut *vt_res = nullptr;
    if (pipe) {
        vt_len = ws.alloc<int32_t>((size_t)npairs * NPIPE);
        vt_off = ws.alloc<int64_t>((size_t)(npairs + 1) * NPIPE);
        vt_scan_tmp = ws.alloc<int64_t>(scan_tmp_count(npairs) * NPIPE);
        vt_pair = ws.alloc<int32_t>(vt_cap * NPIPE);
        vt_pass = ws.alloc<int32_t>(vt_cap * NPIPE);
        vt_prog = ws.alloc<int32_t>(vt_cap * NPIPE);
        vt_res = ws.alloc<PassOut>(vt_cap * NPIPE);
        pair_done = ws.alloc<int32_t>((size_t)npairs * 2);
        if (!vt_len || !vt_off || !vt_scan_tmp || !vt_pair || !vt_pass || !vt_prog || !vt_res || !pair_done) {
            set_error("sw_batch: device workspace allocation failed");
            return DB200_ERR_NOMEM;
        }
        DB200_CUDA_CHECK(cudaMemsetAsync(vt_prog, 0, vt_cap * NPIPE * sizeof(int32_t), st));
        DB200_CUDA_CHECK(cudaMemsetAsync(pair_done, 0, (size_t)npairs * 2 * sizeof(int32_t), st));
    }
    const int wide_blocks = std::min(e->sm_count * 4, 1024);
    const size_t slab_bytes = (e->trace_slab_bytes / (size_t)wide_blocks) & ~(size_t)255;
    uint8_t *slab = ws.alloc<uint8_t>(slab_bytes * (size_t)wide_blocks);
    int32_t *narrow_list = ws.alloc<int32_t>((size_t)npairs * TRACE_MEDIUM_CLASSES);
    int32_t *medium_list = ws.alloc<int32_t>((size_t)npairs * TRACE_MEDIUM_CLASSES);
    int32_t *wide_list = ws.alloc<int32_t>((size_t)npairs * TRACE_MEDIUM_CLASSES);
    int32_t *slab_scalar_list = ws.alloc<int32_t>((size_t)npairs * TRACE_MEDIUM_CLASSES);
    if (!units || !unit_off || !scan_tmp || !pool || !meta || !rmeta || !state || !fwd || !rev || !tasks_f || !tasks_r || !amb_f ||
        !amb_r || !colmax_off || !cig_off_tmp || !lens || !colmax || !bins || !cigar_pool || !counters || !boundary || !slab ||
        !bnd_pool || !bnd_off ||
        !narrow_list || !medium_list || !wide_list || !slab_scalar_list) {
        set_error("sw_batch: device workspace allocation failed");
        return DB200_ERR_NOMEM;
    }
    int32_t *hist_f = bins, *start_f = bins + nb, *fill_f = bins + 2 * nb;
    int32_t *hist_r = bins + 3 * nb, *start_r = bins + 4 * nb, *fill_r = bins + 5 * nb;
    int32_t *segs = bins + 6 * nb;
    int32_t *seg_begin_f = segs, *seg_count_f = segs + SW_NSEG, *seg_begin_r = segs + 2 * SW_NSEG, *seg_count_r = segs + 3 * SW_NSEG;
    int32_t *amb_count_f = segs + 4 * SW_NSEG, *amb_count_r = segs + 5 * SW_NSEG;
    // counters: [128] unit cursor, [130..131] trace cursor (u64), [132..133] cigar cursor (u64), [134] overflow flag, ...;
    // [256 ..] work counters of the pass launches (at most 5 per configuration)
    DB200_CUDA_CHECK(cudaMemsetAsync(bins, 0, (nb * 6 + 1024) * sizeof(int32_t), st));
    DB200_CUDA_CHECK(cudaMemsetAsync(counters, 0, 4096 * sizeof(int32_t), st));
    uint32_t *unit_cursor = reinterpret_cast<uint32_t *>(counters + 128);
    unsigned long long *cigar_cursor = reinterpret_cast<unsigned long long *>(counters + 132);
    int32_t *overflow_flag = counters + 134;

    const int TB = 256;
    const unsigned gseq = (unsigned)((nseq + TB - 1) / TB), gpair = (unsigned)((npairs + TB - 1) / TB);
    const unsigned gwarp_seq = (unsigned)((nseq * 32 + TB - 1) / TB), gwarp_pair = (unsigned)((npairs * 32 + TB - 1) / TB);

    // 1. pack ---------------------------------------------------------------------------------------
    const uint32_t *pool_f = pool;   // where the forward sequences live
    prof_mark(e, st, ST_BEGIN);
    sw_units_kernel<<<gseq, TB, 0, st>>>(d_qoff, d_toff, d_qlen, d_tlen, npairs, units);
    count_launch(e);
    exclusive_scan_i32_i64(e, units, unit_off, nseq, scan_tmp, st);
    if (d_packed) {
        // the caller's batch is already in pool format ([all queries][all targets], 32 bases per 16-byte unit): one device copy
        // instead of the pack kernel; q_bytes / t_bytes are base counts, so the units are bounded by seq_units
        if (packed_units < 0 || (size_t)packed_units > seq_units) { set_error("sw_batch: packed unit count does not match the lengths"); return DB200_ERR_ARG; }
        // 16-byte aligned: the kernels read the caller's batch in place (the forward sequences are never written; reversed
        // sequences go to the workspace pool); otherwise one device copy into the workspace pool
        if ((reinterpret_cast<uintptr_t>(d_packed) & 15u) == 0 && !getenv("DB200_SW_COPY_PACKED")) pool_f = d_packed;
        else DB200_CUDA_CHECK(cudaMemcpyAsync(pool, d_packed, (size_t)packed_units * 16, cudaMemcpyDeviceToDevice, st));
        sw_meta_kernel<<<gpair, TB, 0, st>>>(d_qlen, d_tlen, npairs, unit_off, meta);
    } else {
        sw_pack_kernel<<<gwarp_seq, TB, 0, st>>>(d_qbuf, d_qoff, d_tbuf, d_toff, d_qlen, d_tlen, npairs, unit_off, pool, meta);
    }
    count_launch(e);

    prof_mark(e, st, ST_PACK);
    // 2. bucket -------------------------------------------------------------------------------------
    InitArgs ia;
    ia.meta = meta; ia.mask_len = d_mask_len; ia.state = state; ia.npairs = npairs; ia.max_match = max_match; ia.flag = prm->flag; ia.sel = sel;
    sw_init_kernel<<<gpair, TB, 0, st>>>(ia, hist_f);
    sw_binscan_kernel<<<1, SW_BINSCAN_THREADS, 0, st>>>(hist_f, start_f, seg_begin_f, seg_count_f);
    sw_scatter_kernel<<<gpair, TB, 0, st>>>(meta, state, npairs, 0, sel, start_f, fill_f, tasks_f);
    sw_colmax_len_kernel<<<gpair, TB, 0, st>>>(meta, state, npairs, lens);
    count_launch(e, 4);
    exclusive_scan_i32_i64(e, lens, colmax_off, npairs, scan_tmp, st);
    DB200_CUDA_CHECK(cudaMemsetAsync(fwd, 0, sizeof(PassOut) * npairs, st));
    DB200_CUDA_CHECK(cudaMemsetAsync(rev, 0, sizeof(PassOut) * npairs, st));

    prof_mark(e, st, ST_BUCKET);
    PassArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.pool = reinterpret_cast<const uint4 *>(pool_f);
    pa.colmax = colmax; pa.colmax_off = colmax_off;
    pa.boundary = boundary; pa.boundary_stride = bstride;
    pa.bnd_pool = bnd_pool; pa.bnd_cap = (unsigned long long)bnd_cap;
    pa.bnd_cursor = reinterpret_cast<unsigned long long *>(counters + 144);
    pa.bnd_off = bnd_off;
    pa.k65536 = 65536u;
    pa.sc = sc;
    int counter_idx = 256;
    const bool can_multi = max_query_len > 0;

    // column-tiled segment `seg` of the task list in `pa`, pipelined: expand it into (pair, pass) units, then one warp per unit
    auto launch_piped = [&](int which, bool colmax, PassArgs pp, cudaStream_t ls) -> int {
        int32_t *len_w = vt_len + (size_t)npairs * which;
        int64_t *off_w = vt_off + (size_t)(npairs + 1) * which;
        sw_npass_kernel<<<gpair, TB, 0, ls>>>(pp.tasks, pp.seg_begin, pp.seg_count, pp.seg, pp.meta, npairs, len_w);
        count_launch(e);
        exclusive_scan_i32_i64(e, len_w, off_w, npairs, vt_scan_tmp + scan_tmp_count(npairs) * which, ls);
        ExpandArgs xa;
        xa.tasks = pp.tasks; xa.seg_begin = pp.seg_begin; xa.seg_count = pp.seg_count; xa.seg = pp.seg; xa.meta = pp.meta;
        xa.vt_off = off_w; xa.vt_pair = vt_pair + vt_cap * which; xa.vt_pass = vt_pass + vt_cap * which;
        xa.bnd_off = pp.bnd_off; xa.bnd_cursor = pp.bnd_cursor; xa.bnd_cap = pp.bnd_cap;
        xa.pipe_max_tasks = pipe_max_tasks;
        sw_expand_kernel<<<gpair, TB, 0, ls>>>(xa);
        count_launch(e);
        pp.vt_pair = xa.vt_pair; pp.vt_pass = xa.vt_pass; pp.vt_total = off_w + npairs;
        pp.vt_prog = vt_prog + vt_cap * which; pp.vt_res = vt_res + vt_cap * which;
        pp.pair_done = pair_done + (which == 2 ? npairs : 0);
        pp.abort_flag = counters + 148;
        return colmax ? launch_pass<32, 32, true, true, false, false, true>(e, pp, ls)
                      : launch_pass<32, 32, false, true, false, false, true>(e, pp, ls);
    };

    // 3. forward pass per segment, then the exact re-run of pairs with an ambiguous end row -----------
    { const int rc = aux_fork(e, st); if (rc) return rc; }
    int aux_i = 0;
    auto tag_sc = [&](int cfg) -> const SwScoring & { return cfgs[cfg].K <= 16 ? sc16 : sc32; };
    for (int oi = 0; oi < 2 * SW_NCFG; ++oi) {
        const int cfg = order[oi >> 1], cmx = 1 - (oi & 1), seg = 2 * cfg + cmx;
        if (!launchable[cfg]) continue;
        if (cfg == SW_CFG_MULTI && !can_multi) continue;
        cudaStream_t ls = e->aux[e->aux_set][(aux_i++) % Engine::NAUX];
        pa.meta = meta; pa.tasks = tasks_f; pa.seg_begin = seg_begin_f; pa.seg_count = seg_count_f; pa.seg = seg;
        pa.counter = counters + (counter_idx++);
        pa.out = fwd;
        pa.amb_list = amb_f; pa.amb_count = amb_count_f + 2 * cfg; pa.amb_base_seg = 2 * cfg;
        int rc;
        if (cfg == SW_CFG_MULTI && pipe) { pa.sc = sc; rc = launch_piped(cmx, cmx != 0, pa, ls); }
        else if (cmx && tag_ok[cfg]) { pa.sc = tag_sc(cfg); rc = launch_cfg<true, false, true>(e, cfg, pa, ls); }
        else if (cmx) { pa.sc = sc; rc = launch_cfg<true, false, false>(e, cfg, pa, ls); }
        else if (tag_ok[cfg]) { pa.sc = tag_sc(cfg); rc = launch_cfg<false, false, true>(e, cfg, pa, ls); }
        else { pa.sc = sc; rc = launch_cfg<false, false, false>(e, cfg, pa, ls); }
        if (rc) return rc;
    }
    { const int rc = aux_join(e, st); if (rc) return rc; }
    pa.sc = sc;
    prof_mark(e, st, ST_FORWARD);
    for (int cfg = 0; cfg < SW_NCFG; ++cfg) {
        if (!launchable[cfg] || (cfg == SW_CFG_MULTI && !can_multi)) continue;
        pa.meta = meta; pa.tasks = amb_f; pa.seg_begin = seg_begin_f; pa.seg_count = amb_count_f; pa.seg = 2 * cfg;
        pa.counter = counters + (counter_idx++);
        pa.out = fwd;
        if (tag_ok[cfg]) continue;   // tagged passes are never ambiguous
        const int rc = launch_cfg<false, true, false>(e, cfg, pa, st);
        if (rc) return rc;
    }

    prof_mark(e, st, ST_FORWARD_WATCH);
    // 4. reversed problems ----------------------------------------------------------------------------
    MidArgs ma;
    ma.fwd = fwd; ma.meta = meta; ma.rmeta = rmeta; ma.state = state; ma.npairs = npairs;
    ma.bias = bias; ma.score_size = prm->score_size; ma.flag = prm->flag; ma.filters = prm->filters;
    ma.max_match = max_match; ma.gap_open = prm->gap_open; ma.gap_extend = prm->gap_extend;
    ma.unit_cursor = unit_cursor; ma.unit_base = (uint32_t)seq_units; ma.sel = sel;
    ma.pool = pool_f; ma.rev = rev; ma.uniform_match = (!prm->mat && prm->match > 0 && prm->mismatch < prm->match && !getenv("DB200_SW_NO_DIAG_SHORTCUT")) ? 1 : 0;
    sw_mid_kernel<<<gpair, TB, 0, st>>>(ma, hist_r);
    sw_reverse_extract_kernel<<<gwarp_pair, TB, 0, st>>>(meta, rmeta, state, fwd, npairs, pool_f, pool);
    sw_binscan_kernel<<<1, SW_BINSCAN_THREADS, 0, st>>>(hist_r, start_r, seg_begin_r, seg_count_r);
    sw_scatter_kernel<<<gpair, TB, 0, st>>>(rmeta, state, npairs, 1, sel, start_r, fill_r, tasks_r);
    count_launch(e, 4);
    prof_mark(e, st, ST_REVERSE_PREP);
    pa.bnd_off = bnd_off + npairs;   // the reversed problems keep their own boundary columns
    pa.pool = reinterpret_cast<const uint4 *>(pool);   // ... and live in the workspace pool
    { const int rc = aux_fork(e, st); if (rc) return rc; }
    for (int oi = 0; oi < SW_NCFG; ++oi) {
        const int cfg = order[oi];
        if (!launchable[cfg] || (cfg == SW_CFG_MULTI && !can_multi)) continue;
        cudaStream_t ls = e->aux[e->aux_set][(aux_i++) % Engine::NAUX];
        pa.meta = rmeta; pa.tasks = tasks_r; pa.seg_begin = seg_begin_r; pa.seg_count = seg_count_r; pa.seg = 2 * cfg;
        pa.counter = counters + (counter_idx++);
        pa.out = rev;
        pa.amb_list = amb_r; pa.amb_count = amb_count_r + 2 * cfg; pa.amb_base_seg = 2 * cfg;
        int rc;
        if (cfg == SW_CFG_MULTI && pipe) { pa.sc = sc; rc = launch_piped(2, false, pa, ls); }
        else if (tag_ok[cfg]) { pa.sc = tag_sc(cfg); rc = launch_cfg<false, false, true>(e, cfg, pa, ls); }
        else { pa.sc = sc; rc = launch_cfg<false, false, false>(e, cfg, pa, ls); }
        if (rc) return rc;
    }
    { const int rc = aux_join(e, st); if (rc) return rc; }
    prof_mark(e, st, ST_REVERSE);
    for (int cfg = 0; cfg < SW_NCFG; ++cfg) {
        if (!launchable[cfg] || (cfg == SW_CFG_MULTI && !can_multi)) continue;
        pa.meta = rmeta; pa.tasks = amb_r; pa.seg_begin = seg_begin_r; pa.seg_count = amb_count_r; pa.seg = 2 * cfg;
        pa.counter = counters + (counter_idx++);
        pa.out = rev;
        pa.sc = sc;
        if (tag_ok[cfg]) continue;   // tagged passes are never ambiguous
        const int rc = launch_cfg<false, true, false>(e, cfg, pa, st);
        if (rc) return rc;
    }

    prof_mark(e, st, ST_REVERSE_WATCH);
    // 5. finalize + traceback + cigar compaction ---------------------------------------------------------
    FinalArgs fa;
    fa.fwd = fwd; fa.rev = rev; fa.meta = meta; fa.state = state; fa.colmax = colmax; fa.colmax_off = colmax_off;
    fa.results = d_results; fa.status_out = nullptr; fa.npairs = npairs;
    fa.flag = prm->flag; fa.filters = prm->filters; fa.filterd = prm->filterd;
    fa.pipe_abort = pipe ? counters + 148 : nullptr; fa.tiled_from = SW_SINGLE_PASS_MAX;
    fa.defer_second_from = max_target_len > 512 ? 512 : 0;
    sw_finalize_kernel<<<gpair, TB, 0, st>>>(fa);
    if (fa.defer_second_from > 0) { sw_second_best_kernel<<<gwarp_pair, TB, 0, st>>>(fa); count_launch(e); }
    prof_mark(e, st, ST_FINALIZE);
    TraceArgs ta;
    ta.pool = reinterpret_cast<const uint4 *>(pool_f); ta.meta = meta; ta.state = state; ta.results = d_results; ta.npairs = npairs;
    ta.sc = sc;
    ta.cigar_pool = cigar_pool; ta.cigar_cursor = cigar_cursor; ta.cigar_cap = (unsigned long long)cig_pool_cap;
    ta.overflow_flag = overflow_flag;
    ta.debug = getenv("DB200_DEBUG") ? 1 : 0;
    ta.slab = slab; ta.slab_total = (unsigned long long)slab_bytes * (unsigned long long)wide_blocks;
    ta.medium_cursor = reinterpret_cast<unsigned long long *>(counters + 146);
    ta.list[0] = narrow_list; ta.list[1] = medium_list; ta.list[2] = wide_list; ta.list[3] = slab_scalar_list;
    for (int l = 0; l < 3; ++l) { ta.count[l] = counters + 136 + l; ta.counter[l] = counters + 140 + l; }
    ta.count[3] = counters + 139;
    ta.class_count = counters + 160;   // [4][TRACE_MEDIUM_CLASSES]
    ta.medium_min_jobs = TRACE_MEDIUM_MIN_JOBS;
    if (const char *mj = getenv("DB200_TRACE_MEDIUM_MIN_JOBS")) ta.medium_min_jobs = atoi(mj);
    const int warp_blocks = e->sm_count * 24;
    ta.coop_blocks[0] = 0; ta.coop_blocks[1] = warp_blocks; ta.coop_blocks[2] = wide_blocks;
    sw_trace_classify_kernel<<<gpair, TB, 0, st>>>(ta);
    {
        // level 0: one thread per pair, band rows and directions in a shared-memory slice
        const size_t sm0 = (size_t)TRACE_NARROW_THREADS * TRACE_NARROW_WORDS * sizeof(uint32_t);
        static bool attr0_done[16] = {false};
        if (!attr0_done[e->device & 15]) {
            DB200_CUDA_CHECK(cudaFuncSetAttribute(sw_trace_narrow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm0));
            attr0_done[e->device & 15] = true;
        }
        const unsigned gn = (unsigned)((npairs + TRACE_NARROW_THREADS - 1) / TRACE_NARROW_THREADS);
        sw_trace_narrow_kernel<<<gn, TRACE_NARROW_THREADS, sm0, st>>>(ta, 0);
        // the same kernel over the jobs of up to TRACE_MEDIUM_CELLS cells, directions in private regions of the slab
        const size_t sm3 = (size_t)TRACE_NARROW_THREADS * TRACE_MEDIUM_WORDS * sizeof(uint32_t);
        sw_trace_narrow_kernel<<<gn, TRACE_NARROW_THREADS, sm3, st>>>(ta, 3);
        count_launch(e);
    }
    {
        // level 1: one warp per pair (state for 2*62+1 diagonals, 8 KB of walk staging)
        constexpr int W1 = 126, S1 = 8192;
        const size_t sm1 = (((size_t)3 * (W1 + 2) * sizeof(int16_t) + 15) & ~(size_t)15) + S1;
        sw_trace_coop_kernel<32, 1, W1, 0, S1><<<warp_blocks, 32, sm1, st>>>(ta);
        // level 2: one block per pair (4094 diagonals, both sub-sequences staged, 48 KB of walk staging shared with nothing)
        constexpr int W2 = TRACE_WMAX, Q2 = TRACE_SEQ_WORDS, S2 = 24576;
        const size_t sm2 = (((size_t)3 * (W2 + 2) * sizeof(int16_t) + (size_t)Q2 * 4 + 15) & ~(size_t)15) + S2;
        static bool attr_done[16] = {false};
        if (!attr_done[e->device & 15]) {
            DB200_CUDA_CHECK(cudaFuncSetAttribute(sw_trace_coop_kernel<128, 2, W2, Q2, S2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
            attr_done[e->device & 15] = true;
        }
        sw_trace_coop_kernel<128, 2, W2, Q2, S2><<<wide_blocks, 128, sm2, st>>>(ta);
    }
    count_launch(e, 4);
    prof_mark(e, st, ST_TRACEBACK);
    sw_cigar_len_kernel<<<gpair, TB, 0, st>>>(state, npairs, lens);
    count_launch(e, 3);
    if (!link) {
        int64_t *off_out = d_cigar_off ? d_cigar_off : cig_off_tmp;
        exclusive_scan_i32_i64(e, lens, off_out, npairs, scan_tmp, st);
        sw_cigar_gather_kernel<<<gpair, TB, 0, st>>>(state, npairs, off_out, cigar_pool, d_cigar_buf, cigar_cap, d_status, overflow_flag,
                                                     nullptr, nullptr, nullptr);
    } else {
        // sub-batch of a larger call: local offsets, then wait for the predecessor's end before placing the cigars
        exclusive_scan_i32_i64(e, lens, cig_off_tmp, npairs, scan_tmp, st);
        if (link->wait) DB200_CUDA_CHECK(cudaStreamWaitEvent(st, link->wait, 0));
        sw_cigar_gather_kernel<<<gpair, TB, 0, st>>>(state, npairs, cig_off_tmp, cigar_pool, d_cigar_buf, cigar_cap, d_status, overflow_flag,
                                                     link->d_base, d_cigar_off, link->d_end);
        if (link->done) DB200_CUDA_CHECK(cudaEventRecord(link->done, st));
    }
    count_launch(e);
    prof_mark(e, st, ST_CIGAR);
    DB200_CUDA_CHECK(cudaGetLastError());
    if (cigar_total) {
        const int64_t *total_src = link ? link->d_end : (d_cigar_off ? d_cigar_off : cig_off_tmp) + npairs;
        DB200_CUDA_CHECK(cudaMemcpyAsync(cigar_total, total_src, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        DB200_CUDA_CHECK(cudaStreamSynchronize(st));
        if (d_cigar_buf && *cigar_total > cigar_cap) {   // the gather kernel skipped what did not fit (overflow flag 3)
            set_error("sw_batch_device: cigar buffer too small (%lld ops, capacity %lld)", (long long)*cigar_total, (long long)cigar_cap);
            return DB200_ERR_CAPACITY;
        }
    }
    return DB200_OK;
}

}  // namespace db200
