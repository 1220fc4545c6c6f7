// One-launch Smith-Waterman for small batches: the latency path behind the synchronous drop-in call.
//
// An unmodified dysgu calls StripedSmithWaterman(q)(t) one pair at a time (_ssw_wrapper.pyx:613-648 -> ssw_init +
// ssw_align, ssw.c:743-857); through the batch pipeline of sw.cu that is ~50 kernel launches on 24 streams for a single
// pair (580 us per call on B200).  Here one warp owns one pair and runs the whole of ssw_align in ONE kernel: encode
// (_ssw_wrapper.pyx:60-68), forward pass (ssw.c:124-346 / 372-548 semantics: first column holding the maximum, smallest
// row in it, per-column maxima incl. the padded profile rows for the second best), reverse pass (ssw.c:827-840), banded
// traceback with band doubling (ssw.c:550-728) and the cigar.  Inputs are read from, and results written to,
// page-locked host memory mapped into the device address space, so a call is: fill the staging area, one launch, one
// stream synchronisation.
//
// Mapping: the target (columns, <= 512) is cut into 32 strips of CPT = 2 / 4 / 8 / 16 columns, one per lane; lane l works on
// row s - l at step s, so the warp sweeps anti-diagonals of strips and the three values a strip needs from its left
// neighbour (H and the horizontal gap state of the same row, H of the row above) arrive by __shfl_up_sync.  Plain int32
// arithmetic: this kernel is bound by the latency of one dependent chain, not by ALU throughput.
// Pairs that do not fit (query > 8,192, target > 512, traceback state beyond the shared-memory budget, more than
// SW_SMALL_CIGAR_STRIDE cigar ops) are flagged SW_SMALL_FALLBACK and taken by the batch pipeline.
#pragma once
#include "sw_kernels.cuh"

namespace db200 {

constexpr int SW_SMALL_MAX_PAIRS = 64;
constexpr int SW_SMALL_MAX_Q = 8192;
constexpr int SW_SMALL_MAX_T = 512;
constexpr int SW_SMALL_XW = 1022;            // band diagonals of the traceback sweep
constexpr int SW_SMALL_DIR_BYTES = 48 * 1024; // direction bytes of the traceback (rows x padded band width)
constexpr int SW_SMALL_RUNS = 1024;          // cigar runs collected by the walk
constexpr int SW_SMALL_CIGAR_STRIDE = 512;   // cigar ops per pair in the output area
constexpr int SW_SMALL_FALLBACK = 100;       // status: not handled here

struct SmallArgs {
    const uint8_t *seq;          // all queries then all targets, ASCII
    const int32_t *qoff, *toff;  // [n + 1] offsets into seq
    const int32_t *mask_len;     // [n]
    db200_sw_result *results;    // [n]
    int32_t *status;             // [n]
    uint32_t *cigar;             // [n][SW_SMALL_CIGAR_STRIDE]
    volatile int32_t *done;      // [n] mapped host memory: receives `epoch` once pair p's outputs are visible to the host
    int32_t epoch;
    int32_t maxq;                // shared-memory layout: bytes reserved for the query codes (multiple of 16)
    int32_t bias, score_size, flag;
    SwScoring sc;
};

__device__ __forceinline__ uint8_t small_code(uint8_t ch)
{
    switch (ch) {
        case 'A': case 'a': case 'U': case 'u': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return 4;
    }
}

// One scoring pass over rows q[qbase + qstep * i], i < m, and columns t[tbase + tstep * j], j < n.
// Returns (warp-uniform) the maximum, the first column holding it and the smallest row holding it in that column;
// optionally the per-column maxima and the last row.
template <int CPT>
__device__ __forceinline__ void small_pass(const uint8_t *qs, int qbase, int qstep, int m, const uint8_t *ts, int tbase, int tstep, int n,
                                           const int16_t *mat_s, int go, int ge, int lane, int16_t *colmax, int16_t *lastrow,
                                           int &bscore, int &bcol, int &brow)
{
    const unsigned FULL = 0xFFFFFFFFu;
    int Hrow[CPT], Frow[CPT], cmax[CPT], tcode[CPT];
    const int c0 = lane * CPT;
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
        Hrow[k] = 0; Frow[k] = 0; cmax[k] = 0;
        tcode[k] = (c0 + k < n) ? (int)ts[tbase + tstep * (c0 + k)] * 6 : SW_PAD_CODE * 6;
    }
    int out_h = 0, out_e = 0, prev_out_h = 0;
    int bs = 0, bc = 0x7FFFFFFF, br = 0;
    const int steps = m + 31;
    for (int s = 0; s < steps; ++s) {
        // from the strip on the left: the horizontal gap state leaving its last column in this row, and H of the row above
        int in_e = __shfl_up_sync(FULL, out_e, 1), in_d = __shfl_up_sync(FULL, prev_out_h, 1);
        if (lane == 0) { in_e = 0; in_d = 0; }
        const int r = s - lane;
        if (r >= 0 && r < m) {
            const int qc = qs[qbase + qstep * r];
            int el = in_e, diag = in_d, h = 0;
#pragma unroll
            for (int k = 0; k < CPT; ++k) {
                const int up = Hrow[k];
                const int f = Frow[k];
                h = max(max(diag + (int)mat_s[tcode[k] + qc], el), max(f, 0));
                diag = up;
                Hrow[k] = h;
                el = max(max(el - ge, h - go), 0);
                Frow[k] = max(max(f - ge, h - go), 0);
                cmax[k] = max(cmax[k], h);
                if (h > bs || (h == bs && c0 + k < bc)) { bs = h; bc = c0 + k; br = r; }
            }
            prev_out_h = out_h;
            out_h = h;
            out_e = el;
            if (lastrow && r == m - 1) {
#pragma unroll
                for (int k = 0; k < CPT; ++k) if (c0 + k < n) lastrow[c0 + k] = (int16_t)Hrow[k];
            }
        }
    }
    if (colmax) {
#pragma unroll
        for (int k = 0; k < CPT; ++k) if (c0 + k < n) colmax[c0 + k] = (int16_t)cmax[k];
    }
    // (largest score, smallest column, smallest row)
    unsigned long long key = bs > 0 ? (((unsigned long long)(unsigned)bs << 40) | ((unsigned long long)(0xFFFFF - bc) << 20) | (unsigned long long)(0xFFFFF - br)) : 0ull;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(FULL, key, o);
        key = other > key ? other : key;
    }
    bscore = (int)(key >> 40);
    bcol = bscore > 0 ? 0xFFFFF - (int)((key >> 20) & 0xFFFFF) : 0;
    brow = bscore > 0 ? 0xFFFFF - (int)(key & 0xFFFFF) : 0;
    __syncwarp();
}

__device__ __forceinline__ void small_pass_any(int n_cols, const uint8_t *qs, int qbase, int qstep, int m, const uint8_t *ts, int tbase,
                                               int tstep, int n, const int16_t *mat_s, int go, int ge, int lane, int16_t *colmax,
                                               int16_t *lastrow, int &bscore, int &bcol, int &brow)
{
    if (n_cols <= 64) small_pass<2>(qs, qbase, qstep, m, ts, tbase, tstep, n, mat_s, go, ge, lane, colmax, lastrow, bscore, bcol, brow);
    else if (n_cols <= 128) small_pass<4>(qs, qbase, qstep, m, ts, tbase, tstep, n, mat_s, go, ge, lane, colmax, lastrow, bscore, bcol, brow);
    else if (n_cols <= 256) small_pass<8>(qs, qbase, qstep, m, ts, tbase, tstep, n, mat_s, go, ge, lane, colmax, lastrow, bscore, bcol, brow);
    else small_pass<16>(qs, qbase, qstep, m, ts, tbase, tstep, n, mat_s, go, ge, lane, colmax, lastrow, bscore, bcol, brow);
}

__global__ void __launch_bounds__(32) sw_small_kernel(SmallArgs a)
{
    extern __shared__ __align__(16) uint8_t sm_raw[];
    __shared__ int16_t mat_s[36];
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x, p = blockIdx.x;
    uint8_t *qs = sm_raw;
    uint8_t *ts = qs + a.maxq;
    int16_t *colmax = reinterpret_cast<int16_t *>(ts + SW_SMALL_MAX_T);
    int16_t *lastrow = colmax + SW_SMALL_MAX_T;
    int16_t *band = lastrow + SW_SMALL_MAX_T;                       // [3][SW_SMALL_XW + 2]
    uint32_t *runs = reinterpret_cast<uint32_t *>(band + 3 * (SW_SMALL_XW + 2));
    uint8_t *dir = reinterpret_cast<uint8_t *>(runs + SW_SMALL_RUNS);
    for (int k = lane; k < 36; k += 32) mat_s[k] = a.sc.mat[k];

    const int m = a.qoff[p + 1] - a.qoff[p], n = a.toff[p + 1] - a.toff[p];
    db200_sw_result r;
    r.score1 = 0; r.score2 = 0; r.ref_begin1 = -1; r.ref_end1 = -1; r.read_begin1 = -1; r.read_end1 = -1; r.ref_end2 = 0; r.cigar_len = 0;
    int status = DB200_PAIR_OK;
    if (m <= 0 || n <= 0) status = DB200_PAIR_EMPTY;
    else if (m > SW_SMALL_MAX_Q || n > SW_SMALL_MAX_T || m > a.maxq) status = SW_SMALL_FALLBACK;
    else {
        int mm = 0;
        for (int k = 0; k < 25; ++k) mm = max(mm, (int)a.sc.mat[(k / 5) * 6 + k % 5]);
        if ((long long)mm * min(m, n) > 30000) status = DB200_PAIR_SCORE_RANGE;
    }
    if (status != DB200_PAIR_OK) {
        if (lane == 0) { a.results[p] = r; a.status[p] = status; __threadfence_system(); a.done[p] = a.epoch; }
        return;
    }
    // ---- encode (_ssw_wrapper.pyx:60-68) ---------------------------------------------------------------
    {
        const uint8_t *q = a.seq + a.qoff[p], *t = a.seq + a.toff[p];
        for (int i = lane; i < m; i += 32) qs[i] = small_code(q[i]);
        for (int j = lane; j < n; j += 32) ts[j] = small_code(t[j]);
    }
    __syncwarp();
    const int go = a.sc.gap_open, ge = a.sc.gap_extend;
    const int mask_len = a.mask_len[p];
    const bool need_colmax = mask_len >= 15 && n > mask_len - 1;

    // ---- forward pass ------------------------------------------------------------------------------------
    int score, end_t, end_q;
    small_pass_any(n, qs, 0, 1, m, ts, 0, 1, n, mat_s, go, ge, lane, need_colmax ? colmax : nullptr, need_colmax ? lastrow : nullptr, score,
                   end_t, end_q);
    const bool overflow8 = score + a.bias >= 255;
    const bool use_word = (a.score_size == 1) || (a.score_size == 2 && overflow8);
    if (a.score_size == 0 && overflow8) {
        if (lane == 0) { a.results[p] = r; a.status[p] = DB200_PAIR_OVERFLOW; __threadfence_system(); a.done[p] = a.epoch; }
        return;
    }
    r.score1 = score;
    if (score == 0) { r.ref_end1 = use_word ? 0 : -1; r.read_end1 = 0; }   // ssw.c:146 vs ssw.c:389
    else { r.ref_end1 = end_t; r.read_end1 = end_q; }
    // ---- second best (ssw.c:326-341 / 530-543, padded profile rows ssw.c:109 / 364) ---------------------------
    if (mask_len < 15) { r.score2 = 0; r.ref_end2 = -1; }
    else if (need_colmax && score > 0) {
        const int L = use_word ? 8 : 16;
        const int P = ((m + L - 1) / L) * L - m;
        int best = 0, bref = 0;
        const int hi1 = max(r.ref_end1 - mask_len, 0);
        const int lo2 = min(r.ref_end1 + mask_len, n) + (use_word ? 0 : 1);
        for (int pass = 0; pass < 2; ++pass) {
            const int lo = pass == 0 ? 0 : lo2, hi = pass == 0 ? hi1 : n;
            for (int j = lo + lane; j < hi; j += 32) {
                int x = colmax[j];
                for (int k = 0; k < P && j - 1 - k >= 0; ++k) x = max(x, (int)lastrow[j - 1 - k]);
                if (x > best) { best = x; bref = j; }
            }
        }
        unsigned long long key = ((unsigned long long)(unsigned)best << 32) | (unsigned)(0x7FFFFFFF - bref);
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(FULL, key, o);
            key = other > key ? other : key;
        }
        const int b = (int)(key >> 32);
        r.score2 = b;
        r.ref_end2 = b > 0 ? 0x7FFFFFFF - (int)(key & 0xFFFFFFFFu) : 0;
    }
    // ---- begin coordinates: reverse pass (ssw.c:827-840) -------------------------------------------------------
    const int flag = a.flag;
    const bool want_begin = flag != 0;        // filters are not handled here (flag is one of 0, 1, 8, 9)
    bool want_cigar = false;
    if (want_begin) {
        if (score == 0) { r.ref_begin1 = use_word ? 0 : -1; r.read_begin1 = 0; }
        else {
            int rs, rt, rq;
            small_pass_any(end_t + 1, qs, end_q, -1, end_q + 1, ts, end_t, -1, end_t + 1, mat_s, go, ge, lane, nullptr, nullptr, rs, rt, rq);
            r.ref_begin1 = end_t - rt;
            r.read_begin1 = end_q - rq;
        }
        want_cigar = (flag & 7) != 0;
    }
    // ---- banded traceback (ssw.c:550-728) ---------------------------------------------------------------------
    int ncig = 0;
    if (want_cigar) {
        if (score == 0) { if (lane == 0) runs[0] = 16u; ncig = 1; }   // 1x1 rectangle with score 0: 1M (ssw.c:844-847, 690-698)
        else {
            const int refLen = r.ref_end1 - r.ref_begin1 + 1, readLen = r.read_end1 - r.read_begin1 + 1;
            const int rbeg = r.read_begin1, tbeg = r.ref_begin1;
            bool done = false;
            if (refLen == readLen) {   // gap-free shortcut: the diagonal already collects the optimum (see sw_trace_classify_kernel)
                int sum = 0;
                for (int i = lane; i < readLen; i += 32) sum += mat_s[ts[tbeg + i] * 6 + qs[rbeg + i]];
                sum = __reduce_add_sync(FULL, sum);
                if (sum == score) { if (lane == 0) runs[0] = ((uint32_t)readLen) << 4; ncig = 1; done = true; }
            }
            if (!done) {
                int16_t *Hs = band + 1, *Es = band + (SW_SMALL_XW + 2) + 1, *Fs = band + 2 * (SW_SMALL_XW + 2) + 1;
                int bw = abs(refLen - readLen) + 1, wst = 0;
                bool fits = true;
                while (true) {
                    const int OFF = min(bw, readLen - 1);
                    const int XW = OFF + min(bw, refLen - 1) + 1;
                    const int width = min(2 * bw + 1, refLen);
                    wst = width;
                    if (XW > SW_SMALL_XW || (long long)readLen * wst > SW_SMALL_DIR_BYTES) { fits = false; break; }
                    for (int x = lane - 1; x <= XW; x += 32) { Hs[x] = 0; Es[x] = 0; Fs[x] = 0; }
                    __syncwarp();
                    int best = 0;
                    const int last_tau = readLen + refLen - 2;
                    for (int tau = 0; tau <= last_tau; ++tau) {
                        const int ilo = max(max(0, tau - refLen + 1), (tau - bw + 1) >> 1);
                        const int ihi = min(min(readLen - 1, tau), (tau + bw) >> 1);
                        for (int i = ilo + lane; i <= ihi; i += 32) {
                            const int j = tau - i;
                            const int x = j - i + OFF;
                            const int hup = Hs[x + 1], eup = Es[x + 1];
                            const int hleft = Hs[x - 1], fleft = Fs[x - 1];
                            const int hdg = Hs[x];
                            int t1 = (i == 0) ? -go : hup - go;
                            int t2 = (i == 0) ? -ge : eup - ge;
                            const int e = t1 > t2 ? t1 : t2;
                            const int e_open = t1 > t2;
                            t1 = hleft - go;
                            t2 = fleft - ge;
                            const int f = t1 > t2 ? t1 : t2;
                            const int f_open = t1 > t2;
                            const int e1 = e > 0 ? e : 0, f1 = f > 0 ? f : 0;
                            const int g = e1 > f1 ? e1 : f1;
                            const int d = hdg + mat_s[ts[tbeg + j] * 6 + qs[rbeg + i]];
                            const int h = g > d ? g : d;
                            best = max(best, h);
                            const int src = (g <= d) ? 0 : (e1 > f1 ? 1 : 2);
                            dir[i * wst + (j - max(0, i - bw))] = (uint8_t)(src | (e_open << 2) | (f_open << 3));
                            Hs[x] = (int16_t)h; Es[x] = (int16_t)max(e, -32768); Fs[x] = (int16_t)max(f, -32768);
                        }
                        __syncwarp();
                    }
                    best = __reduce_max_sync(FULL, best);
                    if (best >= score) break;
                    bw *= 2;
                }
                if (!fits) status = SW_SMALL_FALLBACK;
                else {
                    int l = 0;
                    if (lane == 0) {
                        const int width = min(2 * bw + 1, refLen);
                        int i = readLen - 1, j = refLen - 1, run = 0, prev = 0, op = 0, state = 2;
                        bool bad = false;
                        while (i > 0) {
                            const int x = j - max(0, i - bw);
                            if (x < 0 || x >= width || j < 0) { bad = true; break; }   // the reference would read out of bounds here
                            switch (trace_step_code(dir[i * wst + x], state)) {
                                case 1: --i; --j; state = 2; op = 0; break;
                                case 2: --i; state = 0; op = 1; break;
                                case 3: --i; state = 2; op = 1; break;
                                case 4: --j; state = 1; op = 2; break;
                                default: --j; state = 2; op = 2; break;
                            }
                            if (op == prev) ++run;
                            else {
                                if (l >= SW_SMALL_RUNS - 2) { bad = true; break; }
                                runs[l++] = ((uint32_t)run << 4) | (uint32_t)prev; prev = op; run = 1;
                            }
                        }
                        if (bad) l = -1;
                        else if (op == 0) runs[l++] = ((uint32_t)(run + 1)) << 4;
                        else { runs[l++] = ((uint32_t)run << 4) | (uint32_t)op; runs[l++] = 16u; }
                        // written in walk order: reversed below
                        for (int x = 0, y = l - 1; x < y; ++x, --y) { const uint32_t t = runs[x]; runs[x] = runs[y]; runs[y] = t; }
                    }
                    l = __shfl_sync(FULL, l, 0);
                    if (l < 0 || l > SW_SMALL_CIGAR_STRIDE) status = SW_SMALL_FALLBACK;
                    else ncig = l;
                }
            }
        }
    }
    __syncwarp();
    if (status == DB200_PAIR_OK && ncig > 0) {
        uint32_t *out = a.cigar + (size_t)p * SW_SMALL_CIGAR_STRIDE;
        for (int k = lane; k < ncig; k += 32) out[k] = runs[k];
        r.cigar_len = ncig;
    }
    __syncwarp();
    if (lane == 0) { a.results[p] = r; a.status[p] = status; __threadfence_system(); a.done[p] = a.epoch; }
}

}  // namespace db200
