// Edit-operation path of one pair (edlib task "path"): which alignment the reference reports, not just its cost.
//
// Behaviour reproduced (dysgu/edlib/src/edlib.cpp): obtainAlignment (1177-1229) solves a sub-problem by plain traceback when
// its state would stay below 1 MB and otherwise splits the target in halves (obtainAlignmentHirschberg, 1247-1412): the
// split row is the first query index whose forward score in the last column of the left half plus the reverse score of
// the row below in the right half equals the optimum, then the two boundary cases, in that order.  The traceback
// (958-1157) walks from the bottom-right corner preferring "up" (insertion), then "left" (deletion), then the diagonal,
// and fills the rest of the first row / column once it leaves the matrix.  edlib evaluates all of this on its banded
// block representation; every cell those rules can select lies on an optimal path, where banded and full values agree,
// so the functions below work on the plain recurrence, evaluated with Myers' bit-vector sweep both for the score columns
// and for the leaves (whose columns keep their vertical delta vectors for the walk back).
//
// The code is plain C++ that also compiles on the host: tests/ed_path_host.cpp builds it with g++ and checks it against
// the reference there, so the device only has to run what the CPU already proved.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define DB200_HD __host__ __device__ inline
#else
#define DB200_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define DB200_POPC64(x) __popcll(x)
#else
#define DB200_POPC64(x) __builtin_popcountll(x)
#endif

namespace db200 {

constexpr int ED_PATH_LEAF_BYTES = 1 << 20;   // (Pv, Mv) of every column of the largest sub-problem below the 1 MB rule
constexpr int ED_PATH_STACK = 64;

struct EdPathScratch {
    uint64_t *Pv, *Mv;        // [ceil(m / 64)] each
    int32_t *left, *right;    // [m] score columns of the two halves (left doubles as the leaf sweep's running last rows)
    uint64_t *dir;            // [ED_PATH_LEAF_BYTES / 8] stored leaf: (Pv, Mv) and last-row value per word and column
    uint8_t *tmp;             // [m + n] operations of a leaf, last first
};

// 64 bits of a Peq row starting at bit `pos` (zero outside the row)
DB200_HD uint64_t path_bits(const uint64_t *row, int W, long pos)
{
    const long w = pos >= 0 ? pos / 64 : -((-pos + 63) / 64);
    const int sh = (int)(pos - w * 64);
    const uint64_t lo = (w >= 0 && w < W) ? row[w] : 0ull;
    const uint64_t hi = (w + 1 >= 0 && w + 1 < W) ? row[w + 1] : 0ull;
    return sh ? ((lo >> sh) | (hi << (64 - sh))) : lo;
}

DB200_HD uint64_t path_brev64(uint64_t x)
{
    x = ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
    x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
    x = ((x >> 8) & 0x00FF00FF00FF00FFull) | ((x & 0x00FF00FF00FF00FFull) << 8);
    x = ((x >> 16) & 0x0000FFFF0000FFFFull) | ((x & 0x0000FFFF0000FFFFull) << 16);
    return (x >> 32) | (x << 32);
}

// Global (NW) scores of every query prefix against `ncol` target symbols: forward = query[q0..q1) against target[t0..],
// reverse = reversed query[q0..q1) against the target read backwards from t1 - 1.  out[r] = distance of the first r + 1
// query symbols (in the sweep's orientation) to all ncol columns.
DB200_HD void path_score_column(const uint64_t *peq, const uint8_t *tcodes, int W, int q0, int q1, int tfirst, int ncol, bool reverse,
                                uint64_t *Pv, uint64_t *Mv, int32_t *out)
{
    const int qlen = q1 - q0;
    const int Wq = (qlen + 63) >> 6;
    for (int w = 0; w < Wq; ++w) { Pv[w] = ~0ull; Mv[w] = 0ull; }
    for (int j = 0; j < ncol; ++j) {
        const uint64_t *row = peq + (long)tcodes[reverse ? tfirst - j : tfirst + j] * W;
        uint64_t hp = 1, hm = 0;   // the row above the matrix grows by one per column
        for (int w = 0; w < Wq; ++w) {
            uint64_t Eq = reverse ? path_brev64(path_bits(row, W, (long)q1 - 64L * (w + 1))) : path_bits(row, W, (long)q0 + 64L * w);
            const uint64_t pv = Pv[w], mv = Mv[w];
            const uint64_t Xv = Eq | mv;
            Eq |= hm;
            const uint64_t Xh = (((Eq & pv) + pv) ^ pv) | Eq;
            uint64_t Ph = mv | ~(Xh | pv);
            uint64_t Mh = pv & Xh;
            const uint64_t op = Ph >> 63, om = Mh >> 63;
            Ph = (Ph << 1) | hp;
            Mh = (Mh << 1) | hm;
            Pv[w] = Mh | ~(Xv | Ph);
            Mv[w] = Ph & Xv;
            hp = op; hm = om;
        }
    }
    int s = ncol;
    for (int r = 0; r < qlen; ++r) {
        s += (int)((Pv[r >> 6] >> (r & 63)) & 1ull) - (int)((Mv[r >> 6] >> (r & 63)) & 1ull);
        out[r] = s;
    }
}

// Stored leaf: for every column j and 64-row word w the vertical delta vectors PM[(j * Wq + w) * 2 + {0, 1}] = (Pv, Mv) and,
// behind them, the value of the word's last row E[j * Wq + w] = D[64 * w + 63][j] - the reference keeps the same three
// things per block (edlib.cpp:1204-1206: 2 words + 1 int), so its 1 MB rule bounds this layout by ED_PATH_LEAF_BYTES.
// D[i][j] = the end of i's word minus the deltas of the rows below i in that word (rows past the query's end continue
// the recurrence consistently, whatever their match bits are).
DB200_HD int path_leaf_cell(const uint64_t *PM, const int32_t *E, int Wq, int i, int j)
{
    const long k = (long)j * Wq + (i >> 6);
    const uint64_t below = (i & 63) == 63 ? 0ull : (~0ull << ((i & 63) + 1));
    return E[k] - DB200_POPC64(PM[2 * k] & below) + DB200_POPC64(PM[2 * k + 1] & below);
}

// The walk back over the stored columns of a leaf (edlib.cpp:958-1157): from the bottom-right corner, preferring "up"
// (insertion), then "left" (deletion), then the diagonal; the rest of the first row / column once it leaves the matrix.
DB200_HD int path_leaf_walk(const uint64_t *PM, int qlen, int tlen, uint8_t *tmp, uint8_t *out)
{
    const int Wq = (qlen + 63) >> 6;
    const int32_t *E = reinterpret_cast<const int32_t *>(PM + (long)tlen * Wq * 2);
    int i = qlen - 1, j = tlen - 1, n = 0;
    int cur = path_leaf_cell(PM, E, Wq, i, j);
    int left = j > 0 ? path_leaf_cell(PM, E, Wq, i, j - 1) : i + 1;   // D[i][j-1]; column -1 holds i + 1
    while (i >= 0 && j >= 0) {
        const uint64_t *pc = PM + (long)j * Wq * 2 + 2 * (i >> 6);
        const int sh = i & 63;
        const int vc = (int)((pc[0] >> sh) & 1ull) - (int)((pc[1] >> sh) & 1ull);
        int vl = 1;
        if (j > 0) {
            const uint64_t *pl = pc - (long)Wq * 2;
            vl = (int)((pl[0] >> sh) & 1ull) - (int)((pl[1] >> sh) & 1ull);
        }
        const int up = cur - vc, diag = left - vl;
        const int op = (up + 1 == cur) ? 1 : ((left + 1 == cur) ? 2 : (diag == cur ? 0 : 3));
        tmp[n++] = (uint8_t)op;
        if (op == 1) { --i; cur = up; left = diag; }
        else {
            cur = (op == 2) ? left : diag;
            if (op != 2) --i;
            --j;
            if (i >= 0 && j >= 0) left = j > 0 ? path_leaf_cell(PM, E, Wq, i, j - 1) : i + 1;
        }
    }
    for (; j >= 0; --j) tmp[n++] = 2;                           // the rest of the first row: deletions
    for (; i >= 0; --i) tmp[n++] = 1;                           // the rest of the first column: insertions
    for (int k = 0; k < n; ++k) out[k] = tmp[n - 1 - k];
    return n;
}

// Traceback of a sub-problem small enough for the reference's direct method (edlib.cpp:1204-1221, 958-1157).  The sweep
// keeps the vertical delta vectors (Pv, Mv) and the last-row value of every word of every column, and the walk rebuilds the
// three neighbour values from them: up and diagonal from one bit each, left from one population count whenever the
// column changes.  Returns the number of operations written to
// out (in alignment order).
DB200_HD int path_leaf(const uint64_t *peq, const uint8_t *tcodes, int W, int q0, int q1, int t0, int t1, const EdPathScratch &s, uint8_t *out)
{
    const int qlen = q1 - q0, tlen = t1 - t0;
    const int Wq = (qlen + 63) >> 6;
    uint64_t *PM = s.dir;                                           // [tlen][Wq][2], then the words' last rows [tlen][Wq]
    int32_t *E = reinterpret_cast<int32_t *>(PM + (long)tlen * Wq * 2);
    int32_t *end = s.left;                                          // running value of every word's last row (column -1: 64 w + 64)
    for (int w = 0; w < Wq; ++w) { s.Pv[w] = ~0ull; s.Mv[w] = 0ull; end[w] = 64 * (w + 1); }
    for (int j = 0; j < tlen; ++j) {
        const uint64_t *row = peq + (long)tcodes[t0 + j] * W;
        uint64_t *pm = PM + (long)j * Wq * 2;
        uint64_t hp = 1, hm = 0;
        for (int w = 0; w < Wq; ++w) {
            uint64_t Eq = path_bits(row, W, (long)q0 + 64L * w);
            const uint64_t pv = s.Pv[w], mv = s.Mv[w];
            const uint64_t Xv = Eq | mv;
            Eq |= hm;
            const uint64_t Xh = (((Eq & pv) + pv) ^ pv) | Eq;
            uint64_t Ph = mv | ~(Xh | pv);
            uint64_t Mh = pv & Xh;
            const uint64_t op = Ph >> 63, om = Mh >> 63;
            Ph = (Ph << 1) | hp;
            Mh = (Mh << 1) | hm;
            const uint64_t npv = Mh | ~(Xv | Ph), nmv = Ph & Xv;
            s.Pv[w] = npv; s.Mv[w] = nmv;
            pm[2 * w] = npv; pm[2 * w + 1] = nmv;
            end[w] += (int)op - (int)om;
            E[(long)j * Wq + w] = end[w];
            hp = op; hm = om;
        }
    }
    return path_leaf_walk(PM, qlen, tlen, s.tmp, out);
}

// Operations (0 match, 1 insertion, 2 deletion, 3 mismatch) aligning the whole query of m symbols to target[t0..t1) at
// the optimal cost `best`.  Returns their number, or -1 if no split consistent with `best` exists (a wrong `best`).
DB200_HD int ed_path_pair(const uint64_t *peq, const uint8_t *tcodes, int W, int m, int t0, int t1, int best, const EdPathScratch &s, uint8_t *out)
{
    int stack[ED_PATH_STACK][5];
    int sp = 0, n = 0;
    stack[sp][0] = 0; stack[sp][1] = m; stack[sp][2] = t0; stack[sp][3] = t1; stack[sp][4] = best; ++sp;
    while (sp > 0) {
        --sp;
        const int q0 = stack[sp][0], q1 = stack[sp][1], a0 = stack[sp][2], a1 = stack[sp][3], score = stack[sp][4];
        const int qlen = q1 - q0, tlen = a1 - a0;
        if (qlen == 0 || tlen == 0) {
            for (int k = 0; k < tlen; ++k) out[n++] = 2;
            for (int k = 0; k < qlen; ++k) out[n++] = 1;
            continue;
        }
        const long long blocks = (qlen + 63) >> 6;
        const long long state = (2ll * 8 + 4) * blocks * tlen + 2ll * 4 * tlen;
        if (state < 1024 * 1024) {
            n += path_leaf(peq, tcodes, W, q0, q1, a0, a1, s, out + n);
            continue;
        }
        const int lw = tlen / 2, rw = tlen - lw;
        path_score_column(peq, tcodes, W, q0, q1, a0, lw, false, s.Pv, s.Mv, s.left);
        path_score_column(peq, tcodes, W, q0, q1, a1 - 1, rw, true, s.Pv, s.Mv, s.right);
        // s.right[r] = cost of the last r + 1 query symbols against the right half
        int idx = -2, ls = 0, rs = 0;
        for (int qi = 0; qi <= qlen - 2; ++qi) {
            const int l = s.left[qi], r = s.right[qlen - 2 - qi];
            if (l + r == score) { idx = qi; ls = l; rs = r; break; }
        }
        if (idx == -2 && lw + s.right[qlen - 1] == score) { idx = -1; ls = lw; rs = s.right[qlen - 1]; }
        if (idx == -2 && s.left[qlen - 1] + rw == score) { idx = qlen - 1; ls = s.left[qlen - 1]; rs = rw; }
        if (idx == -2 || sp + 2 > ED_PATH_STACK) return -1;
        const int qs = q0 + idx + 1;
        // lower-right part below, upper-left part on top: it is expanded (and written) first
        stack[sp][0] = qs; stack[sp][1] = q1; stack[sp][2] = a0 + lw; stack[sp][3] = a1; stack[sp][4] = rs; ++sp;
        stack[sp][0] = q0; stack[sp][1] = qs; stack[sp][2] = a0; stack[sp][3] = a0 + lw; stack[sp][4] = ls; ++sp;
    }
    return n;
}

}  // namespace db200
