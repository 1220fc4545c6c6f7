/*
 * TEST INFRASTRUCTURE ONLY - never linked, imported or executed by the product path.
 *
 * Plain scalar restatement of edlib's task "path" (reference: dysgu/edlib/src/edlib.cpp): which of the
 * equal-cost alignments the reference reports.  Deliberately NOT bit-parallel and NOT sharing code with
 * the product (dysgu_b200/csrc/ed_path.cuh): integer matrices and one rolling column, so that the CUDA
 * path, its host build and this file are three routes to the same answer.
 *
 *   driver           edlib.cpp:268-285    global alignment of the query to target[start0 .. end0] of the
 *                                          first location; nothing when the distance is above k or a
 *                                          sequence is empty (161-179 return before this step)
 *   obtainAlignment  edlib.cpp:1177-1229  direct traceback while (2*8+4)*ceil(m/64)*n + 2*4*n < 1 MB,
 *                                          else the halving below; an empty side = insertions / deletions
 *   traceback        edlib.cpp:958-1157   from the bottom-right corner: up (insertion) if it explains the
 *                                          cell, else left (deletion), else the diagonal (match / mismatch)
 *   Hirschberg       edlib.cpp:1247-1412  halve the target; the split row is the first query index whose
 *                                          forward value in the left half's last column plus the reverse
 *                                          value of the row below in the right half equals the optimum,
 *                                          then the boundary above the matrix, then the one below
 *
 * edlib evaluates these on banded blocks; every cell the rules can select lies on an optimal path, where
 * banded and full values agree, so the plain recurrence gives the same alignment.
 *
 * Parity status: PINNED - tests/test_oracle_golden.py checks it against tests/golden/ed_path.json (vectors
 * from the reference's compiled wrapper) and oracle/validate_oracle.py against oracle/_ref on random pairs.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "ed_oracle.h"

typedef struct { unsigned char eq[256][256]; } eqfull;

static void eqfull_init(eqfull *m, const char *pairs, int npairs)
{
    memset(m, 0, sizeof(*m));
    for (int a = 0; a < 256; ++a) m->eq[a][a] = 1;
    for (int i = 0; i < npairs; ++i) { /* symmetric, not transitive: edlib.cpp:72-80 */
        unsigned char a = (unsigned char)pairs[2 * i], b = (unsigned char)pairs[2 * i + 1];
        m->eq[a][b] = 1;
        m->eq[b][a] = 1;
    }
}

/* col[r] = NW distance of the first r + 1 symbols of q (walked with qstep) to ncol symbols of t (walked with tstep) */
static void score_column(const eqfull *eq, const unsigned char *q, int qlen, int qstep, const unsigned char *t, int ncol, int tstep,
                         int *col)
{
    for (int r = 0; r < qlen; ++r) col[r] = r + 1;
    for (int j = 0; j < ncol; ++j) {
        const unsigned char tc = t[(long)j * tstep];
        int diag = j, up = j + 1;
        for (int r = 0; r < qlen; ++r) {
            const int left = col[r];
            int cur = diag + (eq->eq[q[(long)r * qstep]][tc] ? 0 : 1);
            if (up + 1 < cur) cur = up + 1;
            if (left + 1 < cur) cur = left + 1;
            diag = left;
            col[r] = cur;
            up = cur;
        }
    }
}

/* direct traceback of q[0..qlen) against t[0..tlen); returns the number of operations appended to out */
static int leaf(const eqfull *eq, const unsigned char *q, int qlen, const unsigned char *t, int tlen, unsigned char *out)
{
    const long stride = (long)tlen + 1;
    int *D = (int *)malloc(sizeof(int) * (size_t)((qlen + 1) * stride));
    unsigned char *rev = (unsigned char *)malloc((size_t)qlen + (size_t)tlen + 1);
    if (!D || !rev) { free(D); free(rev); return -1; }
    for (int j = 0; j <= tlen; ++j) D[j] = j;
    for (int i = 1; i <= qlen; ++i) {
        int *row = D + i * stride, *prev = row - stride;
        row[0] = i;
        for (int j = 1; j <= tlen; ++j) {
            int cur = prev[j - 1] + (eq->eq[q[i - 1]][t[j - 1]] ? 0 : 1);
            if (prev[j] + 1 < cur) cur = prev[j] + 1;
            if (row[j - 1] + 1 < cur) cur = row[j - 1] + 1;
            row[j] = cur;
        }
    }
    int i = qlen, j = tlen, n = 0;
    while (i > 0 && j > 0) {
        const int cur = D[i * stride + j];
        if (D[(i - 1) * stride + j] + 1 == cur) { rev[n++] = 1; --i; }             /* up: insertion */
        else if (D[i * stride + j - 1] + 1 == cur) { rev[n++] = 2; --j; }          /* left: deletion */
        else { rev[n++] = (unsigned char)(D[(i - 1) * stride + j - 1] == cur ? 0 : 3); --i; --j; }
    }
    while (j > 0) { rev[n++] = 2; --j; }
    while (i > 0) { rev[n++] = 1; --i; }
    for (int k = 0; k < n; ++k) out[k] = rev[n - 1 - k];
    free(D);
    free(rev);
    return n;
}

static int solve(const eqfull *eq, const unsigned char *q, int qlen, const unsigned char *t, int tlen, int best, unsigned char *out)
{
    if (qlen == 0 || tlen == 0) {
        for (int k = 0; k < tlen + qlen; ++k) out[k] = (unsigned char)(qlen == 0 ? 2 : 1);
        return tlen + qlen;
    }
    const long long blocks = (qlen + 63) / 64;
    const long long state = (2ll * 8 + 4) * blocks * tlen + 2ll * 4 * tlen;
    if (state < 1024 * 1024) return leaf(eq, q, qlen, t, tlen, out);
    const int lw = tlen / 2, rw = tlen - lw;
    int *left = (int *)malloc(sizeof(int) * (size_t)qlen), *right = (int *)malloc(sizeof(int) * (size_t)qlen);
    if (!left || !right) { free(left); free(right); return -1; }
    score_column(eq, q, qlen, 1, t, lw, 1, left);                              /* left[r]: q[0..r] vs the left half */
    score_column(eq, q + qlen - 1, qlen, -1, t + tlen - 1, rw, -1, right);     /* right[r]: the last r + 1 symbols vs the right half */
    int idx = -2, ls = 0, rs = 0;
    for (int qi = 0; qi <= qlen - 2; ++qi)
        if (left[qi] + right[qlen - 2 - qi] == best) { idx = qi; ls = left[qi]; rs = right[qlen - 2 - qi]; break; }
    if (idx == -2 && lw + right[qlen - 1] == best) { idx = -1; ls = lw; rs = right[qlen - 1]; }
    if (idx == -2 && left[qlen - 1] + rw == best) { idx = qlen - 1; ls = left[qlen - 1]; rs = rw; }
    free(left);
    free(right);
    if (idx == -2) return -1;
    const int ul = idx + 1;
    const int n1 = solve(eq, q, ul, t, lw, ls, out);
    if (n1 < 0) return -1;
    const int n2 = solve(eq, q + ul, qlen - ul, t + lw, rw, rs, out + n1);
    return n2 < 0 ? -1 : n1 + n2;
}

/*
 * edlib.align(query, target, mode, task="path", k, additionalEqualities): the fields of ed_oracle_align plus the edit
 * operations (0 match, 1 insertion, 2 deletion, 3 mismatch) in ops (capacity m + n); *n_ops = 0 where the reference
 * returns no alignment.  Returns 0, or -1 on an internal failure (allocation, no split consistent with the distance).
 */
int ed_path_oracle_align(const char *query, int32_t m, const char *target, int32_t n, int32_t mode, int32_t k,
                         const char *eq_pairs, int32_t n_eq_pairs, ed_oracle_result *r, unsigned char *ops, int32_t *n_ops)
{
    *n_ops = 0;
    if (ed_oracle_align(query, m, target, n, mode, ED_TASK_PATH, k, eq_pairs, n_eq_pairs, r) != 0) return -1;
    if (r->status != 0 || r->editDistance < 0 || r->numLocations <= 0 || m == 0 || n == 0) return 0;
    const int start = r->startLocations ? r->startLocations[0] : 0, end = r->endLocations[0];
    eqfull *eq = (eqfull *)malloc(sizeof(eqfull));
    if (!eq) return -1;
    eqfull_init(eq, eq_pairs, eq_pairs ? n_eq_pairs : 0);
    const int len = solve(eq, (const unsigned char *)query, m, (const unsigned char *)target + start, end - start + 1, r->editDistance, ops);
    free(eq);
    if (len < 0) return -1;
    *n_ops = len;
    return 0;
}

/* CSR batch in the layout of the reference driver (ref_driver.c: ref_ed_path_batch): pair p writes its operations at
 * path_buf[(qoff[p] - qoff[0]) + (toff[p] - toff[0]) ..), path_len[p] of them. */
int ed_path_oracle_batch(const char *qbuf, const int64_t *qoff, const char *tbuf, const int64_t *toff, int64_t npairs, int32_t mode,
                         const int32_t *k, const char *eq_pairs, int32_t n_eq_pairs, int32_t *out_fixed, int32_t *loc_buf,
                         int64_t loc_cap, int64_t *loc_off, unsigned char *path_buf, int32_t *path_len)
{
    int64_t pos = 0;
    for (int64_t p = 0; p < npairs; ++p) {
        ed_oracle_result r;
        int32_t nops = 0;
        const int32_t m = (int32_t)(qoff[p + 1] - qoff[p]), n = (int32_t)(toff[p + 1] - toff[p]);
        if (ed_path_oracle_align(qbuf + qoff[p], m, tbuf + toff[p], n, mode, k ? k[p] : -1, eq_pairs, n_eq_pairs, &r,
                                 path_buf + (qoff[p] - qoff[0]) + (toff[p] - toff[0]), &nops) != 0) return -1;
        path_len[p] = nops;
        int32_t *o = out_fixed + p * 4;
        o[0] = r.status; o[1] = r.editDistance; o[2] = r.alphabetLength; o[3] = r.numLocations;
        loc_off[p] = pos;
        for (int a = 0; a < r.numLocations; ++a) {
            if (pos >= loc_cap) { ed_oracle_free(&r); return -2; }
            loc_buf[2 * pos] = r.startLocations ? r.startLocations[a] : INT32_MIN;
            loc_buf[2 * pos + 1] = r.endLocations[a];
            ++pos;
        }
        ed_oracle_free(&r);
    }
    loc_off[npairs] = pos;
    return 0;
}
