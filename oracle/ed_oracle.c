/*
 * TEST INFRASTRUCTURE ONLY - never linked, imported or executed by the product path.
 *
 * Plain full-matrix dynamic-programming restatement of the observable results of dysgu's
 * vendored edlib (reference: dysgu/edlib/src/edlib.cpp).  It is deliberately NOT a bit-vector
 * implementation: the reference's Myers/Ukkonen machinery only skips cells that are provably
 * larger than k, so its outputs equal those of the unbanded unit-cost DP written here
 * (edlib.cpp:546-709 for HW/SHW, 735-944 for NW, 141-296 for the driver).
 *
 * Parity status: PINNED.  oracle/validate_oracle.py compares every field with
 * oracle/_ref/libedlib_ref.so (the reference's edlib.cpp compiled unmodified) on randomized and
 * adversarial pairs; tests/golden/ holds vectors generated from that compiled reference.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "ed_oracle.h"

static inline int imin(int a, int b) { return a < b ? a : b; }

typedef struct { unsigned char eq[256][256 / 8]; } eqmat;

static void eq_init(eqmat *m, const char *pairs, int npairs)
{
    memset(m, 0, sizeof(*m));
    for (int a = 0; a < 256; ++a) m->eq[a][a >> 3] |= (unsigned char)(1u << (a & 7));
    for (int i = 0; i < npairs; ++i) { /* symmetric, not transitive: edlib.cpp:72-80 */
        unsigned char a = (unsigned char)pairs[2 * i], b = (unsigned char)pairs[2 * i + 1];
        m->eq[a][b >> 3] |= (unsigned char)(1u << (b & 7));
        m->eq[b][a >> 3] |= (unsigned char)(1u << (a & 7));
    }
}
static inline int eq_test(const eqmat *m, unsigned char a, unsigned char b) { return (m->eq[a][b >> 3] >> (b & 7)) & 1; }

/*
 * Last DP row.  row[p+1] = D[m][p] for p = -1..n-1.
 * free_start != 0: D[0][p] = 0 (HW, gap before the query is free, edlib.cpp:583).
 * free_start == 0: D[0][p] = p+1 (SHW / NW).
 */
static int *last_row(const unsigned char *q, int m, const unsigned char *t, int n, int free_start, const eqmat *eq)
{
    int *prev = (int *)malloc(sizeof(int) * (size_t)(n + 1));
    int *cur = (int *)malloc(sizeof(int) * (size_t)(n + 1));
    for (int p = 0; p <= n; ++p) prev[p] = free_start ? 0 : p;
    for (int i = 1; i <= m; ++i) {
        cur[0] = i;
        unsigned char qc = q[i - 1];
        for (int p = 1; p <= n; ++p) {
            int sub = prev[p - 1] + (eq_test(eq, qc, t[p - 1]) ? 0 : 1);
            int del = cur[p - 1] + 1;
            int ins = prev[p] + 1;
            cur[p] = imin(sub, imin(del, ins));
        }
        int *x = prev; prev = cur; cur = x;
    }
    free(cur);
    return prev;
}

int ed_oracle_align(const char *query, int32_t m, const char *target, int32_t n,
                    int32_t mode, int32_t task, int32_t k,
                    const char *eq_pairs, int32_t n_eq_pairs, ed_oracle_result *r)
{
    const unsigned char *q = (const unsigned char *)query, *t = (const unsigned char *)target;
    memset(r, 0, sizeof(*r));
    r->editDistance = -1;

    /* alphabet: distinct bytes of query and target (edlib.cpp:1433-1472) */
    unsigned char seen[256];
    memset(seen, 0, sizeof(seen));
    int alpha = 0;
    for (int i = 0; i < m; ++i) if (!seen[q[i]]) { seen[q[i]] = 1; ++alpha; }
    for (int i = 0; i < n; ++i) if (!seen[t[i]]) { seen[t[i]] = 1; ++alpha; }
    r->alphabetLength = alpha;

    if (m == 0 || n == 0) { /* edlib.cpp:161-179; k is not consulted and starts stay unset */
        r->endLocations = (int32_t *)malloc(sizeof(int32_t));
        r->numLocations = 1;
        if (mode == ED_MODE_NW) { r->editDistance = m > n ? m : n; r->endLocations[0] = n - 1; }
        else if (mode == ED_MODE_SHW || mode == ED_MODE_HW) { r->editDistance = m; r->endLocations[0] = -1; }
        else { free(r->endLocations); r->endLocations = NULL; r->numLocations = 0; r->status = 1; }
        return 0;
    }
    if (mode != ED_MODE_NW && mode != ED_MODE_SHW && mode != ED_MODE_HW) { r->status = 1; return 0; }

    eqmat *eq = (eqmat *)malloc(sizeof(eqmat));
    eq_init(eq, eq_pairs, n_eq_pairs);
    const int W = ((m + 63) / 64) * 64 - m; /* padding cells in the last 64-bit block, edlib.cpp:183 */

    int *row = last_row(q, m, t, n, mode == ED_MODE_HW, eq);
    if (mode == ED_MODE_NW) {
        int d = row[n];
        if (k < 0 || d <= k) {
            r->editDistance = d;
            r->endLocations = (int32_t *)malloc(sizeof(int32_t));
            r->endLocations[0] = n - 1;
            r->numLocations = 1;
        }
    } else {
        /* column -1 is only visible through the wildcard padding of the last block
         * (reported at column c-W, edlib.cpp:663-676), i.e. only when W > 0 */
        int first = (W > 0) ? 0 : 1;
        int best = row[first];
        for (int p = first; p <= n; ++p) best = imin(best, row[p]);
        if (k < 0 || best <= k) {
            int cnt = 0;
            for (int p = first; p <= n; ++p) if (row[p] == best) ++cnt;
            r->editDistance = best;
            r->endLocations = (int32_t *)malloc(sizeof(int32_t) * (size_t)cnt);
            r->numLocations = cnt;
            cnt = 0;
            for (int p = first; p <= n; ++p) if (row[p] == best) r->endLocations[cnt++] = p - 1;
        }
    }
    free(row);

    if (r->editDistance >= 0 && (task == ED_TASK_LOC || task == ED_TASK_PATH)) {
        r->startLocations = (int32_t *)malloc(sizeof(int32_t) * (size_t)r->numLocations);
        if (mode == ED_MODE_HW) {
            /* edlib.cpp:225-258: prefix search of the reversed query in the reversed target
             * prefix; the LAST best position gives the smallest start */
            unsigned char *rq = (unsigned char *)malloc((size_t)m);
            for (int i = 0; i < m; ++i) rq[i] = q[m - 1 - i];
            for (int a = 0; a < r->numLocations; ++a) {
                int e = r->endLocations[a];
                if (e == -1) { r->startLocations[a] = 0; continue; }
                unsigned char *rt = (unsigned char *)malloc((size_t)(e + 1));
                for (int i = 0; i <= e; ++i) rt[i] = t[e - i];
                int *rrow = last_row(rq, m, rt, e + 1, 0, eq);
                int first = (W > 0) ? 0 : 1;
                int best = rrow[first], lastp = first;
                for (int p = first; p <= e + 1; ++p) best = imin(best, rrow[p]);
                for (int p = first; p <= e + 1; ++p) if (rrow[p] == best) lastp = p;
                r->startLocations[a] = e - (lastp - 1);
                free(rrow); free(rt);
            }
            free(rq);
        } else {
            for (int a = 0; a < r->numLocations; ++a) r->startLocations[a] = 0;
        }
    }
    free(eq);
    return 0;
}

void ed_oracle_free(ed_oracle_result *r)
{
    free(r->endLocations);
    free(r->startLocations);
    r->endLocations = r->startLocations = NULL;
    r->numLocations = 0;
}

/*
 * Batch helper for tests.  out_fixed holds 4 int32 per pair: status, editDistance,
 * alphabetLength, numLocations; locations are appended as (start, end) int32 pairs to loc_buf
 * (start = INT32_MIN when the reference leaves startLocations NULL), offsets in loc_off[npairs+1].
 */
int ed_oracle_batch(const char *qbuf, const int64_t *qoff, const char *tbuf, const int64_t *toff,
                    int64_t npairs, int32_t mode, int32_t task, const int32_t *k,
                    int32_t *out_fixed, int32_t *loc_buf, int64_t loc_cap, int64_t *loc_off)
{
    int64_t pos = 0;
    for (int64_t p = 0; p < npairs; ++p) {
        ed_oracle_result r;
        ed_oracle_align(qbuf + qoff[p], (int32_t)(qoff[p + 1] - qoff[p]), tbuf + toff[p],
                        (int32_t)(toff[p + 1] - toff[p]), mode, task, k ? k[p] : -1, NULL, 0, &r);
        int32_t *o = out_fixed + p * 4;
        o[0] = r.status; o[1] = r.editDistance; o[2] = r.alphabetLength; o[3] = r.numLocations;
        loc_off[p] = pos;
        if (pos + r.numLocations > loc_cap) { ed_oracle_free(&r); return -1; }
        for (int a = 0; a < r.numLocations; ++a) {
            loc_buf[2 * (pos + a)] = r.startLocations ? r.startLocations[a] : INT32_MIN;
            loc_buf[2 * (pos + a) + 1] = r.endLocations[a];
        }
        pos += r.numLocations;
        ed_oracle_free(&r);
    }
    loc_off[npairs] = pos;
    return 0;
}
