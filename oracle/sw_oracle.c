/*
 * TEST INFRASTRUCTURE ONLY - never linked, imported or executed by the product path.
 *
 * Scalar CPU restatement of the *observable* semantics of dysgu's vendored striped
 * Smith-Waterman (reference: dysgu/scikitbio/ssw.c, SSW library 0.1.4), written from the
 * behaviour of that code, not from its text.  No SIMD, no striping: a textbook Gotoh
 * recurrence plus the tie-breaks, padding rows and band rules the reference exposes.
 *
 * Parity status: PINNED.  oracle/validate_oracle.py compares every field produced here with
 * oracle/_ref/libssw_ref.so (the reference's own ssw.c compiled unmodified by
 * oracle/build_ref.sh) on >1e6 randomized + adversarial pairs, and tests/golden/ holds
 * vectors generated from that compiled reference.
 *
 * Naming follows the reference: "read" = profiled sequence (wrapper's query, DP rows i),
 * "ref" = the sequence given at call time (wrapper's target, DP columns j).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "sw_oracle.h"

static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }

typedef struct {
    int score;   /* best H */
    int ref;     /* column of the best cell */
    int read;    /* row of the best cell */
    int score2;  /* second-best column maximum outside the mask window */
    int ref2;
    int overflow;
} pass_result;

/*
 * One scoring pass.  Follows sw_sse2_byte (ssw.c:124-346) when lanes==16 and sw_sse2_word
 * (ssw.c:372-548) when lanes==8:
 *  - rows are padded to a multiple of `lanes` with cells that score 0 against everything
 *    (the striped profile's padding, ssw.c:109 / ssw.c:364); the padded rows take part in
 *    the per-column maxima (ssw.c:218,295 / 450,500)
 *  - columns are visited upward (ref_dir==0) or downward (ref_dir==1, ssw.c:182-186)
 *  - the best column is the one where the running maximum last rose strictly
 *    (ssw.c:284-291 / 492-496); the best row is the smallest row holding that value in the
 *    saved column (ssw.c:301-309 / 505-513)
 *  - byte mode stops and reports 255 once max+bias >= 255 (ssw.c:286,318)
 *  - the scan stops after a column whose maximum equals `terminate` (ssw.c:297 / 501)
 *  - second best: ssw.c:326-341 (byte; right scan starts at edge+1) and ssw.c:530-543
 *    (word; right scan starts at edge)
 */
static void sw_pass(const int8_t *read, int readLen, const int8_t *ref, int ref_dir, int refLen,
                    const int8_t *mat, int n, int gapO, int gapE, int terminate, int lanes, int bias,
                    int maskLen, pass_result *out)
{
    const int byte_mode = (lanes == 16);
    const int rows = ((readLen + lanes - 1) / lanes) * lanes;
    int *H = (int *)calloc((size_t)rows + 1, sizeof(int));     /* previous column */
    int *E = (int *)calloc((size_t)rows + 1, sizeof(int));     /* horizontal gap state entering this column */
    int *Hbest = (int *)calloc((size_t)rows + 1, sizeof(int)); /* copy of the column that set the maximum */
    int *colmax = (int *)calloc((size_t)(refLen > 0 ? refLen : 1), sizeof(int));
    int best = 0;
    int end_ref = byte_mode ? -1 : 0; /* ssw.c:146 vs ssw.c:389 */
    int end_read = readLen - 1;
    int overflow = 0;

    int begin = 0, end = refLen, step = 1;
    if (ref_dir == 1) { begin = refLen - 1; end = -1; step = -1; }

    for (int j = begin; j != end; j += step) {
        const int8_t *mrow = mat + (int)ref[j] * n;
        int diag = 0, f = 0, cmax = 0;
        for (int i = 0; i < rows; ++i) {
            int s = (i < readLen) ? mrow[read[i]] : 0;
            int e = E[i];
            int h = diag + s;
            h = imax(h, e);
            h = imax(h, f);
            h = imax(h, 0);
            diag = H[i];
            H[i] = h;
            cmax = imax(cmax, h);
            E[i] = imax(imax(e - gapE, h - gapO), 0);
            f = imax(imax(f - gapE, h - gapO), 0);
        }
        if (cmax > best) {
            best = cmax;
            if (byte_mode && best + bias >= 255) { overflow = 1; break; }
            end_ref = j;
            memcpy(Hbest, H, (size_t)rows * sizeof(int));
        }
        colmax[j] = cmax;
        if (cmax == terminate) break;
    }

    for (int i = 0; i < rows; ++i)
        if (Hbest[i] == best && i < end_read) end_read = i;

    out->overflow = overflow;
    out->score = overflow ? 255 : best;
    out->ref = end_ref;
    out->read = end_read;
    out->score2 = 0;
    out->ref2 = 0;

    int edge = (end_ref - maskLen) > 0 ? (end_ref - maskLen) : 0;
    for (int j = 0; j < edge; ++j)
        if (colmax[j] > out->score2) { out->score2 = colmax[j]; out->ref2 = j; }
    edge = (end_ref + maskLen) > refLen ? refLen : (end_ref + maskLen);
    for (int j = edge + (byte_mode ? 1 : 0); j < refLen; ++j)
        if (colmax[j] > out->score2) { out->score2 = colmax[j]; out->ref2 = j; }

    free(H); free(E); free(Hbest); free(colmax);
}

/*
 * Banded traceback (behaviour of banded_sw, ssw.c:550-728) written as a pure function of
 * (i, j): row i = read, column j = ref, band row i = columns [max(i-bw,0), min(refLen-1,i+bw)].
 * Any neighbour outside the band (or outside the rectangle) reads as 0.  Vertical gap state is
 * called E and horizontal F as in the reference function; neither is clamped at 0.
 * The band doubles until the best cell seen so far reaches `score` (the running best is kept
 * across iterations, ssw.c:562,631).  Traceback walks from the bottom-right corner while i > 0
 * (ssw.c:641) and then adds one closing M (ssw.c:690-707).
 * Direction codes: 1 diagonal; 2/3 vertical extend/open; 4/5 horizontal extend/open.
 */
static int banded_traceback(const int8_t *ref, const int8_t *read, int refLen, int readLen, int score,
                            int gapO, int gapE, int bw, const int8_t *mat, int n,
                            uint32_t **cigar_out, int32_t *cigar_len_out)
{
    int best = 0;
    uint8_t *dir = NULL;
    int width_d = 0;
    int *Hp = (int *)malloc(sizeof(int) * (size_t)(refLen + 2));
    int *Ep = (int *)malloc(sizeof(int) * (size_t)(refLen + 2));
    int *Hc = (int *)malloc(sizeof(int) * (size_t)(refLen + 2));
    int *Ec = (int *)malloc(sizeof(int) * (size_t)(refLen + 2));
    do {
        width_d = bw * 2 + 1;
        size_t need = (size_t)width_d * (size_t)readLen * 3;
        free(dir);
        dir = (uint8_t *)calloc(need ? need : 1, 1);
        if (!dir) { free(Hp); free(Ep); free(Hc); free(Ec); return -1; }
        int pbeg = 0, pend = -1; /* previous row's band */
        for (int i = 0; i < readLen; ++i) {
            int beg = imax(0, i - bw), end = imin(refLen - 1, i + bw);
            int f = 0, hleft = 0;
            uint8_t *dl = dir + (size_t)width_d * (size_t)i * 3;
            for (int j = beg; j <= end; ++j) {
                int up_in = (i > 0 && j >= pbeg && j <= pend);
                int dg_in = (i > 0 && j - 1 >= pbeg && j - 1 <= pend);
                int hup = up_in ? Hp[j] : 0, eup = up_in ? Ep[j] : 0;
                int hdg = dg_in ? Hp[j - 1] : 0;
                int t1 = (i == 0) ? -gapO : hup - gapO;
                int t2 = (i == 0) ? -gapE : eup - gapE;
                int e = t1 > t2 ? t1 : t2;
                uint8_t de = t1 > t2 ? 3 : 2;
                t1 = hleft - gapO;
                t2 = f - gapE;
                f = t1 > t2 ? t1 : t2;
                uint8_t df = t1 > t2 ? 5 : 4;
                int e1 = e > 0 ? e : 0, f1 = f > 0 ? f : 0;
                int g = e1 > f1 ? e1 : f1;
                int d = hdg + mat[(int)ref[j] * n + read[i]];
                int h = g > d ? g : d;
                if (h > best) best = h;
                uint8_t dh = (g <= d) ? 1 : (e1 > f1 ? de : df);
                int x = j - beg;
                dl[x * 3 + 0] = de; dl[x * 3 + 1] = df; dl[x * 3 + 2] = dh;
                Hc[j] = h; Ec[j] = e;
                hleft = h;
            }
            int *t = Hp; Hp = Hc; Hc = t;
            t = Ep; Ep = Ec; Ec = t;
            pbeg = beg; pend = end;
        }
        bw *= 2;
    } while (best < score);
    bw /= 2;
    free(Hp); free(Ep); free(Hc); free(Ec);

    /* traceback */
    int cap = 16, l = 0;
    uint32_t *c = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)cap);
    int i = readLen - 1, j = refLen - 1;
    int run = 0, prev = 0, op = 0, state = 2;
    int status = 0;
    while (i > 0) {
        int beg = imax(0, i - bw);
        int x = j - beg;
        if (x < 0 || x >= width_d) { status = -2; break; } /* reference would read out of bounds */
        uint8_t dcode = dir[((size_t)width_d * (size_t)i + (size_t)x) * 3 + (size_t)state];
        switch (dcode) {
            case 1: --i; --j; state = 2; op = 0; break;
            case 2: --i; state = 0; op = 1; break;
            case 3: --i; state = 2; op = 1; break;
            case 4: --j; state = 1; op = 2; break;
            case 5: --j; state = 2; op = 2; break;
            default: status = -3; break;
        }
        if (status) break;
        if (op == prev) ++run;
        else {
            if (l + 3 >= cap) { cap *= 2; c = (uint32_t *)realloc(c, sizeof(uint32_t) * (size_t)cap); }
            c[l++] = ((uint32_t)run << 4) | (uint32_t)prev;
            prev = op; run = 1;
        }
    }
    free(dir);
    if (status) { free(c); return status; }
    if (l + 3 >= cap) { cap += 4; c = (uint32_t *)realloc(c, sizeof(uint32_t) * (size_t)cap); }
    if (op == 0) c[l++] = ((uint32_t)(run + 1)) << 4;
    else { c[l++] = ((uint32_t)run << 4) | (uint32_t)op; c[l++] = 16; }
    for (int a = 0, b = l - 1; a < b; ++a, --b) { uint32_t t = c[a]; c[a] = c[b]; c[b] = t; }
    *cigar_out = c;
    *cigar_len_out = l;
    return 0;
}

/* Driver: behaviour of ssw_init + ssw_align (ssw.c:743-764, 772-857). */
int sw_oracle_align(const int8_t *read, int32_t readLen, const int8_t *ref, int32_t refLen,
                    const int8_t *mat, int32_t n, int32_t score_size, int32_t gapO, int32_t gapE,
                    int32_t flag, int32_t filters, int32_t filterd, int32_t maskLen, sw_oracle_result *r)
{
    memset(r, 0, sizeof(*r));
    r->ref_begin1 = -1;
    r->read_begin1 = -1;
    int bias = 0;
    for (int i = 0; i < n * n; ++i) if (mat[i] < bias) bias = mat[i];
    bias = -bias;

    pass_result fw;
    int word = 0;
    if (score_size == 0 || score_size == 2) {
        sw_pass(read, readLen, ref, 0, refLen, mat, n, gapO, gapE, -1, 16, bias, maskLen, &fw);
        if (fw.score == 255) {
            if (score_size == 2) {
                sw_pass(read, readLen, ref, 0, refLen, mat, n, gapO, gapE, -1, 8, 0, maskLen, &fw);
                word = 1;
            } else {
                return SW_ORACLE_ERR_OVERFLOW; /* ssw.c:803-806 returns NULL */
            }
        }
    } else if (score_size == 1) {
        sw_pass(read, readLen, ref, 0, refLen, mat, n, gapO, gapE, -1, 8, 0, maskLen, &fw);
        word = 1;
    } else {
        return SW_ORACLE_ERR_ARG;
    }
    r->used_word = word;
    r->score1 = (uint16_t)fw.score;
    r->ref_end1 = fw.ref;
    r->read_end1 = fw.read;
    if (maskLen >= 15) { r->score2 = (uint16_t)fw.score2; r->ref_end2 = fw.ref2; }
    else { r->score2 = 0; r->ref_end2 = -1; }
    if (flag == 0 || (flag == 2 && r->score1 < filters)) return 0;

    /* begin coordinates: reverse pass over read[0..read_end1] reversed, columns ref_end1..0 */
    int rl = r->read_end1 + 1;
    int8_t *rr = (int8_t *)malloc((size_t)(rl > 0 ? rl : 1));
    for (int i = 0; i < rl; ++i) rr[i] = read[r->read_end1 - i];
    pass_result rv;
    sw_pass(rr, rl, ref, 1, r->ref_end1 + 1, mat, n, gapO, gapE, r->score1, word ? 8 : 16, bias, maskLen, &rv);
    free(rr);
    r->ref_begin1 = rv.ref;
    r->read_begin1 = r->read_end1 - rv.read;
    if ((7 & flag) == 0 || ((2 & flag) != 0 && r->score1 < filters) ||
        ((4 & flag) != 0 && (r->ref_end1 - r->ref_begin1 > filterd || r->read_end1 - r->read_begin1 > filterd)))
        return 0;

    int cr = r->ref_end1 - r->ref_begin1 + 1;
    int cq = r->read_end1 - r->read_begin1 + 1;
    int bw = abs(cr - cq) + 1;
    if (r->ref_begin1 < 0) {
        /* score 0 on the byte path: the reference reads ref[-1] here (ssw.c:844-847); the only
         * observable outcome for a 1x1 rectangle with score 0 is the cigar 1M. */
        r->cigar = (uint32_t *)malloc(sizeof(uint32_t));
        r->cigar[0] = 16;
        r->cigar_len = 1;
        return 0;
    }
    int st = banded_traceback(ref + r->ref_begin1, read + r->read_begin1, cr, cq, r->score1, gapO, gapE, bw,
                              mat, n, &r->cigar, &r->cigar_len);
    if (st) return SW_ORACLE_ERR_TRACEBACK;
    return 0;
}

void sw_oracle_free(sw_oracle_result *r)
{
    free(r->cigar);
    r->cigar = NULL;
    r->cigar_len = 0;
}

/* nucleotide encoding of the wrapper (reference: _ssw_wrapper.pyx:60-68 np_nt_table):
 * A/a,U/u -> 0, C/c -> 1, G/g -> 2, T/t -> 3, everything else -> 4. */
void sw_oracle_encode_nt(const char *s, int32_t len, int8_t *out)
{
    for (int i = 0; i < len; ++i) {
        int8_t c = 4;
        switch (s[i]) {
            case 'A': case 'a': case 'U': case 'u': c = 0; break;
            case 'C': case 'c': c = 1; break;
            case 'G': case 'g': c = 2; break;
            case 'T': case 't': c = 3; break;
            default: c = 4;
        }
        out[i] = c;
    }
}

/* 5x5 matrix of the wrapper (reference: _ssw_wrapper.pyx:681-693): N scores 0 with anything. */
void sw_oracle_nt_matrix(int match, int mismatch, int8_t *mat25)
{
    for (int a = 0; a < 5; ++a)
        for (int b = 0; b < 5; ++b)
            mat25[a * 5 + b] = (int8_t)((a == 4 || b == 4) ? 0 : (a == b ? match : mismatch));
}

/*
 * Batch helper for tests: CSR sequences (ASCII), nucleotide defaults.  out_fixed holds 8 int32
 * per pair: score1, score2, ref_begin1, ref_end1, read_begin1, read_end1, ref_end2, cigar_len;
 * cigars are appended to cigar_buf (capacity cigar_cap uint32) with offsets in cigar_off[npairs+1].
 * Returns 0, or -1 if cigar_buf is too small.
 */
int sw_oracle_batch_nt(const char *qbuf, const int64_t *qoff, const char *tbuf, const int64_t *toff,
                       int64_t npairs, int match, int mismatch, int gapO, int gapE, int score_size,
                       int flag, int filters, int filterd, const int32_t *maskLen,
                       int32_t *out_fixed, uint32_t *cigar_buf, int64_t cigar_cap, int64_t *cigar_off,
                       int32_t *status)
{
    int8_t mat[25];
    sw_oracle_nt_matrix(match, mismatch, mat);
    int64_t cpos = 0;
    for (int64_t p = 0; p < npairs; ++p) {
        int32_t ql = (int32_t)(qoff[p + 1] - qoff[p]), tl = (int32_t)(toff[p + 1] - toff[p]);
        int8_t *q = (int8_t *)malloc((size_t)(ql > 0 ? ql : 1));
        int8_t *t = (int8_t *)malloc((size_t)(tl > 0 ? tl : 1));
        sw_oracle_encode_nt(qbuf + qoff[p], ql, q);
        sw_oracle_encode_nt(tbuf + toff[p], tl, t);
        sw_oracle_result r;
        int st = sw_oracle_align(q, ql, t, tl, mat, 5, score_size, gapO, gapE, flag, filters, filterd, maskLen[p], &r);
        status[p] = st;
        int32_t *o = out_fixed + p * 8;
        o[0] = r.score1; o[1] = r.score2; o[2] = r.ref_begin1; o[3] = r.ref_end1;
        o[4] = r.read_begin1; o[5] = r.read_end1; o[6] = r.ref_end2; o[7] = r.cigar_len;
        cigar_off[p] = cpos;
        if (cpos + r.cigar_len > cigar_cap) { free(q); free(t); sw_oracle_free(&r); return -1; }
        if (r.cigar_len) memcpy(cigar_buf + cpos, r.cigar, sizeof(uint32_t) * (size_t)r.cigar_len);
        cpos += r.cigar_len;
        sw_oracle_free(&r);
        free(q); free(t);
    }
    cigar_off[npairs] = cpos;
    return 0;
}
