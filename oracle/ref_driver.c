/*
 * TEST INFRASTRUCTURE ONLY - never linked, imported or executed by the product path.
 *
 * Batch driver around the reference's own compiled cores (oracle/_ref/libssw_ref.so and
 * libedlib_ref.so, built unmodified from the reference checkout by oracle/build_ref.sh).
 * It calls them exactly the way the reference wrappers do - ssw_init + ssw_align per pair
 * (_ssw_wrapper.pyx:603-606, 634-638) and edlibAlign per pair (edlib.pyx:130) - over CSR
 * batches, optionally on several host threads (OpenMP), and is used to (1) pin the oracle,
 * (2) generate tests/golden vectors, (3) time the CPU baseline / reference arm in bench.py.
 * The structs below restate the reference's public ABI (ssw.h:38-48, edlib.h:140-218).
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <pthread.h>
#include <unistd.h>

typedef struct {
    uint16_t score1, score2;
    int32_t ref_begin1, ref_end1, read_begin1, read_end1, ref_end2;
    uint32_t *cigar;
    int32_t cigarLen;
} ref_s_align;

typedef struct { char first, second; } ref_eq_pair;
typedef struct { int k; int mode; int task; const ref_eq_pair *additionalEqualities; int additionalEqualitiesLength; } ref_edlib_config;
typedef struct {
    int status, editDistance;
    int *endLocations, *startLocations;
    int numLocations;
    unsigned char *alignment;
    int alignmentLength, alphabetLength;
} ref_edlib_result;

typedef void *(*fn_ssw_init)(const int8_t *, int32_t, const int8_t *, int32_t, int8_t);
typedef void (*fn_init_destroy)(void *);
typedef ref_s_align *(*fn_ssw_align)(const void *, const int8_t *, int32_t, uint8_t, uint8_t, uint8_t, uint16_t, int32_t, int32_t);
typedef void (*fn_align_destroy)(ref_s_align *);
typedef ref_edlib_result (*fn_edlib_align)(const char *, int, const char *, int, ref_edlib_config);
typedef void (*fn_edlib_free)(ref_edlib_result);

static fn_ssw_init p_ssw_init;
static fn_init_destroy p_init_destroy;
static fn_ssw_align p_ssw_align;
static fn_align_destroy p_align_destroy;
static fn_edlib_align p_edlib_align;
static fn_edlib_free p_edlib_free;

int ref_driver_load(const char *ssw_path, const char *edlib_path)
{
    void *h1 = dlopen(ssw_path, RTLD_NOW | RTLD_LOCAL);
    void *h2 = dlopen(edlib_path, RTLD_NOW | RTLD_LOCAL);
    if (!h1 || !h2) { fprintf(stderr, "ref_driver_load: %s\n", dlerror()); return -1; }
    p_ssw_init = (fn_ssw_init)dlsym(h1, "ssw_init");
    p_init_destroy = (fn_init_destroy)dlsym(h1, "init_destroy");
    p_ssw_align = (fn_ssw_align)dlsym(h1, "ssw_align");
    p_align_destroy = (fn_align_destroy)dlsym(h1, "align_destroy");
    p_edlib_align = (fn_edlib_align)dlsym(h2, "edlibAlign");
    p_edlib_free = (fn_edlib_free)dlsym(h2, "edlibFreeAlignResult");
    if (!p_ssw_init || !p_init_destroy || !p_ssw_align || !p_align_destroy || !p_edlib_align || !p_edlib_free) return -2;
    return 0;
}

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}


/* ---- tiny pthread work-sharing helper (chunks of 64 pairs handed out by an atomic counter) ---- */
typedef void (*pair_fn)(int64_t p, void *ctx, void **scratch);
typedef struct { pair_fn fn; void *ctx; int64_t n; int64_t next; } work_t;
static void *work_thread(void *arg)
{
    work_t *w = (work_t *)arg;
    void *scratch[4] = {0, 0, 0, 0};
    for (;;) {
        int64_t b = __atomic_fetch_add(&w->next, 64, __ATOMIC_RELAXED);
        if (b >= w->n) break;
        int64_t e = b + 64 < w->n ? b + 64 : w->n;
        for (int64_t p = b; p < e; ++p) w->fn(p, w->ctx, scratch);
    }
    for (int i = 0; i < 4; ++i) free(scratch[i]);
    return NULL;
}
static int resolve_threads(int nthreads)
{
    if (nthreads > 0) return nthreads;
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}
static void run_pairs(pair_fn fn, void *ctx, int64_t n, int nthreads)
{
    work_t w = {fn, ctx, n, 0};
    nthreads = resolve_threads(nthreads);
    if (nthreads > 256) nthreads = 256;
    if (nthreads <= 1) { work_thread(&w); return; }
    pthread_t th[256];
    for (int i = 0; i < nthreads; ++i) pthread_create(&th[i], NULL, work_thread, &w);
    for (int i = 0; i < nthreads; ++i) pthread_join(th[i], NULL);
}

/* nucleotide table of the wrapper (_ssw_wrapper.pyx:60-68): A,U->0 C->1 G->2 T->3 (either case), else 4 */
static void encode_nt(const char *s, int32_t len, int8_t *out)
{
    static int8_t table[256];
    static int ready = 0;
    if (!ready) {
        memset(table, 4, sizeof(table));
        table['A'] = table['a'] = table['U'] = table['u'] = 0;
        table['C'] = table['c'] = 1;
        table['G'] = table['g'] = 2;
        table['T'] = table['t'] = 3;
        ready = 1;
    }
    for (int32_t i = 0; i < len; ++i) out[i] = table[(unsigned char)s[i]];
}

/*
 * Same output layout as sw_oracle_batch_nt.  nthreads <= 0: all cores.  Returns wall seconds
 * spent inside the alignment loop (encode + ssw_init + ssw_align + destroy, what each wrapper
 * call pays at C level), or a negative number on error.
 */
typedef struct {
    const char *qbuf, *tbuf;
    const int64_t *qoff, *toff;
    int8_t mat[25];
    int gapO, gapE, score_size, flag, filters, filterd;
    const int32_t *maskLen;
    int32_t *out_fixed, *status;
    uint32_t **cig;
    int keep_cigar;
} sw_ctx;

static void sw_one(int64_t p, void *vctx, void **scratch)
{
    sw_ctx *c = (sw_ctx *)vctx;
    int32_t ql = (int32_t)(c->qoff[p + 1] - c->qoff[p]), tl = (int32_t)(c->toff[p + 1] - c->toff[p]);
    /* scratch[0]/[1]: encode buffers, scratch[2]/[3]: their capacities stored as pointers-to-size */
    size_t *caps = (size_t *)scratch[2];
    if (!caps) { caps = (size_t *)calloc(2, sizeof(size_t)); scratch[2] = caps; }
    if ((size_t)ql + 1 > caps[0]) { caps[0] = (size_t)ql * 2 + 64; scratch[0] = realloc(scratch[0], caps[0]); }
    if ((size_t)tl + 1 > caps[1]) { caps[1] = (size_t)tl * 2 + 64; scratch[1] = realloc(scratch[1], caps[1]); }
    int8_t *q = (int8_t *)scratch[0], *t = (int8_t *)scratch[1];
    encode_nt(c->qbuf + c->qoff[p], ql, q);
    encode_nt(c->tbuf + c->toff[p], tl, t);
    void *prof = p_ssw_init(q, ql, c->mat, 5, (int8_t)c->score_size);
    ref_s_align *a = p_ssw_align(prof, t, tl, (uint8_t)c->gapO, (uint8_t)c->gapE, (uint8_t)c->flag,
                                 (uint16_t)c->filters, c->filterd, c->maskLen[p]);
    int32_t *o = c->out_fixed + p * 8;
    if (!a) {
        c->status[p] = -1;
        memset(o, 0, 8 * sizeof(int32_t));
    } else {
        c->status[p] = 0;
        o[0] = a->score1; o[1] = a->score2; o[2] = a->ref_begin1; o[3] = a->ref_end1;
        o[4] = a->read_begin1; o[5] = a->read_end1; o[6] = a->ref_end2; o[7] = a->cigarLen;
        if (a->cigarLen > 0 && c->keep_cigar) {
            c->cig[p] = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)a->cigarLen);
            memcpy(c->cig[p], a->cigar, sizeof(uint32_t) * (size_t)a->cigarLen);
        }
        p_align_destroy(a);
    }
    p_init_destroy(prof);
}

double ref_sw_batch_nt(const char *qbuf, const int64_t *qoff, const char *tbuf, const int64_t *toff,
                       int64_t npairs, int match, int mismatch, int gapO, int gapE, int score_size,
                       int flag, int filters, int filterd, const int32_t *maskLen, int nthreads,
                       int32_t *out_fixed, uint32_t *cigar_buf, int64_t cigar_cap, int64_t *cigar_off,
                       int32_t *status)
{
    sw_ctx c;
    memset(&c, 0, sizeof(c));
    c.qbuf = qbuf; c.tbuf = tbuf; c.qoff = qoff; c.toff = toff;
    for (int a = 0; a < 5; ++a)
        for (int b = 0; b < 5; ++b)
            c.mat[a * 5 + b] = (int8_t)((a == 4 || b == 4) ? 0 : (a == b ? match : mismatch));
    c.gapO = gapO; c.gapE = gapE; c.score_size = score_size; c.flag = flag; c.filters = filters; c.filterd = filterd;
    c.maskLen = maskLen; c.out_fixed = out_fixed; c.status = status;
    c.cig = (uint32_t **)calloc((size_t)(npairs > 0 ? npairs : 1), sizeof(uint32_t *));
    c.keep_cigar = cigar_buf != NULL;
    encode_nt("", 0, NULL); /* build the table before threads start */
    double t0 = now_s();
    run_pairs(sw_one, &c, npairs, nthreads);
    double dt = now_s() - t0;
    int64_t pos = 0;
    int bad = 0;
    for (int64_t p = 0; p < npairs; ++p) {
        int32_t len = out_fixed[p * 8 + 7];
        if (cigar_off) cigar_off[p] = pos;
        if (c.cig[p]) {
            if (pos + len <= cigar_cap) memcpy(cigar_buf + pos, c.cig[p], sizeof(uint32_t) * (size_t)len);
            else bad = 1;
            free(c.cig[p]);
        }
        pos += len;
    }
    if (cigar_off) cigar_off[npairs] = pos;
    free(c.cig);
    return bad ? -1.0 : dt;
}

/* Same output layout as ed_oracle_batch. */
typedef struct {
    const char *qbuf, *tbuf;
    const int64_t *qoff, *toff;
    int32_t mode, task;
    const int32_t *k;
    int32_t *out_fixed;
    int32_t **locs;
    int keep;
    uint8_t *path_buf;    /* task path: operations of pair p at (qoff[p] - qoff[0]) + (toff[p] - toff[0]) */
    int32_t *path_len;
} ed_ctx;

static void ed_one(int64_t p, void *vctx, void **scratch)
{
    (void)scratch;
    ed_ctx *c = (ed_ctx *)vctx;
    ref_edlib_config cfg;
    cfg.k = c->k ? c->k[p] : -1; cfg.mode = c->mode; cfg.task = c->task;
    cfg.additionalEqualities = NULL; cfg.additionalEqualitiesLength = 0;
    ref_edlib_result r = p_edlib_align(c->qbuf + c->qoff[p], (int)(c->qoff[p + 1] - c->qoff[p]),
                                       c->tbuf + c->toff[p], (int)(c->toff[p + 1] - c->toff[p]), cfg);
    int32_t *o = c->out_fixed + p * 4;
    o[0] = r.status; o[1] = r.editDistance; o[2] = r.alphabetLength; o[3] = r.numLocations;
    if (r.numLocations > 0 && c->keep) {
        c->locs[p] = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)r.numLocations);
        for (int a = 0; a < r.numLocations; ++a) {
            c->locs[p][2 * a] = r.startLocations ? r.startLocations[a] : INT32_MIN;
            c->locs[p][2 * a + 1] = r.endLocations ? r.endLocations[a] : INT32_MIN;
        }
    }
    if (c->path_buf) {
        c->path_len[p] = r.alignment ? r.alignmentLength : 0;
        if (r.alignment)
            memcpy(c->path_buf + (c->qoff[p] - c->qoff[0]) + (c->toff[p] - c->toff[0]), r.alignment, (size_t)r.alignmentLength);
    }
    p_edlib_free(r);
}

/* EdlibAlignResult.alignment of every pair (task path) next to the fields ref_ed_batch reports; path_buf holds
 * total query + target bytes, path_len[p] = alignmentLength (0: no alignment). */
double ref_ed_path_batch(const char *qbuf, const int64_t *qoff, const char *tbuf, const int64_t *toff,
                         int64_t npairs, int32_t mode, int32_t task, const int32_t *k, int nthreads,
                         int32_t *out_fixed, int32_t *loc_buf, int64_t loc_cap, int64_t *loc_off,
                         uint8_t *path_buf, int32_t *path_len);

double ref_ed_batch(const char *qbuf, const int64_t *qoff, const char *tbuf, const int64_t *toff,
                    int64_t npairs, int32_t mode, int32_t task, const int32_t *k, int nthreads,
                    int32_t *out_fixed, int32_t *loc_buf, int64_t loc_cap, int64_t *loc_off)
{
    return ref_ed_path_batch(qbuf, qoff, tbuf, toff, npairs, mode, task, k, nthreads, out_fixed, loc_buf, loc_cap, loc_off, NULL, NULL);
}

double ref_ed_path_batch(const char *qbuf, const int64_t *qoff, const char *tbuf, const int64_t *toff,
                         int64_t npairs, int32_t mode, int32_t task, const int32_t *k, int nthreads,
                         int32_t *out_fixed, int32_t *loc_buf, int64_t loc_cap, int64_t *loc_off,
                         uint8_t *path_buf, int32_t *path_len)
{
    ed_ctx c;
    memset(&c, 0, sizeof(c));
    c.path_buf = path_len ? path_buf : NULL; c.path_len = path_len;
    c.qbuf = qbuf; c.tbuf = tbuf; c.qoff = qoff; c.toff = toff; c.mode = mode; c.task = task; c.k = k;
    c.out_fixed = out_fixed;
    c.locs = (int32_t **)calloc((size_t)(npairs > 0 ? npairs : 1), sizeof(int32_t *));
    c.keep = loc_buf != NULL;
    double t0 = now_s();
    run_pairs(ed_one, &c, npairs, nthreads);
    double dt = now_s() - t0;
    int64_t pos = 0;
    int bad = 0;
    for (int64_t p = 0; p < npairs; ++p) {
        int32_t len = out_fixed[p * 4 + 3];
        if (loc_off) loc_off[p] = pos;
        if (c.locs[p]) {
            if (pos + len <= loc_cap) memcpy(loc_buf + 2 * pos, c.locs[p], sizeof(int32_t) * 2 * (size_t)len);
            else bad = 1;
            free(c.locs[p]);
        }
        pos += len;
    }
    if (loc_off) loc_off[npairs] = pos;
    free(c.locs);
    return bad ? -1.0 : dt;
}

int ref_driver_max_threads(void) { return resolve_threads(0); }
