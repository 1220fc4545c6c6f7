/*
 * dysgu_b200_align.h - C ABI of the B200-native batched alignment engine (libdysgu_b200_align.so).
 *
 * Drop-in boundary for the one data-parallel hot path of kcleal/dysgu: the many independent
 * alignments issued through StripedSmithWaterman(query)(target) and edlib.align(...).
 * Plain pointers and sizes only; no torch / Python types.  Every entry point cites the reference
 * interface it replaces (paths relative to the reference checkout, dysgu v1.9.0).
 *
 * Conventions
 *   - Sequences travel as CSR batches: a byte buffer plus int64 offsets[npairs+1]; sequence i is
 *     buf[off[i] .. off[i+1]).  SW sequences are raw ASCII (encoded on the device exactly like
 *     _ssw_wrapper.pyx:60-68: A/a/U/u=0 C/c=1 G/g=2 T/t=3, anything else=4); edit-distance
 *     sequences are raw bytes compared case-sensitively (edlib.pyx:12-54, edlib.cpp:1433-1472).
 *   - *_host entry points take host pointers and perform H2D / compute / D2H internally.
 *   - *_device entry points take device pointers on `device` and enqueue everything on `stream`
 *     (a cudaStream_t passed as void*; NULL = the legacy default stream); outputs are device
 *     pointers too.  The SW device call returns after enqueueing unless cigar_total is requested; the edit-distance device
 *     calls synchronise the stream once per call (the size of the match-vector pool is read back before the scoring
 *     kernels are enqueued) and again when loc_total / path_total are requested.
 *   - Threads: the batch entry points may be called from several host threads (the reference's cores are
 *     reentrant and edlib.pyx:129 releases the GIL); calls on the same device run one after the other
 *     (a per-device lock held for the whole *_host call and for the enqueue of a *_device call), calls
 *     on different devices run concurrently.  Results of a *_device call must have been consumed
 *     (or the stream synchronised) before the next call on that device - the workspace is reused.
 *     A ticket queue belongs to one thread.
 *   - fork(): a child of a process that has not yet called a compute / init / device-count entry point creates its own
 *     engine lazily.  A child forked AFTER the parent initialised CUDA cannot use the GPU (a CUDA restriction): every
 *     compute entry point then fails with DB200_ERR_FORKED and a message saying so (db200_forked_after_init() == 1).
 *     That is the order `dysgu call -p N` produces (re_map in the parent, cluster.pyx:475-478, then the Pool of
 *     merge_svs.pyx:396); dysgu_b200.patch keeps that pool's work in the parent as device batches and
 *     dysgu_b200/server.py serves forked children from the parent's engine.
 *   - All functions return DB200_OK (0) or a negative DB200_ERR_* code; db200_last_error() gives
 *     a thread-local message.  There is no CPU fallback: without a CUDA device every compute
 *     entry point fails with DB200_ERR_CUDA.
 */
#ifndef DYSGU_B200_ALIGN_H
#define DYSGU_B200_ALIGN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DB200_OK 0
#define DB200_ERR_CUDA -1      /* no device / CUDA runtime error */
#define DB200_ERR_ARG -2       /* invalid argument */
#define DB200_ERR_NOMEM -3     /* device or host allocation failed */
#define DB200_ERR_CAPACITY -4  /* caller-provided output buffer too small */
#define DB200_ERR_INTERNAL -5
#define DB200_ERR_FORKED -6    /* the process was fork()ed after its parent initialised CUDA: no GPU use is possible here */

/* per-pair status values written to the status arrays */
#define DB200_PAIR_OK 0
#define DB200_PAIR_EMPTY 1          /* an empty sequence: the reference reads out of bounds here (ssw.c:195) */
#define DB200_PAIR_OVERFLOW 2       /* score_size==0 and the 8-bit score saturated: ssw_align returns NULL (ssw.c:803-806) */
#define DB200_PAIR_SCORE_RANGE 3    /* best possible score does not fit int16 */
#define DB200_PAIR_UNSUPPORTED 4
#define DB200_PAIR_NOSCRATCH 5       /* device scratch / output pool exhausted for this pair: retry in a smaller batch */

/* ------------------------------------------------------------------------------------------ */
/* engine                                                                                       */
/* ------------------------------------------------------------------------------------------ */
/* Every `device` argument of this header is a CUDA device ordinal, optionally combined with an engine instance:
 * DB200_INSTANCE(dev, k), k = 0..3, names the k-th independent engine on that GPU (its own streams, workspace and call
 * lock).  Calls on different instances run concurrently, so two host threads that alternate over the batches of a long
 * job keep the GPU busy while the next batch is still crossing PCIe (what bench.py's e2e arm does). */
#define DB200_INSTANCE(dev, k) ((dev) | ((k) << 8))
int db200_device_count(void);
/* Creates (lazily, per process) the engine state for `device`; see the fork() note above. */
int db200_init(int device);
/* 1 if this process was fork()ed after its parent initialised CUDA through this library (compute calls fail with DB200_ERR_FORKED). */
int db200_forked_after_init(void);
/* Chunking of this engine's Smith-Waterman *_host entry points: 0 = automatic (a small first chunk, growing: the copies of
 * later chunks hide behind earlier compute inside one call); n > 0 = chunks of n pairs.  A caller that keeps several calls in
 * flight on different engine instances (DB200_INSTANCE) sets n to its batch size, so that the overlap happens between calls
 * and every call runs its kernels once over the whole batch. */
int db200_set_host_chunk_pairs(int device, int64_t pairs);
/* Device used by the legacy single-pair entry points (their reference signatures carry no device): the environment
 * variable DB200_DEVICE, else LOCAL_RANK, else 0. */
int db200_default_device(void);
void db200_shutdown(void);
const char *db200_last_error(void);
const char *db200_version(void);
/* Bytes currently held by the grow-only device workspace of `device`. */
int64_t db200_workspace_bytes(int device);
/* Kernel launches issued by this library on `device` since the last call with reset != 0. */
int64_t db200_kernel_launches(int device, int reset);
/* Stage profiling: CUDA events between pipeline stages on the launching stream.  `collect` synchronises
 * the device and returns the accumulated milliseconds per stage (see enum Stage in csrc/engine.cuh:
 * 1 pack, 2 bucket, 3 forward pass, 4 forward re-run, 5 reverse prep, 6 reverse pass, 7 reverse re-run,
 * 8 finalize, 9 traceback, 10 cigar, 11 ed prep, 12 ed forward, 13 ed starts, 14 ed output, 15 ed path). */
/* Myers word steps (one 64-row block update) the edit-distance kernels evaluated since the last reset; synchronises. */
int64_t db200_ed_word_steps(int device, int reset);
int db200_profile_enable(int device, int on);
int db200_profile_collect(int device, float *ms, int nstages);
/* Page-locked host memory for the *_host entry points (plain malloc'd buffers work too, slower). */
void *db200_host_alloc(int64_t bytes);
void db200_host_free(void *p);

/* ------------------------------------------------------------------------------------------ */
/* Smith-Waterman (replaces ssw_init + ssw_align + align_destroy: ssw.h:72-125, ssw.c:743-862)  */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int32_t match;       /* used when mat == NULL: 5x5 ACGTN matrix, N scores 0 (_ssw_wrapper.pyx:681-693) */
    int32_t mismatch;
    int32_t gap_open;    /* weight_gapO of ssw_align (positive) */
    int32_t gap_extend;  /* weight_gapE of ssw_align (positive) */
    int32_t score_size;  /* ssw_init score_size: 0 = 8-bit only, 1 = 16-bit only, 2 = both */
    int32_t flag;        /* ssw_align flag (ssw.h:91-100); the wrapper passes 1 for every dysgu call site */
    int32_t filters;     /* ssw_align filters */
    int32_t filterd;     /* ssw_align filterd */
    const int8_t *mat;   /* optional host pointer to a 5x5 substitution matrix in ACGTN order, or NULL */
    int32_t n;           /* edge length of mat (must be 5 when mat != NULL) */
} db200_sw_params;

/*
 * Fixed-size result record, one per pair; mirrors s_align (ssw.h:38-48).
 * Wrapper view (_ssw_wrapper.pyx:130-175): score1 = optimal_alignment_score, score2 =
 * suboptimal_alignment_score, read_* = query_*, ref_* = target_*.
 */
typedef struct {
    int32_t score1;
    int32_t score2;
    int32_t ref_begin1;
    int32_t ref_end1;
    int32_t read_begin1;
    int32_t read_end1;
    int32_t ref_end2;
    int32_t cigar_len;
} db200_sw_result;

/*
 * Align npairs (query_i, target_i).  query = the profiled sequence ("read"), target = the sequence
 * passed at call time ("ref").  mask_len[i] is the maskLen argument of ssw_align (NULL: the
 * wrapper's mask_auto rule max(len(query)/2, 15), _ssw_wrapper.pyx:583-588).
 * Cigars (BAM encoding len<<4|op, op M/I/D = 0/1/2) are written back to back into cigar_buf in pair
 * order; cigar_off[i]..cigar_off[i+1] delimits pair i.  cigar_buf may be NULL when no cigar is
 * wanted (then cigar_len is still reported and cigar_off filled).
 */
int db200_sw_batch_host(int device, const db200_sw_params *params,
                        const uint8_t *qbuf, const int64_t *qoff,
                        const uint8_t *tbuf, const int64_t *toff, int64_t npairs,
                        const int32_t *mask_len,
                        db200_sw_result *results, uint32_t *cigar_buf, int64_t cigar_cap,
                        int64_t *cigar_off, int32_t *status);

/*
 * Same with every buffer resident in device memory.  q_bytes / t_bytes are the total buffer
 * lengths (qoff[npairs], toff[npairs]) and max_query_len / max_target_len upper bounds of the sequence lengths -
 * known to the caller, needed to size the workspace without a device round trip (0 disables
 * the column-tiled kernel, i.e. targets longer than 1024 are then reported as unsupported).
 * Everything is enqueued on `stream`; the call only synchronises when cigar_total is non-NULL
 * (it then receives the number of cigar ops, and a cigar_cap that was too small is reported as
 * DB200_ERR_CAPACITY; without cigar_total an overflowing pair keeps its cigar_len but its ops are not written).
 */
int db200_sw_batch_device(int device, const db200_sw_params *params,
                          const uint8_t *d_qbuf, const int64_t *d_qoff, int64_t q_bytes,
                          const uint8_t *d_tbuf, const int64_t *d_toff, int64_t t_bytes, int64_t npairs,
                          int32_t max_query_len, int32_t max_target_len, const int32_t *d_mask_len,
                          db200_sw_result *d_results, uint32_t *d_cigar_buf, int64_t cigar_cap,
                          int64_t *d_cigar_off, int32_t *d_status, int64_t *cigar_total, void *stream);

/*
 * Other host input forms of the same call (results, cigars and status exactly as db200_sw_batch_host):
 *
 * db200_sw_batch_host_slices - sequences are (start, length) views of shared source buffers that travel to the device once
 *   per call: the shape dysgu's remapping has, one reference region fetched per group of events and many windows cut out
 *   of it, overlapping freely (re_map.py:429-434, 192-199).  A side whose length array is NULL is plain CSR (its start
 *   array is then offsets[npairs + 1]).
 *
 * db200_sw_batch_host_packed - both sides pre-encoded in the engine's pool format: 4-bit codes (A/a/U/u 0, C/c 1, G/g 2,
 *   T/t 3, other 4; padding 5 - the table of _ssw_wrapper.pyx:60-68), eight bases per 32-bit word, every sequence padded
 *   to a 16-byte unit of 32 bases.  The reference's C boundary takes encoded sequences too (int8 codes, ssw.h:72-125; the
 *   wrapper encodes, _ssw_wrapper.pyx:668-679); this encoding is half a byte per base, so a batch crosses PCIe at half the
 *   bytes of ASCII and lands in the device pool without a pack kernel.  unit_off[npairs + 1] count 16-byte units and must
 *   be dense (unit_off[i + 1] - unit_off[i] == ceil(len[i] / 32)); db200_sw_pack_host writes this format from ASCII
 *   (CSR when len == NULL, else slices; `units` == NULL only sizes: unit_off and len_out are filled).
 */
int db200_sw_batch_host_slices(int device, const db200_sw_params *params,
                               const uint8_t *qsrc, int64_t qsrc_bytes, const int64_t *qstart, const int32_t *qlen,
                               const uint8_t *tsrc, int64_t tsrc_bytes, const int64_t *tstart, const int32_t *tlen,
                               int64_t npairs, const int32_t *mask_len,
                               db200_sw_result *results, uint32_t *cigar_buf, int64_t cigar_cap,
                               int64_t *cigar_off, int32_t *status);
int db200_sw_batch_host_packed(int device, const db200_sw_params *params,
                               const uint32_t *qunits, const int64_t *qunit_off, const int32_t *qlen,
                               const uint32_t *tunits, const int64_t *tunit_off, const int32_t *tlen,
                               int64_t npairs, const int32_t *mask_len,
                               db200_sw_result *results, uint32_t *cigar_buf, int64_t cigar_cap,
                               int64_t *cigar_off, int32_t *status);
/* The pre-packed form with everything resident in device memory: d_units = [all queries][all targets], dense; q_bases /
 * t_bases the sums of the lengths; otherwise as db200_sw_batch_device.  A 16-byte aligned d_units is read in place by every
 * kernel of the call (scoring, reverse preparation, traceback) and never written: it must stay unchanged until the work
 * enqueued on `stream` has finished, like the other device inputs.  An unaligned one is copied into the workspace first. */
int db200_sw_batch_device_packed(int device, const db200_sw_params *params, const uint32_t *d_units, int64_t n_units,
                                 const int32_t *d_qlen, const int32_t *d_tlen, int64_t q_bases, int64_t t_bases, int64_t npairs,
                                 int32_t max_query_len, int32_t max_target_len, const int32_t *d_mask_len,
                                 db200_sw_result *d_results, uint32_t *d_cigar_buf, int64_t cigar_cap,
                                 int64_t *d_cigar_off, int32_t *d_status, int64_t *cigar_total, void *stream);
int db200_sw_pack_host(const uint8_t *buf, const int64_t *off, const int32_t *len, int64_t nseq,
                       uint32_t *units, int64_t units_cap, int64_t *unit_off, int32_t *len_out, int32_t threads);

/*
 * Legacy single-pair entry points with the reference's signatures (ssw.h:72-125), implemented as
 * "batch of one" so that an unmodified Cython wrapper can link against this library.
 * `read` / `ref` are already-encoded int8 codes (0..4) as produced by the wrapper.
 */
typedef struct {
    uint16_t score1;
    uint16_t score2;
    int32_t ref_begin1;
    int32_t ref_end1;
    int32_t read_begin1;
    int32_t read_end1;
    int32_t ref_end2;
    uint32_t *cigar;
    int32_t cigarLen;
} db200_s_align; /* layout-identical to s_align */
typedef struct db200_s_profile db200_s_profile;

db200_s_profile *db200_ssw_init(const int8_t *read, int32_t readLen, const int8_t *mat, int32_t n, int8_t score_size);
void db200_init_destroy(db200_s_profile *p);
db200_s_align *db200_ssw_align(const db200_s_profile *prof, const int8_t *ref, int32_t refLen,
                               uint8_t weight_gapO, uint8_t weight_gapE, uint8_t flag,
                               uint16_t filters, int32_t filterd, int32_t maskLen);
void db200_align_destroy(db200_s_align *a);

/* ------------------------------------------------------------------------------------------ */
/* Edit distance (replaces edlibAlign + edlibFreeAlignResult: edlib.h:146-271, edlib.cpp:141-296)*/
/* ------------------------------------------------------------------------------------------ */
#define DB200_ED_MODE_NW 0   /* EDLIB_MODE_NW  */
#define DB200_ED_MODE_SHW 1  /* EDLIB_MODE_SHW */
#define DB200_ED_MODE_HW 2   /* EDLIB_MODE_HW  */
#define DB200_ED_TASK_DISTANCE 0
#define DB200_ED_TASK_LOC 1
#define DB200_ED_TASK_PATH 2 /* EDLIB_TASK_PATH: through db200_ed_path_batch_* (no dysgu caller requests it) */
#define DB200_ED_NO_START INT32_MIN /* start location left unset (edlib.pyx:140: None) */

typedef struct {
    int32_t mode;
    int32_t task;
    int32_t k;                   /* -1: unlimited, as EdlibAlignConfig.k */
    const char *eq_pairs;        /* additionalEqualities as 2*n_eq_pairs bytes (first, second), or NULL */
    int32_t n_eq_pairs;
} db200_ed_params;

/* mirrors EdlibAlignResult minus the pointers (edlib.h:162-218) */
typedef struct {
    int32_t status;          /* EDLIB_STATUS_OK / EDLIB_STATUS_ERROR */
    int32_t edit_distance;   /* -1 if larger than k */
    int32_t alphabet_length;
    int32_t num_locations;
} db200_ed_result;

/*
 * locations are written as (start, end) int32 pairs, back to back in pair order;
 * loc_off[i]..loc_off[i+1] delimits pair i (units: locations).  start == DB200_ED_NO_START where
 * edlib leaves startLocations NULL.  k may be NULL (params->k applies to every pair).
 */
int db200_ed_batch_host(int device, const db200_ed_params *params,
                        const uint8_t *qbuf, const int64_t *qoff,
                        const uint8_t *tbuf, const int64_t *toff, int64_t npairs,
                        const int32_t *k,
                        db200_ed_result *results, int32_t *loc_buf, int64_t loc_cap,
                        int64_t *loc_off);

int db200_ed_batch_device(int device, const db200_ed_params *params,
                          const uint8_t *d_qbuf, const int64_t *d_qoff, int64_t q_bytes,
                          const uint8_t *d_tbuf, const int64_t *d_toff, int64_t t_bytes, int64_t npairs,
                          int32_t max_query_len, int32_t max_target_len, const int32_t *d_k,
                          db200_ed_result *d_results, int32_t *d_loc_buf, int64_t loc_cap,
                          int64_t *d_loc_off, int64_t *loc_total, void *stream);

/*
 * task "path" (EDLIB_TASK_PATH; edlib.cpp:268-285 driver, 1177-1229 obtainAlignment, 958-1157 traceback, 1247-1412
 * Hirschberg): everything db200_ed_batch_* reports for task "locations", plus the edit operations (EDLIB_EDOP_*:
 * 0 match, 1 insertion, 2 deletion, 3 mismatch) of the global alignment of the query to
 * target[start .. end] of the pair's FIRST location - the alignment the reference reports, not just one of its cost.
 * params->task must be DB200_ED_TASK_PATH.  The operations are written back to back into path_buf in pair order;
 * path_off[i]..path_off[i+1] delimits pair i (CSR like the locations; path_off has npairs + 1 entries).  A pair has at
 * most len(query) + len(target) operations, so path_cap = total query bytes + total target bytes always suffices; a
 * smaller buffer that overflows is reported with DB200_ERR_CAPACITY.  A pair has no operations where the reference
 * returns no alignment (distance above k, an empty sequence: edlib.cpp:161-179).
 * db200_edlibAlignmentToCigar formats one pair's operations as the cigar string of edlib.pyx:143-147.
 */
int db200_ed_path_batch_host(int device, const db200_ed_params *params,
                             const uint8_t *qbuf, const int64_t *qoff,
                             const uint8_t *tbuf, const int64_t *toff, int64_t npairs,
                             const int32_t *k,
                             db200_ed_result *results, int32_t *loc_buf, int64_t loc_cap, int64_t *loc_off,
                             uint8_t *path_buf, int64_t path_cap, int64_t *path_off);

/* Device-resident variant; synchronises once for the match-vector pool size and again when loc_total or path_total is
 * non-NULL (they receive the totals, and capacity / internal errors are only reported then). */
int db200_ed_path_batch_device(int device, const db200_ed_params *params,
                               const uint8_t *d_qbuf, const int64_t *d_qoff, int64_t q_bytes,
                               const uint8_t *d_tbuf, const int64_t *d_toff, int64_t t_bytes, int64_t npairs,
                               int32_t max_query_len, int32_t max_target_len, const int32_t *d_k,
                               db200_ed_result *d_results, int32_t *d_loc_buf, int64_t loc_cap,
                               int64_t *d_loc_off, int64_t *loc_total,
                               uint8_t *d_path_buf, int64_t path_cap, int64_t *d_path_off, int64_t *path_total, void *stream);

/* ------------------------------------------------------------------------------------------ */
/* Several GPUs of one box behind one call.  Pairs are independent: no collective.  */
/* The batch is cut into one contiguous range per device at equal shares of WORK (cells for     */
/* Smith-Waterman, 64-row word steps for edit distance - not equal pair counts), every range    */
/* runs through the single-device host entry point on its own host thread, and all outputs      */
/* come back in the caller's pair order.  Same arguments as db200_sw_batch_host /               */
/* db200_ed_batch_host with `device` replaced by a device list; pairs_per_device (may be NULL)  */
/* receives how many pairs each device took.  What dysgu merge's all-vs-all job lists            */
/* (merge_svs.pyx:395-416) hand over when several GPUs are present.                             */
/* ------------------------------------------------------------------------------------------ */
int db200_sw_batch_host_multi(const int32_t *devices, int32_t ndevices, const db200_sw_params *params,
                              const uint8_t *qbuf, const int64_t *qoff, const uint8_t *tbuf, const int64_t *toff,
                              int64_t npairs, const int32_t *mask_len, db200_sw_result *results, uint32_t *cigar_buf,
                              int64_t cigar_cap, int64_t *cigar_off, int32_t *status, int64_t *pairs_per_device);
int db200_ed_batch_host_multi(const int32_t *devices, int32_t ndevices, const db200_ed_params *params,
                              const uint8_t *qbuf, const int64_t *qoff, const uint8_t *tbuf, const int64_t *toff,
                              int64_t npairs, const int32_t *k, db200_ed_result *results, int32_t *loc_buf,
                              int64_t loc_cap, int64_t *loc_off, int64_t *pairs_per_device);

/* Legacy single-pair entry point with edlib's signature (edlib.h:146-271); task EDLIB_TASK_PATH fills
 * alignment / alignmentLength like the reference. */
typedef struct { char first; char second; } db200_EdlibEqualityPair;
typedef struct {
    int k;
    int mode;
    int task;
    const db200_EdlibEqualityPair *additionalEqualities;
    int additionalEqualitiesLength;
} db200_EdlibAlignConfig; /* layout-identical to EdlibAlignConfig */
typedef struct {
    int status;
    int editDistance;
    int *endLocations;
    int *startLocations;
    int numLocations;
    unsigned char *alignment;
    int alignmentLength;
    int alphabetLength;
} db200_EdlibAlignResult; /* layout-identical to EdlibAlignResult */

db200_EdlibAlignConfig db200_edlibNewAlignConfig(int k, int mode, int task,
                                                 const db200_EdlibEqualityPair *additionalEqualities,
                                                 int additionalEqualitiesLength);
db200_EdlibAlignConfig db200_edlibDefaultAlignConfig(void);
db200_EdlibAlignResult db200_edlibAlign(const char *query, int queryLength, const char *target, int targetLength,
                                        const db200_EdlibAlignConfig config);
void db200_edlibFreeAlignResult(db200_EdlibAlignResult result);
/* edlibAlignmentToCigar (edlib.h:257-271, edlib.cpp:298-347): 0 = EDLIB_CIGAR_STANDARD, 1 = EDLIB_CIGAR_EXTENDED.
 * Host-side formatting of an edit-operation array; the caller frees the returned string. */
char *db200_edlibAlignmentToCigar(const unsigned char *alignment, int alignmentLength, int cigarFormat);

/* ------------------------------------------------------------------------------------------ */
/* Batch text formatting on the host.  The strings dysgu's callers read from the   */
/* result objects, for every pair of a batch in one call, from the arrays the batch entry       */
/* points return.  No CUDA involved.  Output is CSR text: pair i is out[out_off[i]..out_off[i+1]) */
/* (not NUL-terminated); with out == NULL only out_off is filled (sizes), so a caller sizes its   */
/* buffer with a first call.                                                                    */
/* ------------------------------------------------------------------------------------------ */
/* AlignmentStructure.cigar (_ssw_wrapper.pyx:240-275): "%d%c" per BAM-encoded op, M/I/D. */
int db200_sw_cigar_text(const uint32_t *cigar_buf, const int64_t *cigar_off, int64_t npairs,
                        char *out, int64_t out_cap, int64_t *out_off);
/* aligned_query_sequence (which = 0; seq = the queries, '-' where the cigar deletes) or
 * aligned_target_sequence (which = 1; seq = the targets, '-' where it inserts) with zero-based
 * coordinates (_ssw_wrapper.pyx:302-364, including "our sequence end is sometimes beyond the cigar"). */
int db200_sw_aligned_text(int which, const uint8_t *seq_buf, const int64_t *seq_off,
                          const db200_sw_result *results, const uint32_t *cigar_buf, const int64_t *cigar_off,
                          int64_t npairs, char *out, int64_t out_cap, int64_t *out_off);
/* edlibAlignmentToCigar for every pair of a task "path" batch (format 0 standard, 1 extended;
 * edlib.cpp:298-347); a pair without operations gives an empty string. */
int db200_ed_cigar_text(const uint8_t *path_buf, const int64_t *path_off, int64_t npairs, int format,
                        char *out, int64_t out_cap, int64_t *out_off);

/* ------------------------------------------------------------------------------------------ */
/* Minimizer sketches: what graph.pyx:59-84 sliding_window_minimum(k, m, s, found) */
/* leaves in `found` for every sequence of a batch - the minimum over each window of k            */
/* consecutive m-mer hashes plus the first and the last m-mer hash, hashes = XXH64(m-mer, seed)   */
/* read as signed 64-bit (map_set_utils.pxd:28-29; dysgu: k = 16, m = 7, seed 42).  The set of     */
/* sequence i comes back sorted ascending in out[out_off[i] .. out_off[i + 1]).  A sequence has at */
/* most len - m + 1 values, so out_cap = total bytes always suffices.  db200_xxh64 is the host    */
/* side of the same hash.                                                                        */
/* ------------------------------------------------------------------------------------------ */
int db200_minimizers_batch_host(int device, const uint8_t *buf, const int64_t *off, int64_t nseq, int32_t k, int32_t m,
                                uint64_t seed, int64_t *out, int64_t out_cap, int64_t *out_off);
uint64_t db200_xxh64(const uint8_t *data, int32_t len, uint64_t seed);

/* ------------------------------------------------------------------------------------------ */
/* Ticket queue: submit / flush / fetch (SURVEY 8b "new batched ABI").                          */
/* The reference's call sites are synchronous per pair - StripedSmithWaterman(q)(t)             */
/* (_ssw_wrapper.pyx:613-652) and edlib.align (edlib.pyx:57-156).  A caller that collects its   */
/* candidates first submits them here; pairs with identical parameters form one device batch.   */
/* fetch flushes implicitly when its ticket is still pending, so submit + fetch of one pair     */
/* behaves like the legacy call.  Sequences are copied at submit time.  Pointers returned by    */
/* fetch stay valid until the next flush, reset or destroy.  Not thread-safe: one queue per      */
/* thread.                                                                                      */
/* ------------------------------------------------------------------------------------------ */
typedef struct db200_queue db200_queue;
db200_queue *db200_queue_create(int device);      /* NULL without a CUDA device */
void db200_queue_destroy(db200_queue *q);
/* mask_len < 0: the wrapper's automatic rule max(len(query)/2, 15).  Returns a ticket >= 0 or DB200_ERR_*. */
int64_t db200_queue_submit_sw(db200_queue *q, const db200_sw_params *params, const uint8_t *query, int32_t qlen,
                              const uint8_t *target, int32_t tlen, int32_t mask_len);
/* params->k applies to this pair. */
int64_t db200_queue_submit_ed(db200_queue *q, const db200_ed_params *params, const uint8_t *query, int32_t qlen,
                              const uint8_t *target, int32_t tlen);
int64_t db200_queue_pending(const db200_queue *q);
int db200_queue_flush(db200_queue *q);
/* cigar: result->cigar_len BAM-encoded ops (NULL if none); status: DB200_PAIR_* */
int db200_queue_fetch_sw(db200_queue *q, int64_t ticket, db200_sw_result *result, const uint32_t **cigar, int32_t *status);
/* locations: result->num_locations (start, end) int32 pairs (NULL if none) */
int db200_queue_fetch_ed(db200_queue *q, int64_t ticket, db200_ed_result *result, const int32_t **locations);
/* task "path" tickets: the edit operations (NULL / 0 where the reference returns no alignment or the ticket's task was another) */
int db200_queue_fetch_ed_path(db200_queue *q, int64_t ticket, const uint8_t **path, int32_t *path_len);
/* Forget every ticket and result (pending pairs are dropped). */
int db200_queue_reset(db200_queue *q);

#ifdef __cplusplus
}
#endif
#endif /* DYSGU_B200_ALIGN_H */
