// Shared engine plumbing for libdysgu_b200_align.so: per-device state, grow-only workspace arena,
// error reporting, launch counting and a small device-side exclusive scan.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <algorithm>
#include <mutex>
#include <vector>

#include "../../include/dysgu_b200_align.h"

namespace db200 {

void set_error(const char *fmt, ...);

#define DB200_CUDA_CHECK(expr)                                                                       \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess) {                                                                     \
            db200::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return DB200_ERR_CUDA;                                                                   \
        }                                                                                            \
    } while (0)

// Grow-only device arena made of cudaMalloc'd blocks that never move.  A batch call resets the
// cursor and carves regions; a request that does not fit the remaining blocks adds a block
// (device-synchronising cudaMalloc), which only happens while batch sizes are still increasing.
struct Arena {
    struct Block { char *ptr; size_t cap; };
    Block blocks[64];
    int nblocks = 0;
    int cur = 0;
    size_t used = 0;
    size_t min_block = (size_t)64 << 20;

    void reset() { cur = 0; used = 0; }
    size_t total() const { size_t t = 0; for (int i = 0; i < nblocks; ++i) t += blocks[i].cap; return t; }
    void *alloc_bytes(size_t bytes);
    void release();
    template <typename T>
    T *alloc(size_t count) { return reinterpret_cast<T *>(alloc_bytes(count * sizeof(T))); }
};

struct Engine {
    int device = -1;
    bool ready = false;
    std::recursive_mutex call_mu;                 // held by every batch entry point for the length of the call
    cudaStream_t stream = nullptr;       // used by the *_host entry points
    static constexpr int NSLOT = 3;      // host chunks in flight: copy of chunk i+2 / compute of chunk i+1 / drain of chunk i
    cudaStream_t slot_stream[NSLOT] = {}; // compute stream per slot ([0] is `stream`): consecutive chunks overlap their stage tails
    cudaStream_t copy_stream = nullptr;  // H2D prefetch of the next chunk
    static constexpr int NAUX = 24;      // side streams: independent length buckets run concurrently so that the
    cudaStream_t aux[NSLOT][NAUX] = {};  // drain phase of one bucket overlaps the bulk of the next; one set per
    cudaEvent_t ev_fork[NSLOT] = {};     // in-flight host chunk so that two chunks can fill the GPU together
    cudaEvent_t ev_join[NSLOT][NAUX] = {};
    int aux_set = 0;                     // set used by the batch being enqueued
    static constexpr int NSUB_MAX = 16;  // sub-batches of one device call (db200_sw_batch_device)
    cudaEvent_t ev_sub[NSUB_MAX] = {};   // cigar placement of sub-batch i done
    cudaEvent_t ev_sub_fork = nullptr, ev_sub_join[NSLOT] = {};
    int64_t *d_sub_words = nullptr;      // [NSUB_MAX + 1] cigar cursor between sub-batches; [NSUB_MAX + 1] Myers word steps evaluated
    cudaEvent_t ev_copy[NSLOT] = {};
    cudaEvent_t ev_done[NSLOT] = {};
    cudaEvent_t ev_chain = nullptr;      // kernels of this engine's latest whole-batch host call have finished
    cudaEvent_t ev_cig[NSLOT] = {};      // cigar block of the slot's previous chunk has left its io arena
    Arena arena[NSLOT];                  // compute workspace (one per in-flight chunk)
    Arena io[NSLOT];                     // device copies of host inputs / outputs (one per in-flight chunk)
    Arena src;                           // slices form: the shared source buffers of a host call, uploaded once
    int64_t launches = 0;
    int sm_count = 0;
    void *pinned = nullptr;              // small pinned scratch (counters read back from the device)
    size_t pinned_cap = 0;
    void *pinned_out[NSLOT] = {};
    size_t pinned_out_cap[NSLOT] = {};
    int64_t host_chunk_pairs = 0;        // db200_set_host_chunk_pairs: 0 = automatic chunk schedule of the host entry points
    void *small_host = nullptr, *small_dev = nullptr;   // mapped page-locked staging of the one-launch small-batch path (sw_small.cuh)
    size_t small_cap = 0;
    int32_t small_epoch = 0;
    size_t trace_slab_bytes = (size_t)4 << 30;  // per-block slabs of the wide banded traceback (DB200_TRACE_SLAB_MB)
    // optional stage profiling (CUDA events recorded on the launching stream between pipeline stages)
    bool profiling = false;
    struct Mark { int stage; cudaEvent_t ev; };
    std::vector<Mark> marks;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
};

Engine *get_engine(int device);  // nullptr + error set on failure
bool forked_after_cuda();        // this process was forked after its parent initialised CUDA (no GPU use possible)

inline void count_launch(Engine *e, int n = 1) { e->launches += n; }

// stage ids reported by db200_profile_collect (a mark closes the stage named by the PREVIOUS mark)
enum Stage { ST_BEGIN = 0, ST_PACK, ST_BUCKET, ST_FORWARD, ST_FORWARD_WATCH, ST_REVERSE_PREP, ST_REVERSE, ST_REVERSE_WATCH,
             ST_FINALIZE, ST_TRACEBACK, ST_CIGAR, ST_ED_PREP, ST_ED_FORWARD, ST_ED_STARTS, ST_ED_OUTPUT, ST_ED_PATH, ST_COUNT };
void prof_mark(Engine *e, cudaStream_t st, int stage);
// fork: side streams wait for everything enqueued on `st` so far; join: `st` waits for all side streams
int aux_fork(Engine *e, cudaStream_t st);
int aux_join(Engine *e, cudaStream_t st);

// Exclusive scan of int32 counts into int64 offsets (offsets[n] = total).  n up to 2^31.
// tmp must hold ceil(n / 1024) + 1 int64 values.
void exclusive_scan_i32_i64(Engine *e, const int32_t *in, int64_t *out, int64_t n, int64_t *tmp, cudaStream_t st);
inline size_t scan_tmp_count(int64_t n) { return (size_t)((n + 1023) / 1024 + 2); }

// Chaining of the sub-batches of one device call: they run on different streams and share one cigar CSR, so the
// only ordered step is the placement of the cigars (each sub-batch starts where its predecessor ended).
struct SwLink {
    const int64_t *d_base;   // device scalar: first cigar slot of this sub-batch (nullptr: 0)
    int64_t *d_end;          // device scalar: receives d_base + the cigar ops of this sub-batch
    cudaEvent_t wait;        // recorded by the predecessor after it wrote its d_end (nullptr: none)
    cudaEvent_t done;        // recorded after d_end is written (nullptr: not needed)
};

int sw_batch_device_impl(Engine *e, Arena &ws, const db200_sw_params *params, const uint8_t *d_qbuf, const int64_t *d_qoff,
                         int64_t q_bytes, const uint8_t *d_tbuf, const int64_t *d_toff, int64_t t_bytes, int64_t npairs,
                         int32_t max_query_len, int32_t max_target_len, const int32_t *d_mask_len, db200_sw_result *d_results, uint32_t *d_cigar_buf, int64_t cigar_cap,
                         int64_t *d_cigar_off, int32_t *d_status, int64_t *cigar_total, cudaStream_t st, const SwLink *link = nullptr,
                         const int32_t *d_qlen = nullptr, const int32_t *d_tlen = nullptr,   // slices form: q_bytes / t_bytes are the sums of the lengths
                         const uint32_t *d_packed = nullptr, int64_t packed_units = 0);   // pre-packed form: pool units of [all queries][all targets], lengths in d_qlen / d_tlen

int sw_small_host(Engine *e, const db200_sw_params *prm, const uint8_t *qbuf, const int64_t *qoff, const uint8_t *tbuf,
                  const int64_t *toff, int64_t npairs, const int32_t *mask_len, db200_sw_result *results, uint32_t *cigar_buf,
                  int64_t cigar_cap, int64_t *cigar_off, int32_t *status, int *handled);

int ed_batch_device_impl(Engine *e, Arena &ws, const db200_ed_params *params, const uint8_t *d_qbuf, const int64_t *d_qoff,
                         int64_t q_bytes, const uint8_t *d_tbuf, const int64_t *d_toff, int64_t t_bytes, int64_t npairs,
                         int32_t max_query_len, int32_t max_target_len, const int32_t *d_k, db200_ed_result *d_results, int32_t *d_loc_buf, int64_t loc_cap,
                         int64_t *d_loc_off, int64_t *loc_total, cudaStream_t st,
                         uint8_t *d_path_buf = nullptr, int64_t path_cap = 0, int64_t *d_path_off = nullptr, int64_t *path_total = nullptr);

}  // namespace db200
