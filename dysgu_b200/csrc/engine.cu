// Engine plumbing and the C ABI of libdysgu_b200_align.so (include/dysgu_b200_align.h).
#include "engine.cuh"

#include <stdarg.h>
#include <stdlib.h>
#include <unistd.h>
#include <pthread.h>
#include <new>
#include <algorithm>
#include <mutex>
#include <vector>

namespace db200 {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---- arena --------------------------------------------------------------------------------------
void *Arena::alloc_bytes(size_t bytes)
{
    bytes = (bytes + 255) & ~size_t(255);
    if (bytes == 0) bytes = 256;
    while (cur < nblocks) {
        if (used + bytes <= blocks[cur].cap) {
            void *p = blocks[cur].ptr + used;
            used += bytes;
            return p;
        }
        ++cur;
        used = 0;
    }
    if (nblocks >= 64) return nullptr;
    size_t cap = std::max(bytes, min_block);
    char *p = nullptr;
    if (cudaMalloc(&p, cap) != cudaSuccess) {
        cudaGetLastError();
        if (cap == bytes || cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        cap = bytes;
    }
    blocks[nblocks].ptr = p;
    blocks[nblocks].cap = cap;
    cur = nblocks++;
    used = bytes;
    return p;
}

void Arena::release()
{
    for (int i = 0; i < nblocks; ++i) cudaFree(blocks[i].ptr);
    nblocks = 0;
    cur = 0;
    used = 0;
}

// ---- engines (per process) ------------------------------------------------------------------------------
// fork(): a child of a process that has NOT touched CUDA through this library creates its own engines lazily (the pid
// changes, the table is cleared).  A child forked AFTER the parent initialised CUDA cannot use the GPU at all - the
// driver state it inherited is unusable and cannot be re-initialised (this is the order dysgu call -p N produces:
// re_map.remap_soft_clips in the parent, cluster.pyx:475-478, then the Pool of merge_svs.pyx:396).  That condition is
// detected here and reported as DB200_ERR_FORKED instead of a misleading "no CUDA device"; the Python layer routes such
// children to the parent's engine (dysgu_b200/server.py) and dysgu_b200.patch keeps the pool's work in the parent.
static std::mutex g_mu;
// [instance][device]: a second / third / fourth independent engine on one GPU (own streams, arenas and call lock) lets
// host threads run whole batches side by side - H2D of one step under the kernels of another (DB200_INSTANCE(dev, k))
static Engine *g_engines[4 * 64] = {nullptr};
static pid_t g_pid = 0;
static bool g_cuda_touched = false;        // this process (or the ancestor it was forked from) has called into the CUDA runtime
static bool g_forked_after_cuda = false;   // set in the child by the atfork handler
static std::once_flag g_atfork_once;

static void atfork_child()
{
    if (g_cuda_touched) g_forked_after_cuda = true;
    new (&g_mu) std::mutex();              // the parent may have held it in another thread at fork time
}

static const char *FORK_MSG =
    "this process was fork()ed after its parent initialised CUDA: the GPU cannot be used here. Keep the alignment batches in "
    "the parent (dysgu_b200.patch / dysgu_b200.server) or start the workers with the 'spawn' method";

bool forked_after_cuda() { return g_forked_after_cuda; }

static int cuda_device_count()
{
    std::call_once(g_atfork_once, [] { pthread_atfork(nullptr, nullptr, atfork_child); });
    if (g_forked_after_cuda) return -1;
    g_cuda_touched = true;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) { cudaGetLastError(); return 0; }
    return count;
}

Engine *get_engine(int device)
{
    std::lock_guard<std::mutex> lock(g_mu);
    if (g_pid != getpid()) { // first use, or a fork()ed child of a parent that had not touched CUDA
        for (auto &e : g_engines) e = nullptr;
        g_pid = getpid();
    }
    const int count = cuda_device_count();
    if (count < 0) { set_error("%s", FORK_MSG); return nullptr; }
    if (count == 0) {
        set_error("no CUDA device available: libdysgu_b200_align has no CPU fallback");
        return nullptr;
    }
    const int inst = device >= 0 ? (device >> 8) : 0;
    if (device >= 0) device &= 0xFF;
    if (device < 0 || device >= count || device >= 64 || inst > 3) { set_error("invalid device %d (have %d)", device, count); return nullptr; }
    Engine *&slot = g_engines[inst * 64 + device];
    if (slot) {
        cudaSetDevice(device);
        return slot;
    }
    if (cudaSetDevice(device) != cudaSuccess) { set_error("cudaSetDevice(%d) failed", device); return nullptr; }
    Engine *e = new Engine();
    e->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { set_error("cudaGetDeviceProperties failed"); delete e; return nullptr; }
    e->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
        set_error("cudaStreamCreate failed");
        delete e;
        return nullptr;
    }
    e->slot_stream[0] = e->stream;
    for (int k = 1; k < Engine::NSLOT; ++k) cudaStreamCreateWithFlags(&e->slot_stream[k], cudaStreamNonBlocking);
    for (int k = 0; k < Engine::NSLOT; ++k) {
        for (int i = 0; i < Engine::NAUX; ++i) {
            cudaStreamCreateWithFlags(&e->aux[k][i], cudaStreamNonBlocking);
            cudaEventCreateWithFlags(&e->ev_join[k][i], cudaEventDisableTiming);
        }
        cudaEventCreateWithFlags(&e->ev_fork[k], cudaEventDisableTiming);
    }
    cudaEventCreateWithFlags(&e->ev_chain, cudaEventDisableTiming);
    for (int i = 0; i < Engine::NSLOT; ++i) {
        cudaEventCreateWithFlags(&e->ev_copy[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&e->ev_done[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&e->ev_cig[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&e->ev_sub_join[i], cudaEventDisableTiming);
    }
    for (int i = 0; i < Engine::NSUB_MAX; ++i) cudaEventCreateWithFlags(&e->ev_sub[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&e->ev_sub_fork, cudaEventDisableTiming);
    if (cudaMalloc(&e->d_sub_words, sizeof(int64_t) * (Engine::NSUB_MAX + 2)) != cudaSuccess) {
        cudaGetLastError();
        set_error("cudaMalloc failed");
        delete e;
        return nullptr;
    }
    cudaMemset(e->d_sub_words, 0, sizeof(int64_t) * (Engine::NSUB_MAX + 2));
    const char *sb = getenv("DB200_TRACE_SLAB_MB");
    e->trace_slab_bytes = (size_t)(sb ? atol(sb) : 4096) << 20;
    if (e->trace_slab_bytes < ((size_t)64 << 20)) e->trace_slab_bytes = (size_t)64 << 20;
    e->ready = true;
    slot = e;
    return e;
}

int aux_fork(Engine *e, cudaStream_t st)
{
    const int k = e->aux_set;
    DB200_CUDA_CHECK(cudaEventRecord(e->ev_fork[k], st));
    for (int i = 0; i < Engine::NAUX; ++i) DB200_CUDA_CHECK(cudaStreamWaitEvent(e->aux[k][i], e->ev_fork[k], 0));
    return DB200_OK;
}

int aux_join(Engine *e, cudaStream_t st)
{
    const int k = e->aux_set;
    for (int i = 0; i < Engine::NAUX; ++i) {
        DB200_CUDA_CHECK(cudaEventRecord(e->ev_join[k][i], e->aux[k][i]));
        DB200_CUDA_CHECK(cudaStreamWaitEvent(st, e->ev_join[k][i], 0));
    }
    return DB200_OK;
}

// ---- stage profiling ---------------------------------------------------------------------------------
void prof_mark(Engine *e, cudaStream_t st, int stage)
{
    if (!e->profiling) return;
    if (e->ev_used == e->ev_pool.size()) {
        cudaEvent_t ev;
        if (cudaEventCreate(&ev) != cudaSuccess) return;
        e->ev_pool.push_back(ev);
    }
    cudaEvent_t ev = e->ev_pool[e->ev_used++];
    cudaEventRecord(ev, st);
    e->marks.push_back({stage, ev});
}

// ---- exclusive scan (int32 -> int64), three small kernels ---------------------------------------------
__global__ void scan_block_sums(const int32_t *in, int64_t n, int64_t *sums)
{
    __shared__ long long sh[32];
    const int64_t base = (int64_t)blockIdx.x * 1024;
    long long v = 0;
    const int64_t i = base + threadIdx.x;
    if (i < n) v = in[i];
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        long long w = sh[threadIdx.x];
        for (int o = 16; o >= 1; o >>= 1) w += __shfl_xor_sync(0xFFFFFFFFu, w, o);
        if (threadIdx.x == 0) sums[blockIdx.x] = w;
    }
}

__global__ void scan_sums(int64_t *sums, int64_t nblocks)
{
    // single thread block, sequential over chunks of 1024 block sums
    __shared__ long long sh[1024];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < nblocks; base += 1024) {
        const int64_t i = base + threadIdx.x;
        long long v = i < nblocks ? sums[i] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            long long add = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
            __syncthreads();
            sh[threadIdx.x] += add;
            __syncthreads();
        }
        if (i < nblocks) sums[i] = carry + sh[threadIdx.x] - v; // exclusive
        __syncthreads();
        if (threadIdx.x == 1023) carry += sh[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) sums[nblocks] = carry; // grand total
}

__global__ void scan_apply(const int32_t *in, int64_t n, const int64_t *sums, int64_t nblocks, int64_t *out)
{
    __shared__ long long sh[1024];
    const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    long long v = i < n ? in[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        long long add = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
        __syncthreads();
        sh[threadIdx.x] += add;
        __syncthreads();
    }
    if (i < n) out[i] = sums[blockIdx.x] + sh[threadIdx.x] - v;
    if (i == n - 1) out[n] = sums[nblocks];
}

void exclusive_scan_i32_i64(Engine *e, const int32_t *in, int64_t *out, int64_t n, int64_t *tmp, cudaStream_t st)
{
    if (n <= 0) return;
    const int64_t nblocks = (n + 1023) / 1024;
    scan_block_sums<<<(unsigned)nblocks, 1024, 0, st>>>(in, n, tmp);
    scan_sums<<<1, 1024, 0, st>>>(tmp, nblocks);
    scan_apply<<<(unsigned)nblocks, 1024, 0, st>>>(in, n, tmp, nblocks, out);
    count_launch(e, 3);
}

}  // namespace db200

// ===================================================================================================
// C ABI
// ===================================================================================================
using namespace db200;

extern "C" {

int db200_device_count(void)
{
    const int count = cuda_device_count();
    return count < 0 ? 0 : count;
}

int db200_forked_after_init(void) { return forked_after_cuda() ? 1 : 0; }

// Chunking of the *_host Smith-Waterman entry points of one engine (instance): 0 = the automatic schedule (small first chunk,
// growing - the copies of later chunks hide behind earlier compute inside ONE call); n > 0 = chunks of n pairs.  A caller that
// keeps several calls in flight on different engine instances (DB200_INSTANCE) sets n to its batch size: the overlap then
// happens between calls and every call runs its kernels once over the whole batch.
int db200_set_host_chunk_pairs(int device, int64_t pairs)
{
    Engine *e = get_engine(device);
    if (!e) return forked_after_cuda() ? DB200_ERR_FORKED : DB200_ERR_CUDA;
    e->host_chunk_pairs = pairs > 0 ? pairs : 0;
    return DB200_OK;
}

int db200_init(int device) { return get_engine(device) ? DB200_OK : (forked_after_cuda() ? DB200_ERR_FORKED : DB200_ERR_CUDA); }

// Device of the legacy single-pair entry points (their signatures have no device argument): DB200_DEVICE, else
// LOCAL_RANK, else 0 - the rule of the Python layer (batch.default_device).
int db200_default_device(void)
{
    const char *v = getenv("DB200_DEVICE");
    if (!v || !*v) v = getenv("LOCAL_RANK");
    const int d = (v && *v) ? atoi(v) : 0;
    return d < 0 ? 0 : d;
}

void db200_shutdown(void)
{
    std::lock_guard<std::mutex> lock(g_mu);
    if (g_pid != getpid()) return;
    for (auto &e : g_engines) {
        if (!e) continue;
        cudaSetDevice(e->device);
        cudaDeviceSynchronize();
        for (int i = 0; i < Engine::NSLOT; ++i) { e->arena[i].release(); e->io[i].release(); }
        e->src.release();
        for (int i = 0; i < Engine::NSLOT; ++i) if (e->pinned_out[i]) cudaFreeHost(e->pinned_out[i]);
        if (e->pinned) cudaFreeHost(e->pinned);
        for (int k = 0; k < Engine::NSLOT; ++k) {
            for (int i = 0; i < Engine::NAUX; ++i) { cudaStreamDestroy(e->aux[k][i]); cudaEventDestroy(e->ev_join[k][i]); }
            cudaEventDestroy(e->ev_fork[k]); cudaEventDestroy(e->ev_copy[k]); cudaEventDestroy(e->ev_done[k]); cudaEventDestroy(e->ev_cig[k]);
            cudaEventDestroy(e->ev_sub_join[k]);
            cudaStreamDestroy(e->slot_stream[k]);   // [0] is e->stream
        }
        cudaStreamDestroy(e->copy_stream);
        for (int i = 0; i < Engine::NSUB_MAX; ++i) cudaEventDestroy(e->ev_sub[i]);
        cudaEventDestroy(e->ev_sub_fork);
        cudaFree(e->d_sub_words);
        delete e;
        e = nullptr;
    }
}

const char *db200_last_error(void) { return g_err; }
const char *db200_version(void) { return "dysgu_b200_align 0.1.0 (sm_100a)"; }

int64_t db200_workspace_bytes(int device)
{
    Engine *e = get_engine(device);
    if (!e) return -1;
    size_t total = 0;
    for (int i = 0; i < Engine::NSLOT; ++i) total += e->arena[i].total() + e->io[i].total();
    return (int64_t)total;
}

int64_t db200_kernel_launches(int device, int reset)
{
    Engine *e = get_engine(device);
    if (!e) return -1;
    const int64_t n = e->launches;
    if (reset) e->launches = 0;
    return n;
}

// 64-row word steps (one Myers block update) the edit-distance kernels evaluated on `device` since the last reset:
// the work actually done, next to the full-matrix cell count GCUPS is quoted on.  Synchronises the device.
int64_t db200_ed_word_steps(int device, int reset)
{
    Engine *e = get_engine(device);
    if (!e) return -1;
    int64_t v = 0;
    if (cudaDeviceSynchronize() != cudaSuccess ||
        cudaMemcpy(&v, e->d_sub_words + Engine::NSUB_MAX + 1, sizeof(int64_t), cudaMemcpyDeviceToHost) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    if (reset) cudaMemset(e->d_sub_words + Engine::NSUB_MAX + 1, 0, sizeof(int64_t));
    return v;
}

int db200_profile_enable(int device, int on)
{
    Engine *e = get_engine(device);
    if (!e) return forked_after_cuda() ? DB200_ERR_FORKED : DB200_ERR_CUDA;
    e->profiling = on != 0;
    e->marks.clear();
    e->ev_used = 0;
    return DB200_OK;
}

// Accumulated milliseconds per stage since the last collect; stage i of `ms` is the time between the
// mark that opened stage i and the next mark.  Synchronises the device.
int db200_profile_collect(int device, float *ms, int nstages)
{
    Engine *e = get_engine(device);
    if (!e) return forked_after_cuda() ? DB200_ERR_FORKED : DB200_ERR_CUDA;
    DB200_CUDA_CHECK(cudaDeviceSynchronize());
    for (int i = 0; i < nstages; ++i) ms[i] = 0.f;
    for (size_t i = 1; i < e->marks.size(); ++i) {
        const int stage = e->marks[i].stage;      // the mark names the stage that just finished
        if (stage == ST_BEGIN) continue;
        float t = 0.f;
        if (cudaEventElapsedTime(&t, e->marks[i - 1].ev, e->marks[i].ev) == cudaSuccess && stage < nstages) ms[stage] += t;
    }
    e->marks.clear();
    e->ev_used = 0;
    return DB200_OK;
}

void *db200_host_alloc(int64_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, (size_t)std::max<int64_t>(bytes, 1), cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        set_error("cudaHostAlloc(%lld) failed", (long long)bytes);
        return nullptr;
    }
    return p;
}

void db200_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

// ---- Smith-Waterman -------------------------------------------------------------------------------
int db200_sw_batch_device(int device, const db200_sw_params *params, const uint8_t *d_qbuf, const int64_t *d_qoff, int64_t q_bytes,
                          const uint8_t *d_tbuf, const int64_t *d_toff, int64_t t_bytes, int64_t npairs, int32_t max_query_len,
                          int32_t max_target_len, const int32_t *d_mask_len, db200_sw_result *d_results, uint32_t *d_cigar_buf, int64_t cigar_cap,
                          int64_t *d_cigar_off, int32_t *d_status, int64_t *cigar_total, void *stream)
{
    Engine *e = get_engine(device);
    if (!e) return forked_after_cuda() ? DB200_ERR_FORKED : DB200_ERR_CUDA;
    std::lock_guard<std::recursive_mutex> call_lock(e->call_mu);   // one batch call per device at a time (arenas, streams and slots are per engine)
    cudaStream_t ust = (cudaStream_t)stream;
    // DB200_DEVICE_SUBBATCHES=n cuts the batch into n sub-batches on the engine's slot streams (they share the caller's
    // output arrays; only the placement of the cigars is ordered, SwLink).  Off by default: measured on B200 with the
    // 524,288-pair step, 1 / 2 / 3 / 4 / 6 / 8 sub-batches take 22.1 / 22.9 / 23.4 / 23.9 / 25.2 / 27.6 ms - the stage
    // tails that would overlap are shorter than the per-launch drain each extra set of 24 bucket kernels adds.
    int nsub = 1;
    if (const char *sp = getenv("DB200_DEVICE_SUBBATCHES")) nsub = atoi(sp);
    nsub = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(nsub, Engine::NSUB_MAX), npairs / 4096));
    if (nsub <= 1 || !d_qoff || !d_toff || !d_results) {
        e->arena[0].reset();
        e->aux_set = 0;
        return sw_batch_device_impl(e, e->arena[0], params, d_qbuf, d_qoff, q_bytes, d_tbuf, d_toff, t_bytes, npairs, max_query_len,
                                    max_target_len, d_mask_len, d_results, d_cigar_buf, cigar_cap, d_cigar_off, d_status, cigar_total, ust);
    }
    constexpr int NS = Engine::NSLOT;
    DB200_CUDA_CHECK(cudaMemsetAsync(e->d_sub_words, 0, sizeof(int64_t), ust));
    DB200_CUDA_CHECK(cudaEventRecord(e->ev_sub_fork, ust));
    for (int k = 0; k < NS && k < nsub; ++k) DB200_CUDA_CHECK(cudaStreamWaitEvent(e->slot_stream[k], e->ev_sub_fork, 0));
    for (int k = 0; k < NS; ++k) e->arena[k].reset();
    for (int s = 0; s < nsub; ++s) {
        const int64_t p0 = npairs * s / nsub, p1 = npairs * (s + 1) / nsub, n = p1 - p0;
        const int slot = s % NS;
        // the byte split is only known on the device: bound the slice by its pair count (sizes the workspace only)
        const int64_t qb = max_query_len > 0 ? std::min<int64_t>(q_bytes, n * (int64_t)max_query_len) : q_bytes;
        const int64_t tb = max_target_len > 0 ? std::min<int64_t>(t_bytes, n * (int64_t)max_target_len) : t_bytes;
        SwLink link;
        link.d_base = e->d_sub_words + s;
        link.d_end = e->d_sub_words + s + 1;
        link.wait = s > 0 ? e->ev_sub[s - 1] : nullptr;
        link.done = e->ev_sub[s];
        if (s >= NS) e->arena[slot].reset();   // stream order on the slot stream makes the reuse safe
        e->aux_set = slot;
        const int rc = sw_batch_device_impl(e, e->arena[slot], params, d_qbuf, d_qoff + p0, qb, d_tbuf, d_toff + p0, tb, n, max_query_len,
                                            max_target_len, d_mask_len ? d_mask_len + p0 : nullptr, d_results + p0, d_cigar_buf, cigar_cap,
                                            d_cigar_off ? d_cigar_off + p0 : nullptr, d_status ? d_status + p0 : nullptr, nullptr,
                                            e->slot_stream[slot], &link);
        if (rc) return rc;
    }
    for (int k = 0; k < NS && k < nsub; ++k) {
        DB200_CUDA_CHECK(cudaEventRecord(e->ev_sub_join[k], e->slot_stream[k]));
        DB200_CUDA_CHECK(cudaStreamWaitEvent(ust, e->ev_sub_join[k], 0));
    }
    if (cigar_total) {
        DB200_CUDA_CHECK(cudaMemcpyAsync(cigar_total, e->d_sub_words + nsub, sizeof(int64_t), cudaMemcpyDeviceToHost, ust));
        DB200_CUDA_CHECK(cudaStreamSynchronize(ust));
    }
    return DB200_OK;
}

// Device-resident batch already in pool format: d_units holds [all queries][all targets] (dense, 16-byte units of 32 bases),
// d_qlen / d_tlen the lengths in bases.  No pack kernel: one device copy puts the units next to the engine's reversed-prefix
// area.  q_bases / t_bases are the sums of the lengths (they size the workspace).
int db200_sw_batch_device_packed(int device, const db200_sw_params *params, const uint32_t *d_units, int64_t n_units, const int32_t *d_qlen,
                                 const int32_t *d_tlen, int64_t q_bases, int64_t t_bases, int64_t npairs, int32_t max_query_len,
                                 int32_t max_target_len, const int32_t *d_mask_len, db200_sw_result *d_results, uint32_t *d_cigar_buf,
                                 int64_t cigar_cap, int64_t *d_cigar_off, int32_t *d_status, int64_t *cigar_total, void *stream)
{
    Engine *e = get_engine(device);
    if (!e) return forked_after_cuda() ? DB200_ERR_FORKED : DB200_ERR_CUDA;
    std::lock_guard<std::recursive_mutex> call_lock(e->call_mu);
    if (!d_units || !d_qlen || !d_tlen || n_units < 0) { set_error("sw_batch_device_packed: bad arguments"); return DB200_ERR_ARG; }
    e->arena[0].reset();
    e->aux_set = 0;
    return sw_batch_device_impl(e, e->arena[0], params, nullptr, nullptr, q_bases, nullptr, nullptr, t_bases, npairs, max_query_len, max_target_len,
                                d_mask_len, d_results, d_cigar_buf, cigar_cap, d_cigar_off, d_status, cigar_total, (cudaStream_t)stream, nullptr,
                                d_qlen, d_tlen, d_units, n_units);
}

// Host entry point.  The batch is cut into chunks that flow through a pipeline of Engine::NSLOT slots: while chunk i is being
// aligned on the compute stream, chunk i+1 is copied to the device on the copy stream (pinned host buffers make
// the copies asynchronous).  Fixed-size results travel back right behind the kernels; the variable-size cigar
// block of a chunk is fetched once its length is known, one chunk later, so the pipeline never drains.
struct SwSlot {
    int64_t c0 = 0, n = 0;
    uint32_t *d_cig = nullptr;
    int64_t cig_bound = 0;
    bool busy = false;
};

static int sw_finalize_slot(Engine *e, int slot, SwSlot &sl, uint32_t *cigar_buf, int64_t cigar_cap, int64_t *cigar_off,
                            int64_t &cig_written, cudaStream_t st)
{
    if (!sl.busy) return DB200_OK;
    DB200_CUDA_CHECK(cudaEventSynchronize(e->ev_done[slot]));
    const int64_t total = reinterpret_cast<int64_t *>(e->pinned_out[slot])[0];
    if (cigar_buf) {
        if (total > sl.cig_bound || cig_written + total > cigar_cap) {
            cudaStreamSynchronize(st);
            set_error("sw_batch_host: cigar buffer too small");
            return DB200_ERR_CAPACITY;
        }
        if (total > 0) {
            DB200_CUDA_CHECK(cudaMemcpyAsync(cigar_buf + cig_written, sl.d_cig, sizeof(uint32_t) * (size_t)total, cudaMemcpyDeviceToHost, st));
            // the slot's io arena is re-carved for the next chunk right after this returns: its H2D (copy stream) must not
            // overtake this asynchronous read of the old cigar block
            DB200_CUDA_CHECK(cudaEventRecord(e->ev_cig[slot], st));
            DB200_CUDA_CHECK(cudaStreamWaitEvent(e->copy_stream, e->ev_cig[slot], 0));
        }
    }
    // entries [c0, c0+n) only: entry c0+n belongs to the next chunk's copy (or is set at the very end)
    if (cigar_off && cig_written) for (int64_t i = 0; i < sl.n; ++i) cigar_off[sl.c0 + i] += cig_written;
    cig_written += total;
    sl.busy = false;
    return DB200_OK;
}

// compute order between whole-batch calls of different engine instances on one device (see sw_batch_host_body)
static std::mutex g_chain_mu;
static cudaEvent_t g_chain_last[64] = {nullptr};

// One side of a host batch: CSR (len == nullptr: sequence i = buf[off[i], off[i + 1])) or slices of a shared source buffer of
// src_bytes bytes (sequence i = buf[off[i], off[i] + len[i])).
struct HostSeqs {
    const uint8_t *buf;
    const int64_t *off;
    const int32_t *len;
    int64_t src_bytes;
    bool packed = false;   // pre-packed form: buf holds pool units (16 bytes = 32 bases), off[] counts units, len[] bases
    int64_t length(int64_t i) const { return len ? (int64_t)len[i] : off[i + 1] - off[i]; }
    const uint8_t *ptr(int64_t i) const { return buf + off[i]; }
};

static int sw_batch_host_body(Engine *e, int device, const db200_sw_params *params, const HostSeqs &Q, const HostSeqs &T, int64_t npairs,
                              const int32_t *mask_len, db200_sw_result *results, uint32_t *cigar_buf, int64_t cigar_cap, int64_t *cigar_off,
                              int32_t *status);

static int sw_batch_host_locked(int device, const db200_sw_params *params, const HostSeqs &Q, const HostSeqs &T, int64_t npairs,
                                const int32_t *mask_len, db200_sw_result *results, uint32_t *cigar_buf, int64_t cigar_cap, int64_t *cigar_off,
                                int32_t *status)
{
    Engine *e = get_engine(device);
    if (!e) return forked_after_cuda() ? DB200_ERR_FORKED : DB200_ERR_CUDA;
    std::lock_guard<std::recursive_mutex> call_lock(e->call_mu);   // one batch call per device at a time (arenas, streams and slots are per engine)
    const int rc = sw_batch_host_body(e, device, params, Q, T, npairs, mask_len, results, cigar_buf, cigar_cap, cigar_off, status);
    if (rc != DB200_OK) {
        // chunks of other slots may still be copying into the caller's arrays: nothing of this call is left in flight on return
        for (int k = 0; k < Engine::NSLOT; ++k) cudaStreamSynchronize(e->slot_stream[k]);
        cudaStreamSynchronize(e->copy_stream);
        cudaGetLastError();
    }
    return rc;
}

int db200_sw_batch_host(int device, const db200_sw_params *params, const uint8_t *qbuf, const int64_t *qoff, const uint8_t *tbuf,
                        const int64_t *toff, int64_t npairs, const int32_t *mask_len, db200_sw_result *results, uint32_t *cigar_buf,
                        int64_t cigar_cap, int64_t *cigar_off, int32_t *status)
{
    if (npairs < 0 || !params || !qoff || !toff || !results) { set_error("sw_batch_host: bad arguments"); return DB200_ERR_ARG; }
    const HostSeqs Q{qbuf, qoff, nullptr, 0}, T{tbuf, toff, nullptr, 0};
    return sw_batch_host_locked(device, params, Q, T, npairs, mask_len, results, cigar_buf, cigar_cap, cigar_off, status);
}

// Pre-packed form: both sides arrive in the engine's own pool format - 4-bit codes (A/U 0, C 1, G 2, T 3, other 4, padding 5),
// eight bases per 32-bit word, every sequence padded to a 16-byte unit of 32 bases - as db200_sw_pack_host writes it.  The
// reference's C boundary also takes encoded sequences (ssw_init / ssw_align read int8 codes, ssw.h:72-125; the wrapper
// encodes, _ssw_wrapper.pyx:668-679); here the encoded form is half a byte per base, so a batch crosses PCIe at half the
// bytes of ASCII and lands in the device pool without a pack kernel.  qunit_off / tunit_off [npairs + 1] count units and
// must be dense: unit_off[i + 1] - unit_off[i] == ceil(len[i] / 32).
int db200_sw_batch_host_packed(int device, const db200_sw_params *params, const uint32_t *qunits, const int64_t *qunit_off, const int32_t *qlen,
                               const uint32_t *tunits, const int64_t *tunit_off, const int32_t *tlen, int64_t npairs, const int32_t *mask_len,
                               db200_sw_result *results, uint32_t *cigar_buf, int64_t cigar_cap, int64_t *cigar_off, int32_t *status)
{
    if (npairs < 0 || !params || !qunit_off || !tunit_off || !qlen || !tlen || !results) { set_error("sw_batch_host_packed: bad arguments"); return DB200_ERR_ARG; }
    for (int64_t i = 0; i < npairs; ++i)
        if (qlen[i] < 0 || tlen[i] < 0 || qunit_off[i + 1] - qunit_off[i] != ((int64_t)qlen[i] + 31) / 32 || tunit_off[i + 1] - tunit_off[i] != ((int64_t)tlen[i] + 31) / 32) {
            set_error("sw_batch_host_packed: unit offsets of pair %lld do not match its lengths", (long long)i);
            return DB200_ERR_ARG;
        }
    HostSeqs Q{reinterpret_cast<const uint8_t *>(qunits), qunit_off, qlen, 0}, T{reinterpret_cast<const uint8_t *>(tunits), tunit_off, tlen, 0};
    Q.packed = T.packed = true;
    return sw_batch_host_locked(device, params, Q, T, npairs, mask_len, results, cigar_buf, cigar_cap, cigar_off, status);
}

// Slices form: sequences are (start, length) views of shared source buffers that travel to the device ONCE per call - the
// shape dysgu's remapping really has: one reference region fetched per group of events (re_map.py:429-434) and many
// windows cut out of it, overlapping freely.  A side whose length array is NULL is plain CSR (start = offsets[npairs + 1]).
int db200_sw_batch_host_slices(int device, const db200_sw_params *params, const uint8_t *qsrc, int64_t qsrc_bytes, const int64_t *qstart,
                               const int32_t *qlen, const uint8_t *tsrc, int64_t tsrc_bytes, const int64_t *tstart, const int32_t *tlen,
                               int64_t npairs, const int32_t *mask_len, db200_sw_result *results, uint32_t *cigar_buf, int64_t cigar_cap,
                               int64_t *cigar_off, int32_t *status)
{
    if (npairs < 0 || !params || !qstart || !tstart || !results || qsrc_bytes < 0 || tsrc_bytes < 0) { set_error("sw_batch_host_slices: bad arguments"); return DB200_ERR_ARG; }
    for (int side = 0; side < 2; ++side) {
        const int64_t *st = side ? tstart : qstart;
        const int32_t *ln = side ? tlen : qlen;
        const int64_t cap = side ? tsrc_bytes : qsrc_bytes;
        if (!ln) continue;
        for (int64_t i = 0; i < npairs; ++i)
            if (st[i] < 0 || ln[i] < 0 || st[i] + ln[i] > cap) { set_error("sw_batch_host_slices: slice %lld leaves its source buffer", (long long)i); return DB200_ERR_ARG; }
    }
    const HostSeqs Q{qsrc, qstart, qlen, qsrc_bytes}, T{tsrc, tstart, tlen, tsrc_bytes};
    return sw_batch_host_locked(device, params, Q, T, npairs, mask_len, results, cigar_buf, cigar_cap, cigar_off, status);
}

// pairs `idx` of (Q, T) as two plain CSR batches (small groups only: the latency path, retries)
static void append_sequence(const HostSeqs &S, int64_t i, std::vector<uint8_t> &out)
{
    const int64_t len = S.length(i);
    if (!S.packed) { out.insert(out.end(), S.ptr(i), S.ptr(i) + len); return; }
    const uint32_t *w = reinterpret_cast<const uint32_t *>(S.buf) + S.off[i] * 4;   // pool units back to letters
    for (int64_t k = 0; k < len; ++k) out.push_back((uint8_t)"ACGTNN??????????"[(w[k >> 3] >> ((k & 7) * 4)) & 15u]);
}

static void materialise(const HostSeqs &Q, const HostSeqs &T, const std::vector<int64_t> &idx, const int32_t *mask_len, std::vector<uint8_t> &qb,
                        std::vector<int64_t> &qo, std::vector<uint8_t> &tb, std::vector<int64_t> &to, std::vector<int32_t> &ml)
{
    qo.assign(1, 0); to.assign(1, 0);
    for (int64_t i : idx) {
        append_sequence(Q, i, qb); qo.push_back((int64_t)qb.size());
        append_sequence(T, i, tb); to.push_back((int64_t)tb.size());
        if (mask_len) ml.push_back(mask_len[i]);
    }
    if (qb.empty()) qb.push_back(0);
    if (tb.empty()) tb.push_back(0);
}

static int sw_batch_host_body(Engine *e, int device, const db200_sw_params *params, const HostSeqs &Q, const HostSeqs &T, int64_t npairs,
                              const int32_t *mask_len, db200_sw_result *results, uint32_t *cigar_buf, int64_t cigar_cap, int64_t *cigar_off,
                              int32_t *status)
{
    const uint8_t *qbuf = Q.buf, *tbuf = T.buf;
    if (cigar_off) cigar_off[0] = 0;
    // ---- latency path: a few short pairs run in one launch (sw_small.cuh); what it cannot hold goes through the pipeline below
    static thread_local bool small_inner = false;
    if (!small_inner && npairs > 0 && npairs <= 64 && qbuf && tbuf && Q.packed == T.packed) {
        std::vector<int32_t> st_local;
        int32_t *st = status;
        if (!st) { st_local.assign((size_t)npairs, 0); st = st_local.data(); }
        std::vector<int64_t> coff_local;
        int64_t *coff = cigar_off;
        if (!coff) { coff_local.assign((size_t)npairs + 1, 0); coff = coff_local.data(); }
        int handled = 0;
        std::vector<uint8_t> sq, stb;
        std::vector<int64_t> sqo, sto;
        std::vector<int32_t> sml;
        const int64_t *qoff_csr = Q.off, *toff_csr = T.off;
        const uint8_t *qbuf_csr = qbuf, *tbuf_csr = tbuf;
        if (Q.len || T.len) {   // the small kernel reads CSR: a few short pairs are cheap to copy
            std::vector<int64_t> all((size_t)npairs);
            for (int64_t i = 0; i < npairs; ++i) all[(size_t)i] = i;
            materialise(Q, T, all, nullptr, sq, sqo, stb, sto, sml);
            qoff_csr = sqo.data(); toff_csr = sto.data(); qbuf_csr = sq.data(); tbuf_csr = stb.data();
        }
        int rc = sw_small_host(e, params, qbuf_csr, qoff_csr, tbuf_csr, toff_csr, npairs, mask_len, results, cigar_buf, cigar_cap, coff, st, &handled);
        if (rc) return rc;
        if (handled) {
            std::vector<int64_t> redo;
            for (int64_t i = 0; i < npairs; ++i) if (st[i] == 100) redo.push_back(i);   // SW_SMALL_FALLBACK
            if (redo.empty()) return DB200_OK;
            // the pairs the small kernel could not hold: one call of the batch pipeline, results spliced back in pair order
            std::vector<uint8_t> qb, tb;
            std::vector<int64_t> qo, to;
            std::vector<int32_t> ml;
            materialise(Q, T, redo, mask_len, qb, qo, tb, to, ml);
            const int64_t gn = (int64_t)redo.size();
            std::vector<db200_sw_result> r((size_t)gn);
            std::vector<int32_t> st2((size_t)gn);
            std::vector<int64_t> co((size_t)gn + 1);
            std::vector<uint32_t> cg(cigar_buf ? qb.size() + tb.size() + 2 * (size_t)gn + 16 : 1);
            small_inner = true;
            {
                const HostSeqs Q2{qb.data(), qo.data(), nullptr, 0}, T2{tb.data(), to.data(), nullptr, 0};
                rc = sw_batch_host_body(e, device, params, Q2, T2, gn, mask_len ? ml.data() : nullptr, r.data(),
                                        cigar_buf ? cg.data() : nullptr, cigar_buf ? (int64_t)cg.size() : 0, co.data(), st2.data());
            }
            small_inner = false;
            if (rc) return rc;
            std::vector<uint32_t> merged;
            std::vector<int64_t> noff((size_t)npairs + 1, 0);
            size_t rk = 0;
            for (int64_t i = 0; i < npairs; ++i) {
                noff[(size_t)i] = (int64_t)merged.size();
                if (rk < redo.size() && redo[rk] == i) {
                    results[i] = r[rk];
                    st[i] = st2[rk];
                    if (cigar_buf) merged.insert(merged.end(), cg.begin() + co[rk], cg.begin() + co[rk + 1]);
                    ++rk;
                } else if (cigar_buf) merged.insert(merged.end(), cigar_buf + coff[i], cigar_buf + coff[i + 1]);
            }
            noff[(size_t)npairs] = (int64_t)merged.size();
            if (cigar_buf) {
                if ((int64_t)merged.size() > cigar_cap) { set_error("sw_batch_host: cigar buffer too small"); return DB200_ERR_CAPACITY; }
                if (!merged.empty()) memcpy(cigar_buf, merged.data(), merged.size() * sizeof(uint32_t));
            } else {
                for (int64_t i = 0; i < npairs; ++i) noff[(size_t)i + 1] = noff[(size_t)i] + results[i].cigar_len;
            }
            if (cigar_off) memcpy(cigar_off, noff.data(), noff.size() * sizeof(int64_t));
            return DB200_OK;
        }
    }
    constexpr int NS = Engine::NSLOT;
    cudaStream_t *cs = e->slot_stream, cst = e->copy_stream;
    const int64_t MAX_CHUNK_BYTES = (int64_t)1 << 30;
    // chunk sizes grow geometrically up to about a fifth of the batch: a small first chunk keeps the exposed first copy
    // short, later (larger) chunks run the kernels at better efficiency while their copy hides behind earlier compute,
    // and the last chunk - whose dependent stage tails nothing can hide - stays moderate
    int64_t chunk_pairs = std::min<int64_t>(std::max<int64_t>(npairs / 16, 32768), 131072);
    // cap at 28 % of the batch: 524,288 pairs run as 32k / 64k / 128k / 147k / 147k (27.5 ms per step on B200; the former
    // cap of a fifth gave six chunks and 28.7 ms, one fixed size of 128k 29.1 ms, 256k 32.1 ms)
    const int64_t chunk_cap = std::min<int64_t>(std::max<int64_t>(npairs * 28 / 100, 65536), 262144);
    bool fixed_chunks = false;
    if (e->host_chunk_pairs > 0) { chunk_pairs = e->host_chunk_pairs; fixed_chunks = true; }   // db200_set_host_chunk_pairs
    if (const char *cp = getenv("DB200_HOST_CHUNK_PAIRS")) { const int64_t v = atoll(cp); if (v > 0) { chunk_pairs = v; fixed_chunks = true; } }
    std::vector<int64_t> schedule;   // DB200_HOST_CHUNK_SCHEDULE=a,b,c: explicit chunk sizes (the last one repeats); tuning aid
    if (const char *sp = getenv("DB200_HOST_CHUNK_SCHEDULE")) {
        for (const char *q = sp; *q;) { char *end; const long long v = strtoll(q, &end, 10); if (end == q) break; if (v > 0) schedule.push_back(v); q = (*end == ',') ? end + 1 : end; }
        if (!schedule.empty()) { fixed_chunks = true; chunk_pairs = schedule[0]; }
    }
    // slices form: the shared source buffers travel once, ahead of the first chunk (copy stream order puts them first)
    const uint8_t *d_qsrc = nullptr, *d_tsrc = nullptr;
    const bool packed = Q.packed && T.packed;
    if (!packed && (Q.len || T.len)) {
        e->src.reset();
        if (Q.len) {
            uint8_t *d = e->src.alloc<uint8_t>((size_t)Q.src_bytes + 16);
            if (!d) { set_error("sw_batch_host_slices: device allocation failed"); return DB200_ERR_NOMEM; }
            if (Q.src_bytes > 0) DB200_CUDA_CHECK(cudaMemcpyAsync(d, Q.buf, (size_t)Q.src_bytes, cudaMemcpyHostToDevice, cst));
            d_qsrc = d;
        }
        if (T.len) {
            uint8_t *d = e->src.alloc<uint8_t>((size_t)T.src_bytes + 16);
            if (!d) { set_error("sw_batch_host_slices: device allocation failed"); return DB200_ERR_NOMEM; }
            if (T.src_bytes > 0) DB200_CUDA_CHECK(cudaMemcpyAsync(d, T.buf, (size_t)T.src_bytes, cudaMemcpyHostToDevice, cst));
            d_tsrc = d;
        }
    }
    int64_t cig_written = 0;
    SwSlot slots[NS];
    int iter = 0;
    // DB200_TIMELINE=1: print when each chunk's copy and compute started / ended on the device (diagnostics only)
    static thread_local bool tl_inner = false;
    const bool timeline = getenv("DB200_TIMELINE") != nullptr && !tl_inner;
    struct TlRec { cudaEvent_t ev[4]; int64_t n; double host_ms; };
    std::vector<TlRec> tl;
    cudaEvent_t tl0 = nullptr;
    auto host_now = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
    const double host0 = host_now();
    if (timeline) { cudaEventCreate(&tl0); cudaEventRecord(tl0, cst); }
    for (int64_t c0 = 0; c0 < npairs; ++iter) {
        int64_t c1 = c0, bytes = 0, qb = 0, tb = 0;
        int32_t maxq = 0, maxt = 0;
        int64_t want = chunk_pairs;
        if (npairs - (c0 + want) < want / 2) want = npairs - c0;   // no tiny trailing chunk
        while (c1 < npairs && c1 - c0 < want) {
            const int64_t ql = Q.length(c1), tl = T.length(c1);
            if (ql < 0 || tl < 0 || ql > 0x7FFFFFF0 || tl > 0x7FFFFFF0) { set_error("sw_batch_host: bad offsets at pair %lld", (long long)c1); return DB200_ERR_ARG; }
            if (c1 > c0 && bytes + ql + tl > MAX_CHUNK_BYTES) break;
            bytes += ql + tl;
            qb += ql; tb += tl;
            maxq = std::max<int32_t>(maxq, (int32_t)ql);
            maxt = std::max<int32_t>(maxt, (int32_t)tl);
            ++c1;
        }
        const int slot = iter % NS;
        cudaStream_t st = cs[slot];
        SwSlot &sl = slots[slot];
        int rc = sw_finalize_slot(e, slot, sl, cigar_buf, cigar_cap, cigar_off, cig_written, st); // chunk iter-NS
        if (rc) return rc;
        const int64_t n = c1 - c0;
        // pinned staging of the rebased offsets / slice starts and lengths (+ the slot's cigar-total word in front)
        const size_t stage_bytes = 64 + sizeof(int64_t) * (size_t)(2 * (n + 1)) + sizeof(int32_t) * (size_t)(2 * n);
        if (e->pinned_out_cap[slot] < stage_bytes) {
            if (e->pinned_out[slot]) cudaFreeHost(e->pinned_out[slot]);
            e->pinned_out[slot] = nullptr;
            e->pinned_out_cap[slot] = 0;
            DB200_CUDA_CHECK(cudaHostAlloc(&e->pinned_out[slot], stage_bytes * 2, cudaHostAllocDefault));
            e->pinned_out_cap[slot] = stage_bytes * 2;
        }
        int64_t *total_word = reinterpret_cast<int64_t *>(e->pinned_out[slot]);
        int64_t *offs = total_word + 8;
        int32_t *lens = reinterpret_cast<int32_t *>(offs + 2 * (n + 1));
        if (Q.len) { for (int64_t i = 0; i < n; ++i) { offs[i] = Q.off[c0 + i]; lens[i] = Q.len[c0 + i]; } offs[n] = 0; }
        else for (int64_t i = 0; i <= n; ++i) offs[i] = Q.off[c0 + i] - Q.off[c0];
        if (T.len) { for (int64_t i = 0; i < n; ++i) { offs[n + 1 + i] = T.off[c0 + i]; lens[n + i] = T.len[c0 + i]; } offs[2 * n + 1] = 0; }
        else for (int64_t i = 0; i <= n; ++i) offs[n + 1 + i] = T.off[c0 + i] - T.off[c0];
        Arena &io = e->io[slot];
        io.reset();
        e->arena[slot].reset();
        const int64_t cig_bound = cigar_buf ? std::min<int64_t>(cigar_cap, qb + tb + 2 * n) : 0;
        const uint8_t *d_q = d_qsrc, *d_t = d_tsrc;
        const int64_t qu = packed ? Q.off[c1] - Q.off[c0] : 0, tu = packed ? T.off[c1] - T.off[c0] : 0;   // pool units of this chunk
        uint32_t *d_pk = packed ? io.alloc<uint32_t>((size_t)(qu + tu) * 4 + 16) : nullptr;
        if (packed && !d_pk) { set_error("sw_batch_host_packed: device allocation failed"); return DB200_ERR_NOMEM; }
        uint8_t *d_qc = Q.len ? nullptr : io.alloc<uint8_t>((size_t)qb + 16), *d_tc = T.len ? nullptr : io.alloc<uint8_t>((size_t)tb + 16);
        if (!Q.len) d_q = d_qc;
        if (!T.len) d_t = d_tc;
        int64_t *d_off = io.alloc<int64_t>((size_t)(2 * (n + 1)) + (size_t)n + 2);   // + the two int32 length arrays
        int32_t *d_lens = reinterpret_cast<int32_t *>(d_off + 2 * (n + 1));
        int32_t *d_mask = mask_len ? io.alloc<int32_t>((size_t)n) : nullptr;
        db200_sw_result *d_res = io.alloc<db200_sw_result>((size_t)n);
        int32_t *d_status = io.alloc<int32_t>((size_t)n);
        int64_t *d_coff = io.alloc<int64_t>((size_t)n + 1);
        uint32_t *d_cig = cigar_buf ? io.alloc<uint32_t>((size_t)cig_bound + 16) : nullptr;
        if ((!Q.len && !d_qc) || (!T.len && !d_tc) || !d_off || (mask_len && !d_mask) || !d_res || !d_status || !d_coff || (cigar_buf && !d_cig)) {
            set_error("sw_batch_host: device allocation failed");
            return DB200_ERR_NOMEM;
        }
        TlRec rec{};
        if (timeline) { for (auto &ev : rec.ev) cudaEventCreate(&ev); rec.n = n; cudaEventRecord(rec.ev[0], cst); }
        // H2D on the copy stream
        if (packed) {
            if (qu > 0) DB200_CUDA_CHECK(cudaMemcpyAsync(d_pk, qbuf + Q.off[c0] * 16, (size_t)qu * 16, cudaMemcpyHostToDevice, cst));
            if (tu > 0) DB200_CUDA_CHECK(cudaMemcpyAsync(d_pk + qu * 4, tbuf + T.off[c0] * 16, (size_t)tu * 16, cudaMemcpyHostToDevice, cst));
        }
        if (!Q.len) DB200_CUDA_CHECK(cudaMemcpyAsync(d_qc, qbuf + Q.off[c0], (size_t)qb, cudaMemcpyHostToDevice, cst));
        if (!T.len) DB200_CUDA_CHECK(cudaMemcpyAsync(d_tc, tbuf + T.off[c0], (size_t)tb, cudaMemcpyHostToDevice, cst));
        DB200_CUDA_CHECK(cudaMemcpyAsync(d_off, offs, sizeof(int64_t) * (size_t)(2 * (n + 1)) + sizeof(int32_t) * (size_t)(2 * n), cudaMemcpyHostToDevice, cst));
        if (mask_len) DB200_CUDA_CHECK(cudaMemcpyAsync(d_mask, mask_len + c0, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, cst));
        DB200_CUDA_CHECK(cudaEventRecord(e->ev_copy[slot], cst));
        DB200_CUDA_CHECK(cudaStreamWaitEvent(st, e->ev_copy[slot], 0));
        // Whole-batch calls of different engine instances on one GPU (db200_set_host_chunk_pairs >= the batch): their copies
        // overlap, their kernels do not - two batches' persistent bucket kernels sharing the SMs took 41 ms each instead of 22
        // (B200, 524,288 pairs).  The kernels of a call start when the previous call's kernels on this device have finished.
        const bool whole_batch = fixed_chunks && c0 == 0 && c1 == npairs && e->host_chunk_pairs > 0;
        // The lock is held from the wait to the record: a call that looked at the last event while another call was still
        // enqueuing its kernels would wait for the call before that one and run side by side with it (seen with three
        // threads starting together: their first calls took 64 ms of kernel time each instead of 19.7).
        std::unique_lock<std::mutex> chain_lk(g_chain_mu, std::defer_lock);
        if (whole_batch) {
            chain_lk.lock();
            if (g_chain_last[e->device] && g_chain_last[e->device] != e->ev_chain) DB200_CUDA_CHECK(cudaStreamWaitEvent(st, g_chain_last[e->device], 0));
        }
        if (timeline) { cudaEventRecord(rec.ev[1], cst); cudaEventRecord(rec.ev[2], st); }
        e->aux_set = slot;
        rc = sw_batch_device_impl(e, e->arena[slot], params, d_q, d_off, qb, d_t, d_off + n + 1, tb, n, maxq, maxt, d_mask, d_res, d_cig,
                                  cig_bound, d_coff, d_status, nullptr, st, nullptr, Q.len ? d_lens : nullptr, T.len ? d_lens + n : nullptr, d_pk, qu + tu);
        if (rc) return rc;
        if (whole_batch) {
            DB200_CUDA_CHECK(cudaEventRecord(e->ev_chain, st));
            g_chain_last[e->device] = e->ev_chain;
            chain_lk.unlock();
        }
        DB200_CUDA_CHECK(cudaMemcpyAsync(results + c0, d_res, sizeof(db200_sw_result) * (size_t)n, cudaMemcpyDeviceToHost, st));
        if (status) DB200_CUDA_CHECK(cudaMemcpyAsync(status + c0, d_status, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, st));
        // n entries only: entry c0 + n is the first one of the next chunk's copy (another stream); the chunk's total travels
        // through total_word and cigar_off[npairs] is set at the very end
        if (cigar_off) DB200_CUDA_CHECK(cudaMemcpyAsync(cigar_off + c0, d_coff, sizeof(int64_t) * (size_t)n, cudaMemcpyDeviceToHost, st));
        DB200_CUDA_CHECK(cudaMemcpyAsync(total_word, d_coff + n, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        DB200_CUDA_CHECK(cudaEventRecord(e->ev_done[slot], st));
        if (timeline) { cudaEventRecord(rec.ev[3], st); rec.host_ms = host_now() - host0; tl.push_back(rec); }
        sl.c0 = c0; sl.n = n; sl.d_cig = d_cig; sl.cig_bound = cig_bound; sl.busy = true;
        c0 = c1;
        if (!fixed_chunks) chunk_pairs = std::min<int64_t>(chunk_pairs * 2, chunk_cap);
        if (!schedule.empty()) chunk_pairs = schedule[std::min<size_t>((size_t)iter + 1, schedule.size() - 1)];
    }
    // drain in submission order
    for (int k = 0; k < NS; ++k) {
        const int slot = (iter + k) % NS;
        const int rc = sw_finalize_slot(e, slot, slots[slot], cigar_buf, cigar_cap, cigar_off, cig_written, cs[slot]);
        if (rc) return rc;
    }
    for (int k = 0; k < NS; ++k) DB200_CUDA_CHECK(cudaStreamSynchronize(cs[k]));
    if (timeline) {
        fprintf(stderr, "[db200] timeline (ms since call): host returned from sync at %.2f\n", host_now() - host0);
        for (size_t i = 0; i < tl.size(); ++i) {
            float t[4];
            for (int k = 0; k < 4; ++k) { cudaEventElapsedTime(&t[k], tl0, tl[i].ev[k]); cudaEventDestroy(tl[i].ev[k]); }
            fprintf(stderr, "[db200]   chunk %zu n=%lld  h2d %.2f-%.2f  compute %.2f-%.2f  enqueued by host at %.2f\n", i, (long long)tl[i].n,
                    t[0], t[1], t[2], t[3], tl[i].host_ms);
        }
        cudaEventDestroy(tl0);
    }
    if (cigar_off) cigar_off[npairs] = cig_written;
    // Pairs whose traceback did not fit the scratch pools of a large chunk are re-run in small groups (each group
    // then has the whole pool to itself); their cigars are appended and cigar_off is rebuilt in pair order.
    static thread_local bool in_retry = false;
    if (status && cigar_buf && cigar_off && !in_retry) {
        std::vector<int64_t> redo;
        for (int64_t i = 0; i < npairs; ++i) if (status[i] == DB200_PAIR_NOSCRATCH) redo.push_back(i);
        if (!redo.empty()) {
            if (getenv("DB200_DEBUG")) fprintf(stderr, "[db200] retrying %zu scratch-starved pairs\n", redo.size());
            std::vector<std::vector<uint32_t>> extra(redo.size());
            const size_t GROUP = 32;
            for (size_t g0 = 0; g0 < redo.size(); g0 += GROUP) {
                const size_t g1 = std::min(redo.size(), g0 + GROUP);
                std::vector<uint8_t> qb, tb;
                std::vector<int64_t> qo, to;
                std::vector<int32_t> ml;
                materialise(Q, T, std::vector<int64_t>(redo.begin() + (long)g0, redo.begin() + (long)g1), mask_len, qb, qo, tb, to, ml);
                const int64_t gn = (int64_t)(g1 - g0);
                std::vector<db200_sw_result> r((size_t)gn);
                std::vector<int32_t> st2((size_t)gn);
                std::vector<int64_t> co((size_t)gn + 1);
                std::vector<uint32_t> cg(qb.size() + tb.size() + 2 * (size_t)gn + 16);
                in_retry = true;
                const HostSeqs Q2{qb.data(), qo.data(), nullptr, 0}, T2{tb.data(), to.data(), nullptr, 0};
                const int rc = sw_batch_host_body(e, device, params, Q2, T2, gn, mask_len ? ml.data() : nullptr,
                                                  r.data(), cg.data(), (int64_t)cg.size(), co.data(), st2.data());
                in_retry = false;
                if (getenv("DB200_DEBUG")) { int nb = 0; for (auto v : st2) nb += v != 0; fprintf(stderr, "[db200] retry group of %lld: rc=%d still failing=%d\n", (long long)gn, rc, nb); }
                if (rc) return rc;
                for (int64_t k = 0; k < gn; ++k) {
                    const int64_t i = redo[g0 + (size_t)k];
                    results[i] = r[(size_t)k];
                    status[i] = st2[(size_t)k];
                    extra[g0 + (size_t)k].assign(cg.begin() + co[(size_t)k], cg.begin() + co[(size_t)k + 1]);
                }
            }
            // rebuild the cigar CSR in pair order
            std::vector<uint32_t> merged;
            merged.reserve((size_t)cig_written + 1024);
            std::vector<int64_t> noff((size_t)npairs + 1, 0);
            size_t rk = 0;
            for (int64_t i = 0; i < npairs; ++i) {
                noff[(size_t)i] = (int64_t)merged.size();
                if (rk < redo.size() && redo[rk] == i) { merged.insert(merged.end(), extra[rk].begin(), extra[rk].end()); ++rk; }
                else merged.insert(merged.end(), cigar_buf + cigar_off[i], cigar_buf + cigar_off[i + 1]);
            }
            noff[(size_t)npairs] = (int64_t)merged.size();
            if ((int64_t)merged.size() > cigar_cap) { set_error("sw_batch_host: cigar buffer too small"); return DB200_ERR_CAPACITY; }
            memcpy(cigar_buf, merged.data(), merged.size() * sizeof(uint32_t));
            memcpy(cigar_off, noff.data(), noff.size() * sizeof(int64_t));
        }
    }
    return DB200_OK;
}

// ---- legacy single-pair SW entry points (ssw.h:72-125) --------------------------------------------------
struct db200_s_profile {
    std::vector<uint8_t> read_ascii;
    int8_t mat[25];
    int32_t n;
    int8_t score_size;
};

static const char CODE2ASCII[6] = {'A', 'C', 'G', 'T', 'N', 'N'};

db200_s_profile *db200_ssw_init(const int8_t *read, int32_t readLen, const int8_t *mat, int32_t n, int8_t score_size)
{
    if (!read || !mat || n != 5 || readLen < 0) { set_error("ssw_init: only 5x5 nucleotide matrices are supported"); return nullptr; }
    db200_s_profile *p = new db200_s_profile();
    p->read_ascii.resize((size_t)readLen);
    for (int32_t i = 0; i < readLen; ++i) p->read_ascii[(size_t)i] = (uint8_t)CODE2ASCII[read[i] >= 0 && read[i] < 5 ? read[i] : 4];
    memcpy(p->mat, mat, 25);
    p->n = n;
    p->score_size = score_size;
    return p;
}

void db200_init_destroy(db200_s_profile *p) { delete p; }

db200_s_align *db200_ssw_align(const db200_s_profile *prof, const int8_t *ref, int32_t refLen, uint8_t weight_gapO, uint8_t weight_gapE,
                               uint8_t flag, uint16_t filters, int32_t filterd, int32_t maskLen)
{
    if (!prof || !ref || refLen < 0) { set_error("ssw_align: bad arguments"); return nullptr; }
    std::vector<uint8_t> t((size_t)refLen);
    for (int32_t i = 0; i < refLen; ++i) t[(size_t)i] = (uint8_t)CODE2ASCII[ref[i] >= 0 && ref[i] < 5 ? ref[i] : 4];
    db200_sw_params prm;
    memset(&prm, 0, sizeof(prm));
    prm.mat = prof->mat; prm.n = 5;
    prm.gap_open = weight_gapO; prm.gap_extend = weight_gapE; prm.score_size = prof->score_size;
    prm.flag = flag; prm.filters = filters; prm.filterd = filterd;
    int64_t qoff[2] = {0, (int64_t)prof->read_ascii.size()}, toff[2] = {0, refLen};
    db200_sw_result r;
    int32_t status = 0;
    int64_t coff[2] = {0, 0};
    std::vector<uint32_t> cig((size_t)(qoff[1] + toff[1] + 4));
    int rc = db200_sw_batch_host(db200_default_device(), &prm, prof->read_ascii.data(), qoff, t.data(), toff, 1, &maskLen, &r, cig.data(), (int64_t)cig.size(),
                                 coff, &status);
    if (rc != DB200_OK || status != DB200_PAIR_OK) {
        // as ssw_align does on error (ssw.c:803-812): a message on stderr and NULL (the reference wrapper does not check it)
        fprintf(stderr, "db200_ssw_align: %s\n", rc != DB200_OK ? g_err : "pair rejected (empty sequence, score range or 8-bit overflow with score_size 0)");
        return nullptr;
    }
    db200_s_align *a = (db200_s_align *)calloc(1, sizeof(db200_s_align));
    a->score1 = (uint16_t)r.score1; a->score2 = (uint16_t)r.score2;
    a->ref_begin1 = r.ref_begin1; a->ref_end1 = r.ref_end1; a->read_begin1 = r.read_begin1; a->read_end1 = r.read_end1;
    a->ref_end2 = r.ref_end2;
    a->cigarLen = r.cigar_len;
    if (r.cigar_len > 0) {
        a->cigar = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)r.cigar_len);
        memcpy(a->cigar, cig.data(), sizeof(uint32_t) * (size_t)r.cigar_len);
    }
    return a;
}

void db200_align_destroy(db200_s_align *a)
{
    if (!a) return;
    free(a->cigar);
    free(a);
}

// ---- edit distance ------------------------------------------------------------------------------------
int db200_ed_batch_device(int device, const db200_ed_params *params, const uint8_t *d_qbuf, const int64_t *d_qoff, int64_t q_bytes,
                          const uint8_t *d_tbuf, const int64_t *d_toff, int64_t t_bytes, int64_t npairs, int32_t max_query_len,
                          int32_t max_target_len, const int32_t *d_k, db200_ed_result *d_results, int32_t *d_loc_buf, int64_t loc_cap,
                          int64_t *d_loc_off, int64_t *loc_total, void *stream)
{
    Engine *e = get_engine(device);
    if (!e) return forked_after_cuda() ? DB200_ERR_FORKED : DB200_ERR_CUDA;
    std::lock_guard<std::recursive_mutex> call_lock(e->call_mu);   // one batch call per device at a time (arenas, streams and slots are per engine)
    e->arena[0].reset();
    e->aux_set = 0;
    return ed_batch_device_impl(e, e->arena[0], params, d_qbuf, d_qoff, q_bytes, d_tbuf, d_toff, t_bytes, npairs, max_query_len,
                                max_target_len, d_k, d_results, d_loc_buf, loc_cap, d_loc_off, loc_total, (cudaStream_t)stream);
}

int db200_ed_path_batch_device(int device, const db200_ed_params *params, const uint8_t *d_qbuf, const int64_t *d_qoff, int64_t q_bytes,
                               const uint8_t *d_tbuf, const int64_t *d_toff, int64_t t_bytes, int64_t npairs, int32_t max_query_len,
                               int32_t max_target_len, const int32_t *d_k, db200_ed_result *d_results, int32_t *d_loc_buf, int64_t loc_cap,
                               int64_t *d_loc_off, int64_t *loc_total, uint8_t *d_path_buf, int64_t path_cap, int64_t *d_path_off,
                               int64_t *path_total, void *stream)
{
    Engine *e = get_engine(device);
    if (!e) return forked_after_cuda() ? DB200_ERR_FORKED : DB200_ERR_CUDA;
    std::lock_guard<std::recursive_mutex> call_lock(e->call_mu);   // one batch call per device at a time (arenas, streams and slots are per engine)
    if (!params || params->task != DB200_ED_TASK_PATH) { set_error("ed_path_batch: params->task must be DB200_ED_TASK_PATH"); return DB200_ERR_ARG; }
    e->arena[0].reset();
    e->aux_set = 0;
    return ed_batch_device_impl(e, e->arena[0], params, d_qbuf, d_qoff, q_bytes, d_tbuf, d_toff, t_bytes, npairs, max_query_len,
                                max_target_len, d_k, d_results, d_loc_buf, loc_cap, d_loc_off, loc_total, (cudaStream_t)stream,
                                d_path_buf, path_cap, d_path_off, path_total);
}

}  // extern "C"

static int ed_batch_host_impl(int device, const db200_ed_params *params, const uint8_t *qbuf, const int64_t *qoff, const uint8_t *tbuf,
                              const int64_t *toff, int64_t npairs, const int32_t *k, db200_ed_result *results, int32_t *loc_buf,
                              int64_t loc_cap, int64_t *loc_off, uint8_t *path_buf, int64_t path_cap, int64_t *path_off)
{
    Engine *e = get_engine(device);
    if (!e) return forked_after_cuda() ? DB200_ERR_FORKED : DB200_ERR_CUDA;
    std::lock_guard<std::recursive_mutex> call_lock(e->call_mu);   // one batch call per device at a time (arenas, streams and slots are per engine)
    if (npairs < 0 || !params || !qoff || !toff || !results || !loc_buf || !loc_off) { set_error("ed_batch_host: bad arguments"); return DB200_ERR_ARG; }
    const bool want_path = params->task == DB200_ED_TASK_PATH;
    if (want_path && (!path_buf || !path_off)) {
        set_error("ed_batch_host: task 'path' needs the path buffers of db200_ed_path_batch_host");
        return DB200_ERR_ARG;
    }
    if (want_path) path_off[0] = 0;
    loc_off[0] = 0;
    // Chunks flow through two slots: while chunk i is aligned on its slot's stream, chunk i+1 is copied to the device on the
    // copy stream (the match-vector pool size is read back inside each chunk, so the host thread follows the compute closely;
    // the copies are what overlaps).
    struct Chunk { int64_t c0, c1, qb, tb; int32_t maxq, maxt; };
    std::vector<Chunk> chunks;
    const int64_t MAX_CHUNK_BYTES = (int64_t)1 << 30;
    // at least 131072 pairs per chunk: the banded / long-query kernels are latency-bound per launch (thread or warp per
    // pair), so cutting a few hundred thousand long pairs into chunks multiplies that latency (200,000 HiFi insertion
    // pairs: 58 ms in one chunk, 142 ms in four); many short pairs (clip vs window) do profit from the overlapped copies
    const int64_t want = std::min<int64_t>(std::max<int64_t>(npairs / 4, 131072), (int64_t)1 << 20);
    for (int64_t c0 = 0; c0 < npairs;) {
        int64_t c1 = c0, bytes = 0;
        int32_t maxq = 0, maxt = 0;
        int64_t lim = want;
        if (npairs - (c0 + lim) < lim) lim = npairs - c0;   // no trailing chunk below the minimum
        while (c1 < npairs && c1 - c0 < lim) {
            const int64_t ql = qoff[c1 + 1] - qoff[c1], tl = toff[c1 + 1] - toff[c1];
            if (ql < 0 || tl < 0 || ql > 0x7FFFFFF0 || tl > 0x7FFFFFF0) { set_error("ed_batch_host: bad offsets at pair %lld", (long long)c1); return DB200_ERR_ARG; }
            if (c1 > c0 && bytes + ql + tl > MAX_CHUNK_BYTES) break;
            bytes += ql + tl;
            maxq = std::max<int32_t>(maxq, (int32_t)ql);
            maxt = std::max<int32_t>(maxt, (int32_t)tl);
            ++c1;
        }
        chunks.push_back(Chunk{c0, c1, qoff[c1] - qoff[c0], toff[c1] - toff[c0], maxq, maxt});
        c0 = c1;
    }
    struct Staged { uint8_t *d_q, *d_t; int64_t *d_off; int32_t *d_k; db200_ed_result *d_res; int64_t *d_loff; int32_t *d_loc; int64_t cap;
                    uint8_t *d_path; int64_t *d_poff; };
    Staged staged[2] = {};
    cudaStream_t cst = e->copy_stream;
    auto stage = [&](size_t ci) -> int {
        const Chunk &c = chunks[ci];
        const int slot = (int)(ci & 1);
        const int64_t n = c.c1 - c.c0;
        const size_t stage_bytes = sizeof(int64_t) * (size_t)(2 * (n + 1));
        if (e->pinned_out_cap[slot] < stage_bytes) {
            if (e->pinned_out[slot]) cudaFreeHost(e->pinned_out[slot]);
            e->pinned_out[slot] = nullptr;
            e->pinned_out_cap[slot] = 0;
            DB200_CUDA_CHECK(cudaHostAlloc(&e->pinned_out[slot], stage_bytes * 2, cudaHostAllocDefault));
            e->pinned_out_cap[slot] = stage_bytes * 2;
        }
        int64_t *offs = reinterpret_cast<int64_t *>(e->pinned_out[slot]);
        for (int64_t i = 0; i <= n; ++i) { offs[i] = qoff[c.c0 + i] - qoff[c.c0]; offs[n + 1 + i] = toff[c.c0 + i] - toff[c.c0]; }
        Arena &io = e->io[slot];
        io.reset();
        Staged &s = staged[slot];
        s.cap = c.tb + 2 * n + 16;   // locations of a chunk are bounded by its target bytes
        s.d_q = io.alloc<uint8_t>((size_t)c.qb + 16);
        s.d_t = io.alloc<uint8_t>((size_t)c.tb + 16);
        s.d_off = io.alloc<int64_t>((size_t)(2 * (n + 1)));
        s.d_k = k ? io.alloc<int32_t>((size_t)n) : nullptr;
        s.d_res = io.alloc<db200_ed_result>((size_t)n);
        s.d_loff = io.alloc<int64_t>((size_t)n + 1);
        s.d_loc = io.alloc<int32_t>((size_t)(2 * s.cap + 16));
        s.d_path = want_path ? io.alloc<uint8_t>((size_t)(c.qb + c.tb + 16)) : nullptr;
        s.d_poff = want_path ? io.alloc<int64_t>((size_t)n + 1) : nullptr;
        if (!s.d_q || !s.d_t || !s.d_off || (k && !s.d_k) || !s.d_res || !s.d_loff || !s.d_loc || (want_path && (!s.d_path || !s.d_poff))) {
            set_error("ed_batch_host: device allocation failed");
            return DB200_ERR_NOMEM;
        }
        DB200_CUDA_CHECK(cudaMemcpyAsync(s.d_q, qbuf + qoff[c.c0], (size_t)c.qb, cudaMemcpyHostToDevice, cst));
        DB200_CUDA_CHECK(cudaMemcpyAsync(s.d_t, tbuf + toff[c.c0], (size_t)c.tb, cudaMemcpyHostToDevice, cst));
        DB200_CUDA_CHECK(cudaMemcpyAsync(s.d_off, offs, stage_bytes, cudaMemcpyHostToDevice, cst));
        if (k) DB200_CUDA_CHECK(cudaMemcpyAsync(s.d_k, k + c.c0, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, cst));
        DB200_CUDA_CHECK(cudaEventRecord(e->ev_copy[slot], cst));
        return DB200_OK;
    };
    int64_t written = 0, pwritten = 0;
    if (!chunks.empty()) { const int rc = stage(0); if (rc) return rc; }
    for (size_t ci = 0; ci < chunks.size(); ++ci) {
        const Chunk &c = chunks[ci];
        const int slot = (int)(ci & 1);
        cudaStream_t st = e->slot_stream[slot];
        // the other slot's previous chunk was drained at the end of the last iteration: its buffers are free
        if (ci + 1 < chunks.size()) { const int rc = stage(ci + 1); if (rc) return rc; }
        const int64_t n = c.c1 - c.c0;
        Staged &s = staged[slot];
        DB200_CUDA_CHECK(cudaStreamWaitEvent(st, e->ev_copy[slot], 0));
        e->arena[slot].reset();
        e->aux_set = slot;
        int64_t total = 0, ptotal = 0;
        int rc = ed_batch_device_impl(e, e->arena[slot], params, s.d_q, s.d_off, c.qb, s.d_t, s.d_off + n + 1, c.tb, n, c.maxq, c.maxt, s.d_k,
                                      s.d_res, s.d_loc, s.cap, s.d_loff, &total, st, s.d_path, c.qb + c.tb, s.d_poff, &ptotal);
        if (rc) return rc;
        if (want_path && pwritten + ptotal > path_cap) { set_error("ed_path_batch_host: path buffer too small (%lld bytes needed so far)", (long long)(pwritten + ptotal)); return DB200_ERR_CAPACITY; }
        if (written + total > loc_cap) { set_error("ed_batch_host: location buffer too small (%lld needed)", (long long)(written + total)); return DB200_ERR_CAPACITY; }
        DB200_CUDA_CHECK(cudaMemcpyAsync(results + c.c0, s.d_res, sizeof(db200_ed_result) * (size_t)n, cudaMemcpyDeviceToHost, st));
        DB200_CUDA_CHECK(cudaMemcpyAsync(loc_off + c.c0, s.d_loff, sizeof(int64_t) * (size_t)(n + 1), cudaMemcpyDeviceToHost, st));
        if (total > 0) DB200_CUDA_CHECK(cudaMemcpyAsync(loc_buf + 2 * written, s.d_loc, sizeof(int32_t) * 2 * (size_t)total, cudaMemcpyDeviceToHost, st));
        if (want_path) {
            DB200_CUDA_CHECK(cudaMemcpyAsync(path_off + c.c0, s.d_poff, sizeof(int64_t) * (size_t)(n + 1), cudaMemcpyDeviceToHost, st));
            if (ptotal > 0) DB200_CUDA_CHECK(cudaMemcpyAsync(path_buf + pwritten, s.d_path, (size_t)ptotal, cudaMemcpyDeviceToHost, st));
        }
        DB200_CUDA_CHECK(cudaStreamSynchronize(st));
        if (want_path && pwritten) for (int64_t i = 0; i <= n; ++i) path_off[c.c0 + i] += pwritten;
        pwritten += ptotal;
        if (written) for (int64_t i = 0; i <= n; ++i) loc_off[c.c0 + i] += written;
        written += total;
    }
    loc_off[npairs] = written;
    if (want_path) path_off[npairs] = pwritten;
    return DB200_OK;
}

extern "C" {

int db200_ed_batch_host(int device, const db200_ed_params *params, const uint8_t *qbuf, const int64_t *qoff, const uint8_t *tbuf,
                        const int64_t *toff, int64_t npairs, const int32_t *k, db200_ed_result *results, int32_t *loc_buf,
                        int64_t loc_cap, int64_t *loc_off)
{
    if (params && params->task == DB200_ED_TASK_PATH) { set_error("ed_batch_host: task 'path' goes through db200_ed_path_batch_host"); return DB200_ERR_ARG; }
    return ed_batch_host_impl(device, params, qbuf, qoff, tbuf, toff, npairs, k, results, loc_buf, loc_cap, loc_off, nullptr, 0, nullptr);
}

int db200_ed_path_batch_host(int device, const db200_ed_params *params, const uint8_t *qbuf, const int64_t *qoff, const uint8_t *tbuf,
                             const int64_t *toff, int64_t npairs, const int32_t *k, db200_ed_result *results, int32_t *loc_buf,
                             int64_t loc_cap, int64_t *loc_off, uint8_t *path_buf, int64_t path_cap, int64_t *path_off)
{
    if (!params || params->task != DB200_ED_TASK_PATH) { set_error("ed_path_batch_host: params->task must be DB200_ED_TASK_PATH"); return DB200_ERR_ARG; }
    return ed_batch_host_impl(device, params, qbuf, qoff, tbuf, toff, npairs, k, results, loc_buf, loc_cap, loc_off, path_buf, path_cap, path_off);
}

// ---- legacy single-pair edlib entry points (edlib.h:146-271) ------------------------------------------------
db200_EdlibAlignConfig db200_edlibNewAlignConfig(int k, int mode, int task, const db200_EdlibEqualityPair *additionalEqualities,
                                                 int additionalEqualitiesLength)
{
    db200_EdlibAlignConfig c;
    c.k = k; c.mode = mode; c.task = task;
    c.additionalEqualities = additionalEqualities;
    c.additionalEqualitiesLength = additionalEqualitiesLength;
    return c;
}

// edlibAlignmentToCigar (edlib.h:257-271, edlib.cpp:298-347): run-length text of an edit-operation array
// (0 match, 1 insertion, 2 deletion, 3 mismatch); format 0 = standard (M/I/D), 1 = extended (=/I/D/X).
// Host-only formatting; the caller frees the string.  NULL for an unknown format or an operation > 3.
char *db200_edlibAlignmentToCigar(const unsigned char *alignment, int alignmentLength, int cigarFormat)
{
    if (cigarFormat != 0 && cigarFormat != 1) return nullptr;
    if (alignmentLength < 0 || (alignmentLength > 0 && !alignment)) return nullptr;
    const char *letters = cigarFormat == 0 ? "MIDM" : "=IDX";
    std::string out;
    for (int i = 0; i < alignmentLength;) {
        if (alignment[i] > 3) return nullptr;
        const char c = letters[alignment[i]];
        int j = i;
        while (j < alignmentLength && alignment[j] <= 3 && letters[alignment[j]] == c) ++j;
        out += std::to_string(j - i);
        out += c;
        i = j;
    }
    char *r = (char *)malloc(out.size() + 1);
    if (!r) return nullptr;
    memcpy(r, out.c_str(), out.size() + 1);
    return r;
}

db200_EdlibAlignConfig db200_edlibDefaultAlignConfig(void) { return db200_edlibNewAlignConfig(-1, DB200_ED_MODE_NW, DB200_ED_TASK_DISTANCE, nullptr, 0); }

db200_EdlibAlignResult db200_edlibAlign(const char *query, int queryLength, const char *target, int targetLength,
                                        const db200_EdlibAlignConfig config)
{
    db200_EdlibAlignResult r;
    memset(&r, 0, sizeof(r));
    r.status = 1;
    r.editDistance = -1;
    if (queryLength < 0 || targetLength < 0) return r;
    db200_ed_params prm;
    memset(&prm, 0, sizeof(prm));
    prm.mode = config.mode; prm.task = config.task; prm.k = config.k;
    prm.eq_pairs = reinterpret_cast<const char *>(config.additionalEqualities);
    prm.n_eq_pairs = config.additionalEqualities ? config.additionalEqualitiesLength : 0;
    int64_t qoff[2] = {0, queryLength}, toff[2] = {0, targetLength}, loff[2] = {0, 0};
    db200_ed_result res;
    std::vector<int32_t> loc((size_t)(2 * (targetLength + 4)));
    const uint8_t dummy = 0;
    const bool want_path = config.task == DB200_ED_TASK_PATH;
    std::vector<uint8_t> path(want_path ? (size_t)queryLength + (size_t)targetLength + 1 : 0);
    int64_t poff[2] = {0, 0};
    int rc = ed_batch_host_impl(db200_default_device(), &prm, queryLength ? (const uint8_t *)query : &dummy, qoff, targetLength ? (const uint8_t *)target : &dummy,
                                toff, 1, nullptr, &res, loc.data(), targetLength + 4, loff, want_path ? path.data() : nullptr,
                                (int64_t)path.size(), want_path ? poff : nullptr);
    const int32_t path_len = (int32_t)(poff[1] - poff[0]);
    if (rc != DB200_OK) return r;
    r.status = res.status;
    r.editDistance = res.edit_distance;
    r.alphabetLength = res.alphabet_length;
    r.numLocations = res.num_locations;
    if (res.num_locations > 0) {
        r.endLocations = (int *)malloc(sizeof(int) * (size_t)res.num_locations);
        const bool starts = loc[0] != DB200_ED_NO_START;
        if (starts) r.startLocations = (int *)malloc(sizeof(int) * (size_t)res.num_locations);
        for (int i = 0; i < res.num_locations; ++i) {
            r.endLocations[i] = loc[(size_t)(2 * i + 1)];
            if (starts) r.startLocations[i] = loc[(size_t)(2 * i)];
        }
    }
    if (want_path && path_len > 0) {   // edlib.cpp:268-285: the operations of the first location's alignment
        r.alignment = (unsigned char *)malloc((size_t)path_len);
        if (r.alignment) { memcpy(r.alignment, path.data(), (size_t)path_len); r.alignmentLength = path_len; }
    }
    return r;
}

void db200_edlibFreeAlignResult(db200_EdlibAlignResult result)
{
    free(result.endLocations);
    free(result.startLocations);
    free(result.alignment);
}

}  // extern "C"
