// Minimizer sketches of clip sequences on the device (the integer hashing scan next to the alignment path).
//
// What graph.pyx:59-84 (sliding_window_minimum, called per soft clip from ClipScoper._add_m_find_candidates, graph.pyx:146)
// puts into its `found` set for one sequence s with window k and m-mer length m (dysgu: k = 16, m = 7, graph.pyx:1467):
//   h[i] = (signed 64-bit) XXH64(s[i .. i + m), seed 42)            for i in [0, len(s) - m]      (map_set_utils.pxd:28-29)
//   found = { min(h[max(0, i - k + 1) .. i]) : all i } + { h[0], h[len(s) - m] }       (signed comparison, graph.pyx:75-83)
// Here: one thread block per sequence; hashes of all positions, window minima, candidates whose value differs from the
// position before, then a block-wide bitonic sort and a unique pass - the set comes back sorted ascending (as signed
// values), CSR over the batch.  XXH64 is the published algorithm (Yann Collet; the reference vendors Stephan Brumme's
// single-header version, include/xxhash64.h); for inputs shorter than 32 bytes only its tail path runs.
#include "engine.cuh"

namespace db200 {

__host__ __device__ inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

// XXH64 of `len` bytes at p (any length); seed as given.
__host__ __device__ inline uint64_t xxh64(const uint8_t *p, int len, uint64_t seed)
{
    const uint64_t P1 = 11400714785074694791ull, P2 = 14029467366897019727ull, P3 = 1609587929392839161ull, P4 = 9650029242287828579ull,
                   P5 = 2870177450012600261ull;
    auto rd64 = [](const uint8_t *q) { uint64_t v = 0; for (int b = 7; b >= 0; --b) v = (v << 8) | q[b]; return v; };
    auto rd32 = [](const uint8_t *q) { return (uint64_t)q[0] | ((uint64_t)q[1] << 8) | ((uint64_t)q[2] << 16) | ((uint64_t)q[3] << 24); };
    auto round1 = [&](uint64_t acc, uint64_t in) { acc += in * P2; acc = rotl64(acc, 31); return acc * P1; };
    const uint8_t *end = p + len;
    uint64_t h;
    if (len >= 32) {
        uint64_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
        while (p + 32 <= end) {
            v1 = round1(v1, rd64(p)); v2 = round1(v2, rd64(p + 8)); v3 = round1(v3, rd64(p + 16)); v4 = round1(v4, rd64(p + 24));
            p += 32;
        }
        h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
        h = (h ^ round1(0, v1)) * P1 + P4;
        h = (h ^ round1(0, v2)) * P1 + P4;
        h = (h ^ round1(0, v3)) * P1 + P4;
        h = (h ^ round1(0, v4)) * P1 + P4;
    } else {
        h = seed + P5;
    }
    h += (uint64_t)len;
    while (p + 8 <= end) { h ^= round1(0, rd64(p)); h = rotl64(h, 27) * P1 + P4; p += 8; }
    if (p + 4 <= end) { h ^= rd32(p) * P1; h = rotl64(h, 23) * P2 + P3; p += 4; }
    while (p < end) { h ^= (uint64_t)(*p) * P5; h = rotl64(h, 11) * P1; ++p; }
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
    return h;
}

constexpr int MZ_THREADS = 256;
constexpr int MZ_SMEM_ITEMS = 4096;   // candidates sorted in shared memory; longer lists are sorted in the global scratch

// scratch layout per sequence (int64 slots at 2 * off[s]): [0, len) hashes, then [len, 2 len) candidates / sorted set
__global__ void __launch_bounds__(MZ_THREADS) minimizer_kernel(const uint8_t *buf, const int64_t *off, int64_t nseq, int k, int m, uint64_t seed,
                                                              long long *scratch, int32_t *count)
{
    __shared__ long long sh[MZ_SMEM_ITEMS];
    __shared__ int sh_n;
    for (int64_t s = blockIdx.x; s < nseq; s += gridDim.x) {
        const uint8_t *seq = buf + off[s];
        const int len = (int)(off[s + 1] - off[s]);
        const int end = len - m + 1;
        __syncthreads();
        if (threadIdx.x == 0) sh_n = 0;
        __syncthreads();
        if (end <= 0) { if (threadIdx.x == 0) count[s] = 0; continue; }
        long long *h = scratch + 2 * off[s];
        long long *cand = h + len;
        for (int i = threadIdx.x; i < end; i += MZ_THREADS) h[i] = (long long)xxh64(seq + i, m, seed);
        __syncthreads();
        // window minima; a position is a candidate when its minimum differs from the previous position's, the two end
        // m-mers always are (graph.pyx:73-74)
        for (int i0 = 0; i0 < end; i0 += MZ_THREADS) {
            const int i = i0 + threadIdx.x;
            if (i < end) {
                long long mn = h[i], mp = 0x7FFFFFFFFFFFFFFFll;
                for (int j = max(0, i - k + 1); j < i; ++j) mn = min(mn, h[j]);
                if (i > 0) { mp = h[i - 1]; for (int j = max(0, i - k); j < i - 1; ++j) mp = min(mp, h[j]); }
                if (i == 0 || mn != mp) cand[atomicAdd(&sh_n, 1)] = mn;
                if ((i == 0 || i == end - 1) && h[i] != mn) cand[atomicAdd(&sh_n, 1)] = h[i];
            }
        }
        __syncthreads();
        const int n = sh_n;
        // bitonic sort of the candidates (padded with +inf to a power of two), in shared memory when they fit
        int np2 = 1;
        while (np2 < n) np2 <<= 1;
        long long *a = np2 <= MZ_SMEM_ITEMS ? sh : cand;
        if (a == sh) for (int i = threadIdx.x; i < np2; i += MZ_THREADS) sh[i] = i < n ? cand[i] : 0x7FFFFFFFFFFFFFFFll;
        else for (int i = n + threadIdx.x; i < np2 && i < len; i += MZ_THREADS) cand[i] = 0x7FFFFFFFFFFFFFFFll;
        // (the global region holds len >= np2 slots whenever n > MZ_SMEM_ITEMS: n <= end + 2 and np2 < 2 n; guarded below)
        const bool ok = a == sh || np2 <= len;
        __syncthreads();
        if (ok) {
            for (int size = 2; size <= np2; size <<= 1)
                for (int stride = size >> 1; stride > 0; stride >>= 1) {
                    for (int i = threadIdx.x; i < np2; i += MZ_THREADS) {
                        const int j = i ^ stride;
                        if (j > i) {
                            const bool up = (i & size) == 0;
                            const long long x = a[i], y = a[j];
                            if ((x > y) == up) { a[i] = y; a[j] = x; }
                        }
                    }
                    __syncthreads();
                }
        } else if (threadIdx.x == 0) {   // never expected (np2 > len): plain insertion sort keeps the result right
            for (int i = 1; i < n; ++i) { const long long x = cand[i]; int j = i - 1; while (j >= 0 && cand[j] > x) { cand[j + 1] = cand[j]; --j; } cand[j + 1] = x; }
        }
        __syncthreads();
        // unique: one thread walks the sorted list (a few hundred items for a clip)
        if (threadIdx.x == 0) {
            int u = 0;
            long long prev = 0;
            for (int i = 0; i < n; ++i) {
                const long long x = a[i];
                if (i == 0 || x != prev) { cand[u++] = x; prev = x; }   // u <= i: never overwrites an unread item of `cand`
            }
            count[s] = u;
        }
        __syncthreads();
    }
}

__global__ void minimizer_gather_kernel(const int64_t *off, int64_t nseq, const long long *scratch, const int64_t *out_off, long long *out, int64_t cap,
                                        int32_t *overflow)
{
    const int64_t s = blockIdx.x;
    if (s >= nseq) return;
    const int len = (int)(off[s + 1] - off[s]);
    const long long *src = scratch + 2 * off[s] + len;
    const int64_t o = out_off[s], n = out_off[s + 1] - o;
    if (o + n > cap) { if (threadIdx.x == 0) atomicExch(overflow, 1); return; }
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) out[o + i] = src[i];
}

}  // namespace db200

using namespace db200;

extern "C" {

uint64_t db200_xxh64(const uint8_t *data, int32_t len, uint64_t seed) { return xxh64(data, len, seed); }

int db200_minimizers_batch_host(int device, const uint8_t *buf, const int64_t *off, int64_t nseq, int32_t k, int32_t m, uint64_t seed,
                                int64_t *out, int64_t out_cap, int64_t *out_off)
{
    Engine *e = get_engine(device);
    if (!e) return forked_after_cuda() ? DB200_ERR_FORKED : DB200_ERR_CUDA;
    std::lock_guard<std::recursive_mutex> call_lock(e->call_mu);
    if (nseq < 0 || !off || !out_off || k <= 0 || m <= 0 || m > 255 || (nseq > 0 && !buf)) { set_error("minimizers_batch_host: bad arguments"); return DB200_ERR_ARG; }
    out_off[0] = 0;
    if (nseq == 0) return DB200_OK;
    const int64_t bytes = off[nseq] - off[0];
    if (bytes < 0 || nseq > ((int64_t)1 << 30)) { set_error("minimizers_batch_host: bad offsets"); return DB200_ERR_ARG; }
    Arena &ws = e->arena[0];
    ws.reset();
    cudaStream_t st = e->stream;
    uint8_t *d_buf = ws.alloc<uint8_t>((size_t)bytes + 16);
    int64_t *d_off = ws.alloc<int64_t>((size_t)nseq + 1);
    long long *d_scratch = ws.alloc<long long>((size_t)(2 * bytes) + 16);
    int32_t *d_count = ws.alloc<int32_t>((size_t)nseq + 1);
    int64_t *d_out_off = ws.alloc<int64_t>((size_t)nseq + 1);
    int64_t *d_tmp = ws.alloc<int64_t>(scan_tmp_count(nseq));
    long long *d_out = ws.alloc<long long>((size_t)std::max<int64_t>(out_cap, 1));
    int32_t *d_flag = ws.alloc<int32_t>(4);
    if (!d_buf || !d_off || !d_scratch || !d_count || !d_out_off || !d_tmp || !d_out || !d_flag) { set_error("minimizers_batch_host: device allocation failed"); return DB200_ERR_NOMEM; }
    std::vector<int64_t> reb((size_t)nseq + 1);
    for (int64_t i = 0; i <= nseq; ++i) reb[(size_t)i] = off[i] - off[0];
    DB200_CUDA_CHECK(cudaMemcpyAsync(d_buf, buf + off[0], (size_t)bytes, cudaMemcpyHostToDevice, st));
    DB200_CUDA_CHECK(cudaMemcpyAsync(d_off, reb.data(), sizeof(int64_t) * (size_t)(nseq + 1), cudaMemcpyHostToDevice, st));
    DB200_CUDA_CHECK(cudaMemsetAsync(d_flag, 0, 16, st));
    const unsigned grid = (unsigned)std::min<int64_t>(nseq, (int64_t)e->sm_count * 8);
    minimizer_kernel<<<grid, MZ_THREADS, 0, st>>>(d_buf, d_off, nseq, k, m, seed, d_scratch, d_count);
    exclusive_scan_i32_i64(e, d_count, d_out_off, nseq, d_tmp, st);
    minimizer_gather_kernel<<<(unsigned)nseq, 128, 0, st>>>(d_off, nseq, d_scratch, d_out_off, d_out, out_cap, d_flag);
    count_launch(e, 2);
    DB200_CUDA_CHECK(cudaGetLastError());
    int32_t flag = 0;
    DB200_CUDA_CHECK(cudaMemcpyAsync(out_off, d_out_off, sizeof(int64_t) * (size_t)(nseq + 1), cudaMemcpyDeviceToHost, st));
    DB200_CUDA_CHECK(cudaMemcpyAsync(&flag, d_flag, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    DB200_CUDA_CHECK(cudaStreamSynchronize(st));
    if (flag || out_off[nseq] > out_cap) { set_error("minimizers_batch_host: output buffer too small (%lld values)", (long long)out_off[nseq]); return DB200_ERR_CAPACITY; }
    if (out_off[nseq] > 0) {
        DB200_CUDA_CHECK(cudaMemcpyAsync(out, d_out, sizeof(int64_t) * (size_t)out_off[nseq], cudaMemcpyDeviceToHost, st));
        DB200_CUDA_CHECK(cudaStreamSynchronize(st));
    }
    return DB200_OK;
}

}  // extern "C"
