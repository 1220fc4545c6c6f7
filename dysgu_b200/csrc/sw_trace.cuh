// Banded traceback / cigar stage (behaviour of banded_sw, ssw.c:550-728), three kernels:
//   sw_trace_classify_kernel  thread per pair: degenerate and gap-free alignments are emitted directly, the rest is
//                             split by band width into a "narrow" and a "wide" work list
//   sw_trace_narrow_kernel    thread per listed pair: scalar banded DP (bands of a few columns, short soft clips)
//   sw_trace_wide_kernel      block per listed pair: anti-diagonal parallel banded DP (long contigs), state in
//                             shared memory, directions in a per-block scratch slab
// All of them reproduce the reference's band rule (cells outside the band read as 0), its tie-breaks
// (ssw.c:610,615,625-626), the band doubling (ssw.c:571-631) and the walk from the bottom-right corner while
// i > 0 followed by one closing M (ssw.c:641-707).
#pragma once
#include "sw_kernels.cuh"

namespace db200 {

struct TraceArgs {
    const uint4 *pool;
    const PairMeta *meta;
    PairState *state;
    db200_sw_result *results;
    int64_t npairs;
    SwScoring sc;
    uint8_t *slab;                 // cooperative kernels: one slab per block (levels run one after the other)
    uint32_t *cigar_pool;
    unsigned long long *cigar_cursor;
    unsigned long long cigar_cap;
    int32_t *overflow_flag;
    // work lists per level: 0 = scalar (thread per pair, directions in shared memory), 1 = one warp per pair,
    // 2 = one block per pair, 3 = scalar with the directions in the slab (kept apart from level 0 so that the warps of
    // either launch hold jobs of similar size)
    int32_t *list[4];
    int32_t *count[4];
    int32_t *counter[3];           // work counters of the cooperative kernels
    int32_t coop_blocks[3];        // grid size per level (slab = slab_total / blocks)
    unsigned long long slab_total;
    unsigned long long *medium_cursor;   // level 0, jobs too large for the shared-memory slice: regions of the slab handed out so far
    int32_t *class_count;                // [4][TRACE_MEDIUM_CLASSES] fill of the sub-lists of every list (count[l] is their sum)
    int32_t medium_min_jobs;             // fewer slab-backed scalar jobs than this go to the warp-per-pair kernel instead
    int32_t debug;
};

// level 0 (thread per pair) only takes tiny jobs: rows x band width below this many cells per attempt;
// level 1 (warp per pair) takes bands of up to TRACE_WARP_BW columns either side; level 2 (block per pair) the rest
constexpr int TRACE_NARROW_CELLS = 768;
// ... and up to this many cells with the direction codes in a private region of the slab instead (one byte per cell): a
// 300-row contig with a band of 3-17 columns is still a job for one thread - the warp-per-pair sweep of level 1 would run
// it with 3-17 of its 32 lanes and a barrier per anti-diagonal (paired-end contigs: traceback stage 9.5 -> see DESIGN.md)
constexpr int TRACE_MEDIUM_CELLS = 3072;
constexpr int TRACE_MEDIUM_REGION = 4096;      // bytes of slab per such job: nibbles, rows padded to 8 cells (<= 1024 rows of one word)
constexpr int TRACE_MEDIUM_CLASSES = 8;        // sub-lists of list 3 by job size, so that the 32 jobs of a warp are of similar length
constexpr int TRACE_MEDIUM_CLASS_CELLS = TRACE_MEDIUM_CELLS / TRACE_MEDIUM_CLASSES;
constexpr int TRACE_MEDIUM_MIN_JOBS = 16384;   // below this many such jobs in a batch the warp-per-pair kernel is quicker
constexpr int TRACE_NARROW_WIDTH = 30;   // widest band row the scalar kernel keeps in its shared-memory slice
constexpr int TRACE_WARP_BW = 62;      // 125 band diagonals, at most 63 cells on an anti-diagonal: two per lane
// (bands up to 126 either side at this level measured slower: long_contigs traceback 44.7 -> 59.6 ms; the block level
// sweeps a wide anti-diagonal in one slot)
// per-thread slice of the scalar kernel: direction nibbles of TRACE_NARROW_CELLS cells + four rolling int16 rows,
// padded to an odd word count so that the 32 slices of a warp start in 32 different banks
constexpr int TRACE_NARROW_DIR_WORDS = TRACE_NARROW_CELLS / 8;
constexpr int TRACE_NARROW_ROW = TRACE_NARROW_WIDTH + 2;
constexpr int TRACE_NARROW_WORDS = (TRACE_NARROW_DIR_WORDS + 4 * TRACE_NARROW_ROW / 2) | 1;
constexpr int TRACE_MEDIUM_WORDS = (4 * TRACE_NARROW_ROW / 2) | 1;   // list 3 keeps its directions in the slab: rolling rows only
constexpr int TRACE_NARROW_THREADS = 128;
// The work lists are TRACE_MEDIUM_CLASSES sub-lists of `stride` entries each, by job size, taken largest class first:
// the 32 jobs of a warp of the scalar kernel are of similar length, and the cooperative kernels start their longest
// pairs first instead of leaving one for the end of the launch (a 6 kb contig pair sweeps for milliseconds).
__device__ __forceinline__ int trace_class(int lvl, int readLen, int refLen, int width)
{
    if (lvl == 3) return min(TRACE_MEDIUM_CLASSES - 1, (readLen * width) / TRACE_MEDIUM_CLASS_CELLS);
    if (lvl == 0) return min(TRACE_MEDIUM_CLASSES - 1, (readLen * width) / (TRACE_NARROW_CELLS / TRACE_MEDIUM_CLASSES));
    return max(0, min(TRACE_MEDIUM_CLASSES - 1, 24 - __clz(readLen + refLen)));   // < 256 steps: 0 ... >= 16384: 7
}
__device__ __forceinline__ void trace_push(int32_t *const *list, int32_t *const *count, int32_t *ccount, long long stride, int lvl, int p,
                                           int readLen, int refLen, int width)
{
    const int c = trace_class(lvl, readLen, refLen, width);
    atomicAdd(count[lvl], 1);
    list[lvl][(long long)c * stride + atomicAdd(ccount + lvl * TRACE_MEDIUM_CLASSES + c, 1)] = p;
}
// task r of a classed list, largest class first (the classes are complete before the consuming launch starts)
__device__ __forceinline__ int trace_task(const int32_t *list, const int32_t *ccount, long long stride, int lvl, long long r)
{
    const int32_t *cc = ccount + lvl * TRACE_MEDIUM_CLASSES;
    int c = TRACE_MEDIUM_CLASSES - 1;
    while (c > 0 && r >= cc[c]) { r -= cc[c]; --c; }
    return list[(long long)c * stride + r];
}
__device__ __forceinline__ int trace_level(int readLen, int refLen, int bw)
{
    const int width = min(2 * bw + 1, refLen);
    if (width <= TRACE_NARROW_WIDTH && (long long)readLen * width <= TRACE_NARROW_CELLS) return 0;
    if (width <= TRACE_NARROW_WIDTH && (long long)readLen * width <= TRACE_MEDIUM_CELLS &&
        (long long)readLen * ((width + 7) >> 3) * 4 <= TRACE_MEDIUM_REGION) return 3;
    return bw <= TRACE_WARP_BW ? 1 : 2;
}
constexpr int TRACE_WMAX = 4094;       // wide kernel: band diagonals held in shared memory (3 x int16 each)
constexpr int TRACE_SEQ_WORDS = 6144;  // wide kernel: packed codes of both sub-sequences staged in shared memory (48k bases)

// sequential reader of packed 4-bit codes: one 32-bit load per 8 bases
struct NibReader {
    const uint32_t *w;
    int idx;
    uint32_t cur;
    __device__ __forceinline__ NibReader(const uint4 *pool, uint32_t unit, int start)
        : w(reinterpret_cast<const uint32_t *>(pool + unit)), idx(start) { cur = w[idx >> 3]; }
    __device__ __forceinline__ int next()
    {
        const int c = (int)((cur >> ((idx & 7) * 4)) & 15u);
        ++idx;
        if ((idx & 7) == 0) cur = w[idx >> 3];
        return c;
    }
};

__device__ __forceinline__ bool trace_emit_single(const TraceArgs &a, PairState &s, db200_sw_result &res, uint32_t op)
{
    const unsigned long long pos = atomicAdd(a.cigar_cursor, 1ull);
    if (pos + 1 > a.cigar_cap) { atomicExch(a.overflow_flag, 1); s.status = DB200_PAIR_NOSCRATCH; return false; }
    a.cigar_pool[pos] = op;
    s.cigar_pos = (int64_t)pos;
    s.cigar_len = 1;
    res.cigar_len = 1;
    return true;
}

__global__ void __launch_bounds__(256) sw_trace_classify_kernel(TraceArgs a)
{
    __shared__ int16_t mat_s[36];
    if (threadIdx.x < 36) mat_s[threadIdx.x] = a.sc.mat[threadIdx.x];
    __syncthreads();
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= a.npairs) return;
    PairState &s = a.state[p];
    if (s.status != DB200_PAIR_OK || !s.want_cigar) return;
    db200_sw_result &res = a.results[p];
    const PairMeta pm = a.meta[p];
    if (res.score1 == 0) {
        // 1x1 rectangle with score 0: the reference's only observable output is 1M (ssw.c:844-847, 690-698)
        trace_emit_single(a, s, res, 16u);
        return;
    }
    const int refLen = res.ref_end1 - res.ref_begin1 + 1;    // columns j
    const int readLen = res.read_end1 - res.read_begin1 + 1; // rows i
    if (refLen == readLen) {
        // Gap-free shortcut.  If the main diagonal of the rectangle already collects the optimal score, every
        // diagonal cell holds exactly its prefix score (a larger value would lift the corner above the optimum),
        // so the reference's tie rule "diagonal wins when gaps do not beat it" (ssw.c:625) walks the diagonal and
        // the cigar is one run of readLen matches.  No band has to be evaluated.
        NibReader rq(a.pool, pm.q_unit, res.read_begin1), rt(a.pool, pm.t_unit, res.ref_begin1);
        int sum = 0;
        for (int i = 0; i < readLen; ++i) {
            const int qc = rq.next(), tc = rt.next();
            sum += mat_s[tc * 6 + qc];
        }
        if (sum == res.score1) { trace_emit_single(a, s, res, ((uint32_t)readLen) << 4); return; }
    }
    const int bw = abs(refLen - readLen) + 1;
    const int lvl = trace_level(readLen, refLen, bw);
    s.trace_bw = bw;
    trace_push(a.list, a.count, a.class_count, a.npairs, lvl, (int)p, readLen, refLen, min(2 * bw + 1, refLen));
}

// direction nibble: bits 0-1 = H source (0 diagonal, 1 vertical state E, 2 horizontal state F),
// bit 2 = E opened from H (code 3) else extended (2), bit 3 = F opened from H (code 5) else extended (4)
__device__ __forceinline__ int trace_step_code(int nib, int state)
{
    if (state == 2) { const int src = nib & 3; return src == 0 ? 1 : (src == 1 ? ((nib & 4) ? 3 : 2) : ((nib & 8) ? 5 : 4)); }
    if (state == 0) return (nib & 4) ? 3 : 2;
    return (nib & 8) ? 5 : 4;
}

// self = 0: jobs of list 0, directions in the thread's shared-memory slice; self = 3: jobs of list 3, directions in the slab
__global__ void __launch_bounds__(TRACE_NARROW_THREADS) sw_trace_narrow_kernel(TraceArgs a, int self)
{
    extern __shared__ uint32_t narrow_sm[];
    __shared__ int16_t mat_s[36];
    if (threadIdx.x < 36) mat_s[threadIdx.x] = a.sc.mat[threadIdx.x];
    __syncthreads();
    const int64_t li = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (li >= *a.count[self]) return;
    const int p = trace_task(a.list[self], a.class_count, a.npairs, self, li);
    PairState &s = a.state[p];
    db200_sw_result &res = a.results[p];
    const PairMeta pm = a.meta[p];
    const int refLen = res.ref_end1 - res.ref_begin1 + 1;
    const int readLen = res.read_end1 - res.read_begin1 + 1;
    const int score = res.score1;
    const int rbeg = res.read_begin1, tbeg = res.ref_begin1; // local copies: the byte stores below may alias `res`
    const int gapO = a.sc.gap_open, gapE = a.sc.gap_extend;
    int bw = self == 0 ? abs(refLen - readLen) + 1 : s.trace_bw;
    if (self == 3 && *a.count[3] < a.medium_min_jobs) {
        // few of these jobs: one thread each would only add its latency (~0.3 ms) to the batch - the warp-per-pair
        // kernel runs them side by side (what happens to the chunks of the host pipeline on the soft-clip workload)
        trace_push(a.list, a.count, a.class_count, a.npairs, bw <= TRACE_WARP_BW ? 1 : 2, p, readLen, refLen, 0);
        return;
    }
    uint32_t *slice = narrow_sm + threadIdx.x * (self == 3 ? TRACE_MEDIUM_WORDS : TRACE_NARROW_WORDS);
    uint8_t *dir = reinterpret_cast<uint8_t *>(slice);
    uint32_t *gdir = nullptr;  // medium jobs: direction nibbles in a private region of the slab, rows padded to whole words -
    int wpr = 0;               // one 32-bit store per eight cells (a byte per cell was one partial-sector write per cell)
    int width = 0;
    while (true) {
        width = min(2 * bw + 1, refLen);
        {
            int lvl = trace_level(readLen, refLen, bw);
            if (lvl == 3 && self == 3 && !gdir) {
                const unsigned long long region = atomicAdd(a.medium_cursor, 1ull);
                if ((region + 1ull) * (unsigned long long)TRACE_MEDIUM_REGION <= a.slab_total)
                    gdir = reinterpret_cast<uint32_t *>(a.slab + region * (unsigned long long)TRACE_MEDIUM_REGION);
                else lvl = bw <= TRACE_WARP_BW ? 1 : 2;   // no room left: a cooperative kernel takes the pair
            }
            if (lvl != self) { // the band outgrew this launch: hand the pair (and the band to try next) to a later one
                s.trace_bw = bw;
                trace_push(a.list, a.count, a.class_count, a.npairs, lvl, p, readLen, refLen, width);
                return;
            }
        }
        const bool use_g = self == 3;
        wpr = (width + 7) >> 3;
        // this attempt's direction nibbles and the two rolling rows of (H, E), band-relative, live in the thread's slice
        int16_t *Hp = reinterpret_cast<int16_t *>(slice + (self == 3 ? 0 : TRACE_NARROW_DIR_WORDS));
        int16_t *Ep = Hp + TRACE_NARROW_ROW, *Hc = Ep + TRACE_NARROW_ROW, *Ec = Hc + TRACE_NARROW_ROW;
        int best = 0;
        int pbeg = 0, pend = -1;
        for (int i = 0; i < readLen; ++i) {
            const int beg = max(0, i - bw), end = min(refLen - 1, i + bw);
            const int qc = pool_code(a.pool, pm.q_unit, rbeg + i);
            NibReader rt(a.pool, pm.t_unit, tbeg + beg);
            int f = 0, hleft = 0;
            uint32_t acc = 0;
            const unsigned long long rowbase = (unsigned long long)i * (unsigned long long)width;
            for (int j = beg; j <= end; ++j) {
                const bool up_in = (i > 0 && j >= pbeg && j <= pend);
                const bool dg_in = (i > 0 && j - 1 >= pbeg && j - 1 <= pend);
                const int hup = up_in ? Hp[j - pbeg] : 0, eup = up_in ? Ep[j - pbeg] : 0;
                const int hdg = dg_in ? Hp[j - 1 - pbeg] : 0;
                int t1 = (i == 0) ? -gapO : hup - gapO;
                int t2 = (i == 0) ? -gapE : eup - gapE;
                const int e = t1 > t2 ? t1 : t2;
                const int e_open = t1 > t2;
                t1 = hleft - gapO;
                t2 = f - gapE;
                f = t1 > t2 ? t1 : t2;
                const int f_open = t1 > t2;
                const int e1 = e > 0 ? e : 0, f1 = f > 0 ? f : 0;
                const int g = e1 > f1 ? e1 : f1;
                const int tc = rt.next();
                const int d = hdg + mat_s[tc * 6 + qc];
                const int h = g > d ? g : d;
                if (h > best) best = h;
                const int src = (g <= d) ? 0 : (e1 > f1 ? 1 : 2);
                const uint8_t nib = (uint8_t)(src | (e_open << 2) | (f_open << 3));
                const unsigned long long cell = rowbase + (unsigned long long)(j - beg);
                if (use_g) {
                    const int c = j - beg;
                    acc |= (uint32_t)nib << ((c & 7) * 4);
                    if ((c & 7) == 7 || j == end) { gdir[i * wpr + (c >> 3)] = acc; acc = 0; }
                } else {
                    uint8_t &byte = dir[cell >> 1];
                    byte = (cell & 1) ? (uint8_t)((byte & 0x0F) | (nib << 4)) : (uint8_t)((byte & 0xF0) | nib);
                }
                Hc[j - beg] = (int16_t)h; Ec[j - beg] = (int16_t)max(e, -32768);
                hleft = h;
            }
            int16_t *t = Hp; Hp = Hc; Hc = t;
            t = Ep; Ep = Ec; Ec = t;
            pbeg = beg; pend = end;
        }
        if (best >= score) break;
        bw *= 2;
    }
    // ---- walk twice: first count the runs, then write them in final (reversed) order ------------------
    int ncig = 0;
    unsigned long long cpos = 0;
    const bool walk_g = self == 3;
    for (int phase = 0; phase < 2; ++phase) {
        int i = readLen - 1, j = refLen - 1;
        int run = 0, prev = 0, op = 0, state = 2; // state: 0 = in E, 1 = in F, 2 = in H
        int l = 0;
        bool bad = false;
        while (i > 0) {
            const int beg = max(0, i - bw);
            const int x = j - beg;
            if (x < 0 || x >= width || j < 0) { bad = true; break; } // the reference would read out of bounds here
            const unsigned long long cell = (unsigned long long)i * (unsigned long long)width + (unsigned long long)x;
            const int nib = walk_g ? (int)((gdir[i * wpr + (x >> 3)] >> ((x & 7) * 4)) & 15u) : ((dir[cell >> 1] >> ((cell & 1) * 4)) & 15);
            switch (trace_step_code(nib, state)) {
                case 1: --i; --j; state = 2; op = 0; break;
                case 2: --i; state = 0; op = 1; break;
                case 3: --i; state = 2; op = 1; break;
                case 4: --j; state = 1; op = 2; break;
                default: --j; state = 2; op = 2; break;
            }
            if (op == prev) ++run;
            else {
                if (phase == 1) a.cigar_pool[cpos + (unsigned long long)(ncig - 1 - l)] = ((uint32_t)run << 4) | (uint32_t)prev;
                ++l;
                prev = op;
                run = 1;
            }
        }
        if (bad) { s.status = DB200_PAIR_UNSUPPORTED; res.cigar_len = 0; return; }
        if (op == 0) {
            if (phase == 1) a.cigar_pool[cpos + (unsigned long long)(ncig - 1 - l)] = ((uint32_t)(run + 1)) << 4;
            ++l;
        } else {
            if (phase == 1) {
                a.cigar_pool[cpos + (unsigned long long)(ncig - 1 - l)] = ((uint32_t)run << 4) | (uint32_t)op;
                a.cigar_pool[cpos + (unsigned long long)(ncig - 2 - l)] = 16u;
            }
            l += 2;
        }
        if (phase == 0) {
            ncig = l;
            cpos = atomicAdd(a.cigar_cursor, (unsigned long long)ncig);
            if (cpos + (unsigned long long)ncig > a.cigar_cap) { atomicExch(a.overflow_flag, 1); s.status = DB200_PAIR_NOSCRATCH; return; }
        }
    }
    s.cigar_pos = (int64_t)cpos;
    s.cigar_len = ncig;
    res.cigar_len = ncig;
}

// Block per pair.  Cells of one anti-diagonal (i + j = const) are independent; with band diagonal x = j - i + OFF the
// three neighbours of (i, x) are (i-1, x+1) [vertical], (i, x-1) [horizontal] - both on the previous anti-diagonal -
// and (i-1, x) [diagonal] two anti-diagonals back, so one value per band diagonal (H, E, F as int16 in shared memory)
// is all the state the sweep needs.  Slots outside the matrix are never written and stay 0, slots outside the band
// are the zeroed guard entries, which is exactly the reference's "cells outside the band read as 0".
template <int T> __device__ __forceinline__ void coop_sync() { if (T == 32) __syncwarp(); else __syncthreads(); }

// ---- TMA bulk copy (cp.async.bulk, SASS: UBLKCP) global -> shared, completion on an mbarrier -------------------
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase)
{
    uint32_t done;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done)
                     : "r"(smem_addr(bar)), "r"(phase)
                     : "memory");
    } while (!done);
}

// T = 32: level 1 (one warp per pair, __syncwarp between anti-diagonals); T = 128: level 2 (one block per pair)
template <int T, int LEVEL, int WMAX, int SEQW, int STAGEB>
__global__ void __launch_bounds__(T) sw_trace_coop_kernel(TraceArgs a)
{
    extern __shared__ int16_t sm[];   // [3][WMAX + 2] band state, SEQW words of packed codes, STAGEB bytes of walk staging
    __shared__ int16_t mat_s[36];
    __shared__ int sh_task, sh_best, sh_ncig;
    __shared__ unsigned long long sh_cpos;
    __shared__ __align__(8) uint64_t tma_bar;   // completion barrier of the bulk copies below
    uint32_t tma_phase = 0;
    int16_t *Hs = sm + 1, *Es = sm + (WMAX + 2) + 1, *Fs = sm + 2 * (WMAX + 2) + 1; // index -1 .. XW are valid
    for (int k = threadIdx.x; k < 36; k += T) mat_s[k] = a.sc.mat[k];
    if (threadIdx.x == 0) mbar_init(&tma_bar, 1);
    coop_sync<T>();
    // the slab area is divided among the blocks that can have work: few pairs -> few, large slabs
    const int nslabs = min((int)gridDim.x, max(*a.count[LEVEL], 1));
    if ((int)blockIdx.x >= nslabs) return;
    const unsigned long long slab_bytes = (a.slab_total / (unsigned long long)nslabs) & ~255ull;
    uint8_t *slab = a.slab + (unsigned long long)blockIdx.x * slab_bytes;
    const int gapO = a.sc.gap_open, gapE = a.sc.gap_extend;

    while (true) {
        if (threadIdx.x == 0) sh_task = atomicAdd(a.counter[LEVEL], 1);
        coop_sync<T>();
        const int ti = sh_task;
        coop_sync<T>();
        if (ti >= *a.count[LEVEL]) break;
        if (threadIdx.x == 0) sh_task = trace_task(a.list[LEVEL], a.class_count, a.npairs, LEVEL, ti);
        coop_sync<T>();
        const int p = sh_task;
        coop_sync<T>();
        PairState &s = a.state[p];
        db200_sw_result &res = a.results[p];
        const PairMeta pm = a.meta[p];
        const int refLen = res.ref_end1 - res.ref_begin1 + 1;
        const int readLen = res.read_end1 - res.read_begin1 + 1;
        const int score = res.score1;
        const int rbeg = res.read_begin1, tbeg = res.ref_begin1; // local copies: the byte stores below may alias `res`
        const uint32_t *qw = reinterpret_cast<const uint32_t *>(a.pool + pm.q_unit);
        const uint32_t *tw = reinterpret_cast<const uint32_t *>(a.pool + pm.t_unit);
        // stage the codes of both sub-sequences in shared memory (the large carve-out leaves little L1 for them)
        {
            uint32_t *seqs = reinterpret_cast<uint32_t *>(sm + 3 * (WMAX + 2));
            // 16-byte aligned word ranges covering the two sub-sequences (sequences start on 16-byte units)
            const int qw0 = (rbeg >> 3) & ~3, qwn = ((((rbeg + readLen - 1) >> 3) - qw0 + 1) + 3) & ~3;
            const int tw0 = (tbeg >> 3) & ~3, twn = ((((tbeg + refLen - 1) >> 3) - tw0 + 1) + 3) & ~3;
            if (SEQW > 0 && qwn + twn <= SEQW) {
                coop_sync<T>();   // everybody is done with the previous pair's codes
                if (threadIdx.x == 0) {
                    fence_proxy_async();
                    mbar_expect_tx(&tma_bar, (uint32_t)(qwn + twn) * 4u);
                    tma_bulk_g2s(seqs, qw + qw0, (uint32_t)qwn * 4u, &tma_bar);
                    tma_bulk_g2s(seqs + qwn, tw + tw0, (uint32_t)twn * 4u, &tma_bar);
                }
                mbar_wait(&tma_bar, tma_phase);
                tma_phase ^= 1u;
                qw = seqs - qw0;           // same indexing as the global arrays
                tw = seqs + qwn - tw0;
            }
        }
        int bw = s.trace_bw;   // first band, or the one a lower level could not hold
        int wst = 0;
        bool failed = false, escalate = false;
        while (true) {
            if (LEVEL == 1 && bw > TRACE_WARP_BW) { escalate = true; break; }
            const int OFF = min(bw, readLen - 1);
            const int XW = OFF + min(bw, refLen - 1) + 1;            // band diagonals with at least one cell
            // Direction codes are stored anti-diagonal by anti-diagonal, one byte per cell at half its band diagonal (the
            // cells of an anti-diagonal sit on every other band diagonal): the threads of a sweep step write
            // neighbouring bytes - one or two 32-byte sectors per warp store, where row-major storage touched 32.
            // (Two cells per byte, paired up by a shuffle, measured slower: long contigs 44.7 -> 51.8 ms.)
            wst = ((XW >> 1) + 1 + 15) & ~15;                         // 16-byte aligned
            const unsigned long long dir_bytes = (unsigned long long)(readLen + refLen - 1) * (unsigned long long)wst;
            const unsigned long long runs_bytes = (unsigned long long)(readLen + refLen + 4) * 4ull;
            if (XW > WMAX || dir_bytes + runs_bytes + 64 > slab_bytes) {
                if (LEVEL == 1) escalate = true; else failed = true;   // level 2 has the larger slabs
                break;
            }
            for (int x = threadIdx.x - 1; x <= XW; x += T) { Hs[x] = 0; Es[x] = 0; Fs[x] = 0; }
            if (threadIdx.x == 0) sh_best = 0;
            coop_sync<T>();
            int best = 0;
            const int last_tau = readLen + refLen - 2;
            for (int tau = 0; tau <= last_tau; ++tau) {
                // rows on this anti-diagonal: 0 <= j = tau - i < refLen and |j - i| <= bw
                const int ilo = max(max(0, tau - refLen + 1), (tau - bw + 1) >> 1);
                const int ihi = min(min(readLen - 1, tau), (tau + bw) >> 1);
                for (int i = ilo + (int)threadIdx.x; i <= ihi; i += T) {
                    const int j = tau - i;
                    const int x = j - i + OFF;
                    const int hup = Hs[x + 1], eup = Es[x + 1];
                    const int hleft = Hs[x - 1], fleft = Fs[x - 1];
                    const int hdg = Hs[x];
                    int t1 = (i == 0) ? -gapO : hup - gapO;
                    int t2 = (i == 0) ? -gapE : eup - gapE;
                    const int e = t1 > t2 ? t1 : t2;
                    const int e_open = t1 > t2;
                    t1 = hleft - gapO;
                    t2 = fleft - gapE;
                    const int f = t1 > t2 ? t1 : t2;
                    const int f_open = t1 > t2;
                    const int e1 = e > 0 ? e : 0, f1 = f > 0 ? f : 0;
                    const int g = e1 > f1 ? e1 : f1;
                    const int qi = rbeg + i, tj = tbeg + j;
                    const int qc = (int)((qw[qi >> 3] >> ((qi & 7) * 4)) & 15u);
                    const int tc = (int)((tw[tj >> 3] >> ((tj & 7) * 4)) & 15u);
                    const int d = hdg + mat_s[tc * 6 + qc];
                    const int h = g > d ? g : d;
                    best = max(best, h);
                    const int src = (g <= d) ? 0 : (e1 > f1 ? 1 : 2);
                    const uint8_t nib = (uint8_t)(src | (e_open << 2) | (f_open << 3));
#ifndef DB200_TRACE_NOSTORE
                    slab[(unsigned long long)tau * (unsigned long long)wst + (unsigned long long)(x >> 1)] = nib;
#else
                    if (h == -12345) slab[0] = nib;
#endif
                    Hs[x] = (int16_t)h; Es[x] = (int16_t)e; Fs[x] = (int16_t)f;
                }
                coop_sync<T>();
            }
            atomicMax(&sh_best, best);
            coop_sync<T>();
            const int b = sh_best;
            coop_sync<T>();
            if (b >= score) break;
            bw *= 2;
        }
        if (escalate) {
            if (threadIdx.x == 0) { s.trace_bw = bw; trace_push(a.list, a.count, a.class_count, a.npairs, 2, p, readLen, refLen, 0); }
            coop_sync<T>();
            continue;
        }
        if (failed) {
            if (threadIdx.x == 0 && a.debug)
                printf("[db200] traceback level %d: pair %d rows %d cols %d band %d does not fit (slab %llu bytes, %d slabs)\n", LEVEL, p,
                       readLen, refLen, bw, slab_bytes, nslabs);
            if (threadIdx.x == 0) { atomicExch(a.overflow_flag, 2); s.status = DB200_PAIR_NOSCRATCH; }
            coop_sync<T>();
            continue;
        }
        // ---- walk: anti-diagonals are staged through shared memory in chunks (one bulk copy each), one thread follows
        //      the directions inside the chunk; runs are collected behind the direction bytes, then copied reversed
        const unsigned long long dir_total = (unsigned long long)(readLen + refLen - 1) * (unsigned long long)wst;
        uint32_t *runs = reinterpret_cast<uint32_t *>(slab + ((dir_total + 15) & ~15ull));
        {
            uint8_t *stage = reinterpret_cast<uint8_t *>(sm) + (((3u * (WMAX + 2) * sizeof(int16_t) + SEQW * 4u) + 15u) & ~15u);
            const int stage_taus = max(2, (int)((unsigned)STAGEB / (unsigned)wst));
            const int OFFW = min(bw, readLen - 1);
            const int width = min(2 * bw + 1, refLen);
            __shared__ int w_i, w_j, w_run, w_prev, w_op, w_state, w_l, w_bad;
            if (threadIdx.x == 0) { w_i = readLen - 1; w_j = refLen - 1; w_run = 0; w_prev = 0; w_op = 0; w_state = 2; w_l = 0; w_bad = 0; }
            coop_sync<T>();
            while (true) {
                if (w_i <= 0 || w_bad) break;
                const int tau_hi = w_i + w_j;
                const int tau_lo = max(0, tau_hi - stage_taus + 1);
                // one bulk copy (TMA) brings the anti-diagonals of this chunk; the direction bytes were written with
                // ordinary stores by this block, hence the proxy fence after the block-wide barrier
                if (threadIdx.x == 0) {
                    const uint32_t bytes = (uint32_t)(tau_hi - tau_lo + 1) * (uint32_t)wst;
                    fence_proxy_async();
                    mbar_expect_tx(&tma_bar, bytes);
                    tma_bulk_g2s(stage, slab + (unsigned long long)tau_lo * (unsigned long long)wst, bytes, &tma_bar);
                }
                mbar_wait(&tma_bar, tma_phase);
                tma_phase ^= 1u;
                if (threadIdx.x == 0) {
                    int i = w_i, j = w_j, run = w_run, prev = w_prev, op = w_op, state = w_state, l = w_l;
                    while (i > 0 && i + j >= tau_lo) {
                        const int x = j - max(0, i - bw);
                        if (x < 0 || x >= width || j < 0) { w_bad = 1; break; } // the reference would read out of bounds here
                        if (j - i > bw) { w_bad = 1; break; }   // a cell the sweep never evaluated
                        const int nib = stage[(i + j - tau_lo) * wst + ((j - i + OFFW) >> 1)];
                        switch (trace_step_code(nib, state)) {
                            case 1: --i; --j; state = 2; op = 0; break;
                            case 2: --i; state = 0; op = 1; break;
                            case 3: --i; state = 2; op = 1; break;
                            case 4: --j; state = 1; op = 2; break;
                            default: --j; state = 2; op = 2; break;
                        }
                        if (op == prev) ++run;
                        else { runs[l++] = ((uint32_t)run << 4) | (uint32_t)prev; prev = op; run = 1; }
                    }
                    w_i = i; w_j = j; w_run = run; w_prev = prev; w_op = op; w_state = state; w_l = l;
                }
                coop_sync<T>();
            }
            if (threadIdx.x == 0) {
                int l = w_l;
                if (w_bad) { s.status = DB200_PAIR_UNSUPPORTED; res.cigar_len = 0; sh_ncig = -1; }
                else {
                    if (w_op == 0) runs[l++] = ((uint32_t)(w_run + 1)) << 4;
                    else { runs[l++] = ((uint32_t)w_run << 4) | (uint32_t)w_op; runs[l++] = 16u; }
                    const unsigned long long cpos = atomicAdd(a.cigar_cursor, (unsigned long long)l);
                    if (cpos + (unsigned long long)l > a.cigar_cap) { atomicExch(a.overflow_flag, 1); s.status = DB200_PAIR_NOSCRATCH; sh_ncig = -1; }
                    else { sh_ncig = l; sh_cpos = cpos; s.cigar_pos = (int64_t)cpos; s.cigar_len = l; res.cigar_len = l; }
                }
            }
        }
        coop_sync<T>();
        const int ncig = sh_ncig;
        if (ncig > 0) {
            const unsigned long long cpos = sh_cpos;
            for (int k = threadIdx.x; k < ncig; k += T) a.cigar_pool[cpos + (unsigned long long)k] = runs[ncig - 1 - k];
        }
        coop_sync<T>();
    }
}

}  // namespace db200
