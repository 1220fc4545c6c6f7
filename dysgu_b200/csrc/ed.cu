// Batched edit-distance pipeline on the device: alphabet / Peq build -> bucket -> Myers forward pass ->
// best score + end locations -> (HW, task >= locations) reverse passes for the start locations.
//
// Reference behaviour being reproduced (dysgu v1.9.0): edlib.pyx:57-156 (wrapper), edlib.cpp:141-296
// (driver: alphabet, empty sequences, k handling, start locations), 355-381 (Peq), 546-709, 735-944.
#include "ed_kernels.cuh"
#include "ed_path.cuh"

namespace db200 {

// configuration -> (words per thread, threads per pair); capacity in 64-bit words of the query = WL * G.  Steps of at
// most 1.33x above 32 words (one pair per warp), so a 10 kb query (157 words) runs in the 160-word configuration
// instead of a 256-word one with 39 % idle lanes.
static const int ED_CFG_WL[ED_NCFG] = {1, 2, 4, 4, 4, 4, 4, 4, 4, 3, 4, 3, 4, 5, 6, 8, 16, 32};
static const int ED_CFG_G[ED_NCFG] = {1, 1, 1, 2, 3, 4, 6, 8, 10, 16, 16, 32, 32, 32, 32, 32, 32, 32};
__device__ __constant__ int d_ed_cfg_words[ED_NCFG] = {1, 2, 4, 8, 12, 16, 24, 32, 40, 48, 64, 96, 128, 160, 192, 256, 512, 1024};

__device__ __forceinline__ int ed_cfg_for(int W)
{
    int c = 0;
    while (c < ED_NCFG - 1 && d_ed_cfg_words[c] < W) ++c;
    return c;
}

struct EdPrepArgs {
    const uint8_t *qbuf, *tbuf;
    const int64_t *qoff, *toff;
    const int32_t *k;      // may be null
    int32_t k_default;
    int32_t mode, task;
    int64_t npairs;
    EdPair *pairs;
    int32_t *peq_words;    // per pair, for the scan
    int32_t *hlog_words;
    int32_t *hist;         // [ED_NCFG * ED_LBINS]
    db200_ed_result *results;
    uint32_t *tsets;       // [npairs][8] symbol set of every target (reused by ed_build_kernel)
    int32_t band;          // NW: route pairs through the banded rounds (segments = rounds instead of configurations)
    int32_t *fb_tasks;     // NW + band: pairs that go straight to the full-column kernel, per configuration
    int32_t *fb_count;
    int64_t fb_stride;
};

// 256-bit symbol set of a byte range, accumulated per lane in registers (no atomics: DNA has four symbols, every lane
// would hit the same shared-memory word) and OR-reduced over the warp; lane l < 8 returns word l of the set.
__device__ __forceinline__ uint32_t symbol_set_word(const uint8_t *s, int len, int lane)
{
    uint64_t b0 = 0, b1 = 0, b2 = 0, b3 = 0;
    auto add = [&](uint32_t c) {
        const uint64_t bit = 1ull << (c & 63u);
        const uint32_t hi = c >> 6;
        b0 |= hi == 0 ? bit : 0ull;
        b1 |= hi == 1 ? bit : 0ull;
        b2 |= hi == 2 ? bit : 0ull;
        b3 |= hi == 3 ? bit : 0ull;
    };
    // bytes up to the first word boundary, then four symbols per 32-bit load, then the tail
    const int head = min(len, (int)((4u - (unsigned)(reinterpret_cast<uintptr_t>(s) & 3u)) & 3u));
    const int nw4 = (len - head) >> 2;
    if (lane < head) add(s[lane]);
    const uint32_t *s4 = reinterpret_cast<const uint32_t *>(s + head);
    for (int i = lane; i < nw4; i += 32) {
        const uint32_t x = s4[i];
        add(x & 0xFFu); add((x >> 8) & 0xFFu); add((x >> 16) & 0xFFu); add(x >> 24);
    }
    const int tail0 = head + 4 * nw4;
    if (tail0 + lane < len) add(s[tail0 + lane]);
    uint32_t w[8] = {(uint32_t)b0, (uint32_t)(b0 >> 32), (uint32_t)b1, (uint32_t)(b1 >> 32),
                     (uint32_t)b2, (uint32_t)(b2 >> 32), (uint32_t)b3, (uint32_t)(b3 >> 32)};
    uint32_t mine = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t r = __reduce_or_sync(0xFFFFFFFFu, w[k]);
        if (lane == k) mine = r;
    }
    return mine;
}

// one warp per pair: symbol sets, sizes, special cases (edlib.cpp:153-179)
__global__ void ed_prep_kernel(EdPrepArgs a)
{
    const int lane = threadIdx.x & 31;
    const int64_t p = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (p >= a.npairs) return;
    const int64_t q0 = a.qoff[p], t0 = a.toff[p];
    const int m = (int)(a.qoff[p + 1] - q0), n = (int)(a.toff[p + 1] - t0);
    const uint32_t qs = symbol_set_word(a.qbuf + q0, m, lane);
    const uint32_t ts = symbol_set_word(a.tbuf + t0, n, lane);
    if (lane < 8) a.tsets[p * 8 + lane] = ts;
    int au = 0, at = 0;
    if (lane < 8) { au = __popc(qs | ts); at = __popc(ts); }
    for (int o = 4; o >= 1; o >>= 1) { au += __shfl_xor_sync(0xFFFFFFFFu, au, o); at += __shfl_xor_sync(0xFFFFFFFFu, at, o); }
    au = __shfl_sync(0xFFFFFFFFu, au, 0);
    at = __shfl_sync(0xFFFFFFFFu, at, 0);
    if (lane != 0) return;
    EdPair pr;
    memset(&pr, 0, sizeof(pr));
    pr.q_off = q0; pr.t_off = t0; pr.m = m; pr.n = n;
    pr.W = (m + 63) >> 6;
    pr.At = at;
    pr.alphabet = au;
    pr.k = a.k ? a.k[p] : a.k_default;
    pr.best = -1;
    pr.cfg = ed_cfg_for(pr.W);
    db200_ed_result r;
    r.status = 0; r.edit_distance = -1; r.alphabet_length = au; r.num_locations = 0;
    int pw = 0, hw = 0;
    if (m == 0 || n == 0) {
        pr.special = 1;
        pr.nloc = 1;
        if (a.mode == DB200_ED_MODE_NW) pr.best = max(m, n);
        else if (a.mode == DB200_ED_MODE_SHW || a.mode == DB200_ED_MODE_HW) pr.best = m;
        else { pr.nloc = 0; r.status = 1; }
        r.edit_distance = pr.best;
        r.num_locations = pr.nloc;
    } else if (a.mode != DB200_ED_MODE_NW && a.mode != DB200_ED_MODE_SHW && a.mode != DB200_ED_MODE_HW) {
        pr.special = 1; r.status = 1;
    } else if (pr.W > d_ed_cfg_words[ED_NCFG - 1]) {
        pr.special = 1; r.status = 1; // queries beyond 65536 symbols are not provided by the device path
    } else {
        const bool need_rev = (a.mode == DB200_ED_MODE_HW && a.task >= DB200_ED_TASK_LOC);
        pw = at * pr.W * (need_rev ? 2 : 1);
        hw = (a.mode != DB200_ED_MODE_NW) ? ((n + 15) >> 4) : 0;
        const int lb = min(n >> 6, ED_LBINS - 1);
        pr.band_round = -1;
        if (a.band) {
            // first banded round whose cost bound 64 * (words - 1) is not already excluded by the length difference
            const int absd = abs(n - m);
            for (int r = 0; r < ED_BAND_ROUNDS; ++r)
                if (pr.W <= ed_band_words(r) || 64 * (ed_band_words(r) - 1) >= absd) { pr.band_round = r; break; }
            if (pr.band_round >= 0) atomicAdd(&a.hist[pr.band_round * ED_LBINS + (ED_LBINS - 1 - lb)], 1);
            else a.fb_tasks[(int64_t)pr.cfg * a.fb_stride + atomicAdd(&a.fb_count[pr.cfg], 1)] = (int32_t)p;
        } else {
            atomicAdd(&a.hist[pr.cfg * ED_LBINS + (ED_LBINS - 1 - lb)], 1);
        }
    }
    a.pairs[p] = pr;
    a.peq_words[p] = pw;
    a.hlog_words[p] = hw;
    a.results[p] = r;
}

struct EdBuildArgs {
    const uint8_t *qbuf, *tbuf;
    EdPair *pairs;
    const int64_t *peq_off, *hlog_off;
    int64_t npairs;
    uint64_t *peq;
    uint8_t *tcodes;
    int32_t mode, task;
    const uint8_t *eq_pairs;   // device copy of (first, second) bytes
    int32_t n_eq_pairs;
    const int32_t *bin_start;
    int32_t *bin_fill;
    int32_t *tasks;
    int32_t band;
    const uint32_t *tsets;
};

// one warp per pair: target symbols -> dense row ids, re-coded target, match vectors of the query
// (forward and, for HW start locations, reversed), bucket scatter.  edlib.cpp:355-381, 1433-1472, 60-91.
__global__ void ed_build_kernel(EdBuildArgs a)
{
    __shared__ uint8_t rank_s[8][256];
    __shared__ uint32_t tset[8][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t p = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (p >= a.npairs) return;
    EdPair pr = a.pairs[p];
    if (pr.special) return;
    pr.peq_off = a.peq_off[p];
    pr.hlog_off = a.hlog_off[p];
    const uint8_t *q = a.qbuf + pr.q_off, *t = a.tbuf + pr.t_off;
    if (lane < 8) tset[warp][lane] = a.tsets[p * 8 + lane];
    __syncwarp();
    // rank of every byte among the target's symbols (0xFF: not a target symbol)
    for (int c = lane; c < 256; c += 32) {
        const uint32_t wbits = tset[warp][c >> 5];
        int r = 0xFF;
        if ((wbits >> (c & 31)) & 1u) {
            r = __popc(wbits & ((1u << (c & 31)) - 1u));
            for (int w = 0; w < (c >> 5); ++w) r += __popc(tset[warp][w]);
        }
        rank_s[warp][c] = (uint8_t)r;
    }
    __syncwarp();
    // re-coded target: four symbols per word where source and destination (same offset) are word-aligned
    uint8_t *tc = a.tcodes + pr.t_off;
    {
        const int head = min(pr.n, (int)((4u - (unsigned)(reinterpret_cast<uintptr_t>(t) & 3u)) & 3u));
        const int nw4 = (pr.n - head) >> 2;
        if (lane < head) tc[lane] = rank_s[warp][t[lane]];
        const uint32_t *t4 = reinterpret_cast<const uint32_t *>(t + head);
        uint32_t *tc4 = reinterpret_cast<uint32_t *>(tc + head);
        for (int i = lane; i < nw4; i += 32) {
            const uint32_t x = t4[i];
            tc4[i] = (uint32_t)rank_s[warp][x & 0xFFu] | ((uint32_t)rank_s[warp][(x >> 8) & 0xFFu] << 8) |
                     ((uint32_t)rank_s[warp][(x >> 16) & 0xFFu] << 16) | ((uint32_t)rank_s[warp][x >> 24] << 24);
        }
        const int tail0 = head + 4 * nw4;
        if (tail0 + lane < pr.n) tc[tail0 + lane] = rank_s[warp][t[tail0 + lane]];
    }
    const bool need_rev = (a.mode == DB200_ED_MODE_HW && a.task >= DB200_ED_TASK_LOC);
    const int W = pr.W, m = pr.m, At = pr.At;
    uint64_t *peq = a.peq + pr.peq_off;
    // match vectors: the warp takes one 64-row word at a time (lane l holds rows l and l + 32) and one ballot pair per
    // target symbol yields that symbol's word - every entry is written exactly once, nothing is read back
    for (int dir = 0; dir < (need_rev ? 2 : 1); ++dir) {
        uint64_t *tab = peq + (int64_t)dir * At * W;
        for (int w = 0; w < W; ++w) {
            const int i_lo = w * 64 + lane, i_hi = i_lo + 32;
            const int c_lo = i_lo < m ? (int)(dir ? q[m - 1 - i_lo] : q[i_lo]) : -1;
            const int c_hi = i_hi < m ? (int)(dir ? q[m - 1 - i_hi] : q[i_hi]) : -1;
            const int r_lo = c_lo >= 0 ? (int)rank_s[warp][c_lo] : 0x1FF;
            const int r_hi = c_hi >= 0 ? (int)rank_s[warp][c_hi] : 0x1FF;
            for (int r = 0; r < At; ++r) {
                bool h_lo = r_lo == r, h_hi = r_hi == r;
                for (int e = 0; e < a.n_eq_pairs; ++e) {   // additionalEqualities: symmetric, not transitive
                    const int x = a.eq_pairs[2 * e], y = a.eq_pairs[2 * e + 1];
                    const bool xr = rank_s[warp][x] == r, yr = rank_s[warp][y] == r;
                    h_lo = h_lo || (c_lo == x && yr) || (c_lo == y && xr);
                    h_hi = h_hi || (c_hi == x && yr) || (c_hi == y && xr);
                }
                const uint32_t lo = __ballot_sync(0xFFFFFFFFu, h_lo), hi = __ballot_sync(0xFFFFFFFFu, h_hi);
                if (lane == 0) tab[(int64_t)r * W + w] = (uint64_t)lo | ((uint64_t)hi << 32);
            }
        }
    }
    if (lane == 0) {
        a.pairs[p] = pr;
        const int lb = min(pr.n >> 6, ED_LBINS - 1);
        const int seg = a.band ? pr.band_round : pr.cfg;   // banded NW: segments are rounds; round -1 sits in the fallback lists already
        if (seg >= 0) {
            const int key = seg * ED_LBINS + (ED_LBINS - 1 - lb);
            const int pos = a.bin_start[key] + atomicAdd(&a.bin_fill[key], 1);
            a.tasks[pos] = (int32_t)p;
        }
    }
}

__global__ void ed_binscan_kernel(const int32_t *hist, int32_t *bin_start, int32_t *seg_begin, int32_t *seg_count)
{
    __shared__ int32_t tot[ED_NCFG];
    const int tid = threadIdx.x;
    if (tid < ED_NCFG) {
        int s = 0;
        for (int b = 0; b < ED_LBINS; ++b) s += hist[tid * ED_LBINS + b];
        tot[tid] = s;
    }
    __syncthreads();
    if (tid < ED_NCFG) {
        int acc = 0;
        for (int c = 0; c < tid; ++c) acc += tot[c];
        seg_begin[tid] = acc;
        seg_count[tid] = tot[tid];
        for (int b = 0; b < ED_LBINS; ++b) { bin_start[tid * ED_LBINS + b] = acc; acc += hist[tid * ED_LBINS + b]; }
    }
}

struct EdCountArgs {
    EdPair *pairs;
    const uint32_t *hlog;
    const int32_t *nw_score;
    db200_ed_result *results;
    int32_t *nloc;
    int32_t *rev_count;   // [ED_NCFG] start-location tasks per config
    int64_t npairs;
    int32_t mode, task;
};

// thread per pair: best score over the last row and the number of end locations (edlib.cpp:663-698, 930-938)
__global__ void ed_count_kernel(EdCountArgs a)
{
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= a.npairs) return;
    EdPair &pr = a.pairs[p];
    if (pr.special) { a.nloc[p] = pr.nloc; return; }
    int best = -1, cnt = 0, ntask = 0;
    if (a.mode == DB200_ED_MODE_NW) {
        const int d = a.nw_score[p];
        if (pr.k < 0 || d <= pr.k) { best = d; cnt = 1; }
    } else {
        const uint32_t *hl = a.hlog + pr.hlog_off;
        const bool col_m1 = (pr.m & 63) != 0;
        int score = pr.m;
        int mn = col_m1 ? pr.m : 0x7FFFFFFF;
        int c = col_m1 ? 1 : 0;
        uint32_t word = 0;
        for (int j = 0; j < pr.n; ++j) {
            if ((j & 15) == 0) word = hl[j >> 4];
            const uint32_t d = (word >> ((j & 15) * 2)) & 3u;
            score += (int)(d & 1u) - (int)(d >> 1);
            if (score < mn) { mn = score; c = 1; }
            else if (score == mn) ++c;
        }
        if (pr.k < 0 || mn <= pr.k) {
            best = mn; cnt = c;
            ntask = cnt - ((col_m1 && mn == pr.m) ? 1 : 0); // the -1 end location needs no reverse pass
        }
    }
    pr.best = best;
    pr.nloc = cnt;
    a.nloc[p] = cnt;
    a.results[p].edit_distance = best;
    a.results[p].num_locations = cnt;
    if (a.mode == DB200_ED_MODE_HW && a.task >= DB200_ED_TASK_LOC && ntask > 0) atomicAdd(&a.rev_count[pr.cfg], ntask);
}

__global__ void ed_revscan_kernel(const int32_t *rev_count, int32_t *rev_begin)
{
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int c = 0; c < ED_NCFG; ++c) { rev_begin[c] = acc; acc += rev_count[c]; }
    }
}

struct EdEmitArgs {
    EdPair *pairs;
    const uint32_t *hlog;
    const int64_t *loc_off;
    int32_t *loc_buf;
    int64_t loc_cap;
    int32_t *loc_pair;
    const int32_t *rev_begin;
    int32_t *rev_fill;
    int32_t *rev_tasks;
    int32_t *overflow;
    int64_t npairs;
    int32_t mode, task;
};

// thread per pair: end locations in ascending order; start locations that need no DP (edlib.cpp:214-267)
__global__ void ed_emit_kernel(EdEmitArgs a)
{
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xFFFFFFFFu;
    const bool want_rev = (a.mode == DB200_ED_MODE_HW && a.task >= DB200_ED_TASK_LOC);
    // every lane takes part in the reservation below, so nothing returns before it
    bool live = p < a.npairs;
    EdPair *prp = live ? &a.pairs[p] : nullptr;
    int64_t o = 0;
    if (live) {
        o = a.loc_off[p];
        prp->loc_off = o;
        if (prp->nloc <= 0) live = false;
        else if (o + prp->nloc > a.loc_cap) { atomicExch(a.overflow, 1); live = false; }
    }
    const bool col_m1 = live && !prp->special && (prp->m & 63) != 0 && prp->m == prp->best;   // the -1 end location needs no start search
    const int ntask = (live && want_rev && !prp->special) ? prp->nloc - (col_m1 ? 1 : 0) : 0;
    // slots in the start-location task list of the pair's configuration: one atomic per warp and configuration
    // (a slot per location used to cost one returning atomic each on a handful of addresses)
    int slot = 0;
    {
        const int key = ntask > 0 ? prp->cfg : -1;
        unsigned todo = __ballot_sync(FULL, key >= 0);
        while (todo) {
            const int leader = __ffs(todo) - 1;
            const int k = __shfl_sync(FULL, key, leader);
            const bool mine = key == k;
            const int v = mine ? ntask : 0;
            int inc = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(FULL, inc, d); if (lane >= d) inc += t; }
            const int total = __shfl_sync(FULL, inc, 31);
            int b = 0;
            if (lane == leader) b = atomicAdd(&a.rev_fill[k], total);
            b = __shfl_sync(FULL, b, leader);
            if (mine) slot = a.rev_begin[k] + b + inc - v;
            todo &= ~__ballot_sync(FULL, mine);
        }
    }
    if (!live) return;
    EdPair &pr = *prp;
    int32_t *out = a.loc_buf + 2 * o;
    if (pr.special) {
        out[0] = DB200_ED_NO_START;
        out[1] = (a.mode == DB200_ED_MODE_NW) ? pr.n - 1 : -1;
        return;
    }
    const int start_fixed = (a.task >= DB200_ED_TASK_LOC) ? 0 : DB200_ED_NO_START;
    if (a.mode == DB200_ED_MODE_NW) { out[0] = start_fixed; out[1] = pr.n - 1; return; }
    const uint32_t *hl = a.hlog + pr.hlog_off;
    int w = 0;
    int score = pr.m;
    if ((pr.m & 63) != 0 && score == pr.best) { out[0] = start_fixed; out[1] = -1; ++w; }
    uint32_t word = 0;
    for (int j = 0; j < pr.n; ++j) {
        if ((j & 15) == 0) word = hl[j >> 4];
        const uint32_t d = (word >> ((j & 15) * 2)) & 3u;
        score += (int)(d & 1u) - (int)(d >> 1);
        if (score == pr.best) {
            out[2 * w] = start_fixed;
            out[2 * w + 1] = j;
            if (want_rev) {
                const int64_t idx = o + w;
                a.loc_pair[idx] = (int32_t)p;
                a.rev_tasks[slot++] = (int32_t)idx;
            }
            ++w;
        }
    }
}

// ------------------------------------------------------------------------------------------------
template <int WL, int G, bool REV>
static int launch_myers(Engine *e, const EdArgs &a, cudaStream_t st)
{
    static int blocks_cached[16] = {0};
    int &blocks = blocks_cached[e->device & 15];
    const size_t smem = (WL <= 8) ? (size_t)4 * ED_SMEM_SYMS * WL * 32 * sizeof(unsigned long long) : 0;
    if (blocks == 0) {
        int per_sm = 0;
        if (smem > 48 * 1024)
            DB200_CUDA_CHECK(cudaFuncSetAttribute(ed_myers_kernel<WL, G, REV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        DB200_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ed_myers_kernel<WL, G, REV>, 128, smem));
        if (per_sm < 1) per_sm = 1;
        blocks = e->sm_count * per_sm;
    }
    ed_myers_kernel<WL, G, REV><<<blocks, 128, smem, st>>>(a);
    count_launch(e);
    DB200_CUDA_CHECK(cudaGetLastError());
    return DB200_OK;
}

// ---- task "path" (edlib.cpp:268-285, 1177-1412) ---------------------------------------------------------------------------
constexpr int ED_PATH_WARP_MIN_W = 8;     // query words from which a pair goes to the warp kernel
constexpr int ED_PATH_WARP_MAX_W = 256;   // 8 words per lane
struct EdPathArgs {
    const EdPair *pairs;
    const uint64_t *peq;
    const uint8_t *tcodes;
    const db200_ed_result *results;
    const int64_t *loc_off;
    const int32_t *loc_buf;
    uint8_t *path_buf;          // slots: pair p writes at path_buf[q_off + t_off ..), at most m + n operations
    int32_t *path_len;
    int32_t *failed;            // set when a pair found no split consistent with its distance (an internal error)
    int64_t npairs;
    uint8_t *scratch;           // one region of `stride` bytes per thread: leaf columns | Pv | Mv | left | right | tmp
    int64_t stride;
    int32_t leaf_bytes, vec_bytes, col_bytes;
    int32_t *counter;
    int32_t split;              // 1: pairs of ED_PATH_WARP_MIN_W..MAX_W query words belong to ed_path_warp_kernel
    const int32_t *order;       // pair ids sorted by configuration and target length (the forward pass's task list), or null
    const int32_t *seg_begin, *seg_count;
};

// Thread per pair, pairs handed out through a counter (their cost spans five orders of magnitude).  The alignment is the
// global one of the query against target[start .. end] of the pair's first location; pairs without a result (distance
// above k, an empty sequence: edlib.cpp:161-179 returns before the path step) report length 0.  All of the work is in
// ed_path.cuh, which the CPU test suite runs against the reference's vectors.
// SHORT: every query this kernel meets has fewer than ED_PATH_WARP_MIN_W words, so the vectors and score columns sit in
// the thread's local memory (interleaved over the warp's lanes, unlike the per-thread regions of the scratch).
template <bool SHORT>
__global__ void __launch_bounds__(64) ed_path_kernel(EdPathArgs a)
{
    constexpr int SW = ED_PATH_WARP_MIN_W - 1;
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    uint8_t *base = a.scratch + tid * a.stride;
    uint64_t lPv[SHORT ? SW : 1], lMv[SHORT ? SW : 1];
    int32_t lleft[SHORT ? 64 * SW : 1], lright[SHORT ? 64 * SW : 1];
    EdPathScratch s;
    s.dir = reinterpret_cast<uint64_t *>(base);
    s.Pv = SHORT ? lPv : reinterpret_cast<uint64_t *>(base + a.leaf_bytes);
    s.Mv = SHORT ? lMv : reinterpret_cast<uint64_t *>(base + a.leaf_bytes + a.vec_bytes);
    s.left = SHORT ? lleft : reinterpret_cast<int32_t *>(base + a.leaf_bytes + 2 * (int64_t)a.vec_bytes);
    s.right = SHORT ? lright : reinterpret_cast<int32_t *>(base + a.leaf_bytes + 2 * (int64_t)a.vec_bytes + a.col_bytes);
    s.tmp = base + a.leaf_bytes + 2 * (int64_t)a.vec_bytes + 2 * (int64_t)a.col_bytes;
    // A warp takes 32 consecutive entries at a time and meets again before the next 32: in the sorted task list these are
    // pairs of one configuration and one 64-column length bin, so the lanes run sweeps and walks of similar length.
    const int lane = threadIdx.x & 31;
    const int64_t total = a.order ? (int64_t)a.seg_begin[ED_NCFG - 1] + a.seg_count[ED_NCFG - 1] : a.npairs;
    while (true) {
        __syncwarp();
        int first = 0;
        if (lane == 0) first = atomicAdd(a.counter, 32);
        first = __shfl_sync(0xFFFFFFFFu, first, 0);
        if (first >= total) break;
        const int64_t idx = (int64_t)first + lane;
        if (idx >= total) continue;
        const int64_t p = a.order ? a.order[total - 1 - idx] : idx;   // longest queries first
        const EdPair pr = a.pairs[p];
        if (a.split && pr.W >= ED_PATH_WARP_MIN_W && pr.W <= ED_PATH_WARP_MAX_W) continue;   // ed_path_warp_kernel's pairs
        const db200_ed_result r = a.results[p];
        int len = 0;
        if (!pr.special && r.status == 0 && r.edit_distance >= 0 && r.num_locations > 0) {
            const int64_t l = a.loc_off[p];
            const int start = a.loc_buf[2 * l], end = a.loc_buf[2 * l + 1];
            len = ed_path_pair(a.peq + pr.peq_off, a.tcodes + pr.t_off, pr.W, pr.m, start, end + 1, r.edit_distance, s,
                               a.path_buf + pr.q_off + pr.t_off);
        }
        if (len < 0) { *a.failed = 1; len = 0; }
        a.path_len[p] = len;
    }
}

// ---- long pairs: one warp per pair ---------------------------------------------------------------------------------------
// The same steps as ed_path_pair, with the two sweeps - score columns of the halving step, stored columns of a leaf - run
// as a wavefront over the lanes: lane l owns WL consecutive 64-row words in registers and works on column s - l at step s,
// the horizontal delta of a lane boundary travels with __shfl_up_sync.  The walk back over a leaf fetches the stored
// columns 32 at a time (path_leaf_walk_warp).

template <int WL, bool STORE, bool REV>
__device__ __noinline__ void path_warp_sweep(const uint64_t *peq, const uint8_t *tcodes, int W, int q0, int q1, int tfirst, int ncol,
                                             uint64_t *PM, int32_t *out)
{
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    const int qlen = q1 - q0, Wq = (qlen + 63) >> 6;
    const int w0 = lane * WL;
    const int nw = max(0, min(WL, Wq - w0));
    const int nl = (Wq + WL - 1) / WL;   // lanes that own words
    // Match-vector bits of this lane: forward word w covers query positions [q0 + 64 w, + 64), reverse word w covers
    // [q1 - 64 (w + 1), + 64) read backwards - either way 64 bits at one fixed bit offset into WL + 1 consecutive words of a
    // Peq row (zero outside the row), loaded once per column.
    const long pos0 = REV ? (long)q1 - 64L * (w0 + WL) : (long)q0 + 64L * w0;
    const long wb = pos0 >= 0 ? (pos0 >> 6) : -((-pos0 + 63) >> 6);
    const int sh = (int)(pos0 - wb * 64);
    uint32_t vmask = 0;
#pragma unroll
    for (int k = 0; k <= WL; ++k) vmask |= (wb + k >= 0 && wb + k < W) ? (1u << k) : 0u;
    uint64_t Pv[WL], Mv[WL];
    int32_t end[WL];                     // value of each word's last row (column -1: 64 w + 64)
#pragma unroll
    for (int i = 0; i < WL; ++i) { Pv[i] = ~0ull; Mv[i] = 0ull; end[i] = 64 * (w0 + i + 1); }
    int32_t *E = STORE ? reinterpret_cast<int32_t *>(PM + (long)ncol * Wq * 2) : nullptr;
    uint32_t cp = 0, cm = 0;             // deltas leaving this lane's last word at the previous step
    const int steps = ncol + nl - 1;
    for (int s = 0; s < steps; ++s) {
        uint32_t hp = __shfl_up_sync(FULL, cp, 1), hm = __shfl_up_sync(FULL, cm, 1);
        if (lane == 0) { hp = 1; hm = 0; }   // the row above the matrix grows by one per column
        const int c = s - lane;
        if (nw > 0 && c >= 0 && c < ncol) {
            const uint64_t *row = peq + (long)tcodes[REV ? tfirst - c : tfirst + c] * W + wb;
            uint64_t r[WL + 1];
#pragma unroll
            for (int k = 0; k <= WL; ++k) r[k] = ((vmask >> k) & 1u) ? row[k] : 0ull;
            uint64_t hin_p = hp, hin_m = hm;
#pragma unroll
            for (int i = 0; i < WL; ++i) {
                if (i < nw) {
                    const uint64_t lo = REV ? r[WL - 1 - i] : r[i], hi = REV ? r[WL - i] : r[i + 1];
                    uint64_t Eq = sh ? ((lo >> sh) | (hi << (64 - sh))) : lo;
                    if (REV) Eq = __brevll(Eq);
                    const uint64_t pv = Pv[i], mv = Mv[i];
                    const uint64_t Xv = Eq | mv;
                    Eq |= hin_m;
                    const uint64_t Xh = (((Eq & pv) + pv) ^ pv) | Eq;
                    uint64_t Ph = mv | ~(Xh | pv);
                    uint64_t Mh = pv & Xh;
                    const uint64_t op = Ph >> 63, om = Mh >> 63;
                    Ph = (Ph << 1) | hin_p;
                    Mh = (Mh << 1) | hin_m;
                    Pv[i] = Mh | ~(Xv | Ph);
                    Mv[i] = Ph & Xv;
                    if (STORE) {
                        const long k = (long)c * Wq + w0 + i;
                        *reinterpret_cast<ulonglong2 *>(PM + 2 * k) = make_ulonglong2(Pv[i], Mv[i]);
                        end[i] += (int)op - (int)om;
                        E[k] = end[i];
                    }
                    hin_p = op; hin_m = om;
                }
            }
            cp = (uint32_t)hin_p; cm = (uint32_t)hin_m;
        }
    }
    if (!STORE) {
        // out[r] = ncol + vertical deltas of rows 0..r: totals of the earlier lanes by a warp scan, then bit by bit
        int tot = 0;
#pragma unroll
        for (int i = 0; i < WL; ++i) if (i < nw) tot += __popcll(Pv[i]) - __popcll(Mv[i]);
        int incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += v;
        }
        int sc = ncol + incl - tot;
#pragma unroll
        for (int i = 0; i < WL; ++i) {
            if (i < nw) {
                const int r0 = (w0 + i) * 64;
                const int nb = min(64, qlen - r0);
                uint64_t p = Pv[i], mm = Mv[i];
                for (int b = 0; b < nb; ++b) {
                    sc += (int)(p & 1ull) - (int)(mm & 1ull);
                    p >>= 1; mm >>= 1;
                    out[r0 + b] = sc;
                }
            }
        }
    }
    __syncwarp();
}

template <bool STORE, bool REV>
__device__ __forceinline__ void path_warp_sweep_any(const uint64_t *peq, const uint8_t *tcodes, int W, int q0, int q1, int tfirst, int ncol,
                                                    uint64_t *PM, int32_t *out)
{
    // the fewest words per lane that cover the sub-problem: a lane's words run one after the other, idle lanes cost nothing
    switch (((q1 - q0 + 63) >> 6) + 31 >> 5) {
        case 1: path_warp_sweep<1, STORE, REV>(peq, tcodes, W, q0, q1, tfirst, ncol, PM, out); break;
        case 2: path_warp_sweep<2, STORE, REV>(peq, tcodes, W, q0, q1, tfirst, ncol, PM, out); break;
        case 3: path_warp_sweep<3, STORE, REV>(peq, tcodes, W, q0, q1, tfirst, ncol, PM, out); break;
        case 4: path_warp_sweep<4, STORE, REV>(peq, tcodes, W, q0, q1, tfirst, ncol, PM, out); break;
        case 5: path_warp_sweep<5, STORE, REV>(peq, tcodes, W, q0, q1, tfirst, ncol, PM, out); break;
        case 6: path_warp_sweep<6, STORE, REV>(peq, tcodes, W, q0, q1, tfirst, ncol, PM, out); break;
        case 7: path_warp_sweep<7, STORE, REV>(peq, tcodes, W, q0, q1, tfirst, ncol, PM, out); break;
        default: path_warp_sweep<8, STORE, REV>(peq, tcodes, W, q0, q1, tfirst, ncol, PM, out); break;
    }
}

// path_leaf_walk for a whole warp.  The scalar walk waits for three dependent loads per step; here lane l holds the stored
// (Pv, Mv, last-row value) of column jb - l at the current row's word, fetched 32 columns at a time, every lane follows the
// same walk and a step reads its two columns with shuffles.  Column -1 is a lane's constant (all deltas +1, D[r][-1] = r + 1).
__device__ int path_leaf_walk_warp(const uint64_t *PM, int qlen, int tlen, uint8_t *tmp, uint8_t *out)
{
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    const int Wq = (qlen + 63) >> 6;
    const int32_t *E = reinterpret_cast<const int32_t *>(PM + (long)tlen * Wq * 2);
    auto below = [](int i) { return (i & 63) == 63 ? 0ull : (~0ull << ((i & 63) + 1)); };
    int i = qlen - 1, j = tlen - 1, n = 0;
    int cur = 0, left = 0;
    bool first = true;
    while (i >= 0 && j >= 0) {
        const int wi = i >> 6, jb = j;
        const int col = jb - lane;
        uint64_t cP = ~0ull, cM = 0ull;
        int cE = 64 * wi + 64;
        if (col >= 0) {
            const long k = (long)col * Wq + wi;
            const ulonglong2 pm = *reinterpret_cast<const ulonglong2 *>(PM + 2 * k);
            cP = pm.x; cM = pm.y; cE = E[k];
        }
        {   // D[i][j-1] from lane 1 (and D[i][j] from lane 0 at the very start)
            const uint64_t b = below(i);
            const int mine = cE - __popcll(cP & b) + __popcll(cM & b);
            if (first) { cur = __shfl_sync(FULL, mine, 0); first = false; }
            left = __shfl_sync(FULL, mine, 1);
        }
        while (i >= 0 && j >= 0 && (i >> 6) == wi && jb - j <= 30) {
            const int src = jb - j, sh = i & 63;
            const uint64_t pc0 = __shfl_sync(FULL, cP, src), pc1 = __shfl_sync(FULL, cM, src);
            const uint64_t pl0 = __shfl_sync(FULL, cP, src + 1), pl1 = __shfl_sync(FULL, cM, src + 1);
            const int vc = (int)((pc0 >> sh) & 1ull) - (int)((pc1 >> sh) & 1ull);
            const int vl = (int)((pl0 >> sh) & 1ull) - (int)((pl1 >> sh) & 1ull);
            const int up = cur - vc, diag = left - vl;
            const int op = (up + 1 == cur) ? 1 : ((left + 1 == cur) ? 2 : (diag == cur ? 0 : 3));
            if (lane == 0) tmp[n] = (uint8_t)op;
            ++n;
            if (op == 1) { --i; cur = up; left = diag; }
            else {
                cur = (op == 2) ? left : diag;
                if (op != 2) --i;
                --j;
                // the new left neighbour D[i][j-1] while its column is still held; otherwise the next fetch recomputes it
                if (i >= 0 && j >= 0 && (i >> 6) == wi && jb - j <= 30) {
                    const int s2 = jb - j + 1;
                    const uint64_t lP = __shfl_sync(FULL, cP, s2), lM = __shfl_sync(FULL, cM, s2);
                    const int lE = __shfl_sync(FULL, cE, s2);
                    const uint64_t b = below(i);
                    left = lE - __popcll(lP & b) + __popcll(lM & b);
                }
            }
        }
    }
    __syncwarp();
    // the rest of the first row (deletions), then of the first column (insertions); then alignment order
    const int nd = j + 1, ni = i + 1;
    for (int k = lane; k < nd; k += 32) tmp[n + k] = 2;
    for (int k = lane; k < ni; k += 32) tmp[n + nd + k] = 1;
    n += nd + ni;
    __syncwarp();
    for (int k = lane; k < n; k += 32) out[k] = tmp[n - 1 - k];
    __syncwarp();
    return n;
}

// ed_path_pair for a whole warp (every lane follows the same control flow; m <= 64 * ED_PATH_WARP_MAX_W)
__device__ int ed_path_pair_warp(const uint64_t *peq, const uint8_t *tcodes, int W, int m, int t0, int t1, int best, const EdPathScratch &s,
                                 uint8_t *out)
{
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    int stack[ED_PATH_STACK][5];
    int sp = 0, n = 0;
    stack[sp][0] = 0; stack[sp][1] = m; stack[sp][2] = t0; stack[sp][3] = t1; stack[sp][4] = best; ++sp;
    while (sp > 0) {
        __syncwarp();
        --sp;
        const int q0 = stack[sp][0], q1 = stack[sp][1], a0 = stack[sp][2], a1 = stack[sp][3], score = stack[sp][4];
        const int qlen = q1 - q0, tlen = a1 - a0;
        if (qlen == 0 || tlen == 0) {
            for (int k = lane; k < tlen; k += 32) out[n + k] = 2;
            for (int k = lane; k < qlen; k += 32) out[n + tlen + k] = 1;
            n += tlen + qlen;
            continue;
        }
        const long long blocks = (qlen + 63) >> 6;
        const long long state = (2ll * 8 + 4) * blocks * tlen + 2ll * 4 * tlen;   // the reference's 1 MB rule (edlib.cpp:1204-1206)
        if (state < 1024 * 1024) {
            path_warp_sweep_any<true, false>(peq, tcodes, W, q0, q1, a0, tlen, s.dir, nullptr);
            n += path_leaf_walk_warp(s.dir, qlen, tlen, s.tmp, out + n);
            continue;
        }
        const int lw = tlen / 2, rw = tlen - lw;
        path_warp_sweep_any<false, false>(peq, tcodes, W, q0, q1, a0, lw, nullptr, s.left);
        path_warp_sweep_any<false, true>(peq, tcodes, W, q0, q1, a1 - 1, rw, nullptr, s.right);
        // first query index whose two halves add up to the optimum, then the two boundary cases (edlib.cpp:1330-1372)
        int idx = -2, ls = 0, rs = 0;
        for (int base = 0; base <= qlen - 2 && idx == -2; base += 32) {
            const int qi = base + lane;
            const bool hit = qi <= qlen - 2 && s.left[qi] + s.right[qlen - 2 - qi] == score;
            const unsigned b = __ballot_sync(FULL, hit);
            if (b) idx = base + __ffs(b) - 1;
        }
        if (idx >= 0) { ls = s.left[idx]; rs = s.right[qlen - 2 - idx]; }
        if (idx == -2 && lw + s.right[qlen - 1] == score) { idx = -1; ls = lw; rs = s.right[qlen - 1]; }
        if (idx == -2 && s.left[qlen - 1] + rw == score) { idx = qlen - 1; ls = s.left[qlen - 1]; rs = rw; }
        if (idx == -2 || sp + 2 > ED_PATH_STACK) return -1;
        const int qs = q0 + idx + 1;
        stack[sp][0] = qs; stack[sp][1] = q1; stack[sp][2] = a0 + lw; stack[sp][3] = a1; stack[sp][4] = rs; ++sp;
        stack[sp][0] = q0; stack[sp][1] = qs; stack[sp][2] = a0; stack[sp][3] = a0 + lw; stack[sp][4] = ls; ++sp;
    }
    __syncwarp();
    return n;
}

__global__ void __launch_bounds__(128) ed_path_warp_kernel(EdPathArgs a)
{
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    const int64_t wid = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    uint8_t *base = a.scratch + wid * a.stride;
    EdPathScratch s;
    s.dir = reinterpret_cast<uint64_t *>(base);
    s.Pv = nullptr; s.Mv = nullptr;   // the vectors live in registers
    s.left = reinterpret_cast<int32_t *>(base + a.leaf_bytes + 2 * (int64_t)a.vec_bytes);
    s.right = reinterpret_cast<int32_t *>(base + a.leaf_bytes + 2 * (int64_t)a.vec_bytes + a.col_bytes);
    s.tmp = base + a.leaf_bytes + 2 * (int64_t)a.vec_bytes + 2 * (int64_t)a.col_bytes;
    // three rounds over the pairs, the costliest class first (a 10 kb x 10 kb pair keeps its warp busy for milliseconds:
    // started last it would be the tail of the launch)
    int round = 0;
    while (round < 3) {
        __syncwarp();
        int64_t p = 0;
        if (lane == 0) p = atomicAdd(a.counter + round, 1);
        p = __shfl_sync(FULL, p, 0);
        if (p >= a.npairs) { ++round; continue; }
        const EdPair pr = a.pairs[p];
        if (pr.W < ED_PATH_WARP_MIN_W || pr.W > ED_PATH_WARP_MAX_W) continue;   // ed_path_kernel's pairs
        const long long cells = (long long)pr.m * pr.n;
        if ((cells >= (1ll << 25) ? 0 : cells >= (1ll << 21) ? 1 : 2) != round) continue;
        const db200_ed_result r = a.results[p];
        int len = 0;
        if (!pr.special && r.status == 0 && r.edit_distance >= 0 && r.num_locations > 0) {
            const int64_t l = a.loc_off[p];
            const int start = a.loc_buf[2 * l], end = a.loc_buf[2 * l + 1];
            len = ed_path_pair_warp(a.peq + pr.peq_off, a.tcodes + pr.t_off, pr.W, pr.m, start, end + 1, r.edit_distance, s,
                                    a.path_buf + pr.q_off + pr.t_off);
        }
        if (lane == 0) {
            if (len < 0) { *a.failed = 1; len = 0; }
            a.path_len[p] = len;
        }
    }
}

// warp per pair: the operations move from their slot to the pair's place in the caller's buffer (CSR, like the locations)
struct EdPathPackArgs {
    const EdPair *pairs;
    const uint8_t *slots;
    const int32_t *path_len;
    const int64_t *path_off;
    uint8_t *out;
    int64_t cap, npairs;
    int32_t *overflow;
};

__global__ void ed_path_pack_kernel(EdPathPackArgs a)
{
    const int lane = threadIdx.x & 31;
    const int64_t p = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (p >= a.npairs) return;
    const int len = a.path_len[p];
    if (len <= 0) return;
    const int64_t dst = a.path_off[p];
    if (dst + len > a.cap) { if (lane == 0) *a.overflow = 1; return; }
    const uint8_t *src = a.slots + a.pairs[p].q_off + a.pairs[p].t_off;
    for (int k = lane; k < len; k += 32) a.out[dst + k] = src[k];
}

template <bool REV>
static int launch_ed_cfg(Engine *e, int cfg, const EdArgs &a, cudaStream_t st)
{
    switch (cfg) {
        case 0: return launch_myers<1, 1, REV>(e, a, st);
        case 1: return launch_myers<2, 1, REV>(e, a, st);
        case 2: return launch_myers<4, 1, REV>(e, a, st);
        case 3: return launch_myers<4, 2, REV>(e, a, st);
        case 4: return launch_myers<4, 3, REV>(e, a, st);
        case 5: return launch_myers<4, 4, REV>(e, a, st);
        case 6: return launch_myers<4, 6, REV>(e, a, st);
        case 7: return launch_myers<4, 8, REV>(e, a, st);
        case 8: return launch_myers<4, 10, REV>(e, a, st);
        case 9: return launch_myers<3, 16, REV>(e, a, st);
        case 10: return launch_myers<4, 16, REV>(e, a, st);
        case 11: return launch_myers<3, 32, REV>(e, a, st);
        case 12: return launch_myers<4, 32, REV>(e, a, st);
        case 13: return launch_myers<5, 32, REV>(e, a, st);
        case 14: return launch_myers<6, 32, REV>(e, a, st);
        case 15: return launch_myers<8, 32, REV>(e, a, st);
        case 16: return launch_myers<16, 32, REV>(e, a, st);
        default: return launch_myers<32, 32, REV>(e, a, st);
    }
}

template <int R>
static int launch_banded(Engine *e, const EdBandArgs &a, int64_t npairs, cudaStream_t st)
{
    const unsigned blocks = (unsigned)((npairs + 127) / 128);   // upper bound: the lists live on the device
    ed_banded_nw_kernel<ed_band_words(R), R == ED_BAND_ROUNDS - 1><<<blocks, 128, 0, st>>>(a);
    count_launch(e);
    DB200_CUDA_CHECK(cudaGetLastError());
    return DB200_OK;
}

int ed_batch_device_impl(Engine *e, Arena &ws, const db200_ed_params *prm, const uint8_t *d_qbuf, const int64_t *d_qoff,
                         int64_t q_bytes, const uint8_t *d_tbuf, const int64_t *d_toff, int64_t t_bytes, int64_t npairs,
                         int32_t max_query_len, int32_t max_target_len, const int32_t *d_k, db200_ed_result *d_results,
                         int32_t *d_loc_buf, int64_t loc_cap, int64_t *d_loc_off, int64_t *loc_total, cudaStream_t st,
                         uint8_t *d_path_buf, int64_t path_cap, int64_t *d_path_off, int64_t *path_total)
{
    if (npairs < 0 || !prm || !d_results) { set_error("ed_batch: bad arguments"); return DB200_ERR_ARG; }
    const bool want_path = prm->task == DB200_ED_TASK_PATH;
    if (want_path && (!d_path_buf || !d_path_off)) {
        set_error("ed_batch: task 'path' needs the path buffers of db200_ed_path_batch_host / db200_ed_path_batch_device");
        return DB200_ERR_ARG;
    }
    if (want_path && npairs == 0) DB200_CUDA_CHECK(cudaMemsetAsync(d_path_off, 0, sizeof(int64_t), st));
    if (want_path && npairs == 0 && path_total) *path_total = 0;
    if (prm->task < 0 || prm->task > 2) { set_error("ed_batch: bad task"); return DB200_ERR_ARG; }
    if (npairs == 0) {
        if (d_loc_off) DB200_CUDA_CHECK(cudaMemsetAsync(d_loc_off, 0, sizeof(int64_t), st));
        if (loc_total) *loc_total = 0;
        return DB200_OK;
    }
    if (npairs > (int64_t)1 << 30) { set_error("ed_batch: at most 2^30 pairs per call"); return DB200_ERR_ARG; }
    if (!d_loc_buf || !d_loc_off) { set_error("ed_batch: location buffers are required"); return DB200_ERR_ARG; }
    (void)max_target_len;
    const bool need_rev = (prm->mode == DB200_ED_MODE_HW && prm->task >= DB200_ED_TASK_LOC);
    // match-vector storage: rows = distinct target symbols (<= 256, and <= target length), W words each.
    // Upper bound without a device round trip: min(256, n) * ceil(m/64) per pair <= (t_bytes_pair) * W ...
    // we bound rows by min(256, max alphabet) using the caller-provided lengths.
    const int64_t Wmax = ((int64_t)std::max(max_query_len, 1) + 63) / 64;
    const size_t nbins = (size_t)ED_NCFG * ED_LBINS;

    EdPair *pairs = ws.alloc<EdPair>(npairs);
    int32_t *peq_words = ws.alloc<int32_t>(npairs);
    int32_t *hlog_words = ws.alloc<int32_t>(npairs);
    int64_t *peq_off = ws.alloc<int64_t>(npairs + 1);
    int64_t *hlog_off = ws.alloc<int64_t>(npairs + 1);
    int64_t *scan_tmp = ws.alloc<int64_t>(scan_tmp_count(npairs) * 2);
    int32_t *bins = ws.alloc<int32_t>(nbins * 3 + 512);
    int32_t *tasks = ws.alloc<int32_t>(npairs);
    uint32_t *tsets = ws.alloc<uint32_t>((size_t)npairs * 8);
    uint8_t *tcodes = ws.alloc<uint8_t>((size_t)t_bytes + 64);
    uint32_t *hlog = ws.alloc<uint32_t>((size_t)(t_bytes / 16 + npairs + 64));
    int32_t *nw_score = ws.alloc<int32_t>(npairs);
    int32_t *nloc = ws.alloc<int32_t>(npairs);
    int32_t *loc_pair = ws.alloc<int32_t>((size_t)std::max<int64_t>(loc_cap, 1));
    int32_t *rev_tasks = ws.alloc<int32_t>((size_t)std::max<int64_t>(loc_cap, 1));
    uint8_t *eq_dev = ws.alloc<uint8_t>(2 * (size_t)std::max(prm->n_eq_pairs, 1));
    // NW runs the banded rounds first (DB200_ED_BANDED=0: full columns only, the measurement baseline)
    const char *band_env = getenv("DB200_ED_BANDED");
    const bool band = prm->mode == DB200_ED_MODE_NW && !(band_env && atoi(band_env) == 0) && npairs * ED_NCFG < ((int64_t)1 << 31);
    int32_t *carry = band ? ws.alloc<int32_t>((size_t)npairs * ED_BAND_ROUNDS) : nullptr;
    int32_t *fb_tasks = band ? ws.alloc<int32_t>((size_t)npairs * ED_NCFG) : nullptr;
    if (band && (!carry || !fb_tasks)) { set_error("ed_batch: device workspace allocation failed"); return DB200_ERR_NOMEM; }
    if (!pairs || !peq_words || !hlog_words || !peq_off || !hlog_off || !scan_tmp || !bins || !tasks || !tsets || !tcodes || !hlog ||
        !nw_score || !nloc || !loc_pair || !rev_tasks || !eq_dev) {
        set_error("ed_batch: device workspace allocation failed");
        return DB200_ERR_NOMEM;
    }
    int32_t *hist = bins, *bin_start = bins + nbins, *bin_fill = bins + 2 * nbins;
    int32_t *small = bins + 3 * nbins;
    static_assert(ED_NCFG <= 32, "the small arrays below hold 32 entries each");
    int32_t *seg_begin = small, *seg_count = small + 32, *rev_count = small + 64, *rev_begin = small + 96, *rev_fill = small + 128;
    int32_t *counters = small + 160; // 64 work counters
    int32_t *overflow = small + 224;
    int32_t *carry_count = small + 232;   // [ED_BAND_ROUNDS]
    int32_t *fb_count = small + 256;      // [ED_NCFG]
    int32_t *fb_begin = small + 288;      // [ED_NCFG] = cfg * npairs
    DB200_CUDA_CHECK(cudaMemsetAsync(bins, 0, (nbins * 3 + 512) * sizeof(int32_t), st));
    if (prm->n_eq_pairs > 0)
        DB200_CUDA_CHECK(cudaMemcpyAsync(eq_dev, prm->eq_pairs, 2 * (size_t)prm->n_eq_pairs, cudaMemcpyHostToDevice, st));

    const int TB = 256;
    const unsigned gpair = (unsigned)((npairs + TB - 1) / TB), gwarp = (unsigned)((npairs * 32 + TB - 1) / TB);
    prof_mark(e, st, ST_BEGIN);

    EdPrepArgs pa;
    pa.qbuf = d_qbuf; pa.tbuf = d_tbuf; pa.qoff = d_qoff; pa.toff = d_toff; pa.k = d_k; pa.k_default = prm->k;
    pa.mode = prm->mode; pa.task = prm->task; pa.npairs = npairs; pa.pairs = pairs; pa.peq_words = peq_words;
    pa.hlog_words = hlog_words; pa.hist = hist; pa.results = d_results;
    pa.tsets = tsets; pa.band = band ? 1 : 0; pa.fb_tasks = fb_tasks; pa.fb_count = fb_count; pa.fb_stride = npairs;
    ed_prep_kernel<<<gwarp, TB, 0, st>>>(pa);
    count_launch(e);
    exclusive_scan_i32_i64(e, peq_words, peq_off, npairs, scan_tmp, st);
    exclusive_scan_i32_i64(e, hlog_words, hlog_off, npairs, scan_tmp, st);
    ed_binscan_kernel<<<1, 32, 0, st>>>(hist, bin_start, seg_begin, seg_count);
    count_launch(e);

    // The Peq pool size depends on the data (distinct target symbols); it is the one quantity read back.
    int64_t peq_total = 0;
    DB200_CUDA_CHECK(cudaMemcpyAsync(&peq_total, peq_off + npairs, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    DB200_CUDA_CHECK(cudaStreamSynchronize(st));
    uint64_t *peq = ws.alloc<uint64_t>((size_t)peq_total + 64);
    if (!peq) { set_error("ed_batch: device workspace allocation failed (Peq, %lld words)", (long long)peq_total); return DB200_ERR_NOMEM; }

    EdBuildArgs ba;
    ba.qbuf = d_qbuf; ba.tbuf = d_tbuf; ba.pairs = pairs; ba.peq_off = peq_off; ba.hlog_off = hlog_off; ba.npairs = npairs;
    ba.peq = peq; ba.tcodes = tcodes; ba.mode = prm->mode; ba.task = prm->task; ba.eq_pairs = eq_dev; ba.n_eq_pairs = prm->n_eq_pairs;
    ba.bin_start = bin_start; ba.bin_fill = bin_fill; ba.tasks = tasks; ba.band = band ? 1 : 0; ba.tsets = tsets;
    ed_build_kernel<<<gwarp, TB, 0, st>>>(ba);
    count_launch(e);
    prof_mark(e, st, ST_ED_PREP);

    EdArgs ea;
    memset(&ea, 0, sizeof(ea));
    ea.pairs = pairs; ea.peq = peq; ea.tcodes = tcodes; ea.hlog = hlog; ea.mode = prm->mode; ea.nw_score = nw_score;
    ea.loc_pair = loc_pair; ea.loc_buf = d_loc_buf;
    ea.wsteps = reinterpret_cast<unsigned long long *>(e->d_sub_words + Engine::NSUB_MAX + 1);
    int ci = 0;
    if (band) {
        EdBandArgs bd;
        memset(&bd, 0, sizeof(bd));
        bd.pairs = pairs; bd.peq = peq; bd.tcodes = tcodes; bd.tasks = tasks; bd.seg_begin = seg_begin; bd.seg_count = seg_count;
        bd.fb_tasks = fb_tasks; bd.fb_count = fb_count; bd.fb_stride = npairs; bd.nw_score = nw_score; bd.wsteps = ea.wsteps;
        for (int r = 0; r < ED_BAND_ROUNDS; ++r) {
            bd.round = r;
            bd.carry = r > 0 ? carry + (size_t)npairs * r : nullptr;
            bd.carry_count = r > 0 ? carry_count + r : nullptr;
            bd.next = carry + (size_t)npairs * (r + 1 < ED_BAND_ROUNDS ? r + 1 : 0);
            bd.next_count = carry_count + (r + 1 < ED_BAND_ROUNDS ? r + 1 : 0);
            int rc;
            switch (r) {
                case 0: rc = launch_banded<0>(e, bd, npairs, st); break;
                case 1: rc = launch_banded<1>(e, bd, npairs, st); break;
                case 2: rc = launch_banded<2>(e, bd, npairs, st); break;
                case 3: rc = launch_banded<3>(e, bd, npairs, st); break;
                default: rc = launch_banded<4>(e, bd, npairs, st); break;
            }
            if (rc) return rc;
        }
        // what no band resolved: the full-column kernels, one list per configuration, side by side on the side streams
        // (every configuration is a launch of its own whose tail - one warp finishing a 10 kb pair - nothing else hides)
        int32_t fb_begin_h[ED_NCFG];
        for (int cfg = 0; cfg < ED_NCFG; ++cfg) fb_begin_h[cfg] = (int32_t)(npairs * cfg);
        DB200_CUDA_CHECK(cudaMemcpyAsync(fb_begin, fb_begin_h, sizeof(fb_begin_h), cudaMemcpyHostToDevice, st));
        { const int rc = aux_fork(e, st); if (rc) return rc; }
        for (int cfg = ED_NCFG - 1; cfg >= 0; --cfg) {
            ea.tasks = fb_tasks; ea.seg_begin = fb_begin; ea.seg_count = fb_count; ea.seg = cfg; ea.counter = counters + (ci++);
            const int rc = launch_ed_cfg<false>(e, cfg, ea, e->aux[e->aux_set][cfg % Engine::NAUX]);
            if (rc) return rc;
        }
        { const int rc = aux_join(e, st); if (rc) return rc; }
    } else {
        { const int rc = aux_fork(e, st); if (rc) return rc; }
        for (int cfg = ED_NCFG - 1; cfg >= 0; --cfg) {   // longest queries first: their drain overlaps the shorter configurations
            ea.tasks = tasks; ea.seg_begin = seg_begin; ea.seg_count = seg_count; ea.seg = cfg; ea.counter = counters + (ci++);
            const int rc = launch_ed_cfg<false>(e, cfg, ea, e->aux[e->aux_set][cfg % Engine::NAUX]);
            if (rc) return rc;
        }
        { const int rc = aux_join(e, st); if (rc) return rc; }
    }
    prof_mark(e, st, ST_ED_FORWARD);

    EdCountArgs ca;
    ca.pairs = pairs; ca.hlog = hlog; ca.nw_score = nw_score; ca.results = d_results; ca.nloc = nloc; ca.rev_count = rev_count;
    ca.npairs = npairs; ca.mode = prm->mode; ca.task = prm->task;
    ed_count_kernel<<<gpair, TB, 0, st>>>(ca);
    ed_revscan_kernel<<<1, 32, 0, st>>>(rev_count, rev_begin);
    count_launch(e, 2);
    exclusive_scan_i32_i64(e, nloc, d_loc_off, npairs, scan_tmp, st);
    EdEmitArgs ma;
    ma.pairs = pairs; ma.hlog = hlog; ma.loc_off = d_loc_off; ma.loc_buf = d_loc_buf; ma.loc_cap = loc_cap; ma.loc_pair = loc_pair;
    ma.rev_begin = rev_begin; ma.rev_fill = rev_fill; ma.rev_tasks = rev_tasks; ma.overflow = overflow; ma.npairs = npairs;
    ma.mode = prm->mode; ma.task = prm->task;
    ed_emit_kernel<<<gpair, TB, 0, st>>>(ma);
    count_launch(e);
    prof_mark(e, st, ST_ED_OUTPUT);

    if (need_rev) {
        { const int rc = aux_fork(e, st); if (rc) return rc; }
        for (int cfg = ED_NCFG - 1; cfg >= 0; --cfg) {
            ea.tasks = rev_tasks; ea.seg_begin = rev_begin; ea.seg_count = rev_count; ea.seg = cfg; ea.counter = counters + (ci++);
            const int rc = launch_ed_cfg<true>(e, cfg, ea, e->aux[e->aux_set][cfg % Engine::NAUX]);
            if (rc) return rc;
        }
        { const int rc = aux_join(e, st); if (rc) return rc; }
    }
    prof_mark(e, st, ST_ED_STARTS);
    if (want_path) {
        // Pairs of 8..256 query words go to the warp kernel, the rest to the thread kernel.  Scratch per worker: the
        // reference's 1 MB rule bounds the stored columns of a leaf, the rest follows the longest query it can meet.
        // the aligned slice of the infix / prefix modes is never longer than query + distance <= twice the query
        const int64_t maxt = prm->mode == DB200_ED_MODE_NW ? std::max(max_target_len, 1)
                                                           : std::max<int64_t>(1, std::min<int64_t>(max_target_len, 2 * (int64_t)max_query_len));
        const bool split = Wmax >= ED_PATH_WARP_MIN_W;
        const int64_t budget = (int64_t)3 << 30;   // 3 GB of scratch per kernel at most
        auto layout = [&](EdPathArgs &xa, int64_t qcap) {
            const int64_t Wq = (qcap + 63) / 64;
            xa.leaf_bytes = (int32_t)std::min<int64_t>(ED_PATH_LEAF_BYTES, (20 * Wq * maxt + 15) & ~15ll);
            xa.vec_bytes = (int32_t)(8 * Wq);
            xa.col_bytes = (int32_t)((4 * (qcap + 1) + 7) & ~7ll);
            xa.stride = (xa.leaf_bytes + 2 * (int64_t)xa.vec_bytes + 2 * (int64_t)xa.col_bytes + qcap + maxt + 1 + 15) & ~15ll;
        };
        EdPathArgs xa;
        memset(&xa, 0, sizeof(xa));
        xa.pairs = pairs; xa.peq = peq; xa.tcodes = tcodes; xa.results = d_results; xa.loc_off = d_loc_off; xa.loc_buf = d_loc_buf;
        uint8_t *slots = ws.alloc<uint8_t>((size_t)(q_bytes + t_bytes + 16));
        int32_t *plen = ws.alloc<int32_t>((size_t)npairs);
        if (!slots || !plen) { set_error("ed_batch: device workspace allocation failed (path slots)"); return DB200_ERR_NOMEM; }
        DB200_CUDA_CHECK(cudaMemsetAsync(plen, 0, sizeof(int32_t) * (size_t)npairs, st));   // pairs without DP are in no task list
        xa.path_buf = slots; xa.path_len = plen; xa.failed = overflow + 1; xa.npairs = npairs; xa.split = split ? 1 : 0;
        xa.order = band ? nullptr : tasks; xa.seg_begin = seg_begin; xa.seg_count = seg_count;
        {
            const int64_t qmax = std::max(max_query_len, 1);
            const bool short_only = Wmax <= ED_PATH_WARP_MAX_W;   // longer queries than the warp kernel takes come back here
            layout(xa, short_only ? std::min<int64_t>(qmax, 64 * (ED_PATH_WARP_MIN_W - 1)) : qmax);
            int64_t threads = std::min<int64_t>((npairs + 63) / 64 * 64, (int64_t)e->sm_count * 256);
            threads = std::max<int64_t>(64, std::min(threads, budget / xa.stride) / 64 * 64);
            xa.scratch = ws.alloc<uint8_t>((size_t)(threads * xa.stride));
            if (!xa.scratch) { set_error("ed_batch: device workspace allocation failed (path scratch)"); return DB200_ERR_NOMEM; }
            xa.counter = counters + (ci++);
            if (short_only) ed_path_kernel<true><<<(unsigned)(threads / 64), 64, 0, st>>>(xa);
            else ed_path_kernel<false><<<(unsigned)(threads / 64), 64, 0, st>>>(xa);
            count_launch(e);
        }
        if (split) {
            layout(xa, std::min<int64_t>(std::max(max_query_len, 1), 64 * ED_PATH_WARP_MAX_W));
            // four blocks of four warps per SM at 122 registers (capping the registers for a fifth block spilled and ran
            // the HiFi path stage at 40 instead of 36 ms)
            int64_t warps = std::min<int64_t>((npairs + 3) / 4 * 4, (int64_t)e->sm_count * 16);
            warps = std::max<int64_t>(4, std::min(warps, budget / xa.stride) / 4 * 4);
            xa.scratch = ws.alloc<uint8_t>((size_t)(warps * xa.stride));
            if (!xa.scratch) { set_error("ed_batch: device workspace allocation failed (path scratch)"); return DB200_ERR_NOMEM; }
            xa.counter = counters + ci;   // one per round
            ci += 3;
            ed_path_warp_kernel<<<(unsigned)(warps / 4), 128, 0, st>>>(xa);
            count_launch(e);
        }
        exclusive_scan_i32_i64(e, plen, d_path_off, npairs, scan_tmp, st);
        EdPathPackArgs ka;
        ka.pairs = pairs; ka.slots = slots; ka.path_len = plen; ka.path_off = d_path_off; ka.out = d_path_buf; ka.cap = path_cap;
        ka.npairs = npairs; ka.overflow = overflow + 2;
        ed_path_pack_kernel<<<gwarp, TB, 0, st>>>(ka);
        count_launch(e);
        prof_mark(e, st, ST_ED_PATH);
    }
    DB200_CUDA_CHECK(cudaGetLastError());
    if (loc_total || (want_path && path_total)) {
        int32_t ovf[3] = {0, 0, 0};
        int64_t ltot = 0, ptot = 0;
        DB200_CUDA_CHECK(cudaMemcpyAsync(&ltot, d_loc_off + npairs, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        if (want_path) DB200_CUDA_CHECK(cudaMemcpyAsync(&ptot, d_path_off + npairs, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        DB200_CUDA_CHECK(cudaMemcpyAsync(ovf, overflow, sizeof(ovf), cudaMemcpyDeviceToHost, st));
        DB200_CUDA_CHECK(cudaStreamSynchronize(st));
        if (loc_total) *loc_total = ltot;
        if (want_path && path_total) *path_total = ptot;
        if (ovf[0]) { set_error("ed_batch: location buffer too small (%lld needed)", (long long)ltot); return DB200_ERR_CAPACITY; }
        if (ovf[1]) { set_error("ed_batch: no alignment path consistent with the distance of a pair (internal error)"); return DB200_ERR_INTERNAL; }
        if (ovf[2]) { set_error("ed_batch: path buffer too small (%lld bytes needed)", (long long)ptot); return DB200_ERR_CAPACITY; }
    }
    return DB200_OK;
}

}  // namespace db200
