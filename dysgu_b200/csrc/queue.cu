// Ticket queue over the batch entry points (include/dysgu_b200_align.h, "submit / flush / fetch").
//
// dysgu's call sites are synchronous per pair (StripedSmithWaterman(q)(t), edlib.align(q, t): SURVEY 8b); a caller
// that collects its candidates first submits them here, flushes once and reads the results by ticket.  Pairs that
// share their parameters form one group = one device batch; the sequences are staged in page-locked memory as CSR
// buffers while they are submitted, so that a flush is nothing but db200_sw_batch_host / db200_ed_batch_host calls.
#include "engine.cuh"

#include <string>
#include <vector>

using namespace db200;

namespace {

// growable page-locked byte / word buffer
template <typename T>
struct PinnedVec {
    T *p = nullptr;
    size_t n = 0, cap = 0;
    ~PinnedVec() { if (p) cudaFreeHost(p); }
    bool reserve(size_t want)
    {
        if (want <= cap) return true;
        size_t ncap = std::max<size_t>(want, std::max<size_t>(cap * 2, 4096));
        T *q = nullptr;
        if (cudaHostAlloc(&q, ncap * sizeof(T), cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return false; }
        if (n) memcpy(q, p, n * sizeof(T));
        if (p) cudaFreeHost(p);
        p = q;
        cap = ncap;
        return true;
    }
    bool append(const T *src, size_t count)
    {
        if (!reserve(n + count)) return false;
        if (count) memcpy(p + n, src, count * sizeof(T));
        n += count;
        return true;
    }
    bool push(T v) { return append(&v, 1); }
    void clear() { n = 0; }
};

struct Group {
    int kind = 0;                 // 0 = Smith-Waterman, 1 = edit distance
    db200_sw_params swp{};
    int8_t mat[25] = {0};
    bool has_mat = false;
    db200_ed_params edp{};
    std::string eq;
    PinnedVec<uint8_t> qbuf, tbuf;
    PinnedVec<int64_t> qoff, toff;
    PinnedVec<int32_t> aux;       // mask_len (SW) or k (edit distance) per pair
    std::vector<int64_t> tickets;
    int64_t pending() const { return (int64_t)tickets.size(); }
};

struct Ticket {
    int32_t kind;
    int32_t state;    // 0 pending, 1 done, 2 failed
    int64_t slot;     // index into the result store of its kind
};

}  // namespace

struct db200_queue {
    int device = 0;
    std::vector<Group *> groups;
    std::vector<Ticket> tickets;
    // result stores (stable until db200_queue_reset)
    std::vector<db200_sw_result> sw_res;
    std::vector<int32_t> sw_status;
    std::vector<int64_t> sw_cig_off;   // [slot] -> first cigar word; length in sw_res[slot].cigar_len
    std::vector<uint32_t> sw_cig;
    std::vector<db200_ed_result> ed_res;
    std::vector<int64_t> ed_loc_off;
    std::vector<int32_t> ed_loc;       // (start, end) pairs
    std::vector<int64_t> ed_path_off;  // [slot] -> first operation in ed_path (-1: the pair's group had another task)
    std::vector<int32_t> ed_path_len;
    std::vector<uint8_t> ed_path;      // edit operations of task "path" groups
    int64_t npending = 0;
    ~db200_queue() { for (auto g : groups) delete g; }
};

static bool same_sw(const Group &g, const db200_sw_params &p)
{
    if (g.kind != 0) return false;
    const db200_sw_params &a = g.swp;
    if (a.match != p.match || a.mismatch != p.mismatch || a.gap_open != p.gap_open || a.gap_extend != p.gap_extend ||
        a.score_size != p.score_size || a.flag != p.flag || a.filters != p.filters || a.filterd != p.filterd) return false;
    if (g.has_mat != (p.mat != nullptr)) return false;
    if (p.mat && memcmp(g.mat, p.mat, 25) != 0) return false;
    return true;
}

static bool same_ed(const Group &g, const db200_ed_params &p)
{
    if (g.kind != 1) return false;
    if (g.edp.mode != p.mode || g.edp.task != p.task) return false;
    const size_t nb = 2 * (size_t)std::max(p.n_eq_pairs, 0);
    if (g.eq.size() != nb) return false;
    return nb == 0 || memcmp(g.eq.data(), p.eq_pairs, nb) == 0;
}

static int64_t submit_common(db200_queue *q, Group *g, const uint8_t *query, int32_t qlen, const uint8_t *target, int32_t tlen, int32_t aux)
{
    if (g->qoff.n == 0) { if (!g->qoff.push(0) || !g->toff.push(0)) { set_error("queue: pinned allocation failed"); return DB200_ERR_NOMEM; } }
    if (!g->qbuf.append(query, (size_t)qlen) || !g->tbuf.append(target, (size_t)tlen) || !g->qoff.push((int64_t)g->qbuf.n) ||
        !g->toff.push((int64_t)g->tbuf.n) || !g->aux.push(aux)) {
        set_error("queue: pinned allocation failed");
        return DB200_ERR_NOMEM;
    }
    const int64_t ticket = (int64_t)q->tickets.size();
    q->tickets.push_back(Ticket{g->kind, 0, -1});
    g->tickets.push_back(ticket);
    ++q->npending;
    return ticket;
}

extern "C" {

db200_queue *db200_queue_create(int device)
{
    if (!get_engine(device)) return nullptr;   // no device: no queue (there is no CPU fallback)
    db200_queue *q = new db200_queue();
    q->device = device;
    return q;
}

void db200_queue_destroy(db200_queue *q) { delete q; }

int64_t db200_queue_submit_sw(db200_queue *q, const db200_sw_params *params, const uint8_t *query, int32_t qlen, const uint8_t *target,
                              int32_t tlen, int32_t mask_len)
{
    if (!q || !params || qlen < 0 || tlen < 0 || (qlen && !query) || (tlen && !target)) { set_error("queue_submit_sw: bad arguments"); return DB200_ERR_ARG; }
    if (params->mat && params->n != 5) { set_error("queue_submit_sw: only 5x5 nucleotide matrices are supported"); return DB200_ERR_ARG; }
    Group *g = nullptr;
    for (auto c : q->groups) if (same_sw(*c, *params)) { g = c; break; }
    if (!g) {
        g = new Group();
        g->kind = 0;
        g->swp = *params;
        g->has_mat = params->mat != nullptr;
        if (params->mat) memcpy(g->mat, params->mat, 25);
        q->groups.push_back(g);
    }
    if (mask_len < 0) mask_len = std::max(qlen / 2, 15);   // the wrapper's mask_auto rule (_ssw_wrapper.pyx:583-588)
    return submit_common(q, g, query, qlen, target, tlen, mask_len);
}

int64_t db200_queue_submit_ed(db200_queue *q, const db200_ed_params *params, const uint8_t *query, int32_t qlen, const uint8_t *target,
                              int32_t tlen)
{
    if (!q || !params || qlen < 0 || tlen < 0 || (qlen && !query) || (tlen && !target)) { set_error("queue_submit_ed: bad arguments"); return DB200_ERR_ARG; }
    Group *g = nullptr;
    for (auto c : q->groups) if (same_ed(*c, *params)) { g = c; break; }
    if (!g) {
        g = new Group();
        g->kind = 1;
        g->edp = *params;
        if (params->n_eq_pairs > 0) g->eq.assign(params->eq_pairs, 2 * (size_t)params->n_eq_pairs);
        q->groups.push_back(g);
    }
    return submit_common(q, g, query, qlen, target, tlen, params->k);
}

int64_t db200_queue_pending(const db200_queue *q) { return q ? q->npending : 0; }

int db200_queue_flush(db200_queue *q)
{
    if (!q) { set_error("queue_flush: bad arguments"); return DB200_ERR_ARG; }
    int first_err = DB200_OK;
    for (auto g : q->groups) {
        const int64_t n = g->pending();
        if (n == 0) continue;
        int rc;
        if (g->kind == 0) {
            const size_t base = q->sw_res.size();
            q->sw_res.resize(base + (size_t)n);
            q->sw_status.resize(base + (size_t)n);
            std::vector<int64_t> coff((size_t)n + 1);
            const int64_t cap = (int64_t)(g->qbuf.n + g->tbuf.n + 2 * (size_t)n + 16);
            const size_t cbase = q->sw_cig.size();
            q->sw_cig.resize(cbase + (size_t)cap);
            db200_sw_params p = g->swp;
            p.mat = g->has_mat ? g->mat : nullptr;
            p.n = g->has_mat ? 5 : 0;
            rc = db200_sw_batch_host(q->device, &p, g->qbuf.p, g->qoff.p, g->tbuf.p, g->toff.p, n, g->aux.p, q->sw_res.data() + base,
                                     q->sw_cig.data() + cbase, cap, coff.data(), q->sw_status.data() + base);
            if (rc == DB200_OK) {
                q->sw_cig.resize(cbase + (size_t)coff[(size_t)n]);
                q->sw_cig_off.resize(base + (size_t)n);
                for (int64_t i = 0; i < n; ++i) {
                    q->sw_cig_off[base + (size_t)i] = (int64_t)cbase + coff[(size_t)i];
                    Ticket &t = q->tickets[(size_t)g->tickets[(size_t)i]];
                    t.state = 1;
                    t.slot = (int64_t)base + i;
                }
            } else {
                q->sw_res.resize(base);
                q->sw_status.resize(base);
                q->sw_cig.resize(cbase);
            }
        } else {
            const size_t base = q->ed_res.size();
            q->ed_res.resize(base + (size_t)n);
            std::vector<int64_t> loff((size_t)n + 1);
            const int64_t cap = (int64_t)(g->tbuf.n + 2 * (size_t)n + 16);
            const size_t lbase = q->ed_loc.size();
            q->ed_loc.resize(lbase + 2 * (size_t)cap);
            db200_ed_params p = g->edp;
            p.eq_pairs = g->eq.empty() ? nullptr : g->eq.data();
            p.n_eq_pairs = (int32_t)(g->eq.size() / 2);
            p.k = -1;
            const bool want_path = p.task == DB200_ED_TASK_PATH;
            const size_t pbase = q->ed_path.size();
            q->ed_path_off.resize(base + (size_t)n, -1);
            q->ed_path_len.resize(base + (size_t)n, 0);
            if (want_path) {
                const int64_t pcap = (int64_t)(g->qbuf.n + g->tbuf.n);
                q->ed_path.resize(pbase + (size_t)pcap + 1);
                std::vector<int64_t> poff((size_t)n + 1);
                rc = db200_ed_path_batch_host(q->device, &p, g->qbuf.p, g->qoff.p, g->tbuf.p, g->toff.p, n, g->aux.p, q->ed_res.data() + base,
                                              q->ed_loc.data() + lbase, cap, loff.data(), q->ed_path.data() + pbase, pcap + 1, poff.data());
                if (rc == DB200_OK) {
                    for (int64_t i = 0; i < n; ++i) {
                        q->ed_path_off[base + (size_t)i] = (int64_t)pbase + poff[(size_t)i];
                        q->ed_path_len[base + (size_t)i] = (int32_t)(poff[(size_t)i + 1] - poff[(size_t)i]);
                    }
                    q->ed_path.resize(pbase + (size_t)poff[(size_t)n]);
                }
            } else {
                rc = db200_ed_batch_host(q->device, &p, g->qbuf.p, g->qoff.p, g->tbuf.p, g->toff.p, n, g->aux.p, q->ed_res.data() + base,
                                         q->ed_loc.data() + lbase, cap, loff.data());
            }
            if (rc != DB200_OK) { q->ed_path.resize(pbase); q->ed_path_off.resize(base); q->ed_path_len.resize(base); }
            if (rc == DB200_OK) {
                q->ed_loc.resize(lbase + 2 * (size_t)loff[(size_t)n]);
                q->ed_loc_off.resize(base + (size_t)n);
                for (int64_t i = 0; i < n; ++i) {
                    q->ed_loc_off[base + (size_t)i] = (int64_t)lbase + 2 * loff[(size_t)i];
                    Ticket &t = q->tickets[(size_t)g->tickets[(size_t)i]];
                    t.state = 1;
                    t.slot = (int64_t)base + i;
                }
            } else {
                q->ed_res.resize(base);
                q->ed_loc.resize(lbase);
            }
        }
        if (rc != DB200_OK) {
            for (int64_t i = 0; i < n; ++i) q->tickets[(size_t)g->tickets[(size_t)i]].state = 2;
            if (first_err == DB200_OK) first_err = rc;
        }
        q->npending -= n;
        g->tickets.clear();
        g->qbuf.clear(); g->tbuf.clear(); g->qoff.clear(); g->toff.clear(); g->aux.clear();
    }
    return first_err;
}

static int resolve(db200_queue *q, int64_t ticket, int kind, const Ticket **out)
{
    if (!q || ticket < 0 || ticket >= (int64_t)q->tickets.size()) { set_error("queue_fetch: unknown ticket %lld", (long long)ticket); return DB200_ERR_ARG; }
    if (q->tickets[(size_t)ticket].kind != kind) { set_error("queue_fetch: ticket %lld is of the other kind", (long long)ticket); return DB200_ERR_ARG; }
    if (q->tickets[(size_t)ticket].state == 0) {
        const int rc = db200_queue_flush(q);
        if (rc != DB200_OK && q->tickets[(size_t)ticket].state != 1) return rc;
    }
    if (q->tickets[(size_t)ticket].state != 1) { set_error("queue_fetch: the batch of ticket %lld failed", (long long)ticket); return DB200_ERR_INTERNAL; }
    *out = &q->tickets[(size_t)ticket];
    return DB200_OK;
}

int db200_queue_fetch_sw(db200_queue *q, int64_t ticket, db200_sw_result *result, const uint32_t **cigar, int32_t *status)
{
    const Ticket *t = nullptr;
    const int rc = resolve(q, ticket, 0, &t);
    if (rc != DB200_OK) return rc;
    const size_t s = (size_t)t->slot;
    if (result) *result = q->sw_res[s];
    if (cigar) *cigar = q->sw_res[s].cigar_len > 0 ? q->sw_cig.data() + q->sw_cig_off[s] : nullptr;
    if (status) *status = q->sw_status[s];
    return DB200_OK;
}

int db200_queue_fetch_ed(db200_queue *q, int64_t ticket, db200_ed_result *result, const int32_t **locations)
{
    const Ticket *t = nullptr;
    const int rc = resolve(q, ticket, 1, &t);
    if (rc != DB200_OK) return rc;
    const size_t s = (size_t)t->slot;
    if (result) *result = q->ed_res[s];
    if (locations) *locations = q->ed_res[s].num_locations > 0 ? q->ed_loc.data() + q->ed_loc_off[s] : nullptr;
    return DB200_OK;
}

int db200_queue_fetch_ed_path(db200_queue *q, int64_t ticket, const uint8_t **path, int32_t *path_len)
{
    const Ticket *t = nullptr;
    const int rc = resolve(q, ticket, 1, &t);
    if (rc != DB200_OK) return rc;
    const size_t s = (size_t)t->slot;
    const bool has = q->ed_path_off[s] >= 0 && q->ed_path_len[s] > 0;
    if (path) *path = has ? q->ed_path.data() + q->ed_path_off[s] : nullptr;
    if (path_len) *path_len = has ? q->ed_path_len[s] : 0;
    return DB200_OK;
}

int db200_queue_reset(db200_queue *q)
{
    if (!q) { set_error("queue_reset: bad arguments"); return DB200_ERR_ARG; }
    for (auto g : q->groups) {
        g->tickets.clear();
        g->qbuf.clear(); g->tbuf.clear(); g->qoff.clear(); g->toff.clear(); g->aux.clear();
    }
    q->tickets.clear();
    q->sw_res.clear(); q->sw_status.clear(); q->sw_cig_off.clear(); q->sw_cig.clear();
    q->ed_res.clear(); q->ed_loc_off.clear(); q->ed_loc.clear();
    q->ed_path_off.clear(); q->ed_path_len.clear(); q->ed_path.clear();
    q->npending = 0;
    return DB200_OK;
}

}  // extern "C"
