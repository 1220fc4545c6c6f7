// Multi-device host entry points: one call, the pairs of a batch spread over several GPUs of the box.
//
// Every pair is independent, so there is no exchange step and no collective: the batch is cut into one contiguous
// range per device, each range is aligned by its own host thread through the single-device entry point
// (db200_sw_batch_host / db200_ed_batch_host: three-slot copy/compute pipeline, per-device lock), and the outputs
// land where the caller's pair order puts them - fixed-size records directly (results + c0), the variable-size
// cigars / locations through per-range regions of the caller's buffer that are closed up afterwards.
//
// Balance: the cut points are placed on equal shares of WORK, not of pairs - cells (len(query) x len(target)) for
// Smith-Waterman, word steps (ceil(len(query) / 64) x len(target)) for edit distance - taken from a prefix sum over the
// caller's CSR offsets.  A contiguous cut needs no gather of the sequences on the host (the ranges are views of the
// caller's buffers); dealing pairs round-robin by length bucket would balance the same way but costs a host copy of
// every base before the first H2D starts, which is more than the kernels take (842 MB per headline step).
#include "engine.cuh"

#include <thread>
#include <vector>

namespace db200 {

static std::vector<int64_t> cut_by_work(const int64_t *qoff, const int64_t *toff, int64_t npairs, int ndev, bool words)
{
    std::vector<double> acc((size_t)npairs + 1, 0.0);
    for (int64_t i = 0; i < npairs; ++i) {
        const double m = (double)(qoff[i + 1] - qoff[i]), n = (double)(toff[i + 1] - toff[i]);
        const double rows = words ? (double)(((int64_t)m + 63) / 64) : m;
        acc[(size_t)i + 1] = acc[(size_t)i] + rows * n + 64.0;   // + a per-pair constant so that empty pairs spread too
    }
    std::vector<int64_t> cut((size_t)ndev + 1, npairs);
    cut[0] = 0;
    int64_t i = 0;
    for (int d = 1; d < ndev; ++d) {
        const double want = acc[(size_t)npairs] * d / ndev;
        while (i < npairs && acc[(size_t)i + 1] <= want) ++i;
        cut[(size_t)d] = i;
    }
    return cut;
}

}  // namespace db200

using namespace db200;

extern "C" {

int db200_sw_batch_host_multi(const int32_t *devices, int32_t ndevices, const db200_sw_params *params, const uint8_t *qbuf,
                              const int64_t *qoff, const uint8_t *tbuf, const int64_t *toff, int64_t npairs, const int32_t *mask_len,
                              db200_sw_result *results, uint32_t *cigar_buf, int64_t cigar_cap, int64_t *cigar_off, int32_t *status,
                              int64_t *pairs_per_device)
{
    if (!devices || ndevices <= 0 || ndevices > 64 || npairs < 0 || !params || !qoff || !toff || !results) {
        set_error("sw_batch_host_multi: bad arguments");
        return DB200_ERR_ARG;
    }
    if (ndevices == 1 || npairs < 2 * ndevices) {
        if (pairs_per_device) { for (int d = 0; d < ndevices; ++d) pairs_per_device[d] = d == 0 ? npairs : 0; }
        return db200_sw_batch_host(devices[0], params, qbuf, qoff, tbuf, toff, npairs, mask_len, results, cigar_buf, cigar_cap, cigar_off, status);
    }
    const std::vector<int64_t> cut = cut_by_work(qoff, toff, npairs, ndevices, false);
    // cigar regions: a range can emit at most (query bytes + target bytes + 2 per pair) ops - the bound the single-device
    // entry point uses; if the caller's buffer holds the sum of the bounds every range writes into its own region of it
    std::vector<int64_t> bound((size_t)ndevices), region((size_t)ndevices + 1, 0);
    for (int d = 0; d < ndevices; ++d) {
        const int64_t c0 = cut[(size_t)d], c1 = cut[(size_t)d + 1];
        bound[(size_t)d] = (qoff[c1] - qoff[c0]) + (toff[c1] - toff[c0]) + 2 * (c1 - c0);
        region[(size_t)d + 1] = region[(size_t)d] + bound[(size_t)d];
    }
    const bool in_place = cigar_buf && region[(size_t)ndevices] <= cigar_cap;
    std::vector<std::vector<uint32_t>> tmp((size_t)ndevices);
    std::vector<std::vector<int64_t>> coff((size_t)ndevices);
    std::vector<int> rcs((size_t)ndevices, DB200_OK);
    std::vector<std::string> errs((size_t)ndevices);
    std::vector<std::thread> threads;
    for (int d = 0; d < ndevices; ++d) {
        threads.emplace_back([&, d] {
            const int64_t c0 = cut[(size_t)d], c1 = cut[(size_t)d + 1], n = c1 - c0;
            coff[(size_t)d].assign((size_t)n + 1, 0);
            if (n == 0) return;
            uint32_t *cg = nullptr;
            int64_t cap = 0;
            if (cigar_buf) {
                if (in_place) { cg = cigar_buf + region[(size_t)d]; cap = bound[(size_t)d]; }
                else { tmp[(size_t)d].resize((size_t)bound[(size_t)d] + 16); cg = tmp[(size_t)d].data(); cap = (int64_t)tmp[(size_t)d].size(); }
            }
            rcs[(size_t)d] = db200_sw_batch_host(devices[d], params, qbuf, qoff + c0, tbuf, toff + c0, n, mask_len ? mask_len + c0 : nullptr,
                                                 results + c0, cg, cap, coff[(size_t)d].data(), status ? status + c0 : nullptr);
            if (rcs[(size_t)d] != DB200_OK) errs[(size_t)d] = db200_last_error();   // the message is thread-local
        });
    }
    for (auto &t : threads) t.join();
    for (int d = 0; d < ndevices; ++d)
        if (rcs[(size_t)d] != DB200_OK) { set_error("device %d: %s", devices[d], errs[(size_t)d].c_str()); return rcs[(size_t)d]; }
    // close the regions up (ranges are in pair order, so are their cigars) and rebase the offsets
    int64_t pos = 0;
    for (int d = 0; d < ndevices; ++d) {
        const int64_t c0 = cut[(size_t)d], n = cut[(size_t)d + 1] - c0;
        const int64_t total = coff[(size_t)d][(size_t)n];
        if (cigar_buf && total > 0) {
            if (pos + total > cigar_cap) { set_error("sw_batch_host_multi: cigar buffer too small"); return DB200_ERR_CAPACITY; }
            const uint32_t *src = in_place ? cigar_buf + region[(size_t)d] : tmp[(size_t)d].data();
            if (src != cigar_buf + pos) memmove(cigar_buf + pos, src, sizeof(uint32_t) * (size_t)total);
        }
        if (cigar_off) for (int64_t i = 0; i < n; ++i) cigar_off[c0 + i] = pos + coff[(size_t)d][(size_t)i];
        pos += total;
        if (pairs_per_device) pairs_per_device[d] = n;
    }
    if (cigar_off) cigar_off[npairs] = pos;
    return DB200_OK;
}

int db200_ed_batch_host_multi(const int32_t *devices, int32_t ndevices, const db200_ed_params *params, const uint8_t *qbuf,
                              const int64_t *qoff, const uint8_t *tbuf, const int64_t *toff, int64_t npairs, const int32_t *k,
                              db200_ed_result *results, int32_t *loc_buf, int64_t loc_cap, int64_t *loc_off, int64_t *pairs_per_device)
{
    if (!devices || ndevices <= 0 || ndevices > 64 || npairs < 0 || !params || !qoff || !toff || !results || !loc_buf || !loc_off) {
        set_error("ed_batch_host_multi: bad arguments");
        return DB200_ERR_ARG;
    }
    if (ndevices == 1 || npairs < 2 * ndevices) {
        if (pairs_per_device) { for (int d = 0; d < ndevices; ++d) pairs_per_device[d] = d == 0 ? npairs : 0; }
        return db200_ed_batch_host(devices[0], params, qbuf, qoff, tbuf, toff, npairs, k, results, loc_buf, loc_cap, loc_off);
    }
    const std::vector<int64_t> cut = cut_by_work(qoff, toff, npairs, ndevices, true);
    std::vector<int64_t> bound((size_t)ndevices), region((size_t)ndevices + 1, 0);
    for (int d = 0; d < ndevices; ++d) {
        const int64_t c0 = cut[(size_t)d], c1 = cut[(size_t)d + 1];
        bound[(size_t)d] = (toff[c1] - toff[c0]) + 2 * (c1 - c0) + 16;   // locations of a range are bounded by its target bytes
        region[(size_t)d + 1] = region[(size_t)d] + bound[(size_t)d];
    }
    const bool in_place = region[(size_t)ndevices] <= loc_cap;
    std::vector<std::vector<int32_t>> tmp((size_t)ndevices);
    std::vector<std::vector<int64_t>> loff((size_t)ndevices);
    std::vector<int> rcs((size_t)ndevices, DB200_OK);
    std::vector<std::string> errs((size_t)ndevices);
    std::vector<std::thread> threads;
    for (int d = 0; d < ndevices; ++d) {
        threads.emplace_back([&, d] {
            const int64_t c0 = cut[(size_t)d], c1 = cut[(size_t)d + 1], n = c1 - c0;
            loff[(size_t)d].assign((size_t)n + 1, 0);
            if (n == 0) return;
            int32_t *lb;
            int64_t cap;
            if (in_place) { lb = loc_buf + 2 * region[(size_t)d]; cap = bound[(size_t)d]; }
            else { tmp[(size_t)d].resize((size_t)(2 * bound[(size_t)d])); lb = tmp[(size_t)d].data(); cap = bound[(size_t)d]; }
            rcs[(size_t)d] = db200_ed_batch_host(devices[d], params, qbuf, qoff + c0, tbuf, toff + c0, n, k ? k + c0 : nullptr, results + c0, lb,
                                                 cap, loff[(size_t)d].data());
            if (rcs[(size_t)d] != DB200_OK) errs[(size_t)d] = db200_last_error();
        });
    }
    for (auto &t : threads) t.join();
    for (int d = 0; d < ndevices; ++d)
        if (rcs[(size_t)d] != DB200_OK) { set_error("device %d: %s", devices[d], errs[(size_t)d].c_str()); return rcs[(size_t)d]; }
    int64_t pos = 0;
    for (int d = 0; d < ndevices; ++d) {
        const int64_t c0 = cut[(size_t)d], n = cut[(size_t)d + 1] - c0;
        const int64_t total = loff[(size_t)d][(size_t)n];
        if (total > 0) {
            if (pos + total > loc_cap) { set_error("ed_batch_host_multi: location buffer too small"); return DB200_ERR_CAPACITY; }
            const int32_t *src = in_place ? loc_buf + 2 * region[(size_t)d] : tmp[(size_t)d].data();
            if (src != loc_buf + 2 * pos) memmove(loc_buf + 2 * pos, src, sizeof(int32_t) * 2 * (size_t)total);
        }
        for (int64_t i = 0; i < n; ++i) loc_off[c0 + i] = pos + loff[(size_t)d][(size_t)i];
        pos += total;
        if (pairs_per_device) pairs_per_device[d] = n;
    }
    loc_off[npairs] = pos;
    return DB200_OK;
}

}  // extern "C"
