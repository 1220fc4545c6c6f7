// Host-side text formatting for whole batches: the strings dysgu's callers read from the result objects -
// AlignmentStructure.cigar (_ssw_wrapper.pyx:240-275), aligned_query_sequence / aligned_target_sequence (302-364) and the
// cigar of edlib.align(task="path") (edlib.pyx:143-147 -> edlibAlignmentToCigar, edlib.cpp:298-347) - built for every
// pair of a batch in one call from the arrays the batch entry points return, instead of per result in Python.
// No device code and no CUDA call: these run wherever the library loads.
#include "engine.cuh"

namespace {

inline int dec_digits(uint32_t v)
{
    int d = 1;
    while (v >= 10) { v /= 10; ++d; }
    return d;
}

inline char *put_dec(char *p, uint32_t v)
{
    const int d = dec_digits(v);
    for (int i = d - 1; i >= 0; --i) { p[i] = (char)('0' + v % 10); v /= 10; }
    return p + d;
}

// Python's seq[a:b] on a sequence of length L, as (start, stop) with start <= stop
inline void py_slice(int64_t a, int64_t b, int64_t L, int64_t &start, int64_t &stop)
{
    start = a < 0 ? std::max<int64_t>(L + a, 0) : std::min<int64_t>(a, L);
    stop = b < 0 ? std::max<int64_t>(L + b, 0) : std::min<int64_t>(b, L);
    if (stop < start) stop = start;
}

// two passes over the same generator: count, then write
template <typename Gen>
int format_csr(int64_t npairs, char *out, int64_t out_cap, int64_t *out_off, Gen gen, const char *what)
{
    if (npairs < 0 || !out_off) { db200::set_error("%s: bad arguments", what); return DB200_ERR_ARG; }
    int64_t total = 0;
    for (int64_t p = 0; p < npairs; ++p) {
        out_off[p] = total;
        const int64_t n = gen(p, (char *)nullptr);
        if (n < 0) { db200::set_error("%s: malformed input at pair %lld", what, (long long)p); return DB200_ERR_ARG; }
        total += n;
    }
    out_off[npairs] = total;
    if (!out) return DB200_OK;   // sizes only
    if (total > out_cap) { db200::set_error("%s: output buffer too small (%lld bytes needed)", what, (long long)total); return DB200_ERR_CAPACITY; }
    for (int64_t p = 0; p < npairs; ++p) gen(p, out + out_off[p]);
    return DB200_OK;
}

}  // namespace

extern "C" {

int db200_sw_cigar_text(const uint32_t *cigar_buf, const int64_t *cigar_off, int64_t npairs, char *out, int64_t out_cap, int64_t *out_off)
{
    if (!cigar_off || (npairs > 0 && cigar_off[npairs] > cigar_off[0] && !cigar_buf)) { db200::set_error("sw_cigar_text: bad arguments"); return DB200_ERR_ARG; }
    auto gen = [&](int64_t p, char *dst) -> int64_t {
        int64_t n = 0;
        for (int64_t i = cigar_off[p]; i < cigar_off[p + 1]; ++i) {
            const uint32_t c = cigar_buf[i], op = c & 0xFu;
            if (op > 2) return -1;                     // mid_table holds M, I, D only (_ssw_wrapper.pyx:48)
            if (dst) { char *e = put_dec(dst + n, c >> 4); *e = "MID"[op]; }
            n += dec_digits(c >> 4) + 1;
        }
        return n;
    };
    return format_csr(npairs, out, out_cap, out_off, gen, "sw_cigar_text");
}

int db200_sw_aligned_text(int which, const uint8_t *seq_buf, const int64_t *seq_off, const db200_sw_result *results,
                          const uint32_t *cigar_buf, const int64_t *cigar_off, int64_t npairs, char *out, int64_t out_cap, int64_t *out_off)
{
    if ((which != 0 && which != 1) || !seq_off || !results || !cigar_off) { db200::set_error("sw_aligned_text: bad arguments"); return DB200_ERR_ARG; }
    const uint32_t gap_op = which == 0 ? 2u : 1u;      // query: gaps where the cigar deletes, target: where it inserts
    auto gen = [&](int64_t p, char *dst) -> int64_t {
        const db200_sw_result &r = results[p];
        const int64_t begin = which == 0 ? r.read_begin1 : r.ref_begin1, end = which == 0 ? r.read_end1 : r.ref_end1;
        const uint8_t *s = seq_buf + seq_off[p];
        int64_t s0, s1;
        py_slice(begin, end + 1, seq_off[p + 1] - seq_off[p], s0, s1);   // seq = sequence[begin:end + 1]
        const int64_t L = s1 - s0;
        int64_t n = 0, index = 0;
        for (int64_t i = cigar_off[p]; i < cigar_off[p + 1]; ++i) {
            const uint32_t c = cigar_buf[i], op = c & 0xFu;
            const int64_t len = c >> 4;
            if (op > 2) return -1;
            if (op == gap_op) {
                if (dst) memset(dst + n, '-', (size_t)len);
                n += len;
            } else {
                int64_t a, b;
                py_slice(index, index + len, L, a, b);                     // seq[index:index + length]
                if (dst && b > a) memcpy(dst + n, s + s0 + a, (size_t)(b - a));
                n += b - a;
                index += len;
            }
        }
        int64_t a, b;
        py_slice(index, end - begin + 1, L, a, b);                         // "our sequence end is sometimes beyond the cigar"
        if (dst && b > a) memcpy(dst + n, s + s0 + a, (size_t)(b - a));
        return n + (b - a);
    };
    return format_csr(npairs, out, out_cap, out_off, gen, "sw_aligned_text");
}

int db200_ed_cigar_text(const uint8_t *path_buf, const int64_t *path_off, int64_t npairs, int format, char *out, int64_t out_cap, int64_t *out_off)
{
    if (!path_off || (format != 0 && format != 1)) { db200::set_error("ed_cigar_text: bad arguments"); return DB200_ERR_ARG; }
    const char *letters = format == 0 ? "MIDM" : "=IDX";
    auto gen = [&](int64_t p, char *dst) -> int64_t {
        int64_t n = 0;
        for (int64_t i = path_off[p]; i < path_off[p + 1];) {
            if (path_buf[i] > 3) return -1;
            const char c = letters[path_buf[i]];
            int64_t j = i;
            while (j < path_off[p + 1] && path_buf[j] <= 3 && letters[path_buf[j]] == c) ++j;
            if (dst) { char *e = put_dec(dst + n, (uint32_t)(j - i)); *e = c; }
            n += dec_digits((uint32_t)(j - i)) + 1;
            i = j;
        }
        return n;
    };
    return format_csr(npairs, out, out_cap, out_off, gen, "ed_cigar_text");
}

}  // extern "C"

// ---- host-side packing into the engine's pool format (db200_sw_batch_host_packed) ------------------------------------------
// ASCII -> 4-bit codes of _ssw_wrapper.pyx:60-68 (A/a/U/u 0, C/c 1, G/g 2, T/t 3, anything else 4), eight bases per 32-bit word,
// every sequence padded with code 5 to a 16-byte unit of 32 bases.  What the device's sw_pack_kernel does, for callers that
// prepare (or cache) their batches on the host: the packed form is what should cross PCIe.
#include <thread>
extern "C" int db200_sw_pack_host(const uint8_t *buf, const int64_t *off, const int32_t *len, int64_t nseq, uint32_t *units,
                                  int64_t units_cap, int64_t *unit_off, int32_t *len_out, int32_t threads)
{
    if (nseq < 0 || !off || !unit_off) return DB200_ERR_ARG;
    unit_off[0] = 0;
    for (int64_t i = 0; i < nseq; ++i) {
        const int64_t l = len ? (int64_t)len[i] : off[i + 1] - off[i];
        if (l < 0 || l > 0x7FFFFFF0) return DB200_ERR_ARG;
        unit_off[i + 1] = unit_off[i] + (l + 31) / 32;
        if (len_out) len_out[i] = (int32_t)l;
    }
    if (!units) return DB200_OK;                       // sizing call
    if (unit_off[nseq] > units_cap || !buf) return DB200_ERR_CAPACITY;
    static uint8_t code[256];
    static bool init = false;
    if (!init) {
        for (int c = 0; c < 256; ++c) code[c] = 4;
        code[(int)'A'] = code[(int)'a'] = code[(int)'U'] = code[(int)'u'] = 0;
        code[(int)'C'] = code[(int)'c'] = 1; code[(int)'G'] = code[(int)'g'] = 2; code[(int)'T'] = code[(int)'t'] = 3;
        init = true;
    }
    auto work = [&](int64_t i0, int64_t i1) {
        for (int64_t i = i0; i < i1; ++i) {
            const uint8_t *src = buf + off[i];
            const int64_t l = len ? (int64_t)len[i] : off[i + 1] - off[i];
            uint32_t *dst = units + unit_off[i] * 4;
            const int64_t nwords = (unit_off[i + 1] - unit_off[i]) * 4;
            for (int64_t w = 0; w < nwords; ++w) {
                uint32_t x = 0x55555555u;
                const int64_t b0 = w * 8;
                if (b0 < l) {
                    x = 0;
                    for (int k = 0; k < 8; ++k) x |= (uint32_t)(b0 + k < l ? code[src[b0 + k]] : 5u) << (4 * k);
                }
                dst[w] = x;
            }
        }
    };
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(threads > 0 ? threads : 1, nseq / 4096 + 1));
    if (nt == 1) { work(0, nseq); return DB200_OK; }
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; ++t) pool.emplace_back(work, nseq * t / nt, nseq * (t + 1) / nt);
    for (auto &th : pool) th.join();
    return DB200_OK;
}
