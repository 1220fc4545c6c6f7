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
