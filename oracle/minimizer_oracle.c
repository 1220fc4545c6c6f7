/*
 * TEST INFRASTRUCTURE ONLY - never linked, imported or executed by the product path.
 *
 * Scalar restatement of dysgu's minimizer sketch of a clip sequence (reference: dysgu/graph.pyx:59-84
 * sliding_window_minimum, hash = XXHash64::hash(ptr, m, 42) from dysgu/include/xxhash64.h through
 * map_set_utils.pxd:28-29, values handled as signed `long`).  The hash below is XXH64 written from the published
 * algorithm description, not from the vendored header; the window minimum follows the reference's monotone deque step by
 * step (pop larger-or-equal values from the back, push, drop the front once it left the window, record the front).
 *
 * Parity status: PINNED.  oracle/build_ref.sh compiles the reference's own xxhash64.h (through the three-line wrapper
 * oracle/xxh_ref_wrap.cpp) into oracle/_ref/libxxh_ref.so; tests/test_oracle_golden.py compares this hash with it on random
 * inputs of every length 0..80 and checks the committed golden sketches (tests/golden/minimizers.json, produced with the
 * reference hash).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static uint64_t rotl(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
#define P1 11400714785074694791ULL
#define P2 14029467366897019727ULL
#define P3 1609587929392839161ULL
#define P4 9650029242287828579ULL
#define P5 2870177450012600261ULL

static uint64_t rd64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }   /* little endian host */
static uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static uint64_t lane(uint64_t acc, uint64_t in) { acc += in * P2; acc = rotl(acc, 31); return acc * P1; }
static uint64_t merge(uint64_t h, uint64_t v) { h ^= lane(0, v); return h * P1 + P4; }

uint64_t oracle_xxh64(const uint8_t *p, int64_t len, uint64_t seed)
{
    const uint8_t *end = p + len;
    uint64_t h;
    if (len >= 32) {
        uint64_t a = seed + P1 + P2, b = seed + P2, c = seed, d = seed - P1;
        for (; p + 32 <= end; p += 32) { a = lane(a, rd64(p)); b = lane(b, rd64(p + 8)); c = lane(c, rd64(p + 16)); d = lane(d, rd64(p + 24)); }
        h = rotl(a, 1) + rotl(b, 7) + rotl(c, 12) + rotl(d, 18);
        h = merge(h, a); h = merge(h, b); h = merge(h, c); h = merge(h, d);
    } else {
        h = seed + P5;
    }
    h += (uint64_t)len;
    for (; p + 8 <= end; p += 8) { h ^= lane(0, rd64(p)); h = rotl(h, 27) * P1 + P4; }
    if (p + 4 <= end) { h ^= (uint64_t)rd32(p) * P1; h = rotl(h, 23) * P2 + P3; p += 4; }
    for (; p < end; ++p) { h ^= (uint64_t)(*p) * P5; h = rotl(h, 11) * P1; }
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
    return h;
}

static int cmp_i64(const void *a, const void *b)
{
    const int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}

/* The set graph.pyx:59-84 builds for one sequence, sorted ascending (signed); returns its size.  `hash` == NULL: the
 * restated XXH64; otherwise the given function (the reference's own hash when pinning). */
typedef uint64_t (*hash_fn)(const void *, uint64_t, uint64_t);
int64_t minimizer_oracle_one(const uint8_t *s, int64_t len, int k, int m, uint64_t seed, hash_fn hash, int64_t *out)
{
    const int64_t end = len - m + 1;
    if (end <= 0) return 0;
    int64_t *dq_val = (int64_t *)malloc(sizeof(int64_t) * (size_t)end), *dq_idx = (int64_t *)malloc(sizeof(int64_t) * (size_t)end);
    int64_t *found = (int64_t *)malloc(sizeof(int64_t) * (size_t)(end + 2));
    int64_t head = 0, tail = 0, nf = 0;
    for (int64_t i = 0; i < end; ++i) {
        const int64_t hx = (int64_t)(hash ? hash(s + i, (uint64_t)m, seed) : oracle_xxh64(s + i, m, seed));
        if (i == 0 || i == end - 1) found[nf++] = hx;
        while (tail > head && dq_val[tail - 1] >= hx) --tail;
        dq_val[tail] = hx; dq_idx[tail] = i; ++tail;
        while (tail > head && dq_idx[head] <= i - k) ++head;
        found[nf++] = dq_val[head];
    }
    qsort(found, (size_t)nf, sizeof(int64_t), cmp_i64);
    int64_t u = 0;
    for (int64_t i = 0; i < nf; ++i) if (i == 0 || found[i] != found[i - 1]) out[u++] = found[i];
    free(dq_val); free(dq_idx); free(found);
    return u;
}

int minimizer_oracle_batch(const uint8_t *buf, const int64_t *off, int64_t nseq, int k, int m, uint64_t seed, hash_fn hash, int64_t *out,
                           int64_t cap, int64_t *out_off)
{
    int64_t pos = 0;
    for (int64_t i = 0; i < nseq; ++i) {
        out_off[i] = pos;
        const int64_t len = off[i + 1] - off[i];
        if (pos + len + 2 > cap) return -1;
        pos += minimizer_oracle_one(buf + off[i], len, k, m, seed, hash, out + pos);
    }
    out_off[nseq] = pos;
    return 0;
}
