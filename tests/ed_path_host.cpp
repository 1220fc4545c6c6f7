// Host build of dysgu_b200/csrc/ed_path.cuh (the code the device runs for edlib task "path"), for tests only:
// tests/test_ed_path_host.py compiles this file with g++ and compares its output with vectors from the reference.
#include "../dysgu_b200/csrc/ed_path.cuh"

#include <stdlib.h>
#include <string.h>
#include <vector>

using namespace db200;

// Global alignment path of query vs target (raw bytes, compared for equality).  Returns the number of operations written
// to out (capacity m + n), -1 on failure; *distance receives the edit distance.
// eq: neq (first, second) byte pairs that also count as equal (edlib's additionalEqualities: symmetric, not transitive).
extern "C" int ed_path_host(const unsigned char *q, int m, const unsigned char *t, int n, const unsigned char *eq, int neq,
                            unsigned char *out, int *distance)
{
    int rank[256];
    for (int c = 0; c < 256; ++c) rank[c] = -1;
    bool present[256] = {false};
    for (int j = 0; j < n; ++j) present[t[j]] = true;
    int At = 0;
    for (int c = 0; c < 256; ++c) if (present[c]) rank[c] = At++;
    const int W = (m + 63) / 64 > 0 ? (m + 63) / 64 : 1;
    std::vector<uint64_t> peq((size_t)(At > 0 ? At : 1) * W, 0ull);
    for (int i = 0; i < m; ++i) {
        if (rank[q[i]] >= 0) peq[(size_t)rank[q[i]] * W + (i >> 6)] |= 1ull << (i & 63);
        for (int e = 0; e < neq; ++e) {
            const int x = eq[2 * e], y = eq[2 * e + 1];
            if (q[i] == x && rank[y] >= 0) peq[(size_t)rank[y] * W + (i >> 6)] |= 1ull << (i & 63);
            if (q[i] == y && rank[x] >= 0) peq[(size_t)rank[x] * W + (i >> 6)] |= 1ull << (i & 63);
        }
    }
    std::vector<uint8_t> tc((size_t)(n > 0 ? n : 1));
    for (int j = 0; j < n; ++j) tc[j] = (uint8_t)rank[t[j]];
    std::vector<uint64_t> Pv(W), Mv(W);
    std::vector<int32_t> left(m + 1), right(m + 1);
    std::vector<uint64_t> dir(ED_PATH_LEAF_BYTES / 8);
    std::vector<uint8_t> tmp((size_t)m + n + 1);
    EdPathScratch s;
    s.Pv = Pv.data(); s.Mv = Mv.data(); s.left = left.data(); s.right = right.data();
    s.dir = dir.data(); s.tmp = tmp.data();
    int best;
    if (m == 0 || n == 0) best = m > n ? m : n;
    else {
        path_score_column(peq.data(), tc.data(), W, 0, m, 0, n, false, s.Pv, s.Mv, s.left);
        best = s.left[m - 1];
    }
    *distance = best;
    return ed_path_pair(peq.data(), tc.data(), W, m, 0, n, best, s, out);
}
