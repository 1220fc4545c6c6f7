// Host-side launch of one sw_pass_kernel instantiation (persistent grid sized from the occupancy of that instantiation) and the
// table of kernel configurations.  Shared by sw.cu and the translation units that hold the fine-grained configurations
// (sw_fine_g*.cu: one per group size, so that the ~100 instantiations compile in parallel).
#pragma once
#include "sw_kernels.cuh"

namespace db200 {

// ------------------------------------------------------------------------------------------------
// kernel configurations: index -> (K columns per virtual lane, G threads per pair); columns per pass = 2*G*K
// ------------------------------------------------------------------------------------------------
//   0 .. 10   "classic": K = 16, every variant (tagged or not, column maxima, WATCH re-run) exists
//  11 .. 58   "fine":    K = 8 .. 18 for G = 1, K = 10 .. 18 for G = 2 .. 16, K = 10 .. 15 for G = 32 (K = 16 is classic);
//                        tagged variants only, chosen per pair by cost when the scoring leaves room for the tag bits
//  59         column-tiled kernel (K = 32, G = 32, passes of 2048 columns)
// Why fine configurations: a pair occupies G threads x 2 virtual lanes x K columns, and with K fixed at 16 a target is
// padded to a multiple of 32 columns per thread while G = 3, 5, 6, 7, 10 leave lanes of the warp idle; with K free the
// columns are spread evenly over a power-of-two group (a 150-column clip: G = 4, K = 19 ... a 100-column one: G = 4, K = 13).
struct SwCfg { int K, G; };
constexpr int SW_NCLASSIC = 11;
constexpr int SW_CFG_MULTI = SW_NCFG - 1;
constexpr int SW_MULTI_COLS = 2048;   // columns per pass of the column-tiled kernel (K = 32, G = 32)

inline const SwCfg *sw_cfg_table()
{
    static SwCfg tab[SW_NCFG];
    static bool done = false;
    if (!done) {
        const int gs[SW_NCLASSIC] = {1, 2, 3, 4, 5, 6, 7, 8, 10, 16, 32};
        int c = 0;
        for (int i = 0; i < SW_NCLASSIC; ++i) tab[c++] = SwCfg{16, gs[i]};
        for (int g = 1; g <= 32; g <<= 1)
            for (int k = (g == 1 ? 8 : 10); k <= (g == 32 ? 15 : 18); ++k)
                if (k != 16) tab[c++] = SwCfg{k, g};
        tab[c++] = SwCfg{32, 32};
        done = (c == SW_NCFG);
        if (!done) { fprintf(stderr, "dysgu_b200: configuration table holds %d entries, expected %d\n", c, SW_NCFG); abort(); }
    }
    return tab;
}

template <int K, int G, bool COLMAX, bool MULTI, bool WATCH, bool TAG, bool PIPE = false>
static int launch_pass(Engine *e, const PassArgs &a, cudaStream_t st)
{
    const size_t smem = (size_t)(SW_PASS_THREADS / 32) * 6 * (2 * K * 32) * sizeof(uint16_t);
    static int blocks_cached[16] = {0};
    int &blocks = blocks_cached[e->device & 15];
    if (blocks == 0) {
        DB200_CUDA_CHECK(cudaFuncSetAttribute(sw_pass_kernel<K, G, COLMAX, MULTI, WATCH, TAG, PIPE>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        DB200_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sw_pass_kernel<K, G, COLMAX, MULTI, WATCH, TAG, PIPE>, SW_PASS_THREADS, smem));
        if (per_sm < 1) per_sm = 1;
        if (MULTI && per_sm * (SW_PASS_THREADS / 32) > SW_MULTI_MAX_WARPS_PER_SM) per_sm = SW_MULTI_MAX_WARPS_PER_SM / (SW_PASS_THREADS / 32);
        blocks = e->sm_count * per_sm;
    }
    sw_pass_kernel<K, G, COLMAX, MULTI, WATCH, TAG, PIPE><<<blocks, SW_PASS_THREADS, smem, st>>>(a);
    count_launch(e);
    DB200_CUDA_CHECK(cudaGetLastError());
    return DB200_OK;
}

// fine configurations of one group size (sw_fine_g<G>.cu); K outside the group's range is an internal error
int sw_launch_fine_g1(Engine *e, int K, bool colmax, const PassArgs &a, cudaStream_t st);
int sw_launch_fine_g2(Engine *e, int K, bool colmax, const PassArgs &a, cudaStream_t st);
int sw_launch_fine_g4(Engine *e, int K, bool colmax, const PassArgs &a, cudaStream_t st);
int sw_launch_fine_g8(Engine *e, int K, bool colmax, const PassArgs &a, cudaStream_t st);
int sw_launch_fine_g16(Engine *e, int K, bool colmax, const PassArgs &a, cudaStream_t st);
int sw_launch_fine_g32(Engine *e, int K, bool colmax, const PassArgs &a, cudaStream_t st);

#define DB200_FINE_CASE(KK, GG) \
    case KK: return colmax ? launch_pass<KK, GG, true, false, false, true>(e, a, st) : launch_pass<KK, GG, false, false, false, true>(e, a, st);

}  // namespace db200
