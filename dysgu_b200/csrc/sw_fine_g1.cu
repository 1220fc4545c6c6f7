// Fine-grained Smith-Waterman configurations with G = 1 threads per pair (tagged variants; see sw_launch.cuh).
#include "sw_launch.cuh"

namespace db200 {

int sw_launch_fine_g1(Engine *e, int K, bool colmax, const PassArgs &a, cudaStream_t st)
{
    switch (K) {
        DB200_FINE_CASE(8, 1)
        DB200_FINE_CASE(9, 1)
        DB200_FINE_CASE(10, 1)
        DB200_FINE_CASE(11, 1)
        DB200_FINE_CASE(12, 1)
        DB200_FINE_CASE(13, 1)
        DB200_FINE_CASE(14, 1)
        DB200_FINE_CASE(15, 1)
        DB200_FINE_CASE(17, 1)
        DB200_FINE_CASE(18, 1)
        default: set_error("sw_batch: no fine configuration K = %d, G = 1", K); return DB200_ERR_INTERNAL;
    }
}

}  // namespace db200
