// Fine-grained Smith-Waterman configurations with G = 2 threads per pair (tagged variants; see sw_launch.cuh).
#include "sw_launch.cuh"

namespace db200 {

int sw_launch_fine_g2(Engine *e, int K, bool colmax, const PassArgs &a, cudaStream_t st)
{
    switch (K) {
        DB200_FINE_CASE(10, 2)
        DB200_FINE_CASE(11, 2)
        DB200_FINE_CASE(12, 2)
        DB200_FINE_CASE(13, 2)
        DB200_FINE_CASE(14, 2)
        DB200_FINE_CASE(15, 2)
        DB200_FINE_CASE(17, 2)
        DB200_FINE_CASE(18, 2)
        default: set_error("sw_batch: no fine configuration K = %d, G = 2", K); return DB200_ERR_INTERNAL;
    }
}

}  // namespace db200
