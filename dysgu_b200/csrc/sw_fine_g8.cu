// Fine-grained Smith-Waterman configurations with G = 8 threads per pair (tagged variants; see sw_launch.cuh).
#include "sw_launch.cuh"

namespace db200 {

int sw_launch_fine_g8(Engine *e, int K, bool colmax, const PassArgs &a, cudaStream_t st)
{
    switch (K) {
        DB200_FINE_CASE(10, 8)
        DB200_FINE_CASE(11, 8)
        DB200_FINE_CASE(12, 8)
        DB200_FINE_CASE(13, 8)
        DB200_FINE_CASE(14, 8)
        DB200_FINE_CASE(15, 8)
        DB200_FINE_CASE(17, 8)
        DB200_FINE_CASE(18, 8)
        default: set_error("sw_batch: no fine configuration K = %d, G = 8", K); return DB200_ERR_INTERNAL;
    }
}

}  // namespace db200
