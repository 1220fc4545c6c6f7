// Fine-grained Smith-Waterman configurations with G = 4 threads per pair (tagged variants; see sw_launch.cuh).
#include "sw_launch.cuh"

namespace db200 {

int sw_launch_fine_g4(Engine *e, int K, bool colmax, const PassArgs &a, cudaStream_t st)
{
    switch (K) {
        DB200_FINE_CASE(10, 4)
        DB200_FINE_CASE(11, 4)
        DB200_FINE_CASE(12, 4)
        DB200_FINE_CASE(13, 4)
        DB200_FINE_CASE(14, 4)
        DB200_FINE_CASE(15, 4)
        DB200_FINE_CASE(17, 4)
        DB200_FINE_CASE(18, 4)
        default: set_error("sw_batch: no fine configuration K = %d, G = 4", K); return DB200_ERR_INTERNAL;
    }
}

}  // namespace db200
