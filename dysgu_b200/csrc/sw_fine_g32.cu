// Fine-grained Smith-Waterman configurations with G = 32 threads per pair (tagged variants; see sw_launch.cuh).
#include "sw_launch.cuh"

namespace db200 {

int sw_launch_fine_g32(Engine *e, int K, bool colmax, const PassArgs &a, cudaStream_t st)
{
    switch (K) {
        DB200_FINE_CASE(10, 32)
        DB200_FINE_CASE(11, 32)
        DB200_FINE_CASE(12, 32)
        DB200_FINE_CASE(13, 32)
        DB200_FINE_CASE(14, 32)
        DB200_FINE_CASE(15, 32)
        default: set_error("sw_batch: no fine configuration K = %d, G = 32", K); return DB200_ERR_INTERNAL;
    }
}

}  // namespace db200
