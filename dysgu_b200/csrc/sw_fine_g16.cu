// Fine-grained Smith-Waterman configurations with G = 16 threads per pair (tagged variants; see sw_launch.cuh).
#include "sw_launch.cuh"

namespace db200 {

int sw_launch_fine_g16(Engine *e, int K, bool colmax, const PassArgs &a, cudaStream_t st)
{
    switch (K) {
        DB200_FINE_CASE(10, 16)
        DB200_FINE_CASE(11, 16)
        DB200_FINE_CASE(12, 16)
        DB200_FINE_CASE(13, 16)
        DB200_FINE_CASE(14, 16)
        DB200_FINE_CASE(15, 16)
        DB200_FINE_CASE(17, 16)
        DB200_FINE_CASE(18, 16)
        default: set_error("sw_batch: no fine configuration K = %d, G = 16", K); return DB200_ERR_INTERNAL;
    }
}

}  // namespace db200
