// Edit-distance pipeline (placeholder until the Myers kernels land in this file).
#include "ed_kernels.cuh"
namespace db200 {
int ed_batch_device_impl(Engine *, Arena &, const db200_ed_params *, const uint8_t *, const int64_t *, int64_t, const uint8_t *,
                         const int64_t *, int64_t, int64_t, int32_t, int32_t, const int32_t *, db200_ed_result *, int32_t *, int64_t,
                         int64_t *, int64_t *, cudaStream_t)
{
    set_error("edit distance: not built yet");
    return DB200_ERR_INTERNAL;
}
}
