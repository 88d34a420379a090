// placeholder until the tcgen05 kernel lands
#include "common.cuh"
#include "kernels.h"
bool k_umma_supported(const dllm_qweight *, size_t) { return false; }
int32_t k_qlinear_umma(dllm_ctx *ctx, const dllm_qweight *, const void *, size_t, float *, void *) {
    DLLM_FAIL(ctx, DLLM_ERR_UNSUPPORTED, "tcgen05 path not built");
}
