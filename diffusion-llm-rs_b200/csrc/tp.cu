// tp.cu — tensor-parallel plumbing: one process per GPU, NCCL over NVLink 5 / NVSwitch, one
// collective per column->row linear pair at the layer boundary.  The
// communicator is created from a 128-byte unique id that the launcher broadcasts
// (torch.distributed / MPI / a file), so this library needs no process manager of its own.
#include <nccl.h>

#include "common.cuh"
#include "kernels.h"

#define NCCL_TRY(ctx, expr)                                                                   \
    do {                                                                                      \
        ncclResult_t _r = (expr);                                                             \
        if (_r != ncclSuccess) {                                                              \
            DLLM_SET_ERR(ctx, "NCCL error at %s:%d: %s", __FILE__, __LINE__, ncclGetErrorString(_r)); \
            return DLLM_ERR_NCCL;                                                             \
        }                                                                                     \
    } while (0)

static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");

int32_t tp_allreduce(dllm_ctx *ctx, float *buf, size_t n) {
    if (ctx->tp_world <= 1 || n == 0 || ctx->tp_skip_comm) return DLLM_OK;
    if (!ctx->nccl_comm) DLLM_FAIL(ctx, DLLM_ERR_NCCL, "dllm_tp_init has not been called");
    NCCL_TRY(ctx, ncclAllReduce(buf, buf, n, ncclFloat32, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    return DLLM_OK;
}

// bf16 partial sums of a row-parallel linear, reduced in place (half the NVLink bytes of the f32 form; the result
// feeds the next tcgen05 linear, which reads bf16 anyway)
int32_t tp_allreduce_bf16(dllm_ctx *ctx, void *buf, size_t n) {
    if (ctx->tp_world <= 1 || n == 0 || ctx->tp_skip_comm) return DLLM_OK;
    if (!ctx->nccl_comm) DLLM_FAIL(ctx, DLLM_ERR_NCCL, "dllm_tp_init has not been called");
    NCCL_TRY(ctx, ncclAllReduce(buf, buf, n, ncclBfloat16, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    return DLLM_OK;
}

// the same all-reduces on an explicit stream (the overlapped tensor-parallel forward issues them on ctx->comm_stream)
int32_t tp_allreduce_on(dllm_ctx *ctx, void *buf, size_t n, bool bf16, cudaStream_t stream) {
    if (ctx->tp_world <= 1 || n == 0 || ctx->tp_skip_comm) return DLLM_OK;
    if (!ctx->nccl_comm) DLLM_FAIL(ctx, DLLM_ERR_NCCL, "dllm_tp_init has not been called");
    NCCL_TRY(ctx, ncclAllReduce(buf, buf, n, bf16 ? ncclBfloat16 : ncclFloat32, ncclSum, (ncclComm_t)ctx->nccl_comm, stream));
    return DLLM_OK;
}

// params_dev = {scale, zp, min, max} of this rank's shard -> min / max over all ranks of the group (two 1-float all-reduces)
int32_t tp_allreduce_minmax(dllm_ctx *ctx, float *params_dev) {
    if (ctx->tp_world <= 1) return DLLM_OK;
    if (!ctx->nccl_comm) DLLM_FAIL(ctx, DLLM_ERR_NCCL, "dllm_tp_init has not been called");
    NCCL_TRY(ctx, ncclGroupStart());
    NCCL_TRY(ctx, ncclAllReduce(params_dev + 2, params_dev + 2, 1, ncclFloat32, ncclMin, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    NCCL_TRY(ctx, ncclAllReduce(params_dev + 3, params_dev + 3, 1, ncclFloat32, ncclMax, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    NCCL_TRY(ctx, ncclGroupEnd());
    return DLLM_OK;
}

namespace {
// gathered [world][M][n_local] -> out [M][world*n_local]
__global__ void interleave_cols_kernel(const float *__restrict__ gathered, size_t M, size_t n_local, int world,
                                       float *__restrict__ out) {
    const size_t total = (size_t)world * M * n_local;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / (M * n_local), rem = i - r * M * n_local;
        const size_t m = rem / n_local, c = rem - m * n_local;
        out[m * ((size_t)world * n_local) + r * n_local + c] = gathered[i];
    }
}
}  // namespace

int32_t tp_allgather_cols(dllm_ctx *ctx, const float *in, size_t M, size_t n_local, float *out) {
    const size_t n = M * n_local;
    if (ctx->tp_world <= 1) {
        if (n && in != out) CUDA_TRY(ctx, cudaMemcpyAsync(out, in, n * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
        return DLLM_OK;
    }
    if (!ctx->nccl_comm) DLLM_FAIL(ctx, DLLM_ERR_NCCL, "dllm_tp_init has not been called");
    if (n == 0) return DLLM_OK;
    // own scratch: ws[7] is the noise_pred buffer of dllm_denoise_step_dev, i.e. possibly `out` itself
    DLLM_TRY(ensure_buf(ctx, ctx->tp_ws, (size_t)ctx->tp_world * n * sizeof(float)));
    float *gathered = (float *)ctx->tp_ws.p;
    NCCL_TRY(ctx, ncclAllGather(in, gathered, n, ncclFloat32, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    size_t blocks = ((size_t)ctx->tp_world * n + 255) / 256, cap = (size_t)ctx->sm_count * 16;
    interleave_cols_kernel<<<(unsigned)(blocks > cap ? cap : blocks), 256, 0, ctx->stream>>>(gathered, M, n_local, ctx->tp_world, out);
    LAUNCH_CHECK(ctx);
    return DLLM_OK;
}

extern "C" {

int32_t dllm_tp_unique_id(uint8_t id_out[128]) {
    if (!id_out) return DLLM_ERR_NULL;
    ncclUniqueId id;
    if (ncclGetUniqueId(&id) != ncclSuccess) return DLLM_ERR_NCCL;
    memcpy(id_out, &id, 128);
    return DLLM_OK;
}

int32_t dllm_tp_init(dllm_ctx *ctx, const uint8_t id[128], int32_t rank, int32_t world) {
    if (!ctx) return DLLM_ERR_NULL;
    ctx->err[0] = 0;
    if (!id || world < 1 || rank < 0 || rank >= world) DLLM_FAIL(ctx, DLLM_ERR_INVALID_PARAMS, "bad rank %d / world %d", rank, world);
    if (ctx->nccl_comm) DLLM_FAIL(ctx, DLLM_ERR_INVALID_PARAMS, "tensor-parallel group already initialised");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId uid;
    memcpy(&uid, id, 128);
    ncclComm_t comm;
    NCCL_TRY(ctx, ncclCommInitRank(&comm, world, uid, rank));
    ctx->nccl_comm = (void *)comm;
    ctx->tp_rank = rank;
    ctx->tp_world = world;
    return DLLM_OK;
}

int32_t dllm_tp_finalize(dllm_ctx *ctx) {
    if (!ctx) return DLLM_ERR_NULL;
    if (ctx->nccl_comm) {
        ncclCommDestroy((ncclComm_t)ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
    ctx->tp_rank = 0;
    ctx->tp_world = 1;
    return DLLM_OK;
}

int32_t dllm_tp_configure(dllm_ctx *ctx, int32_t chunks, int32_t reserve_sms, int32_t skip_comm) {
    if (!ctx) return DLLM_ERR_NULL;
    ctx->err[0] = 0;
    if (chunks < 0 || chunks > 16 || reserve_sms < -1 || reserve_sms > ctx->sm_count / 2)
        DLLM_FAIL(ctx, DLLM_ERR_INVALID_PARAMS, "chunks in 0..16 (0 = default), reserve_sms in -1..%d (-1 = default)", ctx->sm_count / 2);
    ctx->tp_chunks = chunks;
    ctx->sm_reserve = reserve_sms;
    ctx->tp_skip_comm = skip_comm != 0;
    return DLLM_OK;
}

int32_t dllm_tp_allreduce_dev(dllm_ctx *ctx, float *buf_dev, size_t n) {
    if (!ctx) return DLLM_ERR_NULL;
    ctx->err[0] = 0;
    return tp_allreduce(ctx, buf_dev, n);
}

int32_t dllm_tp_allgather_cols_dev(dllm_ctx *ctx, const float *in_dev, size_t M, size_t n_local, float *out_dev) {
    if (!ctx) return DLLM_ERR_NULL;
    ctx->err[0] = 0;
    return tp_allgather_cols(ctx, in_dev, M, n_local, out_dev);
}

}  // extern "C"
