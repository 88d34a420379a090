// Legacy tensor-path issue rates on sm_100a: HMMA.16816.F32 vs IMMA.16832.U8.S8 (cycles per instruction per SM sub-core).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int KIND, int CHAINS>
__global__ void rate_kernel(int iters, unsigned long long *out, int *sink) {
    uint32_t a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
    float f[CHAINS][4];
    int d[CHAINS][4];
    for (int c = 0; c < CHAINS; ++c) for (int i = 0; i < 4; ++i) { f[c][i] = 0.f; d[c][i] = 0; }
    __syncthreads();
    unsigned long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(f[c][0]), "+f"(f[c][1]), "+f"(f[c][2]), "+f"(f[c][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else
                asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+r"(d[c][0]), "+r"(d[c][1]), "+r"(d[c][2]), "+r"(d[c][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
    }
    unsigned long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    int s = 0;
    for (int c = 0; c < CHAINS; ++c) for (int i = 0; i < 4; ++i) s += d[c][i] + (int)f[c][i];
    if (s == 0x7fffffff) *sink = s;
}

template <int KIND, int CHAINS>
void run(const char *name, int warps) {
    unsigned long long *out; int *sink;
    cudaMalloc(&out, 148 * 8); cudaMalloc(&sink, 4);
    const int iters = 2000;
    rate_kernel<KIND, CHAINS><<<148, warps * 32>>>(iters, out, sink);
    rate_kernel<KIND, CHAINS><<<148, warps * 32>>>(iters, out, sink);
    unsigned long long h[148];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    cudaError_t e = cudaDeviceSynchronize();
    double cyc = (double)h[0];
    double per_subcore = cyc / ((double)iters * CHAINS * warps / 4.0);
    printf("%-6s chains %d warps/SM %2d: %.0f cycles, %.2f cyc per instr per sub-core (%.2f per warp-instr serial)  %s\n", name, CHAINS, warps, cyc,
           per_subcore, cyc / ((double)iters * CHAINS), e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(out); cudaFree(sink);
}

int main() {
    for (int w : {4, 8, 24}) {
        run<0, 1>("HMMA", w); run<0, 4>("HMMA", w);
        run<1, 1>("IMMA", w); run<1, 4>("IMMA", w);
    }
    return 0;
}
