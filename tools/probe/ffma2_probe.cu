// Throughput probe: packed FFMA2 vs scalar FFMA on sm_100a (issue cycles per warp instruction).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float *out, int iters, float s) {
    float2 a[8];
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f - i);
    const float2 w = make_float2(s, s * 0.5f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { a[i].x = fmaf(a[i].x, w.x, w.y); a[i].y = fmaf(a[i].y, w.x, w.y); }
            else if (MODE == 1) a[i] = __ffma2_rn(a[i], w, w);
            else a[i] = __fadd2_rn(a[i], w);
        }
    }
    float r = 0;
    for (int i = 0; i < 8; ++i) r += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> void run(const char *name) {
    float *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    k<MODE><<<148 * 8, 256>>>(d, 100, 1.0001f);
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(d, iters, 1.0001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // warp instructions per SMSP: 8 blocks * 8 warps / 4 SMSP = 16 warps per SMSP
    const double inst = (double)iters * 8 * (MODE == 0 ? 2 : 1) * 16;
    printf("%s: %.3f ms, %.2f ns per warp-instr per SMSP (%.2f cycles at 1.9 GHz), flops %.1f TF\n", name, ms,
           ms * 1e6 / inst, ms * 1e6 / inst * 1.9, (double)iters * 8 * 2 * (MODE == 2 ? 1 : 2) * 148 * 8 * 256 / ms / 1e9);
}
int main() { run<0>("FFMA scalar"); run<1>("FFMA2 packed"); run<2>("FADD2 packed"); return 0; }
