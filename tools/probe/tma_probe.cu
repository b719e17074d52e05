// Stand-alone probe: which 3-D TMA box loads work on this GPU/driver?  (developer tool)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap tmap, float *out, int c0, int c1, int c2, unsigned bytes,
                      int n) {
    extern __shared__ __align__(1024) float buf[];
    __shared__ __align__(8) uint64_t bar;
    uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(&bar);
    uint32_t buf_s = (uint32_t)__cvta_generic_to_shared(buf);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
            ::"r"(buf_s), "l"((uint64_t)&tmap), "r"(c0), "r"(c1), "r"(c2), "r"(bar_s) : "memory");
    }
    __syncthreads();
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(bar_s) : "memory");
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = buf[i];
}

int main() {
    void *sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)sym;
    const int X = 256, Y = 256, Z = 32;
    float *in, *out;
    cudaMalloc(&in, sizeof(float) * X * Y * Z);
    cudaMalloc(&out, sizeof(float) * 65536);
    float *h = (float *)malloc(sizeof(float) * X * Y * Z);
    for (int i = 0; i < X * Y * Z; ++i) h[i] = (float)i;
    cudaMemcpy(in, h, sizeof(float) * X * Y * Z, cudaMemcpyHostToDevice);
    struct Case { int bx, by, bz; CUtensorMapSwizzle sw; int c0, c1, c2; const char *name; } cases[] = {
        {32, 1, 8, CU_TENSOR_MAP_SWIZZLE_128B, 0, 0, 0, "swz128 32x1x8 @0"},
        {32, 4, 8, CU_TENSOR_MAP_SWIZZLE_128B, 0, 0, 0, "swz128 32x4x8 @0"},
        {32, 1, 8, CU_TENSOR_MAP_SWIZZLE_NONE, 0, 0, 0, "none 32x1x8 @0"},
        {68, 12, 6, CU_TENSOR_MAP_SWIZZLE_NONE, 0, 0, 0, "none 68x12x6 @0"},
        {68, 12, 6, CU_TENSOR_MAP_SWIZZLE_NONE, 4, 7, 3, "none 68x12x6 @4,7,3"},
        {68, 12, 6, CU_TENSOR_MAP_SWIZZLE_NONE, -4, 0, 0, "none 68x12x6 @-4,0,0"},
        {68, 12, 6, CU_TENSOR_MAP_SWIZZLE_NONE, 0, -2, -1, "none 68x12x6 @0,-2,-1"},
        {68, 12, 6, CU_TENSOR_MAP_SWIZZLE_NONE, 248, 250, 30, "none 68x12x6 @248,250,30 (high OOB)"},
        {68, 12, 6, CU_TENSOR_MAP_SWIZZLE_NONE, 256, 256, 32, "none 68x12x6 fully OOB"},
        {32, 1, 8, CU_TENSOR_MAP_SWIZZLE_128B, -4, 0, -2, "swz128 32x1x8 @-4,0,-2"},
        {32, 1, 8, CU_TENSOR_MAP_SWIZZLE_128B, 2, 0, 0, "swz128 32x1x8 @2,0,0 (unaligned x)"},
    };
    for (auto &c : cases) {
        CUtensorMap tm;
        cuuint64_t gdim[3] = {X, Y, Z};
        cuuint64_t gstr[2] = {X * 4ull, X * Y * 4ull};
        cuuint32_t box[3] = {(cuuint32_t)c.bx, (cuuint32_t)c.by, (cuuint32_t)c.bz};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult rc = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, in, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          c.sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc != CUDA_SUCCESS) { printf("%-40s encode failed %d\n", c.name, (int)rc); continue; }
        int n = c.bx * c.by * c.bz;
        probe<<<1, 128, n * 4 + 1024>>>(tm, out, c.c0, c.c1, c.c2, (unsigned)n * 4u, n);
        cudaError_t e = cudaDeviceSynchronize();
        float first = -1, last = -1;
        if (e == cudaSuccess) {
            cudaMemcpy(&first, out, 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(&last, out + n - 1, 4, cudaMemcpyDeviceToHost);
        }
        printf("%-40s %s first=%g last=%g\n", c.name, cudaGetErrorString(e), first, last);
        if (e != cudaSuccess) { printf("sticky error, stopping\n"); break; }
    }
    return 0;
}
