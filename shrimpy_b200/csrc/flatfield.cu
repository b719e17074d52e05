// Bright-field flat-field correction for sm_100a: per-pixel median over the scan axis and the
// scale field derived from it.
//
// Reference: _LabelfreePreprocessor._flat_field_BF (shrimpy/preprocessing.py:385-404):
//     static_pattern = volume.quantile(0.5, dim=0)          # per-pixel median over Z, numpy.median semantics
//     return volume / static_pattern * static_pattern.mean()
// Here the correction is expressed as a per-pixel scale  s[y,x] = mean(pattern) / pattern[y,x]  so that it
// can be fused into the deskew kernel (the deskew interpolates along z only, so scaling commutes with it).
//
// median_z_kernel: a CTA stages a [Z][64 x] tile (row pitch padded by one bank), then each warp runs an
// exact radix select per pixel column: every lane keeps ceil(Z/32) values in registers and the warp
// narrows the answer one bit per step with a ballot-free count (popc of per-lane compares + warp add).
// For even Z both middle order statistics are found and averaged (numpy.median / quantile(0.5) linear).
#include "common.cuh"

#include <algorithm>
#include <cfloat>

namespace shrimpy {

constexpr int kMedThreads = 256;
constexpr int kMedTileX = 64;
constexpr int kMedMaxPerLane = 40;   // Z <= 1280

// order-preserving key: uint16 as is; float32 with the usual sign fix-up
__device__ __forceinline__ uint32_t key_of(uint16_t v) { return v; }
__device__ __forceinline__ uint32_t key_of(float v) {
    const uint32_t b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
template <typename T> __device__ __forceinline__ float value_of(uint32_t k);
template <> __device__ __forceinline__ float value_of<uint16_t>(uint32_t k) { return (float)k; }
template <> __device__ __forceinline__ float value_of<float>(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// k-th smallest (0-based) of the warp's keys; lanes hold `cnt` valid keys each in v[0..cnt)
template <int NBITS, int MAXV>
__device__ __forceinline__ uint32_t warp_select(const uint32_t (&v)[MAXV], int k) {
    uint32_t prefix = 0;
#pragma unroll 1
    for (int bit = NBITS - 1; bit >= 0; --bit) {
        // keys that match the decided prefix above `bit` and have this bit clear: (v & mask) == prefix, where
        // mask covers `bit` and everything above it (prefix is still zero at `bit` and below; padding keys are
        // all-ones and never match while a bit is undecided)
        const uint32_t mask = 0xffffffffu << bit;
        int c = 0;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) c += ((v[i] & mask) == prefix) ? 1 : 0;
        c = __reduce_add_sync(0xffffffffu, c);
        if (k >= c) {
            k -= c;
            prefix |= 1u << bit;
        }
    }
    return prefix;
}

template <typename T, int MAXV>   // MAXV = registers per lane holding column values: Z <= 32 * MAXV
__global__ void __launch_bounds__(kMedThreads) median_z_kernel(const T *__restrict__ raw, float *__restrict__ pattern,
                                                               int Z, int Y, int X, long long sz, long long sy,
                                                               int tiles_x, int tile_x, int pitch_elems) {
    extern __shared__ __align__(16) unsigned char smem_med[];
    T *tile = reinterpret_cast<T *>(smem_med);
    const int y = blockIdx.x / tiles_x;
    const int x0 = (blockIdx.x % tiles_x) * tile_x;
    const int nx = min(tile_x, X - x0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // stage: one scan slice row (nx elements) per warp iteration, coalesced along x
    for (int z = warp; z < Z; z += kMedThreads / 32) {
        const T *src = raw + (long long)z * sz + (long long)y * sy + x0;
        for (int i = lane; i < nx; i += 32) tile[z * pitch_elems + i] = __ldg(src + i);
    }
    __syncthreads();

    constexpr int NBITS = sizeof(T) == 2 ? 16 : 32;
    const int k_hi = Z / 2, k_lo = (Z - 1) / 2;      // the two middle ranks (equal for odd Z)
    for (int col = warp; col < nx; col += kMedThreads / 32) {
        uint32_t v[MAXV];
        int cnt = 0;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int z = lane + 32 * i;
            v[i] = 0xffffffffu;
            if (z < Z) {
                v[i] = key_of(tile[z * pitch_elems + col]);
                cnt = i + 1;
            }
        }
        const uint32_t m_lo = warp_select<NBITS, MAXV>(v, k_lo);
        float med = value_of<T>(m_lo);
        if (k_hi != k_lo) {
            // next order statistic: m_lo again if it occurs often enough, else the smallest key above it
            int le = 0;
            uint32_t above = 0xffffffffu;
#pragma unroll
            for (int i = 0; i < MAXV; ++i)
                if (i < cnt) {
                    le += v[i] <= m_lo ? 1 : 0;
                    if (v[i] > m_lo) above = min(above, v[i]);
                }
            le = __reduce_add_sync(0xffffffffu, le);
            above = __reduce_min_sync(0xffffffffu, above);
            const uint32_t m_hi = (le > k_hi) ? m_lo : above;
            // numpy: mean of the two middle values (float32 data stays float32)
            med = (sizeof(T) == 2) ? 0.5f * (value_of<T>(m_lo) + value_of<T>(m_hi))
                                   : __fmul_rn(__fadd_rn(value_of<T>(m_lo), value_of<T>(m_hi)), 0.5f);
        }
        if (lane == 0) pattern[(long long)y * X + x0 + col] = med;
    }
}

// sum of the pattern in float64 (one atomicAdd per CTA), then scale = mean / pattern
__global__ void __launch_bounds__(256) pattern_sum_kernel(const float *__restrict__ pattern, long long n, double *sum) {
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        acc += (double)pattern[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ double part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 8; ++i) t += part[i];
        atomicAdd(sum, t);
    }
}

__global__ void __launch_bounds__(256) pattern_scale_kernel(const float *__restrict__ pattern, long long n,
                                                            const double *sum, float *__restrict__ scale) {
    const float mean = (float)(*sum / (double)n);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        scale[i] = mean / pattern[i];
}

// standalone correction (when the deskew does not follow): out = raw * scale[y,x], float32
template <typename T>
__global__ void __launch_bounds__(256) apply_scale_kernel(const T *__restrict__ raw, const float *__restrict__ scale,
                                                          float *__restrict__ out, long long plane, long long total) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
        __stcs(out + i, (float)__ldg(raw + i) * __ldg(scale + i % plane));
}

}  // namespace shrimpy

using namespace shrimpy;

extern "C" int shrimpy_flatfield_pattern_device(const void *d_raw, int raw_dtype, float *d_pattern, int Z, int Y, int X,
                                                int64_t raw_stride_z, int64_t raw_stride_y, void *stream) {
    if (Z <= 0 || Y <= 0 || X <= 0) return fail(SHRIMPY_EINVAL, "flatfield: bad shape (%d,%d,%d)", Z, Y, X);
    if (raw_dtype != SHRIMPY_U16 && raw_dtype != SHRIMPY_F32) return fail(SHRIMPY_EINVAL, "flatfield: bad dtype");
    if (!d_raw || !d_pattern) return fail(SHRIMPY_EINVAL, "flatfield: null device pointer");
    if (Z > 32 * kMedMaxPerLane)
        return fail(SHRIMPY_EINVAL, "flatfield: Z=%d exceeds the %d slices the median kernel holds in registers", Z,
                    32 * kMedMaxPerLane);
    const long long sy = raw_stride_y ? raw_stride_y : X;
    const long long sz = raw_stride_z ? raw_stride_z : (long long)Y * sy;
    const int es = raw_dtype == SHRIMPY_U16 ? 2 : 4;
    // tile width: 64 pixels when the [Z][tile] slab fits shared memory (two CTAs per SM preferred), else narrower;
    // row pitch: the tile plus one 4-byte bank so that a column walk hits distinct banks
    int tile_x = kMedTileX;
    while (tile_x > 16 && (size_t)Z * (tile_x + 4 / es) * es > 110 * 1024) tile_x >>= 1;
    const int pitch_elems = tile_x + 4 / es;
    const size_t smem = (size_t)Z * pitch_elems * es;
    if (smem > 220 * 1024) return fail(SHRIMPY_EINVAL, "flatfield: Z=%d does not fit the shared-memory tile", Z);
    const int tiles_x = (X + tile_x - 1) / tile_x;
    const long long blocks = (long long)tiles_x * Y;
    if (blocks > 2147483647LL) return fail(SHRIMPY_EINVAL, "flatfield: image too large for the grid");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int per_lane = (Z + 31) / 32;
#define SHRIMPY_MEDIAN(T, V)                                                                                        \
    do {                                                                                                            \
        auto kern = median_z_kernel<T, V>;                                                                          \
        if (smem + 1024 > 48 * 1024)                                                                                \
            SHRIMPY_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
        kern<<<(unsigned)blocks, kMedThreads, smem, s>>>(static_cast<const T *>(d_raw), d_pattern, Z, Y, X, sz, sy, \
                                                          tiles_x, tile_x, pitch_elems);                             \
    } while (0)
    if (raw_dtype == SHRIMPY_U16) {
        if (per_lane <= 8) SHRIMPY_MEDIAN(uint16_t, 8);
        else if (per_lane <= 20) SHRIMPY_MEDIAN(uint16_t, 20);
        else SHRIMPY_MEDIAN(uint16_t, kMedMaxPerLane);
    } else {
        if (per_lane <= 8) SHRIMPY_MEDIAN(float, 8);
        else if (per_lane <= 20) SHRIMPY_MEDIAN(float, 20);
        else SHRIMPY_MEDIAN(float, kMedMaxPerLane);
    }
#undef SHRIMPY_MEDIAN
    count_launch();
    SHRIMPY_CUDA_TRY(cudaGetLastError());
    return SHRIMPY_OK;
}

extern "C" int shrimpy_flatfield_scale_device(const float *d_pattern, int64_t count, float *d_scale, double *d_scratch,
                                              void *stream) {
    if (!d_pattern || !d_scale || !d_scratch || count <= 0) return fail(SHRIMPY_EINVAL, "flatfield scale: bad arguments");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    SHRIMPY_CUDA_TRY(cudaMemsetAsync(d_scratch, 0, sizeof(double), s));
    const int blocks = (int)std::min<long long>((count + 255) / 256, 1184);
    pattern_sum_kernel<<<blocks, 256, 0, s>>>(d_pattern, count, d_scratch);
    pattern_scale_kernel<<<blocks, 256, 0, s>>>(d_pattern, count, d_scratch, d_scale);
    count_launch(2);
    SHRIMPY_CUDA_TRY(cudaGetLastError());
    return SHRIMPY_OK;
}

extern "C" int shrimpy_flatfield_apply_device(const void *d_raw, int raw_dtype, const float *d_scale, float *d_out, int Z,
                                              int Y, int X, void *stream) {
    if (Z <= 0 || Y <= 0 || X <= 0 || !d_raw || !d_scale || !d_out) return fail(SHRIMPY_EINVAL, "flatfield apply: bad arguments");
    if (raw_dtype != SHRIMPY_U16 && raw_dtype != SHRIMPY_F32) return fail(SHRIMPY_EINVAL, "flatfield apply: bad dtype");
    const long long plane = (long long)Y * X, total = plane * Z;
    const int blocks = (int)std::min<long long>((total + 255) / 256, 148LL * 32);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (raw_dtype == SHRIMPY_U16)
        apply_scale_kernel<uint16_t><<<blocks, 256, 0, s>>>(static_cast<const uint16_t *>(d_raw), d_scale, d_out, plane, total);
    else
        apply_scale_kernel<float><<<blocks, 256, 0, s>>>(static_cast<const float *>(d_raw), d_scale, d_out, plane, total);
    count_launch();
    SHRIMPY_CUDA_TRY(cudaGetLastError());
    return SHRIMPY_OK;
}
