// Bright-field flat-field correction for sm_100a: per-pixel median over the scan axis and the
// scale field derived from it.
//
// Reference: _LabelfreePreprocessor._flat_field_BF (shrimpy/preprocessing.py:385-404):
//     static_pattern = volume.quantile(0.5, dim=0)          # per-pixel median over Z, numpy.median semantics
//     return volume / static_pattern * static_pattern.mean()
// Here the correction is expressed as a per-pixel scale  s[y,x] = mean(pattern) / pattern[y,x]  so that it
// can be fused into the deskew kernel (the deskew interpolates along z only, so scaling commutes with it).
//
// median_z_kernel: exact per-pixel median by byte-wise radix selection with shared-memory histograms.  A CTA owns
// 64 consecutive pixels of one image row; per pass every thread walks its quarter of the scan axis (loads are
// coalesced along x: the 64 pixels of a slice are 128 / 256 contiguous bytes) and counts one key byte of the keys
// that still match the decided prefix into hist[byte][pixel] (the 32 lanes of a warp hit 32 different banks, so
// the shared atomics never conflict).  Four threads per pixel then locate the bin holding the wanted rank.
// uint16 needs 2 passes, float32 4; the slab of a CTA (Z x 64 pixels) is re-read from L1/L2, DRAM sees it once.
// For even Z the upper middle value is the same key when it occurs often enough, else the smallest key above it
// (one more min pass, only for CTAs that need it); the two are averaged (numpy.median / quantile(0.5) linear).
// Any Z is accepted (the first version held a column in registers and stopped at Z = 1280).
#include "common.cuh"

#include <algorithm>
#include <cfloat>

namespace shrimpy {

constexpr int kMedThreads = 256;
constexpr int kMedPixels = 64;                      // pixels per CTA
constexpr int kMedQuarters = kMedThreads / kMedPixels;

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

template <typename T> struct Pair;
template <> struct Pair<uint16_t> { typedef ushort2 type; };
template <> struct Pair<float> { typedef float2 type; };

// PAIR: a thread loads two adjacent pixels with one 4 / 8 byte access (needs the alignment the host checks).
// Per-pixel shared arrays are indexed by SLOT: pixel 2j lives in slot j, pixel 2j+1 in slot 32 + j, so that both
// the counting phase (lane = pixel pair) and the search phase (lane = slot) touch 32 different banks per access.
template <typename T, bool PAIR>
__global__ void __launch_bounds__(kMedThreads) median_z_kernel(const T *__restrict__ raw, float *__restrict__ pattern,
                                                               int Z, int Y, int X, long long sz, long long sy, int tiles_x) {
    extern __shared__ __align__(16) unsigned hist[];           // [256 bins][64 slots]
    __shared__ unsigned part[kMedQuarters][kMedPixels];         // counts per quarter of the bins
    __shared__ unsigned s_prefix[kMedPixels];                   // decided high bytes of the wanted key
    __shared__ unsigned s_rank[kMedPixels];                     // rank still to be resolved among the matching keys
    __shared__ unsigned s_le[kMedPixels];                       // keys <= the selected key (after the last pass)
    __shared__ unsigned s_above[kMedPixels];                    // smallest key above the selected one (0 = not known yet)

    constexpr int NB = sizeof(T);                               // key bytes: 2 or 4
    constexpr int PPT = PAIR ? 2 : 1;                           // pixels per thread while counting
    constexpr int kSlices = kMedThreads * PPT / kMedPixels;     // z slices the counting threads split the scan axis into
    const int y = blockIdx.x / tiles_x;
    const int x0 = (blockIdx.x % tiles_x) * kMedPixels;
    const unsigned k_lo = (unsigned)(Z - 1) / 2, k_hi = (unsigned)Z / 2;   // the two middle ranks (equal for odd Z)

    // counting role: pixel(s) cpx.. and z slice cq
    const int cpx = PAIR ? 2 * (threadIdx.x % 32) : threadIdx.x % kMedPixels;
    const int cq = PAIR ? threadIdx.x / 32 : threadIdx.x / kMedPixels;
    const int cslot0 = PAIR ? threadIdx.x % 32 : (cpx >> 1) + 32 * (cpx & 1);
    const bool clive0 = x0 + cpx < X, clive1 = PAIR && x0 + cpx + 1 < X;
    const T *col = raw + (long long)y * sy + x0 + cpx;
    // search role: slot sl (pixel spx), quarter q of the bins
    const int sl = threadIdx.x % kMedPixels, q = threadIdx.x / kMedPixels;
    const int spx = 2 * (sl % 32) + sl / 32;
    const bool slive = x0 + spx < X;

    for (int i = threadIdx.x; i < 256 * kMedPixels; i += kMedThreads) hist[i] = 0;
    if (q == 0) {
        s_prefix[sl] = 0;
        s_rank[sl] = k_lo;
        s_le[sl] = 0;
        s_above[sl] = 0;
    }
    __syncthreads();

#pragma unroll 1
    for (int pass = 0; pass < NB; ++pass) {
        const int shift = 8 * (NB - 1 - pass);
        const unsigned prefix0 = s_prefix[cslot0], prefix1 = PAIR ? s_prefix[cslot0 + 32] : 0u;
        auto count = [&](T v, unsigned prefix, int slot) {
            const uint32_t key = key_of(v);
            if (pass == 0 || (key >> (shift + 8)) == prefix) atomicAdd(&hist[((key >> shift) & 255u) * kMedPixels + slot], 1u);
        };
        if (clive0) {
            // batches of independent loads keep enough bytes in flight (a plain loop was latency-bound)
            constexpr int kBatch = 8;
            int z = cq;
            for (; z + (kBatch - 1) * kSlices < Z; z += kBatch * kSlices) {
                if (PAIR) {
                    typename Pair<T>::type v[kBatch];
#pragma unroll
                    for (int i = 0; i < kBatch; ++i)
                        v[i] = __ldg(reinterpret_cast<const typename Pair<T>::type *>(col + (long long)(z + i * kSlices) * sz));
#pragma unroll
                    for (int i = 0; i < kBatch; ++i) {
                        count(v[i].x, prefix0, cslot0);
                        if (clive1) count(v[i].y, prefix1, cslot0 + 32);
                    }
                } else {
                    T v[kBatch];
#pragma unroll
                    for (int i = 0; i < kBatch; ++i) v[i] = __ldg(col + (long long)(z + i * kSlices) * sz);
#pragma unroll
                    for (int i = 0; i < kBatch; ++i) count(v[i], prefix0, cslot0);
                }
            }
            for (; z < Z; z += kSlices) {
                count(__ldg(col + (long long)z * sz), prefix0, cslot0);
                if (clive1) count(__ldg(col + (long long)z * sz + 1), prefix1, cslot0 + 32);
            }
        }
        __syncthreads();
        // four threads per pixel: counts of their 64 bins, then the quarter holding the rank scans again
        unsigned mine = 0;
        for (int b = 0; b < 64; ++b) mine += hist[(q * 64 + b) * kMedPixels + sl];
        part[q][sl] = mine;
        __syncthreads();
        unsigned before = 0;
        for (int i = 0; i < q; ++i) before += part[i][sl];
        const unsigned rank = s_rank[sl], prefix = s_prefix[sl];
        __syncthreads();                       // everyone has read s_rank / s_prefix before the owner rewrites them
        if (rank >= before && rank < before + mine) {
            unsigned acc = before;
            int bin = q * 64;
            for (;; ++bin) {
                const unsigned c = hist[bin * kMedPixels + sl];
                if (rank < acc + c) {
                    s_prefix[sl] = (prefix << 8) | (unsigned)bin;
                    s_rank[sl] = rank - acc;
                    s_le[sl] += acc + (pass == NB - 1 ? c : 0u);   // keys below the bin (all passes) + the bin itself (last)
                    break;
                }
                acc += c;
            }
            if (pass == NB - 1 && k_hi != k_lo) {
                // the smallest key above the selected one is the next occupied bin of this pass, if there is one
                for (int nb = bin + 1; nb < 256; ++nb)
                    if (hist[nb * kMedPixels + sl]) {
                        s_above[sl] = ((prefix << 8) | (unsigned)nb) + 1u;   // stored + 1: 0 means "not found"
                        break;
                    }
            }
        }
        __syncthreads();
        if (pass + 1 < NB) {
            for (int i = threadIdx.x; i < 256 * kMedPixels; i += kMedThreads) hist[i] = 0;
            __syncthreads();
        }
    }

    if (k_hi != k_lo) {
        // upper middle value: the same key if enough keys are <= it, else the smallest key above it; that one is
        // known from the last histogram unless it differs in a higher byte -- then one min pass over the column
        const bool need0 = clive0 && s_le[cslot0] <= k_hi && s_above[cslot0] == 0;
        const bool need1 = clive1 && s_le[cslot0 + 32] <= k_hi && s_above[cslot0 + 32] == 0;
        if (__syncthreads_or(need0 || need1)) {
            if (q == 0 && s_above[sl] == 0) s_above[sl] = 0xffffffffu;
            __syncthreads();
            if (need0 || need1) {
                const uint32_t m0 = s_prefix[cslot0], m1 = PAIR ? s_prefix[cslot0 + 32] : 0u;
                uint32_t best0 = 0xffffffffu, best1 = 0xffffffffu;
                for (int z = cq; z < Z; z += kSlices) {
                    if (need0) {
                        const uint32_t key = key_of(__ldg(col + (long long)z * sz));
                        if (key > m0) best0 = min(best0, key);
                    }
                    if (need1) {
                        const uint32_t key = key_of(__ldg(col + (long long)z * sz + 1));
                        if (key > m1) best1 = min(best1, key);
                    }
                }
                // stored + 1 like the histogram result (a key of 0xffffffff cannot be "above" anything that needs it)
                if (need0 && best0 != 0xffffffffu) atomicMin(&s_above[cslot0], best0 + 1u);
                if (need1 && best1 != 0xffffffffu) atomicMin(&s_above[cslot0 + 32], best1 + 1u);
            }
            __syncthreads();
        }
    }
    if (q == 0 && slive) {
        const uint32_t m_lo = s_prefix[sl];
        const uint32_t m_hi = (k_hi == k_lo || s_le[sl] > k_hi) ? m_lo : s_above[sl] - 1u;
        // numpy: mean of the two middle values (float32 data stays float32)
        const float lo = value_of<T>(m_lo), hi = value_of<T>(m_hi);
        const float med = (k_hi == k_lo) ? lo : (sizeof(T) == 2) ? 0.5f * (lo + hi) : __fmul_rn(__fadd_rn(lo, hi), 0.5f);
        pattern[(long long)y * X + x0 + spx] = med;
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
    const long long sy = raw_stride_y ? raw_stride_y : X;
    const long long sz = raw_stride_z ? raw_stride_z : (long long)Y * sy;
    const int tiles_x = (X + kMedPixels - 1) / kMedPixels;
    const long long blocks = (long long)tiles_x * Y;
    if (blocks > 2147483647LL) return fail(SHRIMPY_EINVAL, "flatfield: image too large for the grid");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t smem = 256 * kMedPixels * sizeof(unsigned);   // 64 KB of histograms: three CTAs per SM
    const int es = raw_dtype == SHRIMPY_U16 ? 2 : 4;
    // two pixels per load when every pair is naturally aligned: even strides, even X tile origin (always), aligned base
    const bool pair = (reinterpret_cast<uintptr_t>(d_raw) % (2 * es)) == 0 && sy % 2 == 0 && sz % 2 == 0 && X % 2 == 0;
#define SHRIMPY_MEDIAN(T, P)                                                                                         \
    do {                                                                                                             \
        auto kern = median_z_kernel<T, P>;                                                                           \
        SHRIMPY_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
        kern<<<(unsigned)blocks, kMedThreads, smem, s>>>(static_cast<const T *>(d_raw), d_pattern, Z, Y, X, sz, sy, tiles_x); \
    } while (0)
    if (raw_dtype == SHRIMPY_U16) {
        if (pair) SHRIMPY_MEDIAN(uint16_t, true);
        else SHRIMPY_MEDIAN(uint16_t, false);
    } else {
        if (pair) SHRIMPY_MEDIAN(float, true);
        else SHRIMPY_MEDIAN(float, false);
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
