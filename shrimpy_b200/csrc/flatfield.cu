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
// 64 consecutive pixels of one image row; per pass every thread walks its share of the scan axis (loads are
// coalesced along x: the 64 pixels of a slice are 128 / 256 contiguous bytes) and counts one key byte of the keys
// that still match the decided prefix into hist[byte][pixel] (the 32 lanes of a warp hit 32 different banks, so
// the shared atomics never conflict).  uint16 needs 2 passes, float32 4; the slab of a CTA (Z x 64 pixels) is re-read
// from L1/L2, DRAM sees it once.  The bin holding the wanted rank is located by a two-level search on the packed
// counters (warp g sums bins 32g..32g+31 of both pixels of a word at once; the group that holds the rank scans its
// 32 bins).  For even Z the upper middle value is the same key when it occurs often enough, else the smallest key
// above it: the next occupied bin of the last histogram, or -- when that lies in a higher prefix -- the minimum of
// the keys above the prefix's range, which the last counting pass tracks in a register (two instructions per key; the
// first version re-walked the column for it).  The two middle values are averaged (numpy.median / quantile(0.5)).
// Z up to 65535 (the first version held a column in registers and stopped at Z = 1280).
#include "common.cuh"

#include <algorithm>
#include <cfloat>
#include <cstring>

namespace shrimpy {

constexpr int kMedThreads = 256;
constexpr int kMedPixels = 64;                      // pixels per CTA

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

// Shared-memory reductions by address, unconditional and predicated.  Written in PTX because the C++ form
// `if (match) atomicAdd(...)` compiles to a branch per key that re-derives the shared window base inside it
// (9 instructions, 71 M warp instructions per launch); this is one predicated ATOMS.
__device__ __forceinline__ void red_shared(unsigned addr, unsigned val) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(val) : "memory");
}
__device__ __forceinline__ void red_shared_if(bool on, unsigned addr, unsigned val) {
    asm volatile(
        "{\n"
        " .reg .pred q;\n"
        " setp.ne.u32 q, %0, 0;\n"
        " @q red.shared.add.u32 [%1], %2;\n"
        "}" ::"r"((unsigned)on), "r"(addr), "r"(val) : "memory");
}

template <typename T> struct Pair;
template <> struct Pair<uint16_t> { typedef ushort2 type; };
template <> struct Pair<float> { typedef float2 type; };

// PAIR: a thread loads two adjacent pixels with one 4 / 8 byte access (needs the alignment the host checks).
// Histogram counters are 16 bits wide (Z <= 65535), two pixels per 32-bit word -- pixel 2j in the low half of word j,
// pixel 2j+1 in the high half -- so the 256 x 64 counters take 32 KB and six CTAs fit an SM (the kernel is latency-
// bound: more resident warps is what it needs).  A warp's 32 pixel pairs hit 32 different words per access.
template <typename T, bool PAIR>
__global__ void __launch_bounds__(kMedThreads) median_z_kernel(const T *__restrict__ raw, float *__restrict__ pattern,
                                                               int Z, int Y, int X, long long sz, long long sy, int tiles_x) {
    constexpr int kWords = kMedPixels / 2;
    constexpr int kGroups = kMedThreads / kWords;               // 8 warps, each owns 32 bins in the search
    constexpr int kGroupBins = 256 / kGroups;
    extern __shared__ __align__(16) unsigned hist[];           // [256 bins][32 words], two 16-bit counters per word
    __shared__ unsigned part[kGroups][kWords];                  // packed counts of every group of 32 bins
    __shared__ unsigned s_prefix[kMedPixels];                   // decided high bytes of the wanted key
    __shared__ unsigned s_rank[kMedPixels];                     // rank still to be resolved among the matching keys
    __shared__ unsigned s_le[kMedPixels];                       // keys <= the selected key (after the last pass)
    __shared__ unsigned s_next[kMedPixels];                     // next occupied key of the last histogram + 1 (0 = none)
    __shared__ unsigned s_far[kMedPixels];                      // min (key - limit) over keys above the last prefix's range

    constexpr int NB = sizeof(T);                               // key bytes: 2 or 4
    constexpr int PPT = PAIR ? 2 : 1;                           // pixels per thread while counting
    constexpr int kSlices = kMedThreads * PPT / kMedPixels;     // z slices the counting threads split the scan axis into
    const int y = blockIdx.x / tiles_x;
    const int x0 = (blockIdx.x % tiles_x) * kMedPixels;
    const unsigned k_lo = (unsigned)(Z - 1) / 2, k_hi = (unsigned)Z / 2;   // the two middle ranks (equal for odd Z)
    const bool even = k_hi != k_lo;

    // counting role: pixel(s) cpx (and cpx + 1), z slice cq
    const int cpx = PAIR ? 2 * (threadIdx.x % kWords) : threadIdx.x % kMedPixels;
    const int cq = PAIR ? threadIdx.x / kWords : threadIdx.x / kMedPixels;
    const bool clive0 = x0 + cpx < X, clive1 = PAIR && clive0;   // PAIR needs an even X, so pairs are never split by the edge
    const T *col = raw + (long long)y * sy + x0 + cpx;
    // search role: word w (pixels 2w, 2w + 1), bin group g (= the warp)
    const int w = threadIdx.x % kWords, g = threadIdx.x / kWords;

    for (int i = threadIdx.x; i < 256 * kWords / 4; i += kMedThreads) reinterpret_cast<uint4 *>(hist)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x < kMedPixels) {
        s_prefix[threadIdx.x] = 0;
        s_rank[threadIdx.x] = k_lo;
        s_le[threadIdx.x] = 0;
        s_next[threadIdx.x] = 0;
        s_far[threadIdx.x] = 0xffffffffu;
    }
    __syncthreads();

#pragma unroll
    for (int pass = 0; pass < NB; ++pass) {
        constexpr int kLastPass = NB - 1;
        const int shift = 8 * (NB - 1 - pass);
        const bool last = pass == kLastPass;                   // the last pass also tracks the smallest key above the prefix's range
        const bool track = last && even;
        const unsigned prefix0 = s_prefix[cpx], prefix1 = PAIR ? s_prefix[cpx + 1] : 0u;
        const unsigned lim0 = (prefix0 + 1u) << 8, lim1 = (prefix1 + 1u) << 8;   // 0 = no key can lie above (wrapped)
        unsigned far0 = 0xffffffffu, far1 = 0xffffffffu;
        // keys below the limit wrap around to values larger than any true difference, so the minimum needs no test
        const unsigned hbase = smem_u32(hist) + (unsigned)(cpx >> 1) * 4u;      // this thread's word of every bin row
        auto count = [&](T v, unsigned prefix, unsigned lim, unsigned &far, int px) {
            const uint32_t key = key_of(v);
            const unsigned addr = hbase + ((key >> shift) & 255u) * (unsigned)(kWords * 4);
            const unsigned one = 1u << (16 * (px & 1));
            if (pass == 0) red_shared(addr, one);
            else red_shared_if((key >> (shift + 8)) == prefix, addr, one);
            if (last) far = min(far, key - lim);          // (unconditionally: a test on `even` per key costs as much)
        };
        if (clive0) {
            // batches of independent loads keep enough bytes in flight (a plain loop was latency-bound)
            constexpr int kBatch = 8;
            int z = cq;
            for (; z + (kBatch - 1) * kSlices < Z; z += kBatch * kSlices) {
                if (PAIR) {
                    typename Pair<T>::type v[kBatch];
#pragma unroll
                    for (int i = 0; i < kBatch; ++i) {
                        const void *src = col + (long long)(z + i * kSlices) * sz;
                        if constexpr (sizeof(T) == 2) {
                            unsigned w32;   // 256-byte L2 prefetch: the neighbouring CTA reads the other 128 B of the line pair
                            asm volatile("ld.global.nc.L2::256B.b32 %0, [%1];" : "=r"(w32) : "l"(src));
                            memcpy(&v[i], &w32, 4);
                        } else {
                            v[i] = __ldg(reinterpret_cast<const typename Pair<T>::type *>(src));
                        }
                    }
                    if constexpr (sizeof(T) == 2) {
                        // uint16 pairs stay one 32-bit word: key bytes come out with one PRMT each, both prefixes are
                        // compared through one XOR (the generic path spends 7.5 / 11 instructions per key on this)
                        const unsigned both = (prefix0 << 8) | (prefix1 << 24);
                        // the same wrap-around minimum as `count`, on 16-bit halves with one DPX instruction per pair
                        const unsigned neglims = ((0u - lim0) & 0xffffu) | ((0u - lim1) << 16);
                        unsigned far2 = 0xffffffffu;
#pragma unroll
                        for (int i = 0; i < kBatch; ++i) {
                            unsigned word;
                            memcpy(&word, &v[i], 4);
                            if (pass == 0) {
                                red_shared(hbase + __byte_perm(word, 0, 0x4441) * (unsigned)(kWords * 4), 1u);
                                red_shared(hbase + __byte_perm(word, 0, 0x4443) * (unsigned)(kWords * 4), 0x10000u);
                            } else {
                                const unsigned diff = word ^ both;
                                red_shared_if((diff & 0x0000ff00u) == 0u, hbase + (word & 0xffu) * (unsigned)(kWords * 4), 1u);
                                red_shared_if((diff & 0xff000000u) == 0u,
                                              hbase + __byte_perm(word, 0, 0x4442) * (unsigned)(kWords * 4), 0x10000u);
                                far2 = __viaddmin_u16x2(word, neglims, far2);   // both halves: min(key - lim, far) mod 2^16
                            }
                        }
                        far0 = min(far0, far2 & 0xffffu);
                        far1 = min(far1, far2 >> 16);
                    } else {
#pragma unroll
                        for (int i = 0; i < kBatch; ++i) {
                            count(v[i].x, prefix0, lim0, far0, cpx);
                            count(v[i].y, prefix1, lim1, far1, cpx + 1);
                        }
                    }
                } else {
                    T v[kBatch];
#pragma unroll
                    for (int i = 0; i < kBatch; ++i) v[i] = __ldg(col + (long long)(z + i * kSlices) * sz);
#pragma unroll
                    for (int i = 0; i < kBatch; ++i) count(v[i], prefix0, lim0, far0, cpx);
                }
            }
            for (; z < Z; z += kSlices) {
                count(__ldg(col + (long long)z * sz), prefix0, lim0, far0, cpx);
                if (clive1) count(__ldg(col + (long long)z * sz + 1), prefix1, lim1, far1, cpx + 1);
            }
            if (track) {
                if (lim0 != 0u) atomicMin(&s_far[cpx], far0);
                if (clive1 && lim1 != 0u) atomicMin(&s_far[cpx + 1], far1);
            }
        }
        __syncthreads();

        // ---- search: which bin holds the rank?  Level 1: packed sums of the 32 bins of every group.
        unsigned mine = 0;
#pragma unroll 8
        for (int b = 0; b < kGroupBins; ++b) mine += hist[(g * kGroupBins + b) * kWords + w];   // halves cannot carry: <= Z
        const unsigned rank0 = s_rank[2 * w], rank1 = s_rank[2 * w + 1];
        const unsigned pre0 = s_prefix[2 * w], pre1 = s_prefix[2 * w + 1];
        part[g][w] = mine;
        __syncthreads();                       // part complete; everyone holds its copy of s_rank / s_prefix
        unsigned before = 0;
        for (int i = 0; i < g; ++i) before += part[i][w];
        const unsigned b0 = before & 0xffffu, b1 = before >> 16, m0 = mine & 0xffffu, m1 = mine >> 16;
        const bool own0 = rank0 >= b0 && rank0 < b0 + m0, own1 = rank1 >= b1 && rank1 < b1 + m1;
        if (own0 || own1) {
            // Level 2: the group that holds a pixel's rank scans its 32 bins, one pixel of the word after the other
            // (a joint walk with per-pixel flags cost 45 instructions per bin).  The walks run on a stepped pointer to the
            // pixel's 16-bit counter: load, add, compare, branch.
            constexpr int kStep = 2 * kWords;                                  // counters (uint16) from one bin to the next
            const unsigned short *counters = reinterpret_cast<const unsigned short *>(hist) + 2 * w;
            auto select = [&](int half, unsigned rank, unsigned before, unsigned pre) -> int {
                const unsigned short *p = counters + half + g * kGroupBins * kStep;
                unsigned acc = before, c = *p;
                while (rank >= acc + c) {
                    acc += c;
                    p += kStep;
                    c = *p;
                }
                const int bin = (int)((p - counters) / kStep);
                const int px = 2 * w + half;
                s_prefix[px] = (pre << 8) | (unsigned)bin;
                s_rank[px] = rank - acc;
                const unsigned le = s_le[px] + acc + (last ? c : 0u);         // keys below the bin (all passes) + the bin (last)
                s_le[px] = le;
                if (track && le <= k_hi) {
                    // the smallest key above the selected one inside this prefix: the next occupied bin -- the rest of
                    // this group, then the first later group whose packed count is not zero
                    int nb = -1;
                    const unsigned short *q = p + kStep;
                    for (int b = bin + 1; b < (g + 1) * kGroupBins; ++b, q += kStep)
                        if (*q) { nb = b; break; }
                    for (int gg = g + 1; nb < 0 && gg < kGroups; ++gg) {
                        if (((part[gg][w] >> (16 * half)) & 0xffffu) == 0u) continue;
                        q = counters + half + gg * kGroupBins * kStep;
                        int b = gg * kGroupBins;
                        while (*q == 0) { q += kStep; ++b; }
                        nb = b;
                    }
                    if (nb >= 0) s_next[px] = ((pre << 8) | (unsigned)nb) + 1u;
                }
                return bin;
            };
            if (own0) select(0, rank0, b0, pre0);
            if (own1) select(1, rank1, b1, pre1);
        }
        __syncthreads();
        if (pass + 1 < NB) {
            for (int i = threadIdx.x; i < 256 * kWords / 4; i += kMedThreads)
                reinterpret_cast<uint4 *>(hist)[i] = make_uint4(0, 0, 0, 0);
            __syncthreads();
        }
    }

    if (threadIdx.x < kMedPixels && x0 + (int)threadIdx.x < X) {
        const int px = threadIdx.x;
        const uint32_t m_lo = s_prefix[px];
        uint32_t m_hi = m_lo;                                   // odd Z, or enough keys <= m_lo
        if (even && s_le[px] <= k_hi)
            m_hi = s_next[px] ? s_next[px] - 1u : s_far[px] + (((m_lo >> 8) + 1u) << 8);
        // numpy: mean of the two middle values (float32 data stays float32)
        const float lo = value_of<T>(m_lo), hi = value_of<T>(m_hi);
        const float med = !even ? lo : (sizeof(T) == 2) ? 0.5f * (lo + hi) : __fmul_rn(__fadd_rn(lo, hi), 0.5f);
        pattern[(long long)y * X + x0 + px] = med;
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

// standalone correction (when the deskew does not follow): out = raw * scale[y,x], float32.
// A thread owns VEC consecutive pixels (16 bytes of input) and walks a range of slices, so the scale values are read
// once per thread and every access is a full 16-byte vector; VEC = 1 is the unaligned fallback.
template <typename T, int VEC>
__global__ void __launch_bounds__(256) apply_scale_kernel(const T *__restrict__ raw, const float *__restrict__ scale,
                                                          float *__restrict__ out, long long plane, int Z, int z_per_block) {
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (i >= plane) return;
    const int z0 = blockIdx.y * z_per_block, z1 = min(Z, z0 + z_per_block);
    float g[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) g[j] = __ldg(scale + i + j);
    for (int z = z0; z < z1; ++z) {
        const T *src = raw + (long long)z * plane + i;
        float *dst = out + (long long)z * plane + i;
        if (VEC == 1) {
            __stcs(dst, (float)__ldg(src) * g[0]);
        } else if (sizeof(T) == 2) {   // VEC == 8: one uint4 in, two float4 out
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(src));
            const unsigned w[4] = {v.x, v.y, v.z, v.w};
            float r[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                r[2 * j] = (float)(w[j] & 0xffffu) * g[2 * j];
                r[2 * j + 1] = (float)(w[j] >> 16) * g[2 * j + 1];
            }
            __stcs(reinterpret_cast<float4 *>(dst), make_float4(r[0], r[1], r[2], r[3]));
            __stcs(reinterpret_cast<float4 *>(dst) + 1, make_float4(r[4], r[5], r[6], r[7]));
        } else {                       // VEC == 4: one float4 in, one float4 out
            const float4 v = __ldg(reinterpret_cast<const float4 *>(src));
            __stcs(reinterpret_cast<float4 *>(dst), make_float4(v.x * g[0], v.y * g[1], v.z * g[2], v.w * g[3]));
        }
    }
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
    if (Z > 65535) return fail(SHRIMPY_EINVAL, "flatfield: Z=%d exceeds the 16-bit histogram counters", Z);
    const size_t smem = 256 * (kMedPixels / 2) * sizeof(unsigned);   // 32 KB of packed histograms: six CTAs per SM
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
    const long long plane = (long long)Y * X;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int vec = raw_dtype == SHRIMPY_U16 ? 8 : 4;
    const bool aligned = plane % vec == 0 && (reinterpret_cast<uintptr_t>(d_raw) & 15u) == 0 &&
                         (reinterpret_cast<uintptr_t>(d_out) & 15u) == 0 && (reinterpret_cast<uintptr_t>(d_scale) & 15u) == 0;
    const int v = aligned ? vec : 1;
    const long long threads = (plane + v - 1) / v;
    const long long bx = (threads + 255) / 256;
    // enough CTAs to fill the machine a few times over: split the scan axis when the image alone is too small
    int zsplit = (int)std::min<long long>(Z, std::max<long long>(1, (148LL * 16 + bx - 1) / bx));
    const int z_per_block = (Z + zsplit - 1) / zsplit;
    zsplit = (Z + z_per_block - 1) / z_per_block;
    if (bx > 2147483647LL || zsplit > 65535) return fail(SHRIMPY_EINVAL, "flatfield apply: image too large for the grid");
    const dim3 grid((unsigned)bx, (unsigned)zsplit);
    if (raw_dtype == SHRIMPY_U16) {
        if (aligned) apply_scale_kernel<uint16_t, 8><<<grid, 256, 0, s>>>(static_cast<const uint16_t *>(d_raw), d_scale, d_out, plane, Z, z_per_block);
        else apply_scale_kernel<uint16_t, 1><<<grid, 256, 0, s>>>(static_cast<const uint16_t *>(d_raw), d_scale, d_out, plane, Z, z_per_block);
    } else {
        if (aligned) apply_scale_kernel<float, 4><<<grid, 256, 0, s>>>(static_cast<const float *>(d_raw), d_scale, d_out, plane, Z, z_per_block);
        else apply_scale_kernel<float, 1><<<grid, 256, 0, s>>>(static_cast<const float *>(d_raw), d_scale, d_out, plane, Z, z_per_block);
    }
    count_launch();
    SHRIMPY_CUDA_TRY(cudaGetLastError());
    return SHRIMPY_OK;
}
