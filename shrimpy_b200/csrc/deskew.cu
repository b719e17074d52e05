// Light-sheet deskew kernels for sm_100a.
//
// out[p, o1, o2] = mean_{k<n} lerp_z(raw[:, Y-1-o0_k, X-1-o1], z_in(o0_k, o2)),
//   o0_k = min(n*p + k, Y-1),  z_in = (shift + o0*m00) + o2*m02   (float64, no FMA),
//   outside (z_in < 0 or z_in > Z-1, strict)  ->  cval.
// This is the closed form of scipy.ndimage.affine_transform(order=1, mode="constant")
// followed by the edge-padded block mean, i.e. what biahub's deskew computes for
// shrimpy/preprocessing.py:408-413 and scripts/measure_psf.py:239-246.
//
// Two kernels:
//   deskew_direct_kernel  one thread per output voxel, plain global gathers; any shape,
//                         stride and alignment.  Correctness fallback.
//   deskew_tma_kernel     one CTA per (tilt block p, 128-byte raw-x tile, o2 tile).  The n raw
//                         tilt rows of the tile are staged in shared memory by TMA as
//                         [scan slice][128 B of x] boxes with the 128-byte swizzle; each thread
//                         owns one output column o2 (so its scan index and lerp weight are
//                         computed once, in float64) and walks the x tile with 128-bit shared
//                         loads; a warp's 32 lanes write 32 consecutive o2 -> coalesced 128-byte
//                         global stores.  The read is coalesced along raw x, the write along o2:
//                         the (z,x) -> (x,o2) transpose happens in shared memory.
#include "common.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace shrimpy {

struct DeskewParams {
    const void *raw;
    float *out;
    int Z, Y, X, Xp, n;        // geometry of the FULL stack
    int p0, pcount;            // window: tilt blocks
    int cbeg, cend;            // window: output columns [cbeg, cend)
    int y_org, y_cnt;          // raw rows held by the slab
    int z_org, z_cnt;          // raw scan slices held by the slab
    long long raw_sz, raw_sy;  // element strides of raw
    long long out_sp, out_s1;  // element strides of out
    double m00, m02, shift;
    float cval;
    float inv_n;
    const float *scale;  // optional flat-field scale field (Y, X) of the FULL stack, or nullptr
    unsigned *range;     // optional: ordered keys of (min, max) over every voxel written (fused value range), or nullptr
    int T2;        // o2 extent of a tile (32, 64, 128 or 256)
    int nz_cap;    // scan slices per staged row (multiple of 8, <= 256)
    int tiles_x;   // number of raw-x tiles
    int tiles_o2;  // number of o2 tiles
};

// Arithmetic shared by every kernel in this file (so that they agree bit for bit):
//   d_k = b_k - a_k                      the two scan taps of tilt row k (exact for uint16 data)
//   S   = ((a_0 + a_1) + a_2) ...        sequential; an outside row contributes cval, d_k = w_k = 0
//   u_k = w_k * (1/n)                    lerp weight pre-scaled by the block-mean factor
//   W   = fma(u_{n-1}, d_{n-1}, ... fma(u_1, d_1, u_0 * d_0))
//   out = n == 1 ? fma(w_0, d_0, a_0) : fma(S, 1/n, W);   every row outside -> exactly cval
// For uint16 data S is an exact integer, so only W and the final scale round (float32).
// With a flat-field scale field s[y,x] (shrimpy/preprocessing.py:385-404 fused in; the deskew interpolates
// along z only, so the per-pixel scale commutes with it):
//   t_k = fma(w_k, d_k, a_k),  g_k = s[y_k, x] * (1/n)   (outside row: t_k = cval, g_k = 1/n)
//   out = fma(g_{n-1}, t_{n-1}, ... g_0 * t_0);           every row outside -> exactly cval

__device__ __forceinline__ double scan_coord(double base, int o2, double m02) {
    // scipy accumulates shift + o0*M00 first, then + o2*M02, each product and sum rounded separately.
    return __dadd_rn(base, __dmul_rn((double)o2, m02));
}

template <typename T>
__global__ void __launch_bounds__(128) deskew_direct_kernel(const DeskewParams P) {
    const int o2 = P.cbeg + (blockIdx.x % P.tiles_o2) * blockDim.x + threadIdx.x;
    const int o1 = blockIdx.x / P.tiles_o2;
    const int p = P.p0 + blockIdx.y;
    if (o2 >= P.cend) return;
    const T *__restrict__ raw = static_cast<const T *>(P.raw);
    const long long xoff = P.X - 1 - o1;
    const double zmax = (double)(P.Z - 1);
    float S = 0.f, W = 0.f, one = P.cval, G = 0.f;
    int n_in = 0;
    for (int k = 0; k < P.n; ++k) {
        const int o0 = min(P.n * p + k, P.Y - 1);
        const long long yoff = (long long)(P.Y - 1 - o0 - P.y_org) * P.raw_sy;
        const double base = __dadd_rn(P.shift, __dmul_rn((double)o0, P.m00));
        const double z = scan_coord(base, o2, P.m02);
        float a = P.cval, d = 0.f, w = 0.f;
        if (z >= 0.0 && z <= zmax) {
            const double fz = floor(z);
            w = (float)(z - fz);
            const int z0 = (int)fz;
            const int z1 = min(z0 + 1, P.Z - 1);
            a = (float)__ldg(raw + (long long)(z0 - P.z_org) * P.raw_sz + yoff + xoff);
            const float b = (float)__ldg(raw + (long long)(z1 - P.z_org) * P.raw_sz + yoff + xoff);
            d = b - a;
            ++n_in;
        }
        const float u = w * P.inv_n;
        S = (k == 0) ? a : S + a;
        W = (k == 0) ? u * d : fmaf(u, d, W);
        if (k == 0) one = fmaf(w, d, a);
        if (P.scale) {
            const bool in = z >= 0.0 && z <= zmax;
            const float g = in ? __ldg(P.scale + (long long)(P.Y - 1 - o0) * P.X + xoff) * P.inv_n : P.inv_n;
            const float t = fmaf(w, d, a);
            G = (k == 0) ? g * t : fmaf(g, t, G);
        }
    }
    const float r = (n_in == 0) ? P.cval : P.scale ? G : (P.n == 1) ? one : fmaf(S, P.inv_n, W);
    __stcs(P.out + (long long)(p - P.p0) * P.out_sp + (long long)o1 * P.out_s1 + (o2 - P.cbeg), r);
    if (P.range) {   // fallback path of the fused value range: one atomic pair per active thread
        atomicMin(P.range, ordered_key(r));
        atomicMax(P.range + 1, ordered_key(r));
    }
}

// ---- TMA-staged kernel -------------------------------------------------------

template <typename T>
struct Chunk;  // 16 bytes of raw x held in a uint4

// 16 bytes of a staged row through a 32-bit shared-window address.  A tap's address is kept as
// (tile + row offset + swizzle key): rows are 128 bytes long and 128-byte aligned, so the swizzled chunk position
// (chunk << 4) ^ key is the low seven bits of that address XOR-ed with (chunk << 4) -- one LOP3 per load.
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

__device__ __forceinline__ uint32_t word_of(const uint4 &v, int i) {
    return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w;
}

template <>
struct Chunk<uint16_t> {
    static constexpr int kElems = 8;
    static constexpr uint32_t kMagic = 0x4B000000u;  // float 2^23: low mantissa bits hold an integer
    static __device__ __forceinline__ float get(const uint4 &v, int j) {
        const uint32_t word = word_of(v, j >> 1);
        return (float)((j & 1) ? (word >> 16) : (word & 0xffffu));
    }
    // uint16 -> bits of the float 2^23 + value with one byte permute (no integer->float convert)
    static __device__ __forceinline__ uint32_t magic(const uint4 &v, int j) {
        return __byte_perm(word_of(v, j >> 1), kMagic, (j & 1) ? 0x7632 : 0x7610);
    }

    // All n rows inside: S as an integer sum of the magic words (exact), W as an FMA chain.
    template <int NAVG>
    static __device__ __forceinline__ void fast(const uint32_t *a0, const uint32_t *a1, const float *w,
                                                const float *u, uint32_t cbyte, float inv_n, float *out) {
        if constexpr (NAVG >= 4) {
            // four rows per voxel: the scalar sequence keeps fewer values live (measured 3-6 % faster than the packed one)
            uint32_t sum[kElems];
            float W[kElems];
#pragma unroll
            for (int k = 0; k < NAVG; ++k) {
                const uint4 A = lds128(a0[k] ^ cbyte);
                const uint4 B = lds128(a1[k] ^ cbyte);
#pragma unroll
                for (int j = 0; j < kElems; ++j) {
                    const uint32_t am = magic(A, j), bm = magic(B, j);
                    const float d = __uint_as_float(bm) - __uint_as_float(am);  // exact: both are 2^23 + integer
                    if (NAVG == 1) {
                        out[j] = fmaf(w[0], d, __uint_as_float(am) - 8388608.0f);
                    } else {
                        sum[j] = (k == 0) ? am : sum[j] + am;
                        W[j] = (k == 0) ? u[0] * d : fmaf(u[k], d, W[j]);
                    }
                }
            }
            if (NAVG > 1) {
#pragma unroll
                for (int j = 0; j < kElems; ++j) {
                    const float S = __uint_as_float(sum[j] - (uint32_t)(NAVG - 1) * kMagic) - 8388608.0f;
                    out[j] = fmaf(S, inv_n, W[j]);
                }
            }
        } else {
            // Pairs of x positions go through packed float32 arithmetic (FADD2 / FFMA2 on sm_100: half the issue slots,
            // the same rounding per half as the scalar sequence, so the results are bit-identical to every other path).
            uint32_t sum[kElems];
            float2 W2[kElems / 2];
#pragma unroll
            for (int k = 0; k < NAVG; ++k) {
                const uint4 A = lds128(a0[k] ^ cbyte);
                const uint4 B = lds128(a1[k] ^ cbyte);
#pragma unroll
                for (int j = 0; j < kElems; j += 2) {
                    const uint32_t am0 = magic(A, j), am1 = magic(A, j + 1), bm0 = magic(B, j), bm1 = magic(B, j + 1);
                    const float2 a2 = make_float2(__uint_as_float(am0), __uint_as_float(am1));
                    const float2 b2 = make_float2(__uint_as_float(bm0), __uint_as_float(bm1));
                    const float2 d2 = __fadd2_rn(b2, make_float2(-a2.x, -a2.y));   // exact: both are 2^23 + integer
                    if (NAVG == 1) {
                        const float2 r = __ffma2_rn(make_float2(w[0], w[0]), d2, __fadd2_rn(a2, make_float2(-8388608.0f, -8388608.0f)));
                        out[j] = r.x;
                        out[j + 1] = r.y;
                    } else {
                        sum[j] = (k == 0) ? am0 : sum[j] + am0;
                        sum[j + 1] = (k == 0) ? am1 : sum[j + 1] + am1;
                        W2[j / 2] = (k == 0) ? __fmul2_rn(make_float2(u[0], u[0]), d2) : __ffma2_rn(make_float2(u[k], u[k]), d2, W2[j / 2]);
                    }
                }
            }
            if (NAVG > 1) {
#pragma unroll
                for (int j = 0; j < kElems; j += 2) {
                    const float2 m2 = make_float2(__uint_as_float(sum[j] - (uint32_t)(NAVG - 1) * kMagic),
                                                  __uint_as_float(sum[j + 1] - (uint32_t)(NAVG - 1) * kMagic));
                    const float2 S2 = __fadd2_rn(m2, make_float2(-8388608.0f, -8388608.0f));
                    const float2 r = __ffma2_rn(S2, make_float2(inv_n, inv_n), W2[j / 2]);
                    out[j] = r.x;
                    out[j + 1] = r.y;
                }
            }
        }
    }
};

// flat-field variant of the fast path (all rows inside): g points at this chunk's kElems scale values of row k,
// already multiplied by 1/n, in shared memory (same address for every lane: broadcast reads)
template <typename T, int NAVG>
__device__ __forceinline__ void fast_scaled(const uint32_t *a0, const uint32_t *a1, const float *w,
                                            const float *g, int g_row_stride, uint32_t cbyte, float *out);

template <>
struct Chunk<float> {
    static constexpr int kElems = 4;
    static __device__ __forceinline__ float get(const uint4 &v, int j) { return __uint_as_float(word_of(v, j)); }

    template <int NAVG>
    static __device__ __forceinline__ void fast(const uint32_t *a0, const uint32_t *a1, const float *w,
                                                const float *u, uint32_t cbyte, float inv_n, float *out) {
        // packed pairs (FADD2 / FFMA2): the same operations and rounding as the scalar sequence of the other kernels
        float2 S2[kElems / 2], W2[kElems / 2], one2[kElems / 2];
#pragma unroll
        for (int k = 0; k < NAVG; ++k) {
            const uint4 A = lds128(a0[k] ^ cbyte);
            const uint4 B = lds128(a1[k] ^ cbyte);
#pragma unroll
            for (int j = 0; j < kElems; j += 2) {
                const float2 a2 = make_float2(get(A, j), get(A, j + 1)), b2 = make_float2(get(B, j), get(B, j + 1));
                const float2 d2 = __fadd2_rn(b2, make_float2(-a2.x, -a2.y));
                S2[j / 2] = (k == 0) ? a2 : __fadd2_rn(S2[j / 2], a2);
                W2[j / 2] = (k == 0) ? __fmul2_rn(make_float2(u[0], u[0]), d2) : __ffma2_rn(make_float2(u[k], u[k]), d2, W2[j / 2]);
                if (k == 0 && NAVG == 1) one2[j / 2] = __ffma2_rn(make_float2(w[0], w[0]), d2, a2);
            }
        }
#pragma unroll
        for (int j = 0; j < kElems; j += 2) {
            const float2 r = (NAVG == 1) ? one2[j / 2] : __ffma2_rn(S2[j / 2], make_float2(inv_n, inv_n), W2[j / 2]);
            out[j] = r.x;
            out[j + 1] = r.y;
        }
    }
};

template <typename T, int NAVG>
__device__ __forceinline__ void fast_scaled(const uint32_t *a0, const uint32_t *a1, const float *w,
                                            const float *g, int g_row_stride, uint32_t cbyte, float *out) {
    constexpr int EPC = Chunk<T>::kElems;
#pragma unroll
    for (int k = 0; k < NAVG; ++k) {
        const uint4 A = lds128(a0[k] ^ cbyte);
        const uint4 B = lds128(a1[k] ^ cbyte);
        float gk[EPC];
#pragma unroll
        for (int j = 0; j < EPC; j += 4) {
            const float4 v = *reinterpret_cast<const float4 *>(g + k * g_row_stride + j);
            gk[j] = v.x; gk[j + 1] = v.y; gk[j + 2] = v.z; gk[j + 3] = v.w;
        }
        // packed pairs (FADD2 / FFMA2): the same operations and rounding per element as the scalar sequence
#pragma unroll
        for (int j = 0; j < EPC; j += 2) {
            float2 a2, d2;
            if (sizeof(T) == 2) {
                const float2 am = make_float2(__uint_as_float(Chunk<uint16_t>::magic(A, j)), __uint_as_float(Chunk<uint16_t>::magic(A, j + 1)));
                const float2 bm = make_float2(__uint_as_float(Chunk<uint16_t>::magic(B, j)), __uint_as_float(Chunk<uint16_t>::magic(B, j + 1)));
                d2 = __fadd2_rn(bm, make_float2(-am.x, -am.y));
                a2 = __fadd2_rn(am, make_float2(-8388608.0f, -8388608.0f));
            } else {
                a2 = make_float2(Chunk<float>::get(A, j), Chunk<float>::get(A, j + 1));
                const float2 b2 = make_float2(Chunk<float>::get(B, j), Chunk<float>::get(B, j + 1));
                d2 = __fadd2_rn(b2, make_float2(-a2.x, -a2.y));
            }
            const float2 t2 = __ffma2_rn(make_float2(w[k], w[k]), d2, a2);
            const float2 g2 = make_float2(gk[j], gk[j + 1]);
            const float2 r2 = (k == 0) ? __fmul2_rn(g2, t2) : __ffma2_rn(g2, t2, make_float2(out[j], out[j + 1]));
            out[j] = r2.x;
            out[j + 1] = r2.y;
        }
    }
}

constexpr int kTmaThreads = 256;
constexpr int kRowBytes = 128;  // one staged row = 128 B of raw x = one swizzle span
constexpr int kMaxTmaAvg = 4;   // template instantiations exist for n = 1..4

// RANGE: the value range of the written voxels is reduced in the same pass (per thread -> warp shuffle -> CTA ->
// one atomicMin / atomicMax pair per CTA on ordered keys): the min/max pass the tracking step makes right after
// the deskew (shrimpy/dynatrack/tracking.py:583-584) disappears.
//
template <typename T, int NAVG, bool SCALED, bool RANGE>
__global__ void __launch_bounds__(kTmaThreads, 4)
    deskew_tma_kernel(const __grid_constant__ CUtensorMap tmap, const DeskewParams P) {
    constexpr int EPC = Chunk<T>::kElems;
    constexpr int TX = 8 * EPC;
    __shared__ __align__(16) float s_scale[SCALED ? NAVG * TX : 4];   // scale * (1/n) of this tile's rows

    extern __shared__ uint8_t smem_dyn[];
    __shared__ __align__(8) uint64_t bar;

    // The 128-byte swizzle wants 1024-byte aligned tiles.
    const uint32_t pad = (1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u;
    const uint8_t *tile = smem_dyn + pad;
    const uint32_t region_bytes = (uint32_t)P.nz_cap * kRowBytes;

    // raw-x tile fastest: CTAs that run together read adjacent 128-byte segments of the same DRAM rows
    // (measured: the alternative order makes no difference on B200)
    // block order: pairs of raw-x tiles fastest (they share 256-byte DRAM accesses), then the o2 tiles of a row (their
    // 1 KB row segments are neighbours in the output), then the remaining x tiles, then the tilt block
    const int tx = 2 * blockIdx.y + (blockIdx.x & 1);
    const int t2 = blockIdx.x >> 1;
    if (tx >= P.tiles_x) return;
    const int p = P.p0 + blockIdx.z;
    const int x0 = tx * TX;
    const int c0 = P.cbeg + t2 * P.T2;
    const int c_last = min(c0 + P.T2, P.cend) - 1;
    const double zmax = (double)(P.Z - 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // Warp 0, lane k: scan range this tile needs from tilt row k (a pure function of the block
    // index; rounding is monotone in o2, so the tile's first/last columns bound every thread's
    // coordinate), then one TMA box per needed row.  The values are shared through smem.
    __shared__ double s_base[NAVG];
    __shared__ int s_zlo[NAVG];
    __shared__ unsigned s_need;
    if (warp == 0) {
        if (lane == 0) {
            mbar_init(&bar, 1);
            fence_mbar_init();
        }
        bool need = false;
        int zlo = 0, yrow = 0;
        if (lane < NAVG) {
            const int o0 = min(NAVG * p + lane, P.Y - 1);
            yrow = P.Y - 1 - o0 - P.y_org;
            const double base = __dadd_rn(P.shift, __dmul_rn((double)o0, P.m00));
            const double zfirst = scan_coord(base, c0, P.m02);
            const double zlast = scan_coord(base, c_last, P.m02);
            need = !(zlast < 0.0 || zfirst > zmax);
            zlo = __double2int_rd(fmin(fmax(zfirst, 0.0), zmax));
            s_base[lane] = base;
            s_zlo[lane] = zlo;
        }
        const unsigned mask = __ballot_sync(0xffffffffu, need);
        if (lane == 0) {
            s_need = mask;
            if (mask) mbar_arrive_expect_tx(&bar, (uint32_t)__popc(mask) * region_bytes);
        }
        __syncwarp();
        if (need) tma_load_3d(smem_u32(tile) + lane * region_bytes, &tmap, x0, yrow, zlo - P.z_org, &bar);
    }
    if (SCALED) {
        for (int i = threadIdx.x; i < NAVG * TX; i += kTmaThreads) {
            const int k = i / TX, x = x0 + (i - k * TX);
            const int o0 = min(NAVG * p + k, P.Y - 1);
            s_scale[i] = (x < P.X) ? __ldg(P.scale + (long long)(P.Y - 1 - o0) * P.X + x) * P.inv_n : 0.f;
        }
    }
    __syncthreads();  // row ranges, scale tile and the initialised barrier visible to everyone
    const bool any_need = s_need != 0;

    // Per-thread column state while the boxes are in flight.
    const int w2_log2 = 31 - __clz(P.T2 >> 5);   // T2 = 32, 64, 128 or 256 columns: 1, 2, 4 or 8 warps along o2
    const int parts = 8 >> w2_log2;              // how many warps share one o2 span, splitting the x chunks
    const int part = warp >> w2_log2;
    const int o2 = c0 + (warp & ((1 << w2_log2) - 1)) * 32 + lane;
    const bool col_ok = o2 < P.cend;

    float w[NAVG], u[NAVG];           // lerp weight, and the same pre-scaled by 1/n
    uint32_t a0[NAVG], a1[NAVG];      // shared address of chunk 0 of the two tap rows: tile + row offset + swizzle key
    const uint32_t tile_s = smem_u32(tile);
    bool inside[NAVG];
#pragma unroll
    for (int k = 0; k < NAVG; ++k) {
        const int zlo = s_zlo[k];
        const double z = scan_coord(s_base[k], o2, P.m02);
        inside[k] = col_ok && z >= 0.0 && z <= zmax;
        const int z0 = inside[k] ? __double2int_rd(z) : zlo;
        w[k] = inside[k] ? (float)(z - (double)z0) : 0.f;
        u[k] = w[k] * P.inv_n;
        const int r0 = min(z0 - zlo, P.nz_cap - 1);
        const int r1 = min(min(z0 + 1, P.Z - 1) - zlo, P.nz_cap - 1);
        a0[k] = tile_s + k * region_bytes + (uint32_t)r0 * kRowBytes + (((uint32_t)r0 & 7u) << 4);
        a1[k] = tile_s + k * region_bytes + (uint32_t)r1 * kRowBytes + (((uint32_t)r1 & 7u) << 4);
    }

    if (any_need) mbar_wait(&bar, 0);

    // Warp-uniform classification: the interior of the volume takes the branch-free fast path.
    bool all_in = true, none_in = true;
#pragma unroll
    for (int k = 0; k < NAVG; ++k) {
        all_in &= inside[k];
        none_in &= !inside[k];
    }
    const bool warp_all_in = __all_sync(0xffffffffu, all_in);
    const bool warp_none_in = __all_sync(0xffffffffu, none_in);

    char *out_col = reinterpret_cast<char *>(P.out + (long long)(p - P.p0) * P.out_sp + (o2 - P.cbeg));
    const long long row_bytes = P.out_s1 * (long long)sizeof(float);
    float vlo = 3.402823466e38f, vhi = -3.402823466e38f;

    for (int c = part; c < 8; c += parts) {
        float r[EPC];
        if (warp_all_in) {
            if (SCALED) fast_scaled<T, NAVG>(a0, a1, w, s_scale + c * EPC, TX, (uint32_t)c << 4, r);
            else Chunk<T>::template fast<NAVG>(a0, a1, w, u, (uint32_t)c << 4, P.inv_n, r);
        } else if (warp_none_in) {
#pragma unroll
            for (int j = 0; j < EPC; ++j) r[j] = P.cval;
        } else {
            // boundary warps: some rows / lanes outside
            float S[EPC], W[EPC], one[EPC], G[EPC];
#pragma unroll
            for (int k = 0; k < NAVG; ++k) {
                uint4 A = make_uint4(0, 0, 0, 0), B = A;
                if (inside[k]) {
                    A = lds128(a0[k] ^ ((uint32_t)c << 4));
                    B = lds128(a1[k] ^ ((uint32_t)c << 4));
                }
#pragma unroll
                for (int j = 0; j < EPC; ++j) {
                    const float a = inside[k] ? Chunk<T>::get(A, j) : P.cval;
                    const float d = inside[k] ? Chunk<T>::get(B, j) - a : 0.f;
                    S[j] = (k == 0) ? a : S[j] + a;
                    W[j] = (k == 0) ? u[k] * d : fmaf(u[k], d, W[j]);   // w = u = 0 outside
                    if (k == 0) one[j] = fmaf(w[k], d, a);
                    if (SCALED) {
                        const float g = inside[k] ? s_scale[k * TX + c * EPC + j] : P.inv_n;
                        const float t = fmaf(w[k], d, a);
                        G[j] = (k == 0) ? g * t : fmaf(g, t, G[j]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < EPC; ++j)
                r[j] = none_in ? P.cval : SCALED ? G[j] : (NAVG == 1) ? one[j] : fmaf(S[j], P.inv_n, W[j]);
        }
        if (col_ok) {
            const int xc = x0 + c * EPC;
            const int xvalid = P.X - xc;  // elements of this chunk that exist (uniform over the CTA)
            char *ptr = out_col + (long long)(P.X - 1 - xc) * row_bytes;
            if (xvalid >= EPC) {
#pragma unroll
                for (int j = 0; j < EPC; ++j) {
                    __stcs(reinterpret_cast<float *>(ptr), r[j]);   // streaming: outputs are never re-read
                    ptr -= row_bytes;
                    asm volatile("" : "+l"(ptr));  // keep a stepped pointer (2 adds), not base+offset (4)
                    if (RANGE) {
                        vlo = fminf(vlo, r[j]);
                        vhi = fmaxf(vhi, r[j]);
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < EPC; ++j) {
                    if (j < xvalid) {
                        __stcs(reinterpret_cast<float *>(ptr), r[j]);
                        if (RANGE) {
                            vlo = fminf(vlo, r[j]);
                            vhi = fmaxf(vhi, r[j]);
                        }
                    }
                    ptr -= row_bytes;
                }
            }
        }
    }
    if (RANGE) {
        __shared__ float s_lo[kTmaThreads / 32], s_hi[kTmaThreads / 32];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            vlo = fminf(vlo, __shfl_xor_sync(0xffffffffu, vlo, o));
            vhi = fmaxf(vhi, __shfl_xor_sync(0xffffffffu, vhi, o));
        }
        if (lane == 0) {
            s_lo[warp] = vlo;
            s_hi[warp] = vhi;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
#pragma unroll
            for (int i = 1; i < kTmaThreads / 32; ++i) {
                vlo = fminf(vlo, s_lo[i]);
                vhi = fmaxf(vhi, s_hi[i]);
            }
            if (vlo <= vhi) {   // the CTA wrote something
                atomicMin(P.range, ordered_key(vlo));
                atomicMax(P.range + 1, ordered_key(vhi));
            }
        }
    }
}

// ---- staged variant: whole-sector, 16-byte vector stores into rows of ANY pitch ---------------------------------------
//
// The kernel above stores straight from registers: a warp's 32 lanes write 32 consecutive floats of one output row.
// When a row is not a whole number of 32-byte sectors long (Xp = 1279, 1799, 10517: every mantis geometry), every row
// starts at another offset inside a sector, each of those 128-byte warp stores straddles sector boundaries, and the
// write-dominated deskews lose a quarter of their bandwidth to partial-sector writes (measured: the same launch into
// rows padded to whole sectors runs 0.83 -> 0.64 ms; a variant of this kernel with overlapping tiles whose edges fell
// on sector boundaries recovered only 6 % of it, because the seven warp boundaries inside a tile stayed ragged).
//
// Here a chunk's results (EPC output rows x 256 columns) go through shared memory and leave as 16-byte stores that
// start on 32-byte boundaries of global memory, whatever the row pitch:
//   * the staged row pitch Pp is chosen with Pp = -out_s1 (mod 8) and the first row starts k0 = (address / 4) mod 8
//     floats in, so that a voxel's index in the stage buffer is congruent to its global float address mod 8: a
//     16-byte aligned shared-memory read is a 16-byte aligned global store, and 8 of them in a row are one sector;
//   * 64 threads drain one row: the 31 whole sectors inside the tile's 256 columns as 62 float4 (a warp instruction
//     writes 512 contiguous bytes), the ragged ends (the sector a row shares with the neighbouring tile, 0..7 floats
//     on either side) as scalars.  Same grid as the plain kernel, nothing computed twice;
//   * two stage buffers alternate, one __syncthreads per chunk (a split arrive / wait pair of mbarriers per buffer with
//     the next chunk's arithmetic in between was tried and measured 5-15 % slower).
// Arithmetic and results are those of deskew_tma_kernel, bit for bit.
constexpr int kStagePitch = 272;     // floats reserved per staged row: the pitch in use is 264 + (0..7)

template <typename T, int NAVG>
__global__ void __launch_bounds__(kTmaThreads, NAVG <= 2 ? 4 : 3)
    deskew_tma_staged_kernel(const __grid_constant__ CUtensorMap tmap, const DeskewParams P) {
    constexpr int EPC = Chunk<T>::kElems;   // output rows per stage
    constexpr int TX = 8 * EPC;
    constexpr int T2 = 256;
    extern __shared__ uint8_t smem_dyn[];
    __shared__ __align__(8) uint64_t bar;

    const uint32_t pad = (1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u;
    const uint8_t *tile = smem_dyn + pad;
    const uint32_t region_bytes = (uint32_t)P.nz_cap * kRowBytes;
    float *stage = reinterpret_cast<float *>(smem_dyn + pad + NAVG * region_bytes);   // 2 x EPC x kStagePitch floats

    // block order: pairs of raw-x tiles fastest (they share 256-byte DRAM accesses), then the o2 tiles of a row (their
    // 1 KB row segments are neighbours in the output), then the remaining x tiles, then the tilt block
    const int tx = 2 * blockIdx.y + (blockIdx.x & 1);
    const int t2 = blockIdx.x >> 1;
    if (tx >= P.tiles_x) return;
    const int p = P.p0 + blockIdx.z;
    const int x0 = tx * TX;
    const int c0 = P.cbeg + t2 * T2;
    const int ncols = min(T2, P.cend - c0);   // columns of this tile that exist
    const int c_last = c0 + ncols - 1;
    const double zmax = (double)(P.Z - 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    __shared__ double s_base[NAVG];
    __shared__ int s_zlo[NAVG];
    __shared__ unsigned s_need;
    if (warp == 0) {
        if (lane == 0) {
            mbar_init(&bar, 1);
            fence_mbar_init();
        }
        bool need = false;
        int zlo = 0, yrow = 0;
        if (lane < NAVG) {
            const int o0 = min(NAVG * p + lane, P.Y - 1);
            yrow = P.Y - 1 - o0 - P.y_org;
            const double base = __dadd_rn(P.shift, __dmul_rn((double)o0, P.m00));
            const double zfirst = scan_coord(base, c0, P.m02);
            const double zlast = scan_coord(base, c_last, P.m02);
            need = !(zlast < 0.0 || zfirst > zmax);
            zlo = __double2int_rd(fmin(fmax(zfirst, 0.0), zmax));
            s_base[lane] = base;
            s_zlo[lane] = zlo;
        }
        const unsigned mask = __ballot_sync(0xffffffffu, need);
        if (lane == 0) {
            s_need = mask;
            if (mask) mbar_arrive_expect_tx(&bar, (uint32_t)__popc(mask) * region_bytes);
        }
        __syncwarp();
        if (need) tma_load_3d(smem_u32(tile) + lane * region_bytes, &tmap, x0, yrow, zlo - P.z_org, &bar);
    }
    __syncthreads();
    const bool any_need = s_need != 0;

    const int t_local = threadIdx.x;   // one thread = one column of the tile
    const int o2 = c0 + t_local;
    const bool col_ok = t_local < ncols;

    float w[NAVG], u[NAVG];
    uint32_t a0[NAVG], a1[NAVG];      // shared address of chunk 0 of the two tap rows: tile + row offset + swizzle key
    const uint32_t tile_s = smem_u32(tile);
    bool inside[NAVG];
#pragma unroll
    for (int k = 0; k < NAVG; ++k) {
        const int zlo = s_zlo[k];
        const double z = scan_coord(s_base[k], o2, P.m02);
        inside[k] = col_ok && z >= 0.0 && z <= zmax;
        const int z0 = inside[k] ? __double2int_rd(z) : zlo;
        w[k] = inside[k] ? (float)(z - (double)z0) : 0.f;
        u[k] = w[k] * P.inv_n;
        const int r0 = min(z0 - zlo, P.nz_cap - 1);
        const int r1 = min(min(z0 + 1, P.Z - 1) - zlo, P.nz_cap - 1);
        a0[k] = tile_s + k * region_bytes + (uint32_t)r0 * kRowBytes + (((uint32_t)r0 & 7u) << 4);
        a1[k] = tile_s + k * region_bytes + (uint32_t)r1 * kRowBytes + (((uint32_t)r1 & 7u) << 4);
    }

    // stage geometry (uniform over the CTA): pitch and start offset that make stage index == global float address mod 8
    const uint32_t s1m = (uint32_t)(P.out_s1 & 7);
    const uint32_t pitch = 264u + ((8u - s1m) & 7u);
    float *const row0 = P.out + (long long)(p - P.p0) * P.out_sp + (c0 - P.cbeg);   // (row o1 = 0, column c0)
    const uint32_t a_first = (uint32_t)(reinterpret_cast<uintptr_t>(row0 + (long long)(P.X - 1 - x0) * P.out_s1) >> 2);

    // drain role: row (threadIdx >> 6) of each pass of four rows, float4 number (threadIdx & 63) of that row.  What a
    // thread stores is the same in every chunk (a chunk is EPC rows: the alignment pattern repeats every 8 rows, and
    // for EPC = 4 it alternates between two patterns), so offsets and predicates are set up once, per pass and parity.
    const int drow = threadIdx.x >> 6, dk = threadIdx.x & 63;
    constexpr int PASSES = EPC / 4, PAR = EPC == 8 ? 1 : 2;
    uint32_t d_src[PAR][PASSES];      // byte offset of this thread's float4 inside a stage buffer
    int d_col[PAR][PASSES];           // its first column (local); < 0: no float4 for this thread
    int d_rag[PAR][PASSES];           // local column of its ragged-end scalar; < 0: none
    uint32_t d_s0[PAR][PASSES];       // stage index of the row's column c0
#pragma unroll
    for (int par = 0; par < PAR; ++par) {
        const uint32_t k0 = (a_first - (uint32_t)(par * EPC) * (uint32_t)P.out_s1) & 7u;
#pragma unroll
        for (int ps = 0; ps < PASSES; ++ps) {
            const int j = ps * 4 + drow;
            const uint32_t s0 = (uint32_t)j * pitch + k0;
            const int t_lo = (int)((8u - (s0 & 7u)) & 7u);      // first whole sector of the row inside the tile
            const int t = t_lo + 4 * dk;                          // 62 float4 = the 31 sectors that always lie inside
            d_s0[par][ps] = s0;
            d_src[par][ps] = (s0 + (uint32_t)t) * 4u;
            d_col[par][ps] = (dk < 62 && t + 3 < ncols) ? t : -1;
            // ragged ends, one scalar per thread: the head [0, t_lo) by threads 0..6, the tail [t_lo + 248, 256) by
            // threads 8..15, and where the row ends inside the tile its last 1..3 floats by threads 16..18
            const int body_end = t_lo + 4 * min(62, max(0, (ncols - t_lo) >> 2));
            const int e = dk < 8 ? dk : dk < 16 ? t_lo + 248 + (dk - 8) : body_end + (dk - 16);
            const bool ragged = dk < 8 ? dk < min(t_lo, ncols) : dk < 16 ? e < ncols : (dk < 19 && body_end < t_lo + 248 && e < ncols);
            d_rag[par][ps] = ragged ? e : -1;
        }
    }
    // running global pointer of (this thread's drain row of the current chunk, column c0)
    float *d_row = row0 + (long long)(P.X - 1 - x0 - drow) * P.out_s1;

    if (any_need) mbar_wait(&bar, 0);

    bool all_in = true, none_in = true;
#pragma unroll
    for (int k = 0; k < NAVG; ++k) {
        all_in &= inside[k];
        none_in &= !inside[k];
    }
    const bool warp_all_in = __all_sync(0xffffffffu, all_in);
    const bool warp_none_in = __all_sync(0xffffffffu, none_in);

    auto compute = [&](int c, float *r) {
        if (warp_all_in) {
            Chunk<T>::template fast<NAVG>(a0, a1, w, u, (uint32_t)c << 4, P.inv_n, r);
        } else if (warp_none_in) {
#pragma unroll
            for (int j = 0; j < EPC; ++j) r[j] = P.cval;
        } else {
            float S[EPC], W[EPC], one[EPC];
#pragma unroll
            for (int k = 0; k < NAVG; ++k) {
                uint4 A = make_uint4(0, 0, 0, 0), B = A;
                if (inside[k]) {
                    A = lds128(a0[k] ^ ((uint32_t)c << 4));
                    B = lds128(a1[k] ^ ((uint32_t)c << 4));
                }
#pragma unroll
                for (int j = 0; j < EPC; ++j) {
                    const float a = inside[k] ? Chunk<T>::get(A, j) : P.cval;
                    const float d = inside[k] ? Chunk<T>::get(B, j) - a : 0.f;
                    S[j] = (k == 0) ? a : S[j] + a;
                    W[j] = (k == 0) ? u[k] * d : fmaf(u[k], d, W[j]);
                    if (k == 0) one[j] = fmaf(w[k], d, a);
                }
            }
#pragma unroll
            for (int j = 0; j < EPC; ++j) r[j] = none_in ? P.cval : (NAVG == 1) ? one[j] : fmaf(S[j], P.inv_n, W[j]);
        }
    };
    // alignment of (row of chunk c, column c0): rows go DOWN in address as x goes up
    auto k0_of = [&](int c) { return (a_first - (uint32_t)(c * EPC) * (uint32_t)P.out_s1) & 7u; };
    auto stage_chunk = [&](int c, const float *r) {
        if (col_ok) {
            float *dst = stage + (c & 1) * (EPC * kStagePitch) + t_local + k0_of(c);
#pragma unroll
            for (int j = 0; j < EPC; ++j) {
                *dst = r[j];
                dst += pitch;
            }
        }
    };
    auto drain_chunk = [&](int c) {
        const int xc = x0 + c * EPC;
        const int par = EPC == 8 ? 0 : (c & 1);
        const char *buf = reinterpret_cast<const char *>(stage + (c & 1) * (EPC * kStagePitch));
#pragma unroll
        for (int ps = 0; ps < PASSES; ++ps) {
            float *grow = d_row - (long long)(ps * 4) * P.out_s1;
            if (xc + ps * 4 + drow < P.X) {
                // (selects, not a runtime index: the tables stay in registers)
                const int t = (PAR == 2 && par) ? d_col[PAR - 1][ps] : d_col[0][ps];
                if (t >= 0) {
                    const float4 v = *reinterpret_cast<const float4 *>(buf + ((PAR == 2 && par) ? d_src[PAR - 1][ps] : d_src[0][ps]));
                    __stcs(reinterpret_cast<float4 *>(grow + t), v);
                }
                if (dk < 19) {           // the first warp of a row also carries its ragged ends
                    const int e = (PAR == 2 && par) ? d_rag[PAR - 1][ps] : d_rag[0][ps];
                    if (e >= 0) __stcs(grow + e, reinterpret_cast<const float *>(buf)[((PAR == 2 && par) ? d_s0[PAR - 1][ps] : d_s0[0][ps]) + e]);
                }
            }
        }
        d_row -= (long long)EPC * P.out_s1;
    };

    for (int c = 0; c < 8; ++c) {
        float r[EPC];
        compute(c, r);
        stage_chunk(c, r);
        __syncthreads();   // the chunk is staged; the buffer drained two chunks ago is free again after this barrier
        drain_chunk(c);
    }
}

__global__ void range_init_kernel(unsigned *slots) {
    slots[0] = 0xffffffffu;
    slots[1] = 0u;
}

__global__ void range_finish_kernel(unsigned *slots) {
    const float lo = ordered_value(slots[0]), hi = ordered_value(slots[1]);
    reinterpret_cast<float *>(slots)[0] = lo;
    reinterpret_cast<float *>(slots)[1] = hi;
}

// ---- host side -----------------------------------------------------------------

static int env_int(const char *name, int fallback) {
    const char *s = getenv(name);
    return (s && *s) ? atoi(s) : fallback;
}

template <typename T>
static int launch_direct(const DeskewParams &Pin, cudaStream_t stream) {
    DeskewParams P = Pin;
    const int threads = 128;
    P.tiles_o2 = (P.cend - P.cbeg + threads - 1) / threads;
    const long long gx = (long long)P.tiles_o2 * P.X;
    if (gx > 2147483647LL || P.pcount > 65535)
        return fail(SHRIMPY_EINVAL, "deskew: volume too large for the direct kernel grid (%lld x %d)", gx, P.pcount);
    deskew_direct_kernel<T><<<dim3((unsigned)gx, (unsigned)P.pcount), threads, 0, stream>>>(P);
    count_launch();
    SHRIMPY_CUDA_TRY(cudaGetLastError());
    return SHRIMPY_OK;
}

template <typename T, int NAVG>
static int launch_tma_staged_n(const CUtensorMap &tmap, const DeskewParams &P, size_t smem, cudaStream_t stream) {
    auto kern = deskew_tma_staged_kernel<T, NAVG>;
    if (smem + 1024 > 48 * 1024)
        SHRIMPY_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<dim3(2u * (unsigned)P.tiles_o2, (unsigned)((P.tiles_x + 1) / 2), (unsigned)P.pcount), kTmaThreads, smem, stream>>>(tmap, P);
    count_launch();
    SHRIMPY_CUDA_TRY(cudaGetLastError());
    return SHRIMPY_OK;
}

template <typename T, int NAVG>
static int launch_tma_n(const CUtensorMap &tmap, const DeskewParams &P, size_t smem, cudaStream_t stream) {
    auto kern = P.scale ? (P.range ? deskew_tma_kernel<T, NAVG, true, true> : deskew_tma_kernel<T, NAVG, true, false>)
                        : (P.range ? deskew_tma_kernel<T, NAVG, false, true> : deskew_tma_kernel<T, NAVG, false, false>);
    if (smem + 1024 > 48 * 1024)  // static smem counts against the 48 KB default as well
        SHRIMPY_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<dim3(2u * (unsigned)P.tiles_o2, (unsigned)((P.tiles_x + 1) / 2), (unsigned)P.pcount), kTmaThreads, smem, stream>>>(tmap, P);
    count_launch();
    SHRIMPY_CUDA_TRY(cudaGetLastError());
    return SHRIMPY_OK;
}

// Returns SHRIMPY_OK and sets *used = true when the TMA kernel was launched; *used = false
// (still SHRIMPY_OK) when the problem is not eligible and the caller should fall back.
template <typename T>
static int launch_tma(const DeskewParams &Pin, cudaStream_t stream, bool *used, bool required, bool staged = false) {
    *used = false;
    DeskewParams P = Pin;
    constexpr int ES = (int)sizeof(T);
    constexpr int TX = 128 / ES;
    const char *why = nullptr;
    if (P.n > kMaxTmaAvg) why = "average_n_slices > 4";
    else if (!(P.m02 > 0.0) || P.m00 > 0.0) why = "scan coordinate not monotone (need m02 > 0, m00 <= 0)";
    else if ((reinterpret_cast<uintptr_t>(P.raw) & 15u) != 0) why = "raw pointer not 16-byte aligned";
    else if ((P.raw_sy * ES) % 16 != 0 || (P.raw_sz * ES) % 16 != 0) why = "raw strides not multiples of 16 bytes";
    else if (P.raw_sy < P.X || P.raw_sz < (long long)P.y_cnt * P.raw_sy) why = "raw strides overlap";
    else if (P.pcount > 65535) why = "too many tilt blocks for grid.z";
    else if (staged && (P.scale || P.range)) why = "the staged variant is not built for the fused scale / value range";

    if (!why) {
        // Tile extent along o2: the staged scan range must fit nz_cap <= 256 slices and the CTA's
        // shared memory; default 128 columns, overridable for experiments.
        int T2 = env_int("SHRIMPY_DESKEW_T2", 256);
        if (staged) T2 = 256;
        if (T2 != 32 && T2 != 64 && T2 != 128 && T2 != 256) T2 = 256;
        const int smem_budget = env_int("SHRIMPY_DESKEW_SMEM", 72 * 1024);
        const long long stage_bytes = staged ? 2LL * (16 / ES) * kStagePitch * (long long)sizeof(float) : 0;
        for (;; T2 >>= 1) {
            if (T2 < 32 || (staged && T2 != 256)) {
                why = "px_to_scan_ratio too large for a staged tile";
                break;
            }
            const long long nz = (long long)std::ceil((T2 - 1) * P.m02) + 3;
            const long long cap = (nz + 7) / 8 * 8;
            if (cap <= 256 && cap * kRowBytes * P.n + stage_bytes + 1024 <= smem_budget) {
                P.T2 = T2;
                P.nz_cap = (int)cap;
                break;
            }
        }
    }
    if (why) {
        if (required) return fail(SHRIMPY_EINVAL, "deskew: TMA kernel not applicable: %s", why);
        return SHRIMPY_OK;
    }
    P.tiles_x = (P.X + TX - 1) / TX;
    P.tiles_o2 = (P.cend - P.cbeg + P.T2 - 1) / P.T2;
    if (P.tiles_x > 2 * 65535) {
        if (required) return fail(SHRIMPY_EINVAL, "deskew: grid too large");
        return SHRIMPY_OK;
    }

    EncodeTiledFn encode = tensor_map_encoder();
    if (!encode) return fail(SHRIMPY_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
    CUtensorMap tmap;
    const cuuint64_t gdim[3] = {(cuuint64_t)P.X, (cuuint64_t)P.y_cnt, (cuuint64_t)P.z_cnt};
    const cuuint64_t gstride[2] = {(cuuint64_t)P.raw_sy * ES, (cuuint64_t)P.raw_sz * ES};
    const cuuint32_t box[3] = {(cuuint32_t)TX, 1u, (cuuint32_t)P.nz_cap};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    // L2 promotion 256 B: a tile row is 128 B of one scan slice, and the CTA of the neighbouring x tile reads the other
    // half of the same 256 B a moment later -- fetched as one DRAM access, the second CTA hits L2.  Measured on
    // config 2: 0.3025 -> 0.2954 ms (0.914 -> 0.935 of the HBM peak), 1000-launch sustained 0.87 -> 0.915.
    const CUresult rc = encode(&tmap, ES == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                               const_cast<void *>(P.raw), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        if (required) return fail(SHRIMPY_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)rc);
        return SHRIMPY_OK;
    }
    const size_t smem = (size_t)P.nz_cap * kRowBytes * P.n + 1024 +
                        (staged ? 2u * (16 / ES) * kStagePitch * sizeof(float) : 0u);
    int err;
    if (staged) {
        switch (P.n) {
            case 1: err = launch_tma_staged_n<T, 1>(tmap, P, smem, stream); break;
            case 2: err = launch_tma_staged_n<T, 2>(tmap, P, smem, stream); break;
            case 3: err = launch_tma_staged_n<T, 3>(tmap, P, smem, stream); break;
            default: err = launch_tma_staged_n<T, 4>(tmap, P, smem, stream); break;
        }
        if (err == SHRIMPY_OK) *used = true;
        return err;
    }
    switch (P.n) {
        case 1: err = launch_tma_n<T, 1>(tmap, P, smem, stream); break;
        case 2: err = launch_tma_n<T, 2>(tmap, P, smem, stream); break;
        case 3: err = launch_tma_n<T, 3>(tmap, P, smem, stream); break;
        default: err = launch_tma_n<T, 4>(tmap, P, smem, stream); break;
    }
    if (err == SHRIMPY_OK) *used = true;
    return err;
}

template <typename T>
static int deskew_dispatch(const DeskewParams &P, int kernel, cudaStream_t stream) {
    if (kernel != SHRIMPY_KERNEL_DIRECT) {
        bool used = false;
        const bool forced = kernel == SHRIMPY_KERNEL_TMA_STAGED;
        // Measured on B200 (profiles/r02_staged_rows_probe.json): the staged variant wins wherever the writes dominate --
        // n = 1: 0.82 -> 0.60 ms, keep_overhang 1.19 -> 0.84 ms, config 5 7.13 -> 5.18 ms (0.92 of the HBM peak), and still
        // 3 % into rows of a whole-sector pitch -- and loses 3-17 % for n >= 2, where the plain kernel stays.
        // n = 2 into rows of an odd pitch: 0.410 -> 0.396 ms; into whole-sector rows the plain kernel keeps its 7 % lead.
        bool staged = forced || (kernel == SHRIMPY_KERNEL_AUTO && !P.scale && !P.range &&
                                 (P.n == 1 || (P.n == 2 && (P.out_s1 % 8) != 0)) &&
                                 env_int("SHRIMPY_DESKEW_STAGED", 1) != 0);
        int err = launch_tma<T>(P, stream, &used, kernel == SHRIMPY_KERNEL_TMA || forced, staged);
        if (err == SHRIMPY_OK && !used && staged && !forced)      // a 256-column tile does not fit (large r): plain tiles
            err = launch_tma<T>(P, stream, &used, false, false);
        if (err != SHRIMPY_OK || used) return err;
    }
    return launch_direct<T>(P, stream);
}

}  // namespace shrimpy

using namespace shrimpy;

// Rows / slices of the full stack that a window touches (host mirror of the kernels' arithmetic).
static void window_needs(int Z, int Y, int n, double m00, double m02, double shift, int p0, int pcount, int cbeg,
                         int cend, int32_t y_range[2], int32_t z_range[2]) {
    const int o0_first = std::min(n * p0, Y - 1);
    const int o0_last = std::min(n * (p0 + pcount) - 1, Y - 1);
    y_range[0] = Y - 1 - o0_last;
    y_range[1] = Y - o0_first;  // exclusive
    // z_in is increasing in o2 (m02 > 0) and non-increasing in o0 (m00 <= 0) for a light-sheet
    // geometry; take the extremes over all four corners so odd signs stay safe.
    double lo = INFINITY, hi = -INFINITY;
    const int o0s[2] = {o0_first, o0_last};
    const int cs[2] = {cbeg, cend - 1};
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
            volatile double t = (double)o0s[a] * m00;
            volatile double base = shift + t;
            volatile double u = (double)cs[b] * m02;
            const double z = base + u;
            lo = std::fmin(lo, z);
            hi = std::fmax(hi, z);
        }
    if (hi < 0.0 || lo > (double)(Z - 1) || cend <= cbeg) {
        z_range[0] = z_range[1] = 0;  // nothing inside
        return;
    }
    z_range[0] = (int)std::floor(std::fmax(lo, 0.0));
    z_range[1] = (int)std::fmin(std::floor(std::fmin(hi, (double)(Z - 1))) + 1.0, (double)(Z - 1)) + 1;  // exclusive
}

extern "C" int shrimpy_deskew_window_needs(int Z, int Y, int n_avg, double m00, double m02, double shift,
                                           int p_begin, int p_count, int c_begin, int c_count, int32_t y_range[2],
                                           int32_t z_range[2]) {
    if (Z <= 0 || Y <= 0 || n_avg <= 0 || p_begin < 0 || p_count <= 0 || c_begin < 0 || c_count < 0 || !y_range ||
        !z_range)
        return fail(SHRIMPY_EINVAL, "window_needs: bad arguments");
    if (p_begin + p_count > (Y + n_avg - 1) / n_avg) return fail(SHRIMPY_EINVAL, "window_needs: tilt blocks out of range");
    window_needs(Z, Y, n_avg, m00, m02, shift, p_begin, p_count, c_begin, c_begin + c_count, y_range, z_range);
    return SHRIMPY_OK;
}

static int deskew_window_impl(const void *d_raw, int raw_dtype, float *d_out, int Z, int Y, int X, int Xp, int n_avg,
                              double m00, double m02, double shift, float cval, int64_t raw_stride_z,
                              int64_t raw_stride_y, int64_t out_stride_p, int64_t out_stride_1,
                              const shrimpy_window *win, const float *d_scale, float *d_range, int kernel, void *stream) {
    if (Z <= 0 || Y <= 0 || X <= 0 || Xp < 0 || n_avg <= 0)
        return fail(SHRIMPY_EINVAL, "deskew: bad shape Z=%d Y=%d X=%d Xp=%d n=%d", Z, Y, X, Xp, n_avg);
    if (raw_dtype != SHRIMPY_U16 && raw_dtype != SHRIMPY_F32)
        return fail(SHRIMPY_EINVAL, "deskew: raw_dtype must be SHRIMPY_U16 or SHRIMPY_F32, got %d", raw_dtype);
    if (kernel < SHRIMPY_KERNEL_AUTO || kernel > SHRIMPY_KERNEL_TMA_STAGED)
        return fail(SHRIMPY_EINVAL, "deskew: unknown kernel selector %d", kernel);
    if (!std::isfinite(m00) || !std::isfinite(m02) || !std::isfinite(shift))
        return fail(SHRIMPY_EINVAL, "deskew: non-finite affine row");
    const int Yn = (Y + n_avg - 1) / n_avg;
    shrimpy_window full = {0, Yn, 0, Xp, 0, Y, 0, Z};
    const shrimpy_window &w = win ? *win : full;
    if (w.p_begin < 0 || w.p_count < 0 || w.p_begin + w.p_count > Yn || w.c_begin < 0 || w.c_count < 0 ||
        w.c_begin + w.c_count > Xp || w.y_origin < 0 || w.y_count <= 0 || w.y_origin + w.y_count > Y ||
        w.z_origin < 0 || w.z_count <= 0 || w.z_origin + w.z_count > Z)
        return fail(SHRIMPY_EINVAL, "deskew: window outside the stack");
    if (w.p_count == 0 || w.c_count == 0) return SHRIMPY_OK;
    if (!d_raw || !d_out) return fail(SHRIMPY_EINVAL, "deskew: null device pointer");

    int32_t yr[2], zr[2];
    window_needs(Z, Y, n_avg, m00, m02, shift, w.p_begin, w.p_count, w.c_begin, w.c_begin + w.c_count, yr, zr);
    if (yr[0] < w.y_origin || yr[1] > w.y_origin + w.y_count)
        return fail(SHRIMPY_EINVAL, "deskew: slab rows [%d,%d) do not cover the needed [%d,%d)", w.y_origin,
                    w.y_origin + w.y_count, yr[0], yr[1]);
    if (zr[1] > zr[0] && (zr[0] < w.z_origin || zr[1] > w.z_origin + w.z_count))
        return fail(SHRIMPY_EINVAL, "deskew: slab slices [%d,%d) do not cover the needed [%d,%d)", w.z_origin,
                    w.z_origin + w.z_count, zr[0], zr[1]);

    DeskewParams P{};
    P.raw = d_raw;
    P.out = d_out;
    P.Z = Z; P.Y = Y; P.X = X; P.Xp = Xp; P.n = n_avg;
    P.p0 = w.p_begin; P.pcount = w.p_count;
    P.cbeg = w.c_begin; P.cend = w.c_begin + w.c_count;
    P.y_org = w.y_origin; P.y_cnt = w.y_count;
    P.z_org = w.z_origin; P.z_cnt = w.z_count;
    P.raw_sy = raw_stride_y ? raw_stride_y : X;
    P.raw_sz = raw_stride_z ? raw_stride_z : (long long)w.y_count * P.raw_sy;
    P.out_s1 = out_stride_1 ? out_stride_1 : w.c_count;
    P.out_sp = out_stride_p ? out_stride_p : (long long)X * P.out_s1;
    if (P.raw_sy < X || P.out_s1 < w.c_count)
        return fail(SHRIMPY_EINVAL, "deskew: inner strides smaller than the row length");
    P.m00 = m00; P.m02 = m02; P.shift = shift;
    P.cval = cval;
    P.inv_n = 1.0f / (float)n_avg;
    P.scale = d_scale;
    P.range = reinterpret_cast<unsigned *>(d_range);   // the two result floats double as the ordered-key slots
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (d_range) {
        range_init_kernel<<<1, 1, 0, s>>>(P.range);
        count_launch();
    }
    const int err = raw_dtype == SHRIMPY_U16 ? deskew_dispatch<uint16_t>(P, kernel, s) : deskew_dispatch<float>(P, kernel, s);
    if (err == SHRIMPY_OK && d_range) {
        range_finish_kernel<<<1, 1, 0, s>>>(P.range);
        count_launch();
        SHRIMPY_CUDA_TRY(cudaGetLastError());
    }
    return err;
}

extern "C" int shrimpy_deskew_window_device(const void *d_raw, int raw_dtype, float *d_out, int Z, int Y, int X,
                                            int Xp, int n_avg, double m00, double m02, double shift, float cval,
                                            int64_t raw_stride_z, int64_t raw_stride_y, int64_t out_stride_p,
                                            int64_t out_stride_1, const shrimpy_window *win, int kernel,
                                            void *stream) {
    return deskew_window_impl(d_raw, raw_dtype, d_out, Z, Y, X, Xp, n_avg, m00, m02, shift, cval, raw_stride_z,
                              raw_stride_y, out_stride_p, out_stride_1, win, nullptr, nullptr, kernel, stream);
}

extern "C" int shrimpy_deskew_flatfield_device(const void *d_raw, int raw_dtype, const float *d_scale, float *d_out,
                                               int Z, int Y, int X, int Xp, int n_avg, double m00, double m02,
                                               double shift, float cval, int64_t raw_stride_z, int64_t raw_stride_y,
                                               const shrimpy_window *win, int kernel, void *stream) {
    if (!d_scale) return fail(SHRIMPY_EINVAL, "deskew: null flat-field scale field");
    return deskew_window_impl(d_raw, raw_dtype, d_out, Z, Y, X, Xp, n_avg, m00, m02, shift, cval, raw_stride_z,
                              raw_stride_y, 0, 0, win, d_scale, nullptr, kernel, stream);
}

extern "C" int shrimpy_deskew_range_device(const void *d_raw, int raw_dtype, const float *d_scale, float *d_out,
                                           float *d_range2, int Z, int Y, int X, int Xp, int n_avg, double m00,
                                           double m02, double shift, float cval, int64_t raw_stride_z,
                                           int64_t raw_stride_y, int kernel, void *stream) {
    if (!d_range2) return fail(SHRIMPY_EINVAL, "deskew: null value-range pointer");
    return deskew_window_impl(d_raw, raw_dtype, d_out, Z, Y, X, Xp, n_avg, m00, m02, shift, cval, raw_stride_z,
                              raw_stride_y, 0, 0, nullptr, d_scale, d_range2, kernel, stream);
}

extern "C" int shrimpy_deskew_device(const void *d_raw, int raw_dtype, float *d_out, int Z, int Y, int X, int Xp,
                                     int n_avg, double m00, double m02, double shift, float cval,
                                     int64_t raw_stride_z, int64_t raw_stride_y, int64_t out_stride_p,
                                     int64_t out_stride_1, int kernel, void *stream) {
    return shrimpy_deskew_window_device(d_raw, raw_dtype, d_out, Z, Y, X, Xp, n_avg, m00, m02, shift, cval,
                                        raw_stride_z, raw_stride_y, out_stride_p, out_stride_1, nullptr, kernel,
                                        stream);
}
