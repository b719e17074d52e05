// Shared pieces of the affine resample kernels (affine.cu: gather / tile / planar-tile kernels,
// affine_stream.cu: z-streaming planar kernel).
#pragma once
#include "common.cuh"

namespace shrimpy {

struct AffineParams {
    const float *in;
    float *out;
    int iz, iy, ix;
    int oz, oy, ox;
    double M[12];
    float cval;
    int nan_to_zero;
    int tiles_x;  // gather kernel: o2 tiles per row; tile kernel: tiles along o2
    int tiles_y, tiles_z;
    int TZ, TY, TX;         // output tile (tile kernel); TY is a power of two
    int log2TY;
    int BZ, BY, BX, pitch;  // staged input box and its odd row pitch (tile kernel)
    unsigned tma_bytes;     // bytes one TMA box load delivers
    int LA, LB;             // planar kernel: tile extent along the lane axis / the other in-plane axis
    int ring_log2, PB, ZC;  // stream kernel: log2(planes in the ring), floats per ring slot, output steps per CTA
    long long in_sy, in_sz; // element strides of the input rows / planes (stream and tilt kernels; ix, ix*iy when dense)
};

constexpr int kAffThreads = 128;
constexpr int kAffItems = 4;
constexpr int kTileThreads = 256;

// numpy.nan_to_num: nan -> 0, +-inf -> +-FLT_MAX
__device__ __forceinline__ float clean(float v) {
    const uint32_t b = __float_as_uint(v);
    if ((b & 0x7f800000u) == 0x7f800000u) v = (b & 0x007fffffu) ? 0.f : __uint_as_float(b - 1u);
    return v;
}

__device__ __forceinline__ float tap(const float *__restrict__ p, int nan_to_zero) {
    const float v = __ldg(p);
    return nan_to_zero ? clean(v) : v;
}

// One axis of the scipy coordinate: exact float64 value -> (floor, fraction), inside test.
// c >= 0  <=>  floor >= 0;   c <= dim-1  <=>  floor < dim-1 or (floor == dim-1 and fraction == 0).
__device__ __forceinline__ bool split_coord(double c, int dim, int &i0, float &w) {
    i0 = __double2int_rd(c);
    w = (float)(c - (double)i0);
    return i0 >= 0 && (i0 < dim - 1 || (i0 == dim - 1 && w == 0.f && c == (double)i0));
}

// Interior fast path: floor and fraction of a coordinate without 64-bit conversions (F2I.F64,
// I2F.F64 and F2F.F32.F64 issue at 1/8 rate).  Adding 1.5*2^29 with round-down leaves
// floor(c * 2^23) + 2^51 in the mantissa: bits [22:0] of the low word are the fraction (23 bits,
// truncated) and the bits above are floor(c) + 2^28.  Valid for |c| < 2^28 (checked on the host).
__device__ __forceinline__ int split_fast(double c, float &w) {
    const double s = __dadd_rd(c, 805306368.0);
    const unsigned hi = (unsigned)__double2hiint(s), lo = (unsigned)__double2loint(s);
    w = __uint_as_float((lo & 0x007fffffu) | 0x3f800000u) - 1.0f;
    return (int)(__funnelshift_l(lo, hi, 9) - 0x90000000u);
}

// ---- packed float32 pairs (FFMA2 / FADD2 on sm_100: half the issue slots of the scalar sequence, same rounding) ----
__device__ __forceinline__ float2 neg2(float2 v) { return make_float2(-v.x, -v.y); }
__device__ __forceinline__ float2 lerp2(float2 w, float2 a, float2 b) {   // fma(w, b - a, a), per half
    return __ffma2_rn(w, __fadd2_rn(b, neg2(a)), a);
}
// Shared load through a 32-bit shared-window address; volatile keeps it behind the mbarrier wait that precedes it.
template <int IMM>
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(IMM));
    return v;
}
__device__ __forceinline__ bool nonfinite(float v) { return (__float_as_uint(v) & 0x7f800000u) == 0x7f800000u; }

// affine_tilt.cu: z-streaming kernel for general matrices with a small z spread per tile.
int launch_affine_tilt(AffineParams P, int nan_to_zero, cudaStream_t s, bool *launched);

// affine_stream.cu: z-streaming kernel for block-diagonal matrices.  *launched = false when not eligible.
int launch_affine_stream(AffineParams P, int nan_to_zero, cudaStream_t s, bool *launched);

}  // namespace shrimpy
