// Trilinear affine resample (registration step, BASELINE.json configs[2]) for sm_100a.
//
// out[o0,o1,o2] = trilinear(vol, c),  c_a = ((M[a][3] + o0*M[a][0]) + o1*M[a][1]) + o2*M[a][2]
// (float64, products and sums rounded separately: scipy.ndimage.affine_transform's order),
// outside (any c_a < 0 or c_a > dim_a - 1, strict) -> cval; order=1, mode="constant".
//
// affine_gather_kernel: each warp owns 32*kItems consecutive o2 of one output row, lane-major,
// so global stores are coalesced; the (o0,o1) part of the coordinate is computed once per
// thread in float64 and only one multiply-add per axis remains per voxel.  Input taps go
// through the read-only path (L1/L2 absorb the 8-fold tap reuse).
#include "common.cuh"

#include <cmath>

namespace shrimpy {

struct AffineParams {
    const float *in;
    float *out;
    int iz, iy, ix;
    int oz, oy, ox;
    double M[12];
    float cval;
    int nan_to_zero;
    int tiles_x;
};

constexpr int kAffThreads = 128;
constexpr int kAffItems = 4;

__device__ __forceinline__ float tap(const float *__restrict__ p, int nan_to_zero) {
    float v = __ldg(p);
    if (nan_to_zero) {
        // numpy.nan_to_num: nan -> 0, +-inf -> +-FLT_MAX
        if (v != v) v = 0.f;
        else if (isinf(v)) v = copysignf(3.402823466e+38f, v);
    }
    return v;
}

__global__ void __launch_bounds__(kAffThreads) affine_gather_kernel(const AffineParams P) {
    const int o1 = blockIdx.x / P.tiles_x;
    const int xt = blockIdx.x % P.tiles_x;
    const int o0 = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int col0 = (xt * (kAffThreads / 32) + warp) * (32 * kAffItems) + lane;

    double base[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double t = __dadd_rn(P.M[4 * a + 3], __dmul_rn((double)o0, P.M[4 * a + 0]));
        base[a] = __dadd_rn(t, __dmul_rn((double)o1, P.M[4 * a + 1]));
    }
    const double hz = (double)(P.iz - 1), hy = (double)(P.iy - 1), hx = (double)(P.ix - 1);
    const long long sz = (long long)P.iy * P.ix, sy = P.ix;
    float *row = P.out + ((long long)o0 * P.oy + o1) * P.ox;

#pragma unroll
    for (int i = 0; i < kAffItems; ++i) {
        const int o2 = col0 + 32 * i;
        if (o2 >= P.ox) break;
        const double cz = __dadd_rn(base[0], __dmul_rn((double)o2, P.M[2]));
        const double cy = __dadd_rn(base[1], __dmul_rn((double)o2, P.M[6]));
        const double cx = __dadd_rn(base[2], __dmul_rn((double)o2, P.M[10]));
        float r = P.cval;
        if (cz >= 0.0 && cz <= hz && cy >= 0.0 && cy <= hy && cx >= 0.0 && cx <= hx) {
            const double fz = floor(cz), fy = floor(cy), fx = floor(cx);
            const float wz = (float)(cz - fz), wy = (float)(cy - fy), wx = (float)(cx - fx);
            const int z0 = (int)fz, y0 = (int)fy, x0 = (int)fx;
            const int z1 = min(z0 + 1, P.iz - 1), y1 = min(y0 + 1, P.iy - 1), x1 = min(x0 + 1, P.ix - 1);
            const float *p00 = P.in + z0 * sz + y0 * sy;
            const float *p01 = P.in + z0 * sz + y1 * sy;
            const float *p10 = P.in + z1 * sz + y0 * sy;
            const float *p11 = P.in + z1 * sz + y1 * sy;
            const float v000 = tap(p00 + x0, P.nan_to_zero), v001 = tap(p00 + x1, P.nan_to_zero);
            const float v010 = tap(p01 + x0, P.nan_to_zero), v011 = tap(p01 + x1, P.nan_to_zero);
            const float v100 = tap(p10 + x0, P.nan_to_zero), v101 = tap(p10 + x1, P.nan_to_zero);
            const float v110 = tap(p11 + x0, P.nan_to_zero), v111 = tap(p11 + x1, P.nan_to_zero);
            const float a00 = fmaf(wx, v001 - v000, v000);
            const float a01 = fmaf(wx, v011 - v010, v010);
            const float a10 = fmaf(wx, v101 - v100, v100);
            const float a11 = fmaf(wx, v111 - v110, v110);
            const float b0 = fmaf(wy, a01 - a00, a00);
            const float b1 = fmaf(wy, a11 - a10, a10);
            r = fmaf(wz, b1 - b0, b0);
        }
        __stcs(row + o2, r);
    }
}

}  // namespace shrimpy

using namespace shrimpy;

extern "C" int shrimpy_affine_device(const float *d_in, float *d_out, int iz, int iy, int ix, int oz, int oy,
                                     int ox, const double M[12], float cval, int nan_to_zero, void *stream) {
    if (iz <= 0 || iy <= 0 || ix <= 0 || oz < 0 || oy < 0 || ox < 0)
        return fail(SHRIMPY_EINVAL, "affine: bad shape in=(%d,%d,%d) out=(%d,%d,%d)", iz, iy, ix, oz, oy, ox);
    if (!M) return fail(SHRIMPY_EINVAL, "affine: null matrix");
    for (int i = 0; i < 12; ++i)
        if (!std::isfinite(M[i])) return fail(SHRIMPY_EINVAL, "affine: non-finite matrix entry %d", i);
    if ((long long)oz * oy * ox == 0) return SHRIMPY_OK;
    if (!d_in || !d_out) return fail(SHRIMPY_EINVAL, "affine: null device pointer");
    AffineParams P{};
    P.in = d_in; P.out = d_out;
    P.iz = iz; P.iy = iy; P.ix = ix;
    P.oz = oz; P.oy = oy; P.ox = ox;
    for (int i = 0; i < 12; ++i) P.M[i] = M[i];
    P.cval = cval;
    P.nan_to_zero = nan_to_zero;
    const int per_block = kAffThreads * kAffItems;
    P.tiles_x = (ox + per_block - 1) / per_block;
    const long long gx = (long long)P.tiles_x * oy;
    if (gx > 2147483647LL || oz > 65535) return fail(SHRIMPY_EINVAL, "affine: output too large for the grid");
    affine_gather_kernel<<<dim3((unsigned)gx, (unsigned)oz), kAffThreads, 0, static_cast<cudaStream_t>(stream)>>>(P);
    count_launch();
    SHRIMPY_CUDA_TRY(cudaGetLastError());
    return SHRIMPY_OK;
}
