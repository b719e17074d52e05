// Trilinear affine resample (registration step, BASELINE.json configs[2]) for sm_100a.
//
// out[o0,o1,o2] = trilinear(vol, c),  c_a = ((M[a][3] + o0*M[a][0]) + o1*M[a][1]) + o2*M[a][2]
// (float64, products and sums rounded separately: scipy.ndimage.affine_transform's order),
// outside (any c_a < 0 or c_a > dim_a - 1, strict) -> cval; order=1, mode="constant".
//
// Dispatch (shrimpy_affine_device, bottom of this file), fastest first:
//   affine_stream_kernel  (affine_stream.cu) block-diagonal matrices: z-streaming march, plane values reused.
//   affine_tilt_kernel    (affine_tilt.cu)   general matrices whose z spread over a tile is a few planes.
//   affine_tile_kernel    (here) any matrix whose tile footprint fits shared memory: one CTA per output tile
//                         (TZ,TY,TX); the input bounding box of the tile (an affine image of a box is bounded
//                         by its 8 corners) is staged by ONE 3-D TMA box load when a warp's 32 consecutive o2
//                         stay within a few input rows, or by cp.async row copies at an ODD row pitch when
//                         lanes walk input y or z, so the tap reads are bank-conflict-free either way.
//                         nan_to_num is applied lazily: a non-finite result (which every non-finite tap
//                         produces) is recomputed from cleaned taps.  The host picks the tile shape.
//   affine_gather_kernel  (here) one thread per output voxel, taps through the read-only global path: any
//                         matrix, shape and alignment.
#include "affine_common.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace shrimpy {

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
    const long long sz = (long long)P.iy * P.ix, sy = P.ix;
    float *row = P.out + ((long long)o0 * P.oy + o1) * P.ox;

#pragma unroll
    for (int i = 0; i < kAffItems; ++i) {
        const int o2 = col0 + 32 * i;
        if (o2 >= P.ox) break;
        int z0, y0, x0;
        float wz, wy, wx;
        bool in = split_coord(__dadd_rn(base[0], __dmul_rn((double)o2, P.M[2])), P.iz, z0, wz);
        in &= split_coord(__dadd_rn(base[1], __dmul_rn((double)o2, P.M[6])), P.iy, y0, wy);
        in &= split_coord(__dadd_rn(base[2], __dmul_rn((double)o2, P.M[10])), P.ix, x0, wx);
        float r = P.cval;
        if (in) {
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

// ---- shared-memory tiled kernel ------------------------------------------------------------------

__device__ __forceinline__ void cp_async4(uint32_t dst, const float *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}

// Slow, fully general voxel: exact edge rule (c == dim-1 is inside, the second tap folds onto the
// first there) and clamped box indices.  Only voxels on the rim of the input volume come here.
// Slow, fully general voxel, recomputed from the output index with scipy's exact coordinate
// arithmetic: exact edge rule (c == dim-1 is inside, the second tap folds onto the first there),
// clamped box indices, optional nan_to_num of the taps.  Voxels within one input voxel of the rim of
// the input volume, voxels outside it and voxels with a non-finite fast-path result come here.
__device__ __noinline__ float affine_edge_voxel(const float *box, const AffineParams *Pp, int o0, int o1, int o2,
                                                int oz0, int oy0, int ox0, int clean_taps) {
    const AffineParams &P = *Pp;
    double c[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        double t = __dadd_rn(P.M[4 * a + 3], __dmul_rn((double)o0, P.M[4 * a + 0]));
        t = __dadd_rn(t, __dmul_rn((double)o1, P.M[4 * a + 1]));
        c[a] = __dadd_rn(t, __dmul_rn((double)o2, P.M[4 * a + 2]));
    }
    int z0, y0, x0;
    float wz, wy, wx;
    bool in = split_coord(c[0], P.iz, z0, wz);
    in &= split_coord(c[1], P.iy, y0, wy);
    in &= split_coord(c[2], P.ix, x0, wx);
    if (!in) return P.cval;
    const int zs = P.BY * P.pitch;
    const int dz = (z0 + 1 < P.iz) ? zs : 0;
    const int dy = (y0 + 1 < P.iy) ? P.pitch : 0;
    const int dx = (x0 + 1 < P.ix) ? 1 : 0;
    const int bz = min(max(z0 - oz0, 0), P.BZ - 1);
    const int by = min(max(y0 - oy0, 0), P.BY - 1);
    const int bx = min(max(x0 - ox0, 0), P.BX - 1);
    const float *q = box + bz * zs + by * P.pitch + bx;
    float v[8] = {q[0], q[dx], q[dy], q[dy + dx], q[dz], q[dz + dx], q[dz + dy], q[dz + dy + dx]};
    if (clean_taps) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = clean(v[i]);
    }
    const float a00 = fmaf(wx, v[1] - v[0], v[0]);
    const float a01 = fmaf(wx, v[3] - v[2], v[2]);
    const float a10 = fmaf(wx, v[5] - v[4], v[4]);
    const float a11 = fmaf(wx, v[7] - v[6], v[6]);
    const float b0 = fmaf(wy, a01 - a00, a00);
    const float b1 = fmaf(wy, a11 - a10, a10);
    return fmaf(wz, b1 - b0, b0);
}

// ITEMS = TX / 32 output columns per lane and tile row; CLEAN = nan_to_num semantics;
// USE_TMA = stage the box with one TMA load (dense pitch) instead of cp.async rows (odd pitch);
// ZSEP = the input (y, x) of a voxel do not depend on o0 (M[1][0] == M[2][0] == 0: in-plane
// registration, z shift/scale) -> the bilinear value of an input plane is shared by consecutive o0.
template <int ITEMS, bool CLEAN, bool USE_TMA, bool ZSEP>
__global__ void __launch_bounds__(kTileThreads, 2)
    affine_tile_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ AffineParams P) {
    extern __shared__ __align__(128) float box_raw[];
    __shared__ int s_org[3];
    __shared__ __align__(8) uint64_t bar;
    // 128-byte aligned start, computed on the shared-window address so the loads stay LDS
    float *box = box_raw + (((128u - (smem_u32(box_raw) & 127u)) & 127u) >> 2);

    // z tiles fastest, then x, then y: neighbours that share a halo run close together in time,
    // so halo re-reads are L2 hits and DRAM sees each input voxel about once.
    const int tz = blockIdx.x % P.tiles_z;
    const int tx = blockIdx.x / P.tiles_z;
    const int ty = blockIdx.y;
    const int t0z = tz * P.TZ, t0y = ty * P.TY, t0x = tx * (32 * ITEMS);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // Box origin: floor of the minimum corner coordinate, per input axis.
    if (threadIdx.x < 3) {
        const int a = threadIdx.x;
        const double e0 = (double)(min(t0z + P.TZ, P.oz) - 1 - t0z);
        const double e1 = (double)(min(t0y + P.TY, P.oy) - 1 - t0y);
        const double e2 = (double)(min(t0x + 32 * ITEMS, P.ox) - 1 - t0x);
        const double m0 = P.M[4 * a + 0], m1 = P.M[4 * a + 1], m2 = P.M[4 * a + 2];
        double lo = P.M[4 * a + 3] + t0z * m0 + t0y * m1 + t0x * m2;
        lo += fmin(e0 * m0, 0.0) + fmin(e1 * m1, 0.0) + fmin(e2 * m2, 0.0);
        const int dim = a == 0 ? P.iz : a == 1 ? P.iy : P.ix;
        // clamp far-away tiles so the int conversion and the pointer arithmetic below stay defined
        int org = __double2int_rd(fmin(fmax(lo - 1e-6, -4.0), (double)dim));
        // TMA: the box must start on a 16-byte boundary of the innermost axis (an unaligned start
        // raises an illegal-instruction fault on sm_100), so x is rounded down to 4 floats.
        if (USE_TMA && a == 2) org &= ~3;
        s_org[a] = org;
    }
    __syncthreads();
    const int oz0 = s_org[0], oy0 = s_org[1], ox0 = s_org[2];
    const int pitch = P.pitch;
    const int zs = P.BY * pitch;

    if (USE_TMA) {
        // One box load; coordinates outside the volume are zero-filled and never read.
        if (threadIdx.x == 0) {
            mbar_init(&bar, 1);
            fence_mbar_init();
            mbar_arrive_expect_tx(&bar, P.tma_bytes);
            tma_load_3d(smem_u32(box), &tmap, ox0, oy0, oz0, &bar);
        }
        __syncthreads();
        mbar_wait(&bar, 0);
    } else {
        // One (z,y) row per warp iteration, lanes along x (coalesced), all copies in flight at once.
        const int rows = P.BZ * P.BY;
        const int lo = max(0, -ox0), hi = min(P.BX, P.ix - ox0);
        const uint32_t box_s = smem_u32(box);
        for (int r = warp; r < rows; r += kTileThreads / 32) {
            const int bz = r / P.BY, by = r - bz * P.BY;
            const int gz = oz0 + bz, gy = oy0 + by;
            if ((unsigned)gz >= (unsigned)P.iz || (unsigned)gy >= (unsigned)P.iy) continue;
            const float *src = P.in + ((long long)gz * P.iy + gy) * P.ix + ox0;
            const int dst = r * pitch;
#pragma unroll 4
            for (int bx = lo + lane; bx < hi; bx += 32) cp_async4(box_s + 4u * (uint32_t)(dst + bx), src + bx);
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();
    }

    // Resample, column-marching: a thread owns output columns (o1, o2) -- lanes hold consecutive o2, so
    // stores are coalesced -- and walks o0 through the tile.  The coordinate advances by ONE float64 add
    // per axis and voxel (c += M_a0).  That drifts from scipy's ((M_a3 + o0 M_a0) + o1 M_a1) + o2 M_a2 by
    // < 1e-11 voxel, harmless for a voxel whose taps are at least one input voxel away from the rim (the
    // interpolant is continuous); every other voxel is recomputed exactly by affine_edge_voxel.
    const int org_off = oz0 * zs + oy0 * pitch + ox0;
    const unsigned hz = (unsigned)max(P.iz - 3, 0), hy = (unsigned)max(P.iy - 3, 0), hx = (unsigned)max(P.ix - 3, 0);
    const unsigned uz = (unsigned)P.iz, uy = (unsigned)P.iy, ux = (unsigned)P.ix;
    const double sz = P.M[0], sy = P.M[4], sx = P.M[8];   // per-step increments along o0
    const int nz = min(P.TZ, P.oz - t0z);
    const long long plane = (long long)P.oy * P.ox;
    for (int ly = warp; ly < P.TY; ly += kTileThreads / 32) {
        const int o1 = t0y + ly;
        if (o1 >= P.oy) break;
#pragma unroll 1
        for (int i = 0; i < ITEMS; ++i) {
            const int o2 = t0x + lane + 32 * i;
            if (o2 >= P.ox) break;
            double cz = ((P.M[3] + (double)t0z * sz) + (double)o1 * P.M[1]) + (double)o2 * P.M[2];
            double cy = ((P.M[7] + (double)t0z * sy) + (double)o1 * P.M[5]) + (double)o2 * P.M[6];
            double cx = ((P.M[11] + (double)t0z * sx) + (double)o1 * P.M[9]) + (double)o2 * P.M[10];
            float *ptr0 = P.out + ((long long)t0z * P.oy + o1) * P.ox + o2;
            float *ptr = ptr0;
            // Steps that need the exact slow path (near the rim of the input, outside it, non-finite taps) are
            // only RECORDED in the hot loop -- it stays straight-line code with safe addresses -- and patched
            // afterwards by affine_edge_voxel.  TZ <= 32, so one bit per step.
            unsigned rare_mask = 0;
            if (ZSEP) {
                // (y, x) part once per column; per step only the z coordinate moves.  The bilinear value of an
                // input plane at this column's (y, x) is shared by consecutive steps: b0 = plane zc, b1 = plane zc+1.
                float wy, wx;
                const int y0 = split_fast(cy, wy), x0 = split_fast(cx, wx);
                const bool col_fast = (unsigned)(y0 - 1) < hy && (unsigned)(x0 - 1) < hx;
                const bool col_out = (unsigned)(y0 + 1) > uy || (unsigned)(x0 + 1) > ux;   // certainly outside
                const float *colq = col_fast ? box + (y0 * pitch + x0 - org_off) : box - oz0 * zs;  // safe when slow
                int zc = -0x40000000;
                float b0 = 0.f, b1 = 0.f;
#pragma unroll 2
                for (int lz = 0; lz < nz; ++lz) {
                    float wz;
                    const int z0 = split_fast(cz, wz);
                    const bool fast = col_fast && (unsigned)(z0 - 1) < hz;
                    const int zq = fast ? z0 : oz0;                    // oz0: first plane of the box, always valid
                    const float *q = colq + zq * zs;
                    if (zq != zc && zq != zc + 1) {                    // first step, or |z scale| > 1: no reuse
                        const float a0 = fmaf(wx, q[1] - q[0], q[0]);
                        const float a1 = fmaf(wx, q[pitch + 1] - q[pitch], q[pitch]);
                        b1 = fmaf(wy, a1 - a0, a0);
                        zc = zq - 1;
                    }
                    const float *q2 = q + zs;
                    const float a2 = fmaf(wx, q2[1] - q2[0], q2[0]);
                    const float a3 = fmaf(wx, q2[pitch + 1] - q2[pitch], q2[pitch]);
                    const float bnew = fmaf(wy, a3 - a2, a2);
                    const bool same = zq == zc;
                    b0 = same ? b0 : b1;
                    b1 = same ? b1 : bnew;
                    zc = zq;
                    const bool outside = col_out || (unsigned)(z0 + 1) > uz;
                    const float res = outside ? P.cval : fmaf(wz, b1 - b0, b0);
                    const bool bad = !outside && (!fast || (CLEAN && (__float_as_uint(res) & 0x7f800000u) == 0x7f800000u));
                    rare_mask |= (unsigned)bad << lz;
                    __stcs(ptr, res);
                    ptr += plane;
                    cz += sz;
                }
            } else {
#pragma unroll 2
                for (int lz = 0; lz < nz; ++lz) {
                    float wz, wy, wx;
                    const int z0 = split_fast(cz, wz), y0 = split_fast(cy, wy), x0 = split_fast(cx, wx);
                    // 1 <= floor(c) <= dim-3 on every axis: both taps exist, nothing to clamp
                    const bool fast = (unsigned)(z0 - 1) < hz && (unsigned)(y0 - 1) < hy && (unsigned)(x0 - 1) < hx;
                    const float *q = fast ? box + (z0 * zs + y0 * pitch + x0 - org_off) : box;
                    const float *q1 = q + pitch, *q2 = q + zs, *q3 = q2 + pitch;
                    const float v000 = q[0], v001 = q[1], v010 = q1[0], v011 = q1[1];
                    const float v100 = q2[0], v101 = q2[1], v110 = q3[0], v111 = q3[1];
                    const float a00 = fmaf(wx, v001 - v000, v000);
                    const float a01 = fmaf(wx, v011 - v010, v010);
                    const float a10 = fmaf(wx, v101 - v100, v100);
                    const float a11 = fmaf(wx, v111 - v110, v110);
                    const float b0 = fmaf(wy, a01 - a00, a00);
                    const float b1 = fmaf(wy, a11 - a10, a10);
                    // floor(c) <= -2 or >= dim on some axis: certainly outside, no exact test needed
                    const bool outside = (unsigned)(z0 + 1) > uz || (unsigned)(y0 + 1) > uy || (unsigned)(x0 + 1) > ux;
                    const float res = outside ? P.cval : fmaf(wz, b1 - b0, b0);
                    const bool bad = !outside && (!fast || (CLEAN && (__float_as_uint(res) & 0x7f800000u) == 0x7f800000u));
                    rare_mask |= (unsigned)bad << lz;
                    __stcs(ptr, res);
                    ptr += plane;
                    cz += sz;
                    cy += sy;
                    cx += sx;
                }
            }
            while (rare_mask) {
                const int lz = __ffs(rare_mask) - 1;
                rare_mask &= rare_mask - 1;
                __stcs(ptr0 + lz * plane, affine_edge_voxel(box, &P, t0z + lz, o1, o2, oz0, oy0, ox0, CLEAN));
            }
        }
    }
}

// Host: pick the output tile whose staged input box is smallest per output voxel.
static bool choose_tile(AffineParams &P, int smem_limit, bool tma) {
    static const int cand[][3] = {{8, 16, 64}, {4, 16, 64}, {4, 8, 64},  {2, 16, 64},  {8, 8, 64},   {4, 32, 64},
                                  {2, 32, 64}, {1, 32, 64}, {4, 8, 128}, {2, 16, 128}, {2, 8, 128},  {1, 16, 128},
                                  {4, 32, 32}, {2, 32, 32}, {8, 32, 32}, {1, 64, 64},  {1, 32, 128}, {2, 64, 32},
                                  {8, 8, 128}, {4, 16, 128}, {8, 16, 128}, {4, 4, 128}, {8, 4, 128},
                                  {16, 8, 64}, {16, 16, 64}, {16, 4, 64}, {16, 8, 32}, {16, 16, 32}, {32, 8, 32}};
    double best = 1e300;
    bool found = false;
    int forced[3] = {0, 0, 0};
    if (const char *f = getenv("SHRIMPY_AFFINE_TILE")) sscanf(f, "%d,%d,%d", &forced[0], &forced[1], &forced[2]);
    for (const auto &c : cand) {
        if (forced[0] && (c[0] != forced[0] || c[1] != forced[1] || c[2] != forced[2])) continue;
        const int T[3] = {c[0], c[1], c[2]};  // not clipped to the output: TY must stay a power of two
        long long B[3];
        for (int a = 0; a < 3; ++a) {
            double span = 0;
            for (int b = 0; b < 3; ++b) span += std::fabs(P.M[4 * a + b]) * (T[b] - 1);
            B[a] = (long long)std::ceil(span) + 3;
        }
        if (tma) B[2] = (B[2] + 3 + 3) / 4 * 4;           // TMA rows start and end on 16-byte boundaries
        const long long pitch = tma ? B[2] : (B[2] | 1);  // cp.async rows sit at an odd pitch
        const long long bytes = B[0] * B[1] * pitch * 4 + 128;
        if (bytes > smem_limit || B[0] > 4096 || B[1] > 4096 || B[2] > 8192) continue;
        if (tma && (B[0] > 256 || B[1] > 256 || B[2] > 256)) continue;
        const double outputs = (double)T[0] * T[1] * T[2];
        // cost: staged elements per output, short rows are penalised (poor coalescing), and a
        // mild preference for larger tiles (fewer CTAs, per-row float64 set-up amortised)
        double cost = (double)(B[0] * B[1] * B[2]) / outputs;
        if (B[2] < 32) cost *= 32.0 / (double)B[2];
        cost += 256.0 / outputs + (T[2] == 32 ? 0.15 : 0.0);
        if (cost < best) {
            best = cost;
            found = true;
            P.TZ = T[0]; P.TY = T[1]; P.TX = T[2];
            P.log2TY = 0;
            while ((1 << P.log2TY) < P.TY) ++P.log2TY;
            P.BZ = (int)B[0]; P.BY = (int)B[1]; P.BX = (int)B[2];
            P.pitch = (int)pitch;
        }
    }
    return found;
}

}  // namespace shrimpy

using namespace shrimpy;

static int affine_impl(const float *d_in, float *d_out, int iz, int iy, int ix, long long in_sz, long long in_sy, int oz,
                       int oy, int ox, const double M[12], float cval, int nan_to_zero, void *stream) {
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
    const bool dense = in_sy == ix && in_sz == (long long)ix * iy;
    P.in_sy = in_sy;
    P.in_sz = in_sz;
    cudaStream_t s = static_cast<cudaStream_t>(stream);

    const char *force = getenv("SHRIMPY_AFFINE_KERNEL");
    bool want_gather = force && force[0] == 'g';
    // the tiled kernel's fast coordinate split needs |c| < 2^28 everywhere on the output grid
    for (int a = 0; a < 3; ++a) {
        const double reach = std::fabs(M[4 * a + 3]) + std::fabs(M[4 * a]) * oz + std::fabs(M[4 * a + 1]) * oy +
                             std::fabs(M[4 * a + 2]) * ox;
        if (!(reach < 134217728.0)) want_gather = true;
    }
    // Block-diagonal matrix (z <-> z, (y,x) <-> (y,x)): the z-streaming planar kernel (affine_stream.cu).
    if (!want_gather && (!force || force[0] == 's')) {
        bool launched = false;
        const int rc = launch_affine_stream(P, nan_to_zero, s, &launched);
        if (rc != SHRIMPY_OK || launched) return rc;
        if (force && force[1] == '!') return fail(SHRIMPY_EINVAL, "affine: the stream kernel was forced but is not eligible");
    }
    // General matrix with a small z spread per tile (rotations of a few degrees): the marching tilt kernel.
    if (!want_gather && (!force || force[0] == 'm')) {
        bool launched = false;
        const int rc = launch_affine_tilt(P, nan_to_zero, s, &launched);
        if (rc != SHRIMPY_OK || launched) return rc;
        if (force && force[1] == '!') return fail(SHRIMPY_EINVAL, "affine: the tilt kernel was forced but is not eligible");
    }
    if (!dense)   // only the plane-streaming kernels take row / plane strides (through the tensor map)
        return fail(SHRIMPY_EINVAL, "affine: strided input is not eligible for the streaming kernels");
    const int smem_limit = 56 * 1024;  // 4 CTAs per SM: staging of one tile overlaps the maths of others
    // TMA staging (dense pitch) when a warp's 32 consecutive o2 touch only a few input rows and the
    // tensor map constraints hold; otherwise cp.async rows at an odd pitch.
    bool tma = !(force && force[0] == 'c') && std::fabs(M[6]) * 32 <= 4.0 && std::fabs(M[2]) * 32 <= 4.0 &&
               (reinterpret_cast<uintptr_t>(d_in) & 15u) == 0 && ix % 4 == 0 && tensor_map_encoder() != nullptr;
    bool tiled = !want_gather && choose_tile(P, smem_limit, tma);
    if (!tiled && !want_gather && tma) {   // the TMA box did not fit: retry with cp.async staging
        tma = false;
        tiled = choose_tile(P, smem_limit, false);
    }
    if (tiled) {
        P.tiles_x = (ox + P.TX - 1) / P.TX;
        P.tiles_y = (oy + P.TY - 1) / P.TY;
        P.tiles_z = (oz + P.TZ - 1) / P.TZ;
        CUtensorMap tmap{};
        if (tma) {
            const cuuint64_t gdim[3] = {(cuuint64_t)ix, (cuuint64_t)iy, (cuuint64_t)iz};
            const cuuint64_t gstride[2] = {(cuuint64_t)ix * 4, (cuuint64_t)ix * iy * 4};
            const cuuint32_t bdim[3] = {(cuuint32_t)P.BX, (cuuint32_t)P.BY, (cuuint32_t)P.BZ};
            P.tma_bytes = bdim[0] * bdim[1] * bdim[2] * 4u;
            const cuuint32_t estr[3] = {1u, 1u, 1u};
            const CUresult rc = tensor_map_encoder()(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(d_in),
                                                     gdim, gstride, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (rc != CUDA_SUCCESS) return fail(SHRIMPY_ECUDA, "affine: cuTensorMapEncodeTiled failed (%d)", (int)rc);
        }
        if (P.tiles_y <= 65535 && (long long)P.tiles_x * P.tiles_z <= 2147483647LL) {
            const size_t smem = (size_t)P.BZ * P.BY * P.pitch * sizeof(float) + 128;
            void (*kern)(const CUtensorMap, const AffineParams) = nullptr;
            const int items = P.TX / 32;
            const bool zsep = M[4] == 0.0 && M[8] == 0.0 && !getenv("SHRIMPY_AFFINE_NO_ZSEP");
#define SHRIMPY_PICK_I(C, T, Z) \
    (items == 1 ? affine_tile_kernel<1, C, T, Z> : items == 2 ? affine_tile_kernel<2, C, T, Z> : affine_tile_kernel<4, C, T, Z>)
#define SHRIMPY_PICK_Z(C, T) (zsep ? SHRIMPY_PICK_I(C, T, true) : SHRIMPY_PICK_I(C, T, false))
            if (nan_to_zero) kern = tma ? SHRIMPY_PICK_Z(true, true) : SHRIMPY_PICK_Z(true, false);
            else kern = tma ? SHRIMPY_PICK_Z(false, true) : SHRIMPY_PICK_Z(false, false);
#undef SHRIMPY_PICK_Z
#undef SHRIMPY_PICK_I
            if (getenv("SHRIMPY_DEBUG"))
                fprintf(stderr, "[shrimpy] affine tile T=(%d,%d,%d) B=(%d,%d,%d) pitch=%d tma=%d clean=%d smem=%zu grid=(%d,%d)\n",
                        P.TZ, P.TY, P.TX, P.BZ, P.BY, P.BX, P.pitch, (int)tma, nan_to_zero, smem,
                        P.tiles_x * P.tiles_z, P.tiles_y);
            if (smem + 1024 > 48 * 1024)  // static smem and the per-CTA reserve count against the 48 KB default
                SHRIMPY_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<dim3((unsigned)(P.tiles_x * P.tiles_z), (unsigned)P.tiles_y), kTileThreads, smem, s>>>(tmap, P);
            count_launch();
            SHRIMPY_CUDA_TRY(cudaGetLastError());
            return SHRIMPY_OK;
        }
    }
    const int per_block = kAffThreads * kAffItems;
    P.tiles_x = (ox + per_block - 1) / per_block;
    const long long gx = (long long)P.tiles_x * oy;
    if (gx > 2147483647LL || oz > 65535) return fail(SHRIMPY_EINVAL, "affine: output too large for the grid");
    affine_gather_kernel<<<dim3((unsigned)gx, (unsigned)oz), kAffThreads, 0, s>>>(P);
    count_launch();
    SHRIMPY_CUDA_TRY(cudaGetLastError());
    return SHRIMPY_OK;
}

extern "C" int shrimpy_affine_device(const float *d_in, float *d_out, int iz, int iy, int ix, int oz, int oy,
                                     int ox, const double M[12], float cval, int nan_to_zero, void *stream) {
    return affine_impl(d_in, d_out, iz, iy, ix, (long long)ix * iy, ix, oz, oy, ox, M, cval, nan_to_zero, stream);
}

extern "C" int shrimpy_affine_strided_device(const float *d_in, float *d_out, int iz, int iy, int ix,
                                             int64_t in_stride_z, int64_t in_stride_y, int oz, int oy, int ox,
                                             const double M[12], float cval, int nan_to_zero, void *stream) {
    if (in_stride_y < ix || in_stride_z < in_stride_y * iy)
        return fail(SHRIMPY_EINVAL, "affine: input strides smaller than the row / plane");
    return affine_impl(d_in, d_out, iz, iy, ix, in_stride_z, in_stride_y, oz, oy, ox, M, cval, nan_to_zero, stream);
}
