// Z-streaming affine resample for GENERAL matrices with a small z spread per tile (registration step,
// BASELINE.json configs[2]: rotations of a few degrees about every axis, anisotropic scale, translation).
//
// Same machine as affine_stream.cu -- one CTA owns an output tile in (o1, o2) and marches through o0 while input
// planes stream through a shared-memory ring with one 2-D TMA box per plane -- but here the
// z coordinate of a voxel depends on (o1, o2) too, so
//   * a step needs a WINDOW of planes [floor(min cz), floor(max cz) + 1] over the tile; the window slides
//     monotonically with o0 and plane z lives in slot z & (R - 1);
//   * the (y, x) footprint of the tile drifts with o0; the staged box covers the whole march of the launch
//     (the host bounds the march length so that it fits);
//   * every CTA tabulates, per step (one thread per step, float64 corner arithmetic): the planes that must have
//     landed / may be handed back, the coordinate of the tile's first voxel relative to (ring origin, box origin)
//     in 9.23 fixed point, and a CLASS: "fast" (every voxel of the tile strictly inside the input: no per-voxel
//     test at all), "outside" (constant fill) or "mixed" (the tile straddles the rim: voxels are classified one by
//     one with a float32 coordinate and a 0.05 margin -- safely interior ones take the fast arithmetic, safely
//     outside ones the pad value, the thin shell in between goes through tilt_exact_voxel, scipy's exact float64
//     arithmetic and edge rule from the output index).
// A voxel of a fast step is  base(step) + offset(column)  in 32-bit fixed point: one integer add per axis; floor,
// ring slot, in-box offset and the 23-bit lerp weight are bit fields of the sum.  Nothing accumulates: the error
// against scipy's float64 coordinate is < 2e-7 voxel (the weight itself has 23 bits), harmless strictly inside the
// volume because the interpolant is continuous; the inside/outside decision never depends on it.
// Interior arithmetic is issued as packed float32 pairs over two columns of a thread (FFMA2 / FADD2).
#include "affine_common.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace shrimpy {

constexpr int kTiltConsumerWarps = 8;
constexpr int kTiltConsumers = 32 * kTiltConsumerWarps;
// Who issues the plane loads after the first R: a ninth (producer) warp waiting on per-slot "empty" barriers, or the
// last of the eight warps to hand a plane back (a shared counter per slot).  Measured on config 3: lanes along o2
// run 4 % faster with the producer warp (the counter's atomic sits in every warp's critical path), lanes along o1
// (all warps meet at a barrier every step anyway) 4 % faster without it.
__host__ __device__ constexpr int tilt_threads(bool swap) { return swap ? kTiltConsumers : kTiltConsumers + 32; }
constexpr int kTiltMaxSteps = 128;
constexpr int kTiltMaxRing = 16;

struct TiltParams {
    int t0z, nsteps;         // output steps [t0z, t0z + nsteps) of this launch
    int dir;                 // +1: planes are visited upwards, -1: downwards
    int pad;
};

// Exact voxel: scipy's coordinate arithmetic from the output index, exact edge rule, clamped taps, nan_to_num.
__device__ __noinline__ float tilt_exact_voxel(const float *ring, const AffineParams *Pp, int o0, int o1, int o2, int oy0,
                                               int ox0, int clean_taps) {
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
    const int mask = (1 << P.ring_log2) - 1;
    const int z1 = min(z0 + 1, P.iz - 1);
    const float *pa = ring + (size_t)(z0 & mask) * P.PB, *pb = ring + (size_t)(z1 & mask) * P.PB;
    const int dy = (y0 + 1 < P.iy) ? P.pitch : 0;
    const int dx = (x0 + 1 < P.ix) ? 1 : 0;
    const int q = min(max(y0 - oy0, 0), P.BY - 2) * P.pitch + min(max(x0 - ox0, 0), P.BX - 2);
    float v[8] = {pa[q], pa[q + dx], pa[q + dy], pa[q + dy + dx], pb[q], pb[q + dx], pb[q + dy], pb[q + dy + dx]};
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

// Per-step record of a CTA (built in the prologue, one thread per step).
struct StepInfo {
    unsigned bz, by, bx;   // coordinate of the tile's first voxel relative to (ring origin, box origin), 9.23 fixed point
    unsigned ctl;          // bits 0-11: planes [0, need) must have landed; 12-23: planes [0, rel) may be handed back; 24-25: class
};
struct StepCoarse {
    float z, y, x, pad;    // the same voxel's absolute coordinate in float32 (mixed steps classify voxels with it)
};
enum { kStepMixed = 0, kStepFast = 1, kStepOutside = 2 };

__device__ __forceinline__ void tilt_consumer_sync() {
    asm volatile("bar.sync 1, %0;" ::"n"(kTiltConsumers) : "memory");
}

// 23 fraction bits of a 9.23 coordinate -> lerp weight in [0, 1): (c & 0x7fffff) | 0x3f800000 as ONE LOP3 (the
// constant 1.0f sits in a register), minus 1.
__device__ __forceinline__ float frac23(unsigned c, unsigned one_bits) {
    unsigned r;
    asm("lop3.b32 %0, %1, 0x007fffff, %2, 0xEA;" : "=r"(r) : "r"(c), "r"(one_bits));
    return __uint_as_float(r) - 1.0f;
}

// IA = items along the lane axis (tile extent 32*IA), RB = rows per warp (tile extent 8*RB along the other axis).
// Lanes run along o2, or along o1 when input x follows o1 (SWAP: ~90 degree in-plane maps); then a step's results
// are transposed through a double-buffered shared tile so that global stores stay coalesced along o2.
template <int IA, int RB, bool SWAP, bool CLEAN>
__global__ void __launch_bounds__(tilt_threads(SWAP), (IA * RB <= 4) ? 3 : 2)
    affine_tilt_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ AffineParams P,
                       const __grid_constant__ TiltParams Q) {
    constexpr int NC = IA * RB, NC2 = NC / 2;
    static_assert(NC % 2 == 0 && NC <= 8, "columns are processed in packed pairs");
    constexpr int LA = 32 * IA, LB = 8 * RB;
    constexpr int TY = SWAP ? LA : LB, TX = SWAP ? LB : LA;

    extern __shared__ __align__(128) float smem_raw[];
    constexpr bool kProducerWarp = !SWAP;
    __shared__ __align__(8) uint64_t full[kTiltMaxRing], empty[kTiltMaxRing];
    __shared__ unsigned released_by[kTiltMaxRing];   // !kProducerWarp: warps that have handed the slot's current plane back
    __shared__ __align__(16) StepInfo tab[kTiltMaxSteps];
    __shared__ __align__(16) StepCoarse coarse[kTiltMaxSteps];
    __shared__ int s_zmin, s_zmax;
    float *ring = smem_raw + (((128u - (smem_u32(smem_raw) & 127u)) & 127u) >> 2);
    uint32_t full_s = smem_u32(full), empty_s = smem_u32(empty), ring_s = smem_u32(ring), tab_s = smem_u32(tab);
    // Opaque to the compiler from here on: otherwise every use in the march re-derives the shared window base
    // (S2R SR_CgaCtaId + MOV + LEA, a ~25-cycle special-register read on the critical path of every step).
    asm volatile("" : "+r"(full_s), "+r"(empty_s), "+r"(ring_s), "+r"(tab_s));

    const int ring_log2 = P.ring_log2, mask = (1 << ring_log2) - 1;
    const int pitch = P.pitch;
    const unsigned slot_bytes = (unsigned)P.PB * 4u, pitch4 = (unsigned)pitch * 4u;
    const int t0y = blockIdx.y * TY, t0x = blockIdx.x * TX, t0z = Q.t0z;
    const int nsteps = Q.nsteps, dir = Q.dir;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool tile_full = t0y + TY <= P.oy && t0x + TX <= P.ox;

    if (tid == 0) {
        s_zmin = 0x7fffffff;
        s_zmax = -0x7fffffff;
        for (int i = 0; i <= mask; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], kTiltConsumerWarps);
            released_by[i] = 0;
        }
        fence_mbar_init();
    }
    // extents of the (clipped) tile and of the march, as float64 offsets from the tile's first voxel
    const double e0 = (double)(nsteps - 1);
    const double e1 = (double)(min(t0y + TY, P.oy) - 1 - t0y), e2 = (double)(min(t0x + TX, P.ox) - 1 - t0x);
    // (y, x) origin of the staged box: minimum corner of tile x march under the affine map
    int org[2];
#pragma unroll
    for (int a = 1; a < 3; ++a) {
        const double m0 = P.M[4 * a], m1 = P.M[4 * a + 1], m2 = P.M[4 * a + 2];
        const double lo = P.M[4 * a + 3] + t0z * m0 + t0y * m1 + t0x * m2 + fmin(e0 * m0, 0.0) + fmin(e1 * m1, 0.0) +
                          fmin(e2 * m2, 0.0);
        const int dim = a == 1 ? P.iy : P.ix;
        org[a - 1] = __double2int_rd(fmin(fmax(lo - 1e-6, -4.0), (double)dim));
    }
    org[1] &= ~3;   // TMA: the box starts on a 16-byte boundary along x
    const int oy0 = org[0], ox0 = org[1];
    __syncthreads();

    // ---- per-step record: plane window, class and start coordinate of this tile ----------------------------------------
    int lo_k = 0, hi_k = 0, any = 0, cls = kStepMixed;
    double c0[3] = {0.0, 0.0, 0.0};
    if (tid < nsteps) {
        bool all_fast = true, some_out = false;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double m1 = P.M[4 * a + 1], m2 = P.M[4 * a + 2];
            // scipy's value at the tile's first voxel, and bounds over the tile
            c0[a] = __dadd_rn(__dadd_rn(__dadd_rn(P.M[4 * a + 3], __dmul_rn((double)(t0z + tid), P.M[4 * a])),
                                        __dmul_rn((double)t0y, m1)), __dmul_rn((double)t0x, m2));
            const double cmin = c0[a] + fmin(e1 * m1, 0.0) + fmin(e2 * m2, 0.0);
            const double cmax = c0[a] + fmax(e1 * m1, 0.0) + fmax(e2 * m2, 0.0);
            const double dim = (double)(a == 0 ? P.iz : a == 1 ? P.iy : P.ix);
            // corner arithmetic is good to ~1e-12 and the relative coordinates below to 2e-7: margins of 1e-5
            all_fast = all_fast && cmin >= 1e-5 && cmax < dim - 1.0 - 1e-5;   // strictly inside, both taps exist, everywhere
            some_out = some_out || cmax < -1e-5 || cmin > dim - 1.0 + 1e-5;   // c < 0 or c > dim - 1 everywhere
            if (a == 0) {
                const int zlo = __double2int_rd(fmin(fmax(cmin - 1e-6, -1e9), 1e9));
                const int zhi = __double2int_rd(fmin(fmax(cmax + 1e-6, -1e9), 1e9)) + 1;
                any = zhi >= 0 && zlo <= P.iz - 1;
                lo_k = min(max(zlo, 0), P.iz - 1);
                hi_k = min(max(zhi, 0), P.iz - 1);
            }
        }
        cls = some_out ? kStepOutside : all_fast ? kStepFast : kStepMixed;
        if (any && cls != kStepOutside) {
            atomicMin(&s_zmin, lo_k);
            atomicMax(&s_zmax, hi_k);
        } else {
            any = 0;
        }
    }
    __syncthreads();
    const int zmin = s_zmin, zmax = s_zmax;
    const int nseq = zmax >= zmin ? zmax - zmin + 1 : 0;
    const int zstart = dir >= 0 ? zmin : zmax;   // plane of sequence number s is zstart + dir * s
    if (tid < nsteps) {
        unsigned need = 0, rel = 0;
        if (any) {
            need = (unsigned)(dir >= 0 ? hi_k - zmin + 1 : zmax - lo_k + 1);
            rel = (unsigned)(dir >= 0 ? lo_k - zmin : zmax - hi_k);
        }
        StepInfo e;
        // relative 9.23 coordinates: z against a multiple of the ring size (only its low bits select the slot),
        // (y, x) against the box origin; two's-complement wrap-around is harmless, interior voxels lie in [0, 512)
        const double zorg = (double)((zmin == 0x7fffffff ? 0 : zmin) & ~mask);
        e.bz = (unsigned)__double2ll_rd((c0[0] - zorg) * 8388608.0);
        e.by = (unsigned)__double2ll_rd((c0[1] - (double)oy0) * 8388608.0);
        e.bx = (unsigned)__double2ll_rd((c0[2] - (double)ox0) * 8388608.0);
        e.ctl = need | (rel << 12) | ((unsigned)cls << 24);
        tab[tid] = e;
        StepCoarse f;
        f.z = (float)c0[0]; f.y = (float)c0[1]; f.x = (float)c0[2]; f.pad = 0.f;
        coarse[tid] = f;
    }
    __syncthreads();

    // ---- plane loads: one 2-D TMA box per input plane, in march order ----------------------------------------------
    auto load_plane = [&](int seq) {
        const int z = zstart + dir * seq;
        const unsigned slot = (unsigned)z & mask;
        mbar_arrive_expect_tx_s(full_s + 8u * slot, P.tma_bytes);
        tma_load_3d_s(ring_s + slot * slot_bytes, &tmap, ox0, oy0, z, full_s + 8u * slot);
    };
    if (kProducerWarp) {
        if (warp == kTiltConsumerWarps) {
            if (lane == 0) {
                for (int seq = 0; seq < nseq; ++seq) {
                    const unsigned slot = (unsigned)(zstart + dir * seq) & mask;
                    const int round = seq >> ring_log2;
                    // suspended by the hardware until the slot is handed back (a try_wait / nanosleep(300) spin was
                    // 75 M of the general case's 607 M issued warp instructions: profiles/r02_ncu_path.txt, source page)
                    if (round > 0) mbar_wait_suspend_s(empty_s + 8u * slot, (round - 1) & 1, 4000u);
                    load_plane(seq);
                }
            }
            return;
        }
    } else if (tid == 0) {
        for (int seq = 0; seq < min(nseq, mask + 1); ++seq) load_plane(seq);   // the rest is issued by the releasing warps
    }

    // ---- consumers: column offsets from the tile's first voxel, 9.23 fixed point (rounded once, never accumulated) ----
    unsigned lz32[NC], ly32[NC], lx32[NC];
    unsigned live = 0;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const int la = lane + 32 * (c % IA), lb = warp + 8 * (c / IA);
        int l1 = SWAP ? la : lb, l2 = SWAP ? lb : la;
        if (t0y + l1 < P.oy && t0x + l2 < P.ox) live |= 1u << c;
        l1 = min(l1, (int)e1);   // columns beyond the output grid shadow the last live one (computed, never stored)
        l2 = min(l2, (int)e2);
        lz32[c] = (unsigned)__double2ll_rn(((double)l1 * P.M[1] + (double)l2 * P.M[2]) * 8388608.0);
        ly32[c] = (unsigned)__double2ll_rn(((double)l1 * P.M[5] + (double)l2 * P.M[6]) * 8388608.0);
        lx32[c] = (unsigned)__double2ll_rn(((double)l1 * P.M[9] + (double)l2 * P.M[10]) * 8388608.0);
    }
    unsigned one_bits = 0x3f800000u;
    asm volatile("" : "+r"(one_bits));   // keep 1.0f in a register (see frac23)
    float *otile = ring + ((size_t)P.PB << ring_log2);   // SWAP only: 2 x LA x (LB + 1)

    const long long plane = (long long)P.oy * P.ox;
    float *pstep = P.out + (long long)t0z * plane;                         // output plane of the current step
    float *pcol = pstep + (long long)(t0y + warp) * P.ox + t0x + lane;      // !SWAP: this thread's column c = 0
    // SWAP: the drain of a step's transposed tile.  A thread always drains the same o2 column b and rows a, a + R,
    // a + 2R, ... (R = consumers / LB), so its shared offset and global pointer are set up once and only stepped.
    constexpr int kDrainRows = kTiltConsumers / LB, kDrainIters = LA / kDrainRows;
    const int drain_a = tid / LB, drain_b = tid % LB;
    const bool drain_full = tile_full;
    const bool drain_col_ok = t0x + drain_b < P.ox;
    float *pdrain = pstep + (long long)(t0y + drain_a) * P.ox + t0x + drain_b;
    const long long drain_stride = (long long)kDrainRows * P.ox;
    const float cval = P.cval;
    unsigned ready = 0, released = 0;             // planes [0, ready) have landed; planes [0, released) were handed back
    unsigned rslot = (unsigned)zstart & mask;      // slot of plane `ready`
    unsigned eslot = rslot;                        // slot of plane `released`

    for (int lz = 0; lz < nsteps; ++lz, pcol += plane, pstep += plane, pdrain += plane) {
        StepInfo e;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(e.bz), "=r"(e.by), "=r"(e.bx), "=r"(e.ctl) : "r"(tab_s + 16u * (unsigned)lz));
        const unsigned need = e.ctl & 0xfffu, rel = (e.ctl >> 12) & 0xfffu, cls = e.ctl >> 24;
#pragma unroll 1
        while (released < rel) {
            if (lane == 0) {
                if (kProducerWarp) {
                    mbar_arrive_s(empty_s + 8u * eslot);
                } else {
                    // the warp's tap loads of this plane have returned (their values were consumed and stored in
                    // earlier steps), so a relaxed counter is enough
                    const unsigned prior = atomicAdd(&released_by[eslot], 1u);
                    if (prior == kTiltConsumerWarps - 1) {          // every warp is done with this plane
                        released_by[eslot] = 0;
                        const int next = (int)released + mask + 1;   // the plane that takes over the slot
                        if (next < nseq) load_plane(next);   // reads-then-async-write needs no proxy fence (as in any TMA ring)
                    }
                }
            }
            eslot = (eslot + (unsigned)dir) & mask;
            ++released;
        }
#pragma unroll 1
        while (ready < need) {
            mbar_wait_s(full_s + 8u * rslot, (ready >> ring_log2) & 1u);
            rslot = (rslot + (unsigned)dir) & mask;
            ++ready;
        }

        float res[NC];
        if (cls == kStepFast) {
            // every voxel of the tile is interior at this step: no tests, packed arithmetic
#pragma unroll
            for (int j = 0; j < NC2; ++j) {
                unsigned offA[2], offB[2];
                float wz[2], wy[2], wx[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int c = 2 * j + h;
                    const unsigned cz = e.bz + lz32[c], cy = e.by + ly32[c], cx = e.bx + lx32[c];
                    wz[h] = frac23(cz, one_bits);
                    wy[h] = frac23(cy, one_bits);
                    wx[h] = frac23(cx, one_bits);
                    const unsigned inpl = (cy >> 23) * pitch4 + (ring_s + ((cx >> 23) << 2));   // shared address of the first tap in slot 0
                    const unsigned sa = (cz >> 23) & mask, sb = (sa + 1u) & mask;
                    offA[h] = sa * slot_bytes + inpl;
                    offB[h] = sb * slot_bytes + inpl;
                }
                const float2 wx2 = make_float2(wx[0], wx[1]), wy2 = make_float2(wy[0], wy[1]), wz2 = make_float2(wz[0], wz[1]);
                const float2 a00 = make_float2(lds_f32<0>(offA[0]), lds_f32<0>(offA[1]));
                const float2 a01 = make_float2(lds_f32<4>(offA[0]), lds_f32<4>(offA[1]));
                const float2 a10 = make_float2(lds_f32<0>(offA[0] + pitch4), lds_f32<0>(offA[1] + pitch4));
                const float2 a11 = make_float2(lds_f32<4>(offA[0] + pitch4), lds_f32<4>(offA[1] + pitch4));
                const float2 b00 = make_float2(lds_f32<0>(offB[0]), lds_f32<0>(offB[1]));
                const float2 b01 = make_float2(lds_f32<4>(offB[0]), lds_f32<4>(offB[1]));
                const float2 b10 = make_float2(lds_f32<0>(offB[0] + pitch4), lds_f32<0>(offB[1] + pitch4));
                const float2 b11 = make_float2(lds_f32<4>(offB[0] + pitch4), lds_f32<4>(offB[1] + pitch4));
                const float2 va = lerp2(wy2, lerp2(wx2, a00, a01), lerp2(wx2, a10, a11));
                const float2 vb = lerp2(wy2, lerp2(wx2, b00, b01), lerp2(wx2, b10, b11));
                const float2 r = lerp2(wz2, va, vb);
                res[2 * j] = r.x;
                res[2 * j + 1] = r.y;
            }
            if (CLEAN) {
                float2 acc = make_float2(0.f, 0.f);   // stays 0 unless some result is non-finite
#pragma unroll
                for (int j = 0; j < NC2; ++j) acc = __ffma2_rn(make_float2(res[2 * j], res[2 * j + 1]), make_float2(0.f, 0.f), acc);
                if (!(acc.x == 0.f && acc.y == 0.f)) {
#pragma unroll
                    for (int c = 0; c < NC; ++c)
                        if (nonfinite(res[c]))
                            res[c] = tilt_exact_voxel(ring, &P, t0z + lz, t0y + (SWAP ? lane + 32 * (c % IA) : warp + 8 * (c / IA)),
                                                      t0x + (SWAP ? warp + 8 * (c / IA) : lane + 32 * (c % IA)), oy0, ox0, 1);
                }
            }
        } else if (cls == kStepOutside) {
#pragma unroll
            for (int c = 0; c < NC; ++c) res[c] = cval;
        } else {
            // The tile straddles the rim of the input (or of the output grid) at this step.  Each voxel is classified
            // with its coordinate in float32 (good to ~2e-3 for coordinates below 2^13, margin 0.05): safely interior ->
            // the same 9.23 arithmetic as a fast step (scalar); safely outside -> cval; within the margin of the rim ->
            // tilt_exact_voxel (scipy's exact float64 arithmetic and edge rule from the output index).
            const StepCoarse f = coarse[lz];
            const float dz = (float)P.iz - 1.0f, dy = (float)P.iy - 1.0f, dx = (float)P.ix - 1.0f;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                float r = cval;
                if (live >> c & 1u) {
                    const float fz = f.z + (float)(int)lz32[c] * 1.1920928955078125e-7f;
                    const float fy = f.y + (float)(int)ly32[c] * 1.1920928955078125e-7f;
                    const float fx = f.x + (float)(int)lx32[c] * 1.1920928955078125e-7f;
                    const float m = 0.05f + 2e-6f * fmaxf(dz, fmaxf(dy, dx));   // float32 coordinate error grows with the size
                    const bool inner = fz >= m && fz <= dz - m && fy >= m && fy <= dy - m && fx >= m && fx <= dx - m;
                    const bool outer = fz < -m || fz > dz + m || fy < -m || fy > dy + m || fx < -m || fx > dx + m;
                    if (inner) {
                        const unsigned cz = e.bz + lz32[c], cy = e.by + ly32[c], cx = e.bx + lx32[c];
                        const float wz = frac23(cz, one_bits), wy = frac23(cy, one_bits), wx = frac23(cx, one_bits);
                        const unsigned inpl = (cy >> 23) * pitch4 + (ring_s + ((cx >> 23) << 2));
                        const unsigned sa = (cz >> 23) & mask, sb = (sa + 1u) & mask;
                        const unsigned oa = sa * slot_bytes + inpl, ob = sb * slot_bytes + inpl;
                        const float a00 = lds_f32<0>(oa), a01 = lds_f32<4>(oa), a10 = lds_f32<0>(oa + pitch4), a11 = lds_f32<4>(oa + pitch4);
                        const float b00 = lds_f32<0>(ob), b01 = lds_f32<4>(ob), b10 = lds_f32<0>(ob + pitch4), b11 = lds_f32<4>(ob + pitch4);
                        const float xa0 = fmaf(wx, a01 - a00, a00), xa1 = fmaf(wx, a11 - a10, a10);
                        const float xb0 = fmaf(wx, b01 - b00, b00), xb1 = fmaf(wx, b11 - b10, b10);
                        const float va = fmaf(wy, xa1 - xa0, xa0), vb = fmaf(wy, xb1 - xb0, xb0);
                        r = fmaf(wz, vb - va, va);
                    }
                    if ((!inner && !outer) || (CLEAN && inner && nonfinite(r)))
                        r = tilt_exact_voxel(ring, &P, t0z + lz, t0y + (SWAP ? lane + 32 * (c % IA) : warp + 8 * (c / IA)),
                                             t0x + (SWAP ? warp + 8 * (c / IA) : lane + 32 * (c % IA)), oy0, ox0, CLEAN);
                }
                res[c] = r;
            }
        }
        if (SWAP) {
            float *ot = otile + (lz & 1) * (LA * (LB + 1));
#pragma unroll
            for (int c = 0; c < NC; ++c) ot[(lane + 32 * (c % IA)) * (LB + 1) + warp + 8 * (c / IA)] = res[c];
            tilt_consumer_sync();   // one barrier per step: the other buffer is written while this one drains
            {
                const float *src = ot + drain_a * (LB + 1) + drain_b;
                float *dst = pdrain;
#pragma unroll
                for (int i = 0; i < kDrainIters; ++i) {
                    if (drain_full || (drain_col_ok && t0y + drain_a + i * kDrainRows < P.oy)) __stcs(dst, src[i * kDrainRows * (LB + 1)]);
                    dst += drain_stride;
                }
            }
        } else if (tile_full) {
#pragma unroll
            for (int rb = 0; rb < RB; ++rb) {
                float *prow = pcol + (long long)(8 * rb) * P.ox;
#pragma unroll
                for (int ia = 0; ia < IA; ++ia) __stcs(prow + 32 * ia, res[rb * IA + ia]);
            }
        } else {
#pragma unroll
            for (int c = 0; c < NC; ++c)
                if (live >> c & 1u) __stcs(pcol + (long long)(8 * (c / IA)) * P.ox + 32 * (c % IA), res[c]);
        }
    }
}

template <int IA, int RB>
static void (*pick_tilt(bool swap, bool clean))(const CUtensorMap, const AffineParams, const TiltParams) {
    return swap ? (clean ? affine_tilt_kernel<IA, RB, true, true> : affine_tilt_kernel<IA, RB, true, false>)
                : (clean ? affine_tilt_kernel<IA, RB, false, true> : affine_tilt_kernel<IA, RB, false, false>);
}

// Host side: tile, ring depth and march length per launch; per-plane tensor map; one launch per z chunk.
// Returns SHRIMPY_OK with *launched = false when the matrix/shape is not eligible (the caller falls back).
int launch_affine_tilt(AffineParams P, int nan_to_zero, cudaStream_t s, bool *launched) {
    *launched = false;
    const double *M = P.M;
    if ((reinterpret_cast<uintptr_t>(P.in) & 15u) != 0 || P.in_sy % 4 != 0 || P.in_sz % 4 != 0 || (long long)P.oy * P.ox >= 2147483647LL ||
        tensor_map_encoder() == nullptr)
        return SHRIMPY_OK;
    const bool swap = std::fabs(M[9]) > std::fabs(M[10]);   // input x follows o1 more than o2: lanes along o1
    if (!(std::fabs(M[0]) <= 16.0)) return SHRIMPY_OK;   // plane counts of a march are kept in 12 bits

    // (IA, RB): !swap: tile = (8 RB) x (32 IA) in (o1, o2); swap: (32 IA) x (8 RB), RB >= 4 so that drained rows are >= 128 bytes
    static const int cand_plain[][2] = {{2, 2}, {4, 1}, {2, 4}, {4, 2}, {2, 1}};
    static const int cand_swap[][2] = {{1, 4}, {1, 8}, {2, 4}, {0, 0}, {0, 0}};
    int forced[4] = {0, 0, 0, 0};   // IA, RB, ring, march length
    if (const char *f = getenv("SHRIMPY_TILT_CFG")) sscanf(f, "%d,%d,%d,%d", &forced[0], &forced[1], &forced[2], &forced[3]);
    double best = 1e300;
    int bIA = 0, bRB = 0;
    for (int k = 0; k < 5; ++k) {
        const int IA = swap ? cand_swap[k][0] : cand_plain[k][0], RB = swap ? cand_swap[k][1] : cand_plain[k][1];
        if (IA == 0 || (forced[0] && (IA != forced[0] || RB != forced[1]))) continue;
        const int TY = swap ? 32 * IA : 8 * RB, TX = swap ? 8 * RB : 32 * IA;
        const long long otile = swap ? 2LL * (32 * IA) * (8 * RB + 1) : 0;
        const double zspread = std::fabs(M[1]) * (TY - 1) + std::fabs(M[2]) * (TX - 1);
        const int window = (int)std::ceil(zspread) + 3;   // [floor(min) , floor(max) + 1] plus the 1e-6 margins
        const long long budget = 108 * 1024;   // two CTAs per SM at least; three when the ring is small enough
        for (int zc : {128, 64, 32, 16}) {
            if (forced[3] && zc != forced[3]) continue;
            const int ZC = std::min(zc, std::max(P.oz, 1));
            const long long BY = (long long)std::ceil(std::fabs(M[4]) * (ZC - 1) + std::fabs(M[5]) * (TY - 1) + std::fabs(M[6]) * (TX - 1)) + 4;
            long long BX = (long long)std::ceil(std::fabs(M[8]) * (ZC - 1) + std::fabs(M[9]) * (TY - 1) + std::fabs(M[10]) * (TX - 1)) + 4;
            BX = (BX + 3 + 3) / 4 * 4;
            if (BY > 256 || BX > 256) continue;
            const long long PB = (BY * BX + 31) / 32 * 32;
            for (int rl = 4; rl >= 2; --rl) {
                if (forced[2] && (1 << rl) != forced[2]) continue;
                if ((1 << rl) < window + 2) break;   // at least two planes of read-ahead
                const long long bytes = ((PB << rl) + otile) * 4 + 128;
                if (bytes > budget) continue;
                // cost model (measured on config 3): the kernel is issue-bound, so the staged volume per output
                // voxel matters little; every extra launch refills the plane window; 4 columns per thread
                // (3 CTAs per SM) beat 8; a ring without slack stalls the producer
                const int nch = (P.oz + ZC - 1) / ZC;
                double cost = 1.0 + 0.15 * (double)(BY * BX) / (TY * TX) + 0.1 * (nch - 1) * (1.0 + (double)window / ZC);
                cost += ((1 << rl) < window + 4 ? 0.2 : 0.0) + (IA * RB == 8 ? 0.3 : 0.0) + (IA * RB == 2 ? 0.15 : 0.0);
                cost += (IA * RB <= 4 && bytes > 72 * 1024) ? 0.2 : 0.0;   // only two of the three possible CTAs fit
                if (cost < best) {
                    best = cost; bIA = IA; bRB = RB;
                    P.ring_log2 = rl; P.ZC = ZC;
                    P.BY = (int)BY; P.BX = (int)BX; P.pitch = (int)BX; P.PB = (int)PB; P.BZ = 1;
                    P.TY = TY; P.TX = TX; P.LA = 32 * IA; P.LB = 8 * RB;
                }
                break;
            }
        }
    }
    if (best == 1e300) return SHRIMPY_OK;
    const int nchunks = (P.oz + P.ZC - 1) / P.ZC;
    P.ZC = (P.oz + nchunks - 1) / nchunks;   // even chunks
    P.tiles_x = (P.ox + P.TX - 1) / P.TX;
    P.tiles_y = (P.oy + P.TY - 1) / P.TY;
    P.tiles_z = nchunks;
    if (P.tiles_y > 65535) return SHRIMPY_OK;

    CUtensorMap tmap{};
    const cuuint64_t gdim[3] = {(cuuint64_t)P.ix, (cuuint64_t)P.iy, (cuuint64_t)P.iz};
    const cuuint64_t gstride[2] = {(cuuint64_t)P.in_sy * 4, (cuuint64_t)P.in_sz * 4};   // rows beyond ix are zero-filled by TMA
    const cuuint32_t bdim[3] = {(cuuint32_t)P.BX, (cuuint32_t)P.BY, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    P.tma_bytes = bdim[0] * bdim[1] * 4u;
    const CUresult rc = tensor_map_encoder()(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(P.in), gdim,
                                             gstride, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return fail(SHRIMPY_ECUDA, "affine tilt: cuTensorMapEncodeTiled failed (%d)", (int)rc);

    void (*kern)(const CUtensorMap, const AffineParams, const TiltParams) = nullptr;
    const bool cl = nan_to_zero != 0;
    if (bIA == 2 && bRB == 2) kern = pick_tilt<2, 2>(swap, cl);
    else if (bIA == 4 && bRB == 1) kern = pick_tilt<4, 1>(swap, cl);
    else if (bIA == 2 && bRB == 4) kern = pick_tilt<2, 4>(swap, cl);
    else if (bIA == 4 && bRB == 2) kern = pick_tilt<4, 2>(swap, cl);
    else if (bIA == 2 && bRB == 1) kern = pick_tilt<2, 1>(swap, cl);
    else if (bIA == 1 && bRB == 4) kern = pick_tilt<1, 4>(swap, cl);
    else if (bIA == 1 && bRB == 8) kern = pick_tilt<1, 8>(swap, cl);
    else return SHRIMPY_OK;
    const size_t smem = (((size_t)P.PB << P.ring_log2) + (swap ? 2 * (size_t)P.LA * (P.LB + 1) : 0)) * sizeof(float) + 128;
    if (getenv("SHRIMPY_DEBUG"))
        fprintf(stderr, "[shrimpy] affine tilt IA=%d RB=%d swap=%d ring=%d box=(%d,%d) PB=%d ZC=%d smem=%zu grid=(%d,%d) x %d launches\n", bIA,
                bRB, (int)swap, 1 << P.ring_log2, P.BY, P.BX, P.PB, P.ZC, smem, P.tiles_x, P.tiles_y, nchunks);
    if (smem + 4096 > 48 * 1024)
        SHRIMPY_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TiltParams Q{};
    Q.dir = M[0] < 0.0 ? -1 : 1;
    for (int ch = 0; ch < nchunks; ++ch) {
        Q.t0z = ch * P.ZC;
        Q.nsteps = std::min(P.ZC, P.oz - Q.t0z);
        kern<<<dim3((unsigned)P.tiles_x, (unsigned)P.tiles_y), tilt_threads(swap), smem, s>>>(tmap, P, Q);
        count_launch();
    }
    SHRIMPY_CUDA_TRY(cudaGetLastError());
    *launched = true;
    return SHRIMPY_OK;
}

}  // namespace shrimpy
