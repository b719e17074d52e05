// Z-streaming planar affine resample for sm_100a (registration step, BASELINE.json configs[2]).
//
// Block-diagonal matrices (z <-> z, (y,x) <-> (y,x): an in-plane 2-D affine plus a z shift/scale -- the mantis
// label-free -> fluorescence registration family, and every near-identity map without tilt).
//
// One CTA owns an output tile in (o1, o2) and MARCHES through o0.  The input planes the march needs form a
// monotone sequence z_first, z_first +- 1, ...; a producer warp streams them through a ring of shared-memory
// slots with one 2-D TMA box per plane (full/empty mbarrier pair per slot), so loads run several planes ahead
// of the arithmetic and nothing is ever re-read along z.  Eight consumer warps keep, per thread, NC output
// columns in registers: the exact float64 (y, x) coordinate, floor and weights of a column are computed ONCE
// for the whole march, the bilinear value of a column in an input plane is computed once per plane and shared
// by the output steps that touch it (two register sets addressed by the parity of the plane's sequence number,
// so nothing is ever copied), and the arithmetic is issued as packed float32 pairs (FFMA2 / FADD2: half the
// issue slots, same rounding as the scalar sequence).  Per output step only the z lerp and the store remain.
//
// The z coordinate of a step does not depend on (o1, o2): the host tabulates floor, weight, inside test and
// plane sequence numbers of every step with scipy's float64 arithmetic and passes the table as a kernel
// parameter, so the per-step bookkeeping (ring waits, reuse tags, branches) runs on the uniform datapath.
//
// Lanes run along the output axis that walks input x (o2, or o1 for ~90 degree maps: SWAP), so tap reads are
// conflict-free; with SWAP a step's plane of results is transposed through a double-buffered shared tile so
// global stores stay coalesced along o2.
//
// Exactness: every column's (y, x) split uses the exact edge rule (c == dim-1 is inside, its second tap folds
// onto the first); columns at the rim simply carry zero tap strides.  Non-finite results (nan_to_num semantics)
// are recomputed from cleaned taps by planar_exact_voxel.  The float32 lerp sequence (x, then y, then z) is the
// one every kernel of this library uses.
#include "affine_common.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace shrimpy {

constexpr int kStreamConsumerWarps = 8;
constexpr int kStreamConsumers = 32 * kStreamConsumerWarps;
constexpr int kStreamThreads = kStreamConsumers + 32;   // + one producer warp
constexpr int kStreamMaxSteps = 128;
constexpr int kStreamMaxRing = 8;

struct ZStep {
    int sA, sB;   // sequence numbers (load order) of the planes holding floor(z) and floor(z)+1
    float wz;
    int inside;
};

struct ZTable {
    ZStep e[kStreamMaxSteps];
    int t0z, nsteps;   // output steps [t0z, t0z + nsteps) of this launch
    int zfirst, zdir;  // plane of sequence number s is zfirst + s * zdir
    int nseq;
};

__device__ __forceinline__ void consumer_sync() {
    asm volatile("bar.sync 1, %0;" ::"n"(kStreamConsumers) : "memory");
}

// Exact voxel of a planar map given its two (already exact) z planes: scipy's (y, x) coordinate arithmetic, the
// exact edge rule, nan_to_num of the taps.  Only voxels whose fast result is non-finite come here.
__device__ __noinline__ float planar_exact_voxel(const float *pa, const float *pb, float wz, const AffineParams *Pp,
                                                 int o1, int o2, int oy0, int ox0, int clean_taps) {
    const AffineParams &P = *Pp;
    const double cy = __dadd_rn(__dadd_rn(P.M[7], __dmul_rn((double)o1, P.M[5])), __dmul_rn((double)o2, P.M[6]));
    const double cx = __dadd_rn(__dadd_rn(P.M[11], __dmul_rn((double)o1, P.M[9])), __dmul_rn((double)o2, P.M[10]));
    int y0, x0;
    float wy, wx;
    bool in = split_coord(cy, P.iy, y0, wy);
    in &= split_coord(cx, P.ix, x0, wx);
    if (!in) return P.cval;
    const int dy = (y0 + 1 < P.iy) ? P.pitch : 0;
    const int dx = (x0 + 1 < P.ix) ? 1 : 0;
    const int q = (y0 - oy0) * P.pitch + (x0 - ox0);
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

// IA = items along the lane axis (tile extent 32*IA), RB = rows per warp (tile extent 8*RB along the other axis).
template <int IA, int RB, bool SWAP, bool CLEAN>
__global__ void __launch_bounds__(kStreamThreads, (IA * RB <= 4) ? 3 : 2)
    affine_stream_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ AffineParams P,
                         const __grid_constant__ ZTable T) {
    constexpr int NC = IA * RB, NC2 = NC / 2;
    static_assert(NC % 2 == 0 && NC <= 16, "columns are processed in packed pairs");
    constexpr int LA = 32 * IA, LB = 8 * RB;
    constexpr int TY = SWAP ? LA : LB, TX = SWAP ? LB : LA;
    constexpr unsigned ALL = (1u << NC) - 1u;

    extern __shared__ __align__(128) float smem_raw[];
    __shared__ __align__(8) uint64_t full[kStreamMaxRing], empty[kStreamMaxRing];
    float *ring = smem_raw + (((128u - (smem_u32(smem_raw) & 127u)) & 127u) >> 2);
    const uint32_t full_s = smem_u32(full), empty_s = smem_u32(empty);

    const int ring_log2 = P.ring_log2, mask = (1 << ring_log2) - 1;
    const int pitch = P.pitch;
    const unsigned slot_bytes = (unsigned)P.PB * 4u;
    const int t0y = blockIdx.y * TY, t0x = blockIdx.x * TX, t0z = T.t0z;
    const int nsteps = T.nsteps;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i <= mask; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], kStreamConsumerWarps);
        }
        fence_mbar_init();
    }
    // (y, x) origin of the staged box: minimum corner of the tile under the in-plane affine
    int org[2];
    {
        const double e1 = (double)(min(t0y + TY, P.oy) - 1 - t0y), e2 = (double)(min(t0x + TX, P.ox) - 1 - t0x);
#pragma unroll
        for (int a = 1; a < 3; ++a) {
            const double m1 = P.M[4 * a + 1], m2 = P.M[4 * a + 2];
            const double lo = P.M[4 * a + 3] + t0y * m1 + t0x * m2 + fmin(e1 * m1, 0.0) + fmin(e2 * m2, 0.0);
            const int dim = a == 1 ? P.iy : P.ix;
            org[a - 1] = __double2int_rd(fmin(fmax(lo - 1e-6, -4.0), (double)dim));
        }
        org[1] &= ~3;   // TMA: the box starts on a 16-byte boundary along x
    }
    const int oy0 = org[0], ox0 = org[1];
    __syncthreads();

    // ---- producer warp: one TMA box per input plane, in march order ----------------------------------------------
    if (warp == kStreamConsumerWarps) {
        if (lane == 0) {
            const uint32_t ring_s = smem_u32(ring);   // (the consumers keep their own copy)
            for (int seq = 0; seq < T.nseq; ++seq) {
                const int slot = seq & mask, round = seq >> ring_log2;
                if (round > 0) mbar_wait_suspend_s(empty_s + 8u * slot, (round - 1) & 1, 2000u);
                mbar_arrive_expect_tx_s(full_s + 8u * slot, P.tma_bytes);
                tma_load_3d_s(ring_s + slot * slot_bytes, &tmap, ox0, oy0, T.zfirst + seq * T.zdir, full_s + 8u * slot);
            }
        }
        return;
    }

    // ---- consumers: per-column set-up (exact (y, x) coordinate, floor, weights, classification) -------------------
    unsigned qoff[NC];                   // byte offset of the column's first tap inside a plane slot
    float2 wx2[NC2], wy2[NC2];
    unsigned live = 0, valid = 0, xedge = 0, yedge = 0;   // bit c: column exists / is inside / has no x+1 / no y+1 tap
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const int a = lane + 32 * (c % IA), b = warp + 8 * (c / IA);
        const int o1 = t0y + (SWAP ? a : b), o2 = t0x + (SWAP ? b : a);
        float wy = 0.f, wx = 0.f;
        qoff[c] = 0;
        if (o1 < P.oy && o2 < P.ox) {
            live |= 1u << c;
            // + o0 * 0 contributes exactly +-0, so this is scipy's value for every o0
            const double cy = __dadd_rn(__dadd_rn(P.M[7], __dmul_rn((double)o1, P.M[5])), __dmul_rn((double)o2, P.M[6]));
            const double cx = __dadd_rn(__dadd_rn(P.M[11], __dmul_rn((double)o1, P.M[9])), __dmul_rn((double)o2, P.M[10]));
            int y0, x0;
            bool in = split_coord(cy, P.iy, y0, wy);
            in &= split_coord(cx, P.ix, x0, wx);
            if (in) {
                valid |= 1u << c;
                if (x0 + 1 >= P.ix) xedge |= 1u << c;
                if (y0 + 1 >= P.iy) yedge |= 1u << c;
                qoff[c] = (unsigned)((y0 - oy0) * pitch + (x0 - ox0)) * 4u;
            }
        }
        if (c & 1) { wy2[c / 2].y = wy; wx2[c / 2].y = wx; }
        else       { wy2[c / 2].x = wy; wx2[c / 2].x = wx; }
    }
    // lean warp: every column exists, is inside and owns all four taps -> constant tap strides, no predicates
    const bool lean = __all_sync(0xffffffffu, valid == ALL && xedge == 0 && yedge == 0);

    const uint32_t ring_s = smem_u32(ring);
    const unsigned pitch4 = (unsigned)pitch * 4u;
    // bilinear value of every column of this thread in the plane with sequence number seq.  Taps are read through
    // 32-bit shared addresses "uniform plane base + per-column offset (+4)", which fold into the LDS address mode.
    auto fill = [&](int seq, float2(&dst)[NC2]) {
        const uint32_t pl = ring_s + (unsigned)(seq & mask) * slot_bytes;
        if (lean) {
            const uint32_t pl1 = pl + pitch4;
#pragma unroll
            for (int j = 0; j < NC2; ++j) {
                const unsigned qa = qoff[2 * j], qb = qoff[2 * j + 1];
                const float2 lo = make_float2(lds_f32<0>(pl + qa), lds_f32<0>(pl + qb));
                const float2 hi = make_float2(lds_f32<4>(pl + qa), lds_f32<4>(pl + qb));
                const float2 lo1 = make_float2(lds_f32<0>(pl1 + qa), lds_f32<0>(pl1 + qb));
                const float2 hi1 = make_float2(lds_f32<4>(pl1 + qa), lds_f32<4>(pl1 + qb));
                dst[j] = lerp2(wy2[j], lerp2(wx2[j], lo, hi), lerp2(wx2[j], lo1, hi1));
            }
        } else {
#pragma unroll
            for (int j = 0; j < NC2; ++j) {
                float v[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int c = 2 * j + h;
                    const uint32_t q = pl + qoff[c];
                    const unsigned dx = (xedge >> c & 1u) ? 0u : 4u, dy = (yedge >> c & 1u) ? 0u : pitch4;
                    const float v00 = lds_f32<0>(q), v01 = lds_f32<0>(q + dx);
                    const float v10 = lds_f32<0>(q + dy), v11 = lds_f32<0>(q + dy + dx);
                    const float wx = h ? wx2[j].y : wx2[j].x, wy = h ? wy2[j].y : wy2[j].x;
                    const float a0 = fmaf(wx, v01 - v00, v00), a1 = fmaf(wx, v11 - v10, v10);
                    v[h] = fmaf(wy, a1 - a0, a0);
                }
                dst[j] = make_float2(v[0], v[1]);
            }
        }
    };

    float *otile = ring + ((size_t)P.PB << ring_log2);   // SWAP only: 2 x LA x (LB + 1)
    const long long plane = (long long)P.oy * P.ox;
    float *pstep = P.out + (long long)t0z * plane;        // start of output plane t0z (advanced per step)
    float *pcol = pstep + (long long)(t0y + warp) * P.ox + t0x + lane;   // !SWAP: this thread's column c = 0
    // SWAP: the drain of a step's transposed tile.  A thread always drains the same o2 column b and rows a, a + R,
    // a + 2R, ... (R = consumers / LB), so its shared offset and global pointer are set up once and only stepped.
    constexpr int kDrainRows = kStreamConsumers / LB, kDrainIters = LA / kDrainRows;
    const int drain_a = tid / LB, drain_b = tid % LB;
    const bool drain_full = t0y + TY <= P.oy && t0x + TX <= P.ox;
    const bool drain_col_ok = t0x + drain_b < P.ox;
    float *pdrain = pstep + (long long)(t0y + drain_a) * P.ox + t0x + drain_b;
    const long long drain_stride = (long long)kDrainRows * P.ox;

    float2 V0[NC2], V1[NC2];       // plane values: V0 <- planes with even sequence number, V1 <- odd
#pragma unroll
    for (int j = 0; j < NC2; ++j) V0[j] = V1[j] = make_float2(0.f, 0.f);
    int tag0 = -1, tag1 = -1;      // sequence numbers held in V0 / V1
    int ready = 0, released = 0;   // planes [0, ready) have landed; planes [0, released) were handed back
    const float cval = P.cval;
    int outside_steps = 0;

    for (int lz = 0; lz < nsteps; ++lz, pstep += plane, pcol += plane, pdrain += plane) {
        const ZStep e = T.e[lz];
        if (!e.inside) {   // written by the constant-fill pass below
            ++outside_steps;
            if (SWAP) consumer_sync();   // keep the double-buffer cadence of the shared tile
            continue;
        }
        const int lo = min(e.sA, e.sB), hi = max(e.sA, e.sB);
#pragma unroll 1
        while (released < lo) {
            if (lane == 0) mbar_arrive_s(empty_s + 8u * (released & mask));
            ++released;
        }
#pragma unroll 1
        while (ready <= hi) {
            mbar_wait_s(full_s + 8u * (ready & mask), (ready >> ring_log2) & 1);
            ++ready;
        }
        // lo and hi (= lo or lo + 1) have different parities: each lives in its own register set
        const int s_even = (lo & 1) ? hi : lo, s_odd = (lo & 1) ? lo : hi;
        if (!(s_even & 1) && tag0 != s_even) {
            fill(s_even, V0);
            tag0 = s_even;
        }
        if ((s_odd & 1) && tag1 != s_odd) {
            fill(s_odd, V1);
            tag1 = s_odd;
        }
        float res[NC];
        {
            float2 r[NC2];
            const float2 wz2 = make_float2(e.wz, e.wz);
            if (e.sA == e.sB) {
#pragma unroll
                for (int j = 0; j < NC2; ++j) r[j] = (e.sA & 1) ? V1[j] : V0[j];
            } else if (e.sA & 1) {
#pragma unroll
                for (int j = 0; j < NC2; ++j) r[j] = lerp2(wz2, V1[j], V0[j]);
            } else {
#pragma unroll
                for (int j = 0; j < NC2; ++j) r[j] = lerp2(wz2, V0[j], V1[j]);
            }
#pragma unroll
            for (int j = 0; j < NC2; ++j) {
                res[2 * j] = r[j].x;
                res[2 * j + 1] = r[j].y;
            }
        }
        unsigned rare = 0;
        if (CLEAN) {
            float2 acc = make_float2(0.f, 0.f);   // stays 0 unless some result is non-finite
#pragma unroll
            for (int j = 0; j < NC2; ++j) acc = __ffma2_rn(make_float2(res[2 * j], res[2 * j + 1]), make_float2(0.f, 0.f), acc);
            if (!(acc.x == 0.f && acc.y == 0.f)) {
#pragma unroll
                for (int c = 0; c < NC; ++c)
                    if (nonfinite(res[c])) rare |= 1u << c;
                rare &= valid;
            }
        }
        if (!lean) {
#pragma unroll
            for (int c = 0; c < NC; ++c) res[c] = (valid >> c & 1u) ? res[c] : cval;
        }
        if (rare) {   // non-finite taps: exact recomputation from cleaned taps (rare)
            const float *pa = ring + (size_t)(e.sA & mask) * P.PB, *pb = ring + (size_t)(e.sB & mask) * P.PB;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                if (rare >> c & 1u) {
                    const int a = lane + 32 * (c % IA), b = warp + 8 * (c / IA);
                    res[c] = planar_exact_voxel(pa, pb, e.wz, &P, t0y + (SWAP ? a : b), t0x + (SWAP ? b : a), oy0, ox0, 1);
                }
            }
        }

        if (!SWAP) {
            if (lean) {
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
        } else {
            float *ot = otile + (lz & 1) * (LA * (LB + 1));
#pragma unroll
            for (int c = 0; c < NC; ++c) ot[(lane + 32 * (c % IA)) * (LB + 1) + warp + 8 * (c / IA)] = res[c];
            consumer_sync();   // one barrier per step: the other buffer is written while this one drains
            {
                const float *src = ot + drain_a * (LB + 1) + drain_b;
                float *dst = pdrain;
#pragma unroll
                for (int i = 0; i < kDrainIters; ++i) {
                    if (drain_full || (drain_col_ok && t0y + drain_a + i * kDrainRows < P.oy)) __stcs(dst, src[i * kDrainRows * (LB + 1)]);
                    dst += drain_stride;
                }
            }
        }
    }

    // ---- steps whose z lies outside the input: constant fill, rows along o2 -----------------------------------------
    if (outside_steps) {
        for (int lz = 0; lz < nsteps; ++lz) {
            if (T.e[lz].inside) continue;
            float *pz = P.out + (long long)(t0z + lz) * plane;
            for (int idx = tid; idx < TY * TX; idx += kStreamConsumers) {
                const int o1 = t0y + idx / TX, o2 = t0x + idx % TX;
                if (o1 < P.oy && o2 < P.ox) __stcs(pz + (long long)o1 * P.ox + o2, cval);
            }
        }
    }
}

template <int IA, int RB>
static void (*pick_stream(bool swap, bool clean))(const CUtensorMap, const AffineParams, const ZTable) {
    return swap ? (clean ? affine_stream_kernel<IA, RB, true, true> : affine_stream_kernel<IA, RB, true, false>)
                : (clean ? affine_stream_kernel<IA, RB, false, true> : affine_stream_kernel<IA, RB, false, false>);
}

// z table of output steps [t0z, t0z + nsteps): scipy's coordinate ((M03 + o0 M00) + o1*0) + o2*0 in float64,
// exact floor / weight / edge rule (the host twin of split_coord), planes numbered in march order.
static void build_ztable(const AffineParams &P, int t0z, int nsteps, ZTable &T) {
    int z0[kStreamMaxSteps], z1[kStreamMaxSteps];
    int zmin = 0x7fffffff, zmax = -1;
    T.t0z = t0z;
    T.nsteps = nsteps;
    for (int k = 0; k < nsteps; ++k) {
        volatile double prod = (double)(t0z + k) * P.M[0];   // product and sum rounded separately
        const double cz = P.M[3] + prod;
        ZStep &e = T.e[k];
        e.sA = e.sB = 0;
        e.wz = 0.f;
        e.inside = 0;
        if (!(cz >= 0.0) || !(cz <= (double)(P.iz - 1))) continue;   // scipy: strict outside test, no tolerance
        const int f = (int)std::floor(cz);
        const float w = (float)(cz - (double)f);
        if (!(f < P.iz - 1 || (f == P.iz - 1 && w == 0.f && cz == (double)f))) continue;
        z0[k] = f;
        z1[k] = std::min(f + 1, P.iz - 1);
        e.wz = w;
        e.inside = 1;
        zmin = std::min(zmin, z0[k]);
        zmax = std::max(zmax, z1[k]);
    }
    const bool desc = P.M[0] < 0.0;
    T.zdir = desc ? -1 : 1;
    T.nseq = zmax >= zmin ? zmax - zmin + 1 : 0;
    T.zfirst = T.nseq ? (desc ? zmax : zmin) : 0;
    for (int k = 0; k < nsteps; ++k) {
        if (!T.e[k].inside) continue;
        T.e[k].sA = desc ? zmax - z0[k] : z0[k] - zmin;
        T.e[k].sB = desc ? zmax - z1[k] : z1[k] - zmin;
    }
}

// Host side: pick the tile and the ring depth, encode the per-plane tensor map, launch one grid per chunk of
// <= 128 output steps.  Returns SHRIMPY_OK with *launched = false when the matrix/shape is not eligible.
int launch_affine_stream(AffineParams P, int nan_to_zero, cudaStream_t s, bool *launched) {
    *launched = false;
    const double *M = P.M;
    if (!(M[1] == 0.0 && M[2] == 0.0 && M[4] == 0.0 && M[8] == 0.0)) return SHRIMPY_OK;
    if ((reinterpret_cast<uintptr_t>(P.in) & 15u) != 0 || P.in_sy % 4 != 0 || P.in_sz % 4 != 0 || (long long)P.oy * P.ox >= 2147483647LL ||
        tensor_map_encoder() == nullptr)
        return SHRIMPY_OK;
    const bool swap = std::fabs(M[9]) > std::fabs(M[10]);   // input x follows o1 more than o2: lanes along o1
    // (IA, RB): lanes x rows-per-warp.  !swap: tile = (8 RB) x (32 IA) in (o1, o2); swap: (32 IA) x (8 RB), RB >= 4
    // so that drained rows are >= 128 bytes.
    static const int cand_plain[][2] = {{4, 2}, {2, 4}, {2, 2}, {4, 1}, {2, 1}};
    static const int cand_swap[][2] = {{2, 4}, {1, 8}, {1, 4}};
    int forced[3] = {0, 0, 0};
    if (const char *f = getenv("SHRIMPY_STREAM_CFG")) sscanf(f, "%d,%d,%d", &forced[0], &forced[1], &forced[2]);
    double best = 1e300;
    int bIA = 0, bRB = 0, bring = 0;
    for (int k = 0; k < (swap ? 3 : 5); ++k) {
        const int IA = swap ? cand_swap[k][0] : cand_plain[k][0], RB = swap ? cand_swap[k][1] : cand_plain[k][1];
        if (forced[0] && (IA != forced[0] || RB != forced[1])) continue;
        const int LA = 32 * IA, LB = 8 * RB, TY = swap ? LA : LB, TX = swap ? LB : LA;
        const long long BY = (long long)std::ceil(std::fabs(M[5]) * (TY - 1) + std::fabs(M[6]) * (TX - 1)) + 3;
        long long BX = (long long)std::ceil(std::fabs(M[9]) * (TY - 1) + std::fabs(M[10]) * (TX - 1)) + 3;
        BX = (BX + 3 + 3) / 4 * 4;
        if (BY > 256 || BX > 256) continue;
        const long long PB = (BY * BX + 31) / 32 * 32;
        const long long otile = swap ? 2LL * LA * (LB + 1) : 0;
        const int ctas = IA * RB <= 4 ? 3 : 2;
        const long long budget = (ctas == 3 ? 72 : 110) * 1024;
        for (int rl = 3; rl >= 2; --rl) {
            if (forced[2] && (1 << rl) != forced[2]) continue;
            const long long bytes = ((PB << rl) + otile) * 4 + 128;
            if (bytes > budget) continue;
            // cost: staged elements per output voxel; deeper rings and more columns per thread are preferred
            // (measured on config 3: with lanes along o1, 256-byte drained rows beat 128-byte ones by 5 %)
            // and with lanes along o2, 4 columns per thread at 3 CTAs per SM and 512-byte rows beat the rest by 3-4 %)
            const double cost = (double)(BY * BX) / (TY * TX) + (rl == 2 ? 0.15 : 0.0) + (IA * RB < 8 ? 0.05 : 0.0) -
                                (swap && LB >= 64 ? 0.25 : 0.0) + (!swap && IA * RB == 8 ? 0.25 : 0.0) -
                                (!swap && LA >= 128 ? 0.1 : 0.0);
            if (cost < best) {
                best = cost; bIA = IA; bRB = RB; bring = rl;
                P.BY = (int)BY; P.BX = (int)BX; P.pitch = (int)BX; P.PB = (int)PB; P.BZ = 1;
                P.LA = LA; P.LB = LB; P.TY = TY; P.TX = TX;
            }
            break;
        }
    }
    if (best == 1e300) return SHRIMPY_OK;
    P.ring_log2 = bring;
    const int nchunks = (P.oz + kStreamMaxSteps - 1) / kStreamMaxSteps;
    P.ZC = (P.oz + nchunks - 1) / nchunks;
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
    if (rc != CUDA_SUCCESS) return fail(SHRIMPY_ECUDA, "affine stream: cuTensorMapEncodeTiled failed (%d)", (int)rc);

    void (*kern)(const CUtensorMap, const AffineParams, const ZTable) = nullptr;
    const bool cl = nan_to_zero != 0;
    if (bIA == 4 && bRB == 2) kern = pick_stream<4, 2>(swap, cl);
    else if (bIA == 2 && bRB == 4) kern = pick_stream<2, 4>(swap, cl);
    else if (bIA == 2 && bRB == 2) kern = pick_stream<2, 2>(swap, cl);
    else if (bIA == 4 && bRB == 1) kern = pick_stream<4, 1>(swap, cl);
    else if (bIA == 2 && bRB == 1) kern = pick_stream<2, 1>(swap, cl);
    else if (bIA == 1 && bRB == 8) kern = pick_stream<1, 8>(swap, cl);
    else if (bIA == 1 && bRB == 4) kern = pick_stream<1, 4>(swap, cl);
    else return SHRIMPY_OK;
    const size_t smem = (((size_t)P.PB << P.ring_log2) + (swap ? 2 * (size_t)P.LA * (P.LB + 1) : 0)) * sizeof(float) + 128;
    if (getenv("SHRIMPY_DEBUG"))
        fprintf(stderr, "[shrimpy] affine stream IA=%d RB=%d swap=%d ring=%d box=(%d,%d) PB=%d ZC=%d smem=%zu grid=(%d,%d) x %d launches\n",
                bIA, bRB, (int)swap, 1 << P.ring_log2, P.BY, P.BX, P.PB, P.ZC, smem, P.tiles_x, P.tiles_y, nchunks);
    if (smem + 4096 > 48 * 1024)
        SHRIMPY_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int ch = 0; ch < nchunks; ++ch) {
        ZTable T;
        const int t0z = ch * P.ZC;
        build_ztable(P, t0z, std::min(P.ZC, P.oz - t0z), T);
        kern<<<dim3((unsigned)P.tiles_x, (unsigned)P.tiles_y), kStreamThreads, smem, s>>>(tmap, P, T);
        count_launch();
    }
    SHRIMPY_CUDA_TRY(cudaGetLastError());
    *launched = true;
    return SHRIMPY_OK;
}

}  // namespace shrimpy
