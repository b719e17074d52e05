// Post-deskew reductions used by the tracking step that follows the deskew:
//   min/max                     -> range of the histogram          (shrimpy/dynatrack/tracking.py:583-584)
//   256-bin histogram           -> background percentile           (tracking.py:587-595, torch.histc semantics)
//   intensity centre of mass    -> sum w, sum w*z, sum w*y, sum w*x with w = max(v - background, 0)
//                                                                    (tracking.py:626-649)
// Each is one streaming pass over the float32 volume (HBM-bound: 4 bytes per voxel).
#include "common.cuh"

#include <algorithm>
#include <cfloat>

namespace shrimpy {

// A read-only 16-byte load the compiler may not reorder against its siblings or fold into one register set: a run of
// these is issued back to back, which is the point (with plain __ldg the guarded loads were serialised through R8).
__device__ __forceinline__ float4 ldg_in_order(const float4 *p) {
    float4 r;
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

__global__ void minmax_init_kernel(unsigned *slots) {
    slots[0] = 0xffffffffu;
    slots[1] = 0u;
}

__global__ void __launch_bounds__(256) minmax_kernel(const float4 *__restrict__ v4, const float *__restrict__ v,
                                                     long long n4, long long n, unsigned *slots) {
    float lo = FLT_MAX, hi = -FLT_MAX;
    const long long stride = (long long)gridDim.x * blockDim.x;
    auto take = [&](const float4 a) {
        lo = fminf(fminf(lo, a.x), fminf(a.y, fminf(a.z, a.w)));
        hi = fmaxf(fmaxf(hi, a.x), fmaxf(a.y, fmaxf(a.z, a.w)));
    };
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 7 * stride < n4; i += 8 * stride) {          // eight 16-byte loads in flight per thread
        float4 a[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = ldg_in_order(v4 + i + k * stride);
#pragma unroll
        for (int k = 0; k < 8; ++k) take(a[k]);
    }
    for (; i < n4; i += stride) take(__ldg(v4 + i));
    for (long long i = 4 * n4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        lo = fminf(lo, v[i]);
        hi = fmaxf(hi, v[i]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(slots, ordered_key(lo));
        atomicMax(slots + 1, ordered_key(hi));
    }
}

__global__ void minmax_finish_kernel(const unsigned *slots, float *out) {
    out[0] = ordered_value(slots[0]);
    out[1] = ordered_value(slots[1]);
}

// torch.histc: bin = floor((v - min) / (max - min) * nbins), v == max falls in the last bin, values outside ignored.
// One sub-histogram per warp in shared memory keeps the atomics short.
//
// The bin is trunc((x - vmin) / range * 256) with an IEEE division (~10 instructions).  The fast path computes
// k = trunc((x - vmin) * (2^23 / range)) instead: bin = k >> 15, and the low 15 bits are the position inside the bin.
// The product is within 8e-5 of a bin width of the exact quotient (three float32 roundings of a value <= 256 plus the
// truncation to 2^-15), so the two can disagree only when those bits are within 4 units of a bin edge; only such
// voxels (flagged in a mask, ~2.4e-4 of uniform data) take the division, in a second, rarely executed block.
// Bin 256 (x == vmax) has its own counter, folded into 255.
__global__ void __launch_bounds__(256) hist256_kernel(const float *__restrict__ v, long long n, float vmin, float vmax,
                                                      unsigned long long *hist) {
    constexpr int kTrash = 257;                       // counter for voxels the fast path does not bin
    __shared__ unsigned sub[8][258];
    for (int i = threadIdx.x; i < 8 * 258; i += 256) (&sub[0][0])[i] = 0u;
    __syncthreads();
    const float range = vmax - vmin;
    const float scale = 8388608.0f / range;
    unsigned *mine = sub[threadIdx.x >> 5];
    auto exact = [&](float x) {
        if (x >= vmin && x <= vmax) {
            const int b = (int)((x - vmin) / range * 256.0f);
            atomicAdd(mine + min(b, 255), 1u);
        }
    };
    // Branch-free: every voxel issues one shared atomic -- to its bin when it lies in range and clear of a bin edge,
    // else to the trash counter.  Out-of-range values give k4 < 0 or k4 > 2^23 + 8 (or land on the first / last edge),
    // NaN converts to 0 (an edge).  Returns true when the voxel has to be looked at by the exact path.
    auto fast = [&](float x) -> bool {
        const int k4 = (int)((x - vmin) * scale) + 4;
        const bool interior = (k4 & 0x7ff8) != 0;
        const bool ok = interior && (unsigned)k4 <= 8388608u + 8u;
        atomicAdd(mine + (ok ? k4 >> 15 : kTrash), 1u);
        return !interior;
    };
    auto count = [&](float x) {
        if (fast(x)) exact(x);
    };
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    // 16-byte loads, four in flight per thread; scalar head and tail
    const long long head = min(n, (long long)((16u - (unsigned)(reinterpret_cast<uintptr_t>(v) & 15u)) & 15u) / 4);
    const long long n4 = (n - head) / 4;
    const float4 *v4 = reinterpret_cast<const float4 *>(v + head);
    long long i = tid;
    if (i + 3 * stride < n4) {
        // software pipeline: the four loads of the next batch are in flight while this batch is binned
        float4 a = __ldg(v4 + i), b = __ldg(v4 + i + stride), c = __ldg(v4 + i + 2 * stride), d = __ldg(v4 + i + 3 * stride);
        for (;;) {
            const long long nx = i + 4 * stride;
            const bool more = nx + 3 * stride < n4;
            float4 na = a, nb = b, nc = c, nd = d;
            if (more) {
                na = __ldg(v4 + nx); nb = __ldg(v4 + nx + stride); nc = __ldg(v4 + nx + 2 * stride); nd = __ldg(v4 + nx + 3 * stride);
            }
            const float e[16] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w, d.x, d.y, d.z, d.w};
            unsigned todo = 0;
#pragma unroll
            for (int q = 0; q < 16; ++q) todo |= (unsigned)fast(e[q]) << q;
            if (todo) {
#pragma unroll
                for (int q = 0; q < 16; ++q)
                    if ((todo >> q) & 1u) exact(e[q]);
            }
            i = nx;
            if (!more) break;
            a = na; b = nb; c = nc; d = nd;
        }
    }
    for (; i < n4; i += stride) {
        const float4 a = __ldg(v4 + i);
        count(a.x); count(a.y); count(a.z); count(a.w);
    }
    for (long long j = tid; j < head; j += stride) count(__ldg(v + j));
    for (long long j = head + 4 * n4 + tid; j < n; j += stride) count(__ldg(v + j));
    __syncthreads();
    for (int b = threadIdx.x; b < 256; b += 256) {
        unsigned long long t = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += sub[w][b] + (b == 255 ? sub[w][256] : 0u);
        if (t) atomicAdd(hist + b, t);
    }
}

// sums[0..3] += (sum w, sum w*z, sum w*y, sum w*x); one (z, y) row per warp iteration, lanes along x.
// VEC: the row is read as a scalar head (up to the next 16-byte boundary), 16-byte vectors and a scalar tail -- rows of
// a deskewed volume are 4 * 1279 bytes long, so the phase changes from row to row.  Twelve vectors per lane are in flight
// -- a whole 1279-voxel row per warp -- (the scalar version, eight 4-byte loads per lane, stopped at 3.3 TB/s).  Sums are float32 within a row (per lane),
// float64 across rows.
template <bool VEC>
__global__ void __launch_bounds__(256) com_kernel(const float *__restrict__ v, int Z, int Y, int X, float background,
                                                  double *sums) {
    double s = 0.0, sz = 0.0, sy = 0.0, sx = 0.0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long rows = (long long)Z * Y;
    for (long long r = (long long)blockIdx.x * 8 + warp; r < rows; r += (long long)gridDim.x * 8) {
        const int z = (int)(r / Y), y = (int)(r - (long long)z * Y);
        const float *row = v + r * X;
        float w_row = 0.f, wx_row = 0.f;
        auto take = [&](float val, float xf) {
            const float w = fmaxf(val - background, 0.f);
            w_row += w;
            wx_row = fmaf(w, xf, wx_row);
        };
        if (VEC) {
            const int head = min(X, (int)((4u - (unsigned)((r * X) & 3)) & 3u));   // v itself is 16-byte aligned
            const int nvec = (X - head) >> 2;
            const float4 *body = reinterpret_cast<const float4 *>(row + head);
            auto take4 = [&](const float4 a, int j) {
                const float x0 = (float)(head + 4 * j);
                take(a.x, x0); take(a.y, x0 + 1.f); take(a.z, x0 + 2.f); take(a.w, x0 + 3.f);
            };
            // head and tail elements first (their loads overlap the vector loads), then chunks of kChunk vectors per
            // lane, all issued before the first is consumed (slots past the row re-read its last vector and are skipped)
            constexpr int kChunk = 12;
            const int t = head + 4 * nvec + lane;
            const float hv = lane < head ? __ldg(row + lane) : -INFINITY;
            const float tv = t < X ? __ldg(row + t) : -INFINITY;
            for (int j0 = lane; j0 < nvec + lane; j0 += 32 * kChunk) {
                float4 a[kChunk];
#pragma unroll
                for (int k = 0; k < kChunk; ++k) a[k] = ldg_in_order(body + min(j0 + 32 * k, nvec - 1));
#pragma unroll
                for (int k = 0; k < kChunk; ++k)
                    if (j0 + 32 * k < nvec) take4(a[k], j0 + 32 * k);
            }
            take(hv, (float)lane);
            take(tv, (float)t);
        } else {
            int x = lane;
            for (; x + 7 * 32 < X; x += 8 * 32) {
                float v8[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v8[j] = __ldg(row + x + 32 * j);
#pragma unroll
                for (int j = 0; j < 8; ++j) take(v8[j], (float)(x + 32 * j));
            }
            for (; x < X; x += 32) take(__ldg(row + x), (float)x);
        }
        s += (double)w_row;
        sx += (double)wx_row;
        sz += (double)w_row * z;
        sy += (double)w_row * y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        sz += __shfl_xor_sync(0xffffffffu, sz, o);
        sy += __shfl_xor_sync(0xffffffffu, sy, o);
        sx += __shfl_xor_sync(0xffffffffu, sx, o);
    }
    if (lane == 0) {
        atomicAdd(sums + 0, s);
        atomicAdd(sums + 1, sz);
        atomicAdd(sums + 2, sy);
        atomicAdd(sums + 3, sx);
    }
}

// Z max-projection of the background-filtered volume, out[i] = max_z max(v[z][i] - background, 0) over the flattened
// (Y, X) plane (tracking.py:1447-1455: `(img - background).clamp_min(0).amax(dim=0)`).  VEC floats per thread, eight
// planes in flight; max is order-independent, so the result equals torch's bit for bit.
template <int VEC>
__global__ void __launch_bounds__(256) zmax_kernel(const float *__restrict__ v, int Z, long long plane, float background,
                                                   float *__restrict__ out) {
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (i >= plane) return;
    float best[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) best[k] = 0.f;
    auto take = [&](const float *p) {
        if (VEC == 4) {
            const float4 a = __ldg(reinterpret_cast<const float4 *>(p));
            best[0] = fmaxf(best[0], a.x - background);
            best[1 % VEC] = fmaxf(best[1 % VEC], a.y - background);
            best[2 % VEC] = fmaxf(best[2 % VEC], a.z - background);
            best[3 % VEC] = fmaxf(best[3 % VEC], a.w - background);
        } else {
            best[0] = fmaxf(best[0], __ldg(p) - background);
        }
    };
    const float *col = v + i;
    int z = 0;
    for (; z + 8 <= Z; z += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) take(col + (long long)(z + k) * plane);
    }
    for (; z < Z; ++z) take(col + (long long)z * plane);
    if (VEC == 4) {
        *reinterpret_cast<float4 *>(out + i) = make_float4(best[0], best[1 % VEC], best[2 % VEC], best[3 % VEC]);
    } else {
        out[i] = best[0];
    }
}

}  // namespace shrimpy

using namespace shrimpy;

static int grid_for(long long work_items, int per_block) {
    int dev = 0;
    cudaGetDevice(&dev);
    return (int)std::min<long long>((work_items + per_block - 1) / per_block, (long long)sm_count(dev) * 8);
}

extern "C" int shrimpy_minmax_device(const float *d_data, int64_t count, float *d_out2, void *stream) {
    if (!d_data || !d_out2 || count <= 0) return fail(SHRIMPY_EINVAL, "minmax: bad arguments");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    unsigned *slots = reinterpret_cast<unsigned *>(d_out2);   // the result words double as the ordered-key scratch
    const bool vec = (reinterpret_cast<uintptr_t>(d_data) & 15u) == 0;
    const long long n4 = vec ? count / 4 : 0;
    minmax_init_kernel<<<1, 1, 0, s>>>(slots);
    minmax_kernel<<<grid_for(std::max<long long>(n4, count / 4 + 1), 256), 256, 0, s>>>(
        reinterpret_cast<const float4 *>(d_data), d_data, n4, count, slots);
    minmax_finish_kernel<<<1, 1, 0, s>>>(slots, d_out2);
    count_launch(3);
    SHRIMPY_CUDA_TRY(cudaGetLastError());
    return SHRIMPY_OK;
}

extern "C" int shrimpy_hist256_device(const float *d_data, int64_t count, float vmin, float vmax, uint64_t *d_hist,
                                      void *stream) {
    if (!d_data || !d_hist || count <= 0 || !(vmax > vmin)) return fail(SHRIMPY_EINVAL, "hist256: bad arguments");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    SHRIMPY_CUDA_TRY(cudaMemsetAsync(d_hist, 0, 256 * sizeof(uint64_t), s));
    hist256_kernel<<<grid_for(count, 256 * 16), 256, 0, s>>>(d_data, count, vmin, vmax,
                                                              reinterpret_cast<unsigned long long *>(d_hist));
    count_launch();
    SHRIMPY_CUDA_TRY(cudaGetLastError());
    return SHRIMPY_OK;
}

extern "C" int shrimpy_center_of_mass_device(const float *d_data, int Z, int Y, int X, float background, double *d_sums4,
                                             void *stream) {
    if (!d_data || !d_sums4 || Z <= 0 || Y <= 0 || X <= 0) return fail(SHRIMPY_EINVAL, "center_of_mass: bad arguments");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    SHRIMPY_CUDA_TRY(cudaMemsetAsync(d_sums4, 0, 4 * sizeof(double), s));
    const int grid = grid_for((long long)Z * Y, 8 * 4);
    if ((reinterpret_cast<uintptr_t>(d_data) & 15u) == 0 && X >= 64)
        com_kernel<true><<<grid, 256, 0, s>>>(d_data, Z, Y, X, background, d_sums4);
    else
        com_kernel<false><<<grid, 256, 0, s>>>(d_data, Z, Y, X, background, d_sums4);
    count_launch();
    SHRIMPY_CUDA_TRY(cudaGetLastError());
    return SHRIMPY_OK;
}

extern "C" int shrimpy_zmax_projection_device(const float *d_data, int Z, int Y, int X, float background, float *d_out,
                                              void *stream) {
    if (!d_data || !d_out || Z <= 0 || Y <= 0 || X <= 0) return fail(SHRIMPY_EINVAL, "zmax_projection: bad arguments");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const long long plane = (long long)Y * X;
    const bool vec = plane % 4 == 0 && ((reinterpret_cast<uintptr_t>(d_data) | reinterpret_cast<uintptr_t>(d_out)) & 15u) == 0;
    if (vec)
        zmax_kernel<4><<<(unsigned)((plane / 4 + 255) / 256), 256, 0, s>>>(d_data, Z, plane, background, d_out);
    else
        zmax_kernel<1><<<(unsigned)((plane + 255) / 256), 256, 0, s>>>(d_data, Z, plane, background, d_out);
    count_launch();
    SHRIMPY_CUDA_TRY(cudaGetLastError());
    return SHRIMPY_OK;
}
