// Library-wide state, geometry and the min reduction of the shrimpy_b200 C-ABI.
#include "common.cuh"

#include <cfloat>
#include <cmath>
#include <mutex>

namespace shrimpy {

std::atomic<int64_t> g_launches{0};

char *last_error_buffer() {
    static thread_local char buf[kErrLen] = {0};
    return buf;
}

EncodeTiledFn tensor_map_encoder() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    });
    return fn;
}

int sm_count(int device) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return 148;
    return n;
}

// ---- min reduction (cval = min(raw)) -------------------------------------------
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Order-preserving float <-> uint map so one atomicMin on an unsigned word works for any sign.
__device__ __forceinline__ unsigned ordered_bits(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void min_init_kernel(unsigned *slot) { *slot = 0xffffffffu; }

template <typename T>
__global__ void __launch_bounds__(256) min_kernel(const T *__restrict__ data, long long count, unsigned *slot) {
    float m = FLT_MAX;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
        m = fminf(m, (float)__ldg(data + i));
    m = warp_min(m);
    __shared__ float part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < 8 ? part[threadIdx.x] : FLT_MAX;
        m = warp_min(m);
        if (threadIdx.x == 0) atomicMin(slot, ordered_bits(m));
    }
}

__global__ void min_finish_kernel(const unsigned *slot, float *result) {
    const unsigned u = *slot;
    *result = __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

}  // namespace shrimpy

using namespace shrimpy;

extern "C" int shrimpy_abi_version(void) { return SHRIMPY_B200_ABI_VERSION; }

extern "C" const char *shrimpy_last_error(void) { return last_error_buffer(); }

extern "C" int64_t shrimpy_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

// The one place the scalar geometry is computed: the Python host (shrimpy_b200/deskew.py) calls it with numpy's
// cos / sin -- what the upstream Python evaluates -- and shrimpy_deskew_geometry with libm's.  Every product and sum
// is rounded separately (volatile temporaries, -ffp-contract=off), in the order the upstream expressions have.
extern "C" int shrimpy_deskew_geometry_trig(int Z, int Y, int X, double cos_theta, double sin_theta,
                                            double px_to_scan_ratio, int keep_overhang, int average_n_slices,
                                            double pixel_size_um, int64_t out_shape[3], double voxel_size[3],
                                            double row0[3]) {
    if (Z <= 0 || Y <= 0 || X <= 0 || average_n_slices <= 0 || !(px_to_scan_ratio > 0.0))
        return fail(SHRIMPY_EINVAL, "geometry: bad arguments");
    if (!std::isfinite(cos_theta) || !std::isfinite(sin_theta)) return fail(SHRIMPY_EINVAL, "geometry: non-finite angle");
    const double st = sin_theta, ct = cos_theta;
    volatile double zr = (double)Z / px_to_scan_ratio;
    volatile double yc = (double)Y * ct;
    const double xp = keep_overhang ? ceil(zr + yc) : ceil(zr - yc);
    if (out_shape) {
        out_shape[0] = (Y + average_n_slices - 1) / average_n_slices;
        out_shape[1] = X;
        out_shape[2] = xp > 0.0 ? (int64_t)xp : 0;
    }
    if (voxel_size) {
        volatile double nst = (double)average_n_slices * st;
        voxel_size[0] = nst * pixel_size_um;
        voxel_size[1] = pixel_size_um;
        voxel_size[2] = pixel_size_um;
    }
    if (row0) {
        volatile double yct = (double)Y * ct;
        row0[0] = -px_to_scan_ratio * ct;
        row0[1] = px_to_scan_ratio;
        row0[2] = keep_overhang ? 0.0 : floor(yct * px_to_scan_ratio);
    }
    return SHRIMPY_OK;
}

extern "C" int shrimpy_deskew_geometry(int Z, int Y, int X, double ls_angle_deg, double px_to_scan_ratio,
                                       int keep_overhang, int average_n_slices, double pixel_size_um,
                                       int64_t out_shape[3], double voxel_size[3], double row0[3]) {
    volatile double deg_pi = ls_angle_deg * M_PI;
    const double theta = deg_pi / 180.0;
    return shrimpy_deskew_geometry_trig(Z, Y, X, cos(theta), sin(theta), px_to_scan_ratio, keep_overhang,
                                        average_n_slices, pixel_size_um, out_shape, voxel_size, row0);
}

extern "C" int shrimpy_min_device(const void *d_raw, int raw_dtype, int64_t count, float *d_result, void *stream) {
    if (!d_raw || !d_result || count <= 0) return fail(SHRIMPY_EINVAL, "min: bad arguments");
    if (raw_dtype != SHRIMPY_U16 && raw_dtype != SHRIMPY_F32) return fail(SHRIMPY_EINVAL, "min: bad dtype");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // the float result slot doubles as the unsigned scratch word
    unsigned *slot = reinterpret_cast<unsigned *>(d_result);
    int dev = 0;
    SHRIMPY_CUDA_TRY(cudaGetDevice(&dev));
    const int blocks = (int)std::min<long long>((count + 255) / 256, (long long)sm_count(dev) * 8);
    min_init_kernel<<<1, 1, 0, s>>>(slot);
    if (raw_dtype == SHRIMPY_U16)
        min_kernel<uint16_t><<<blocks, 256, 0, s>>>(static_cast<const uint16_t *>(d_raw), count, slot);
    else
        min_kernel<float><<<blocks, 256, 0, s>>>(static_cast<const float *>(d_raw), count, slot);
    min_finish_kernel<<<1, 1, 0, s>>>(slot, d_result);
    count_launch(3);
    SHRIMPY_CUDA_TRY(cudaGetLastError());
    return SHRIMPY_OK;
}
