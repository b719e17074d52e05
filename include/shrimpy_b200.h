/*
 * shrimpy_b200 -- C-ABI of the B200-native light-sheet deskew / affine resample.
 *
 * The reference (czbiohub-sf/shrimPy) is pure Python and has no FFI of its own:
 * the boundary this library replaces is the set of Python call sites into the
 * un-vendored `biahub` package:
 *
 *   shrimpy/preprocessing.py:226-231   biahub.deskew.get_deskewed_data_shape
 *   shrimpy/preprocessing.py:408-413   biahub.deskew.fast_deskew_zyx
 *   scripts/measure_psf.py:230-246     biahub.analysis.deskew.{get_deskewed_data_shape,deskew_data}
 *
 * Every entry point below names the reference call it stands behind.  The
 * signatures carry plain pointers and sizes only (no torch types); the Python
 * host layer (shrimpy_b200/_cabi.py) binds them with ctypes, see INTEGRATION.md.
 *
 * Conventions
 *   - raw stack  raw[z, y, x], shape (Z, Y, X): axis 0 scan, axis 1 tilt,
 *     axis 2 coverslip (scripts/measure_psf.py:91,101).
 *   - deskewed   out[p, o1, o2], shape (ceil(Y/n), X, Xp), float32.
 *   - all device-level calls are asynchronous on the given stream, allocate
 *     nothing and never synchronise; they are thread-safe for distinct
 *     streams/buffers.
 *   - return value: 0 on success, otherwise a SHRIMPY_E* code; a human-readable
 *     message is available from shrimpy_last_error() (thread-local).
 */
#ifndef SHRIMPY_B200_H
#define SHRIMPY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SHRIMPY_B200_ABI_VERSION 1

enum {
    SHRIMPY_OK = 0,
    SHRIMPY_EINVAL = 1,   /* bad argument (shape, dtype, stride, alignment)   */
    SHRIMPY_ECUDA = 2,    /* a CUDA runtime / driver call failed                */
    SHRIMPY_ENOGPU = 3,   /* no sm_100 device visible                           */
    SHRIMPY_ENOMEM = 4    /* device or pinned-host allocation failed            */
};

enum {
    SHRIMPY_U16 = 0,      /* camera frames as acquired (mantis.yaml:8 "16bit") */
    SHRIMPY_F32 = 1       /* what shrimpy/preprocessing.py:316 hands over today */
};

/* Kernel selection for the deskew; AUTO picks TMA when strides/alignment allow. */
enum {
    SHRIMPY_KERNEL_AUTO = 0,
    SHRIMPY_KERNEL_DIRECT = 1, /* plain gather, any shape/stride/alignment       */
    SHRIMPY_KERNEL_TMA = 2,    /* TMA-staged smem tiles, 128-bit smem reads      */
    /* The TMA kernel with its results staged through shared memory and stored as 16-byte vectors on 32-byte
     * boundaries whatever the row pitch (csrc/deskew.cu deskew_tma_staged_kernel); same voxels bit for bit.  AUTO takes
     * it for average_n_slices == 1, where the writes dominate. */
    SHRIMPY_KERNEL_TMA_STAGED = 3
};

int shrimpy_abi_version(void);
const char *shrimpy_last_error(void);

/*
 * Geometry of the deskew (replaces biahub.deskew.get_deskewed_data_shape,
 * called at shrimpy/preprocessing.py:228-231 and scripts/measure_psf.py:230-234).
 *   out_shape[3]  = (ceil(Y/n), X, Xp)
 *   voxel_size[3] = (n*sin(theta)*pixel_size_um, pixel_size_um, pixel_size_um)
 *   row0[3]       = (M00, M02, Z_shift) of the output->input affine row that
 *                   maps (o0, o2) to the scan coordinate:
 *                   z_in = (Z_shift + o0*M00) + o2*M02
 * All arithmetic in float64, libm cos/sin of theta = ls_angle_deg * pi / 180.
 *
 * shrimpy_deskew_geometry_trig is the same computation with cos(theta) and sin(theta) supplied by the caller.  It is
 * the ONE code path of the geometry: the Python host calls it with numpy's cos / sin (what the upstream Python
 * evaluates), shrimpy_deskew_geometry calls it with libm's.  numpy and libm may differ in the last bit of a cosine,
 * and one ulp in M00 can move a voxel across the strict inside/outside rule or change ceil() in Xp; a host that must
 * reproduce the Python path bit for bit passes the trig values the Python path used.
 */
int shrimpy_deskew_geometry(int Z, int Y, int X, double ls_angle_deg, double px_to_scan_ratio,
                            int keep_overhang, int average_n_slices, double pixel_size_um,
                            int64_t out_shape[3], double voxel_size[3], double row0[3]);
int shrimpy_deskew_geometry_trig(int Z, int Y, int X, double cos_theta, double sin_theta, double px_to_scan_ratio,
                                 int keep_overhang, int average_n_slices, double pixel_size_um,
                                 int64_t out_shape[3], double voxel_size[3], double row0[3]);

/*
 * Device-resident deskew of one stack (replaces the body of
 * biahub.deskew.fast_deskew_zyx, called at shrimpy/preprocessing.py:408-413;
 * also absorbs the uint16->float32 convert of preprocessing.py:316 when
 * raw_dtype == SHRIMPY_U16).
 *
 *   d_raw      device pointer, element (z,y,x) at d_raw[z*raw_stride_z + y*raw_stride_y + x]
 *   d_out      device pointer, element (p,o1,o2) at d_out[p*out_stride_p + o1*out_stride_1 + o2]
 *   strides    in ELEMENTS; pass 0 for the contiguous default
 *   m00,m02,shift   row0 from shrimpy_deskew_geometry (or the host's own float64 values)
 *   n_avg      average_n_slices (edge-replicated when Y % n_avg != 0)
 *   cval       value written where z_in < 0 or z_in > Z-1 (strict)
 *   kernel     SHRIMPY_KERNEL_*
 *   stream     cudaStream_t (NULL = legacy default stream)
 */
int shrimpy_deskew_device(const void *d_raw, int raw_dtype, float *d_out, int Z, int Y, int X, int Xp,
                          int n_avg, double m00, double m02, double shift, float cval,
                          int64_t raw_stride_z, int64_t raw_stride_y, int64_t out_stride_p,
                          int64_t out_stride_1, int kernel, void *stream);

/*
 * Windowed form of the same kernel: computes out[p_begin:p_begin+p_count, :, c_begin:c_begin+c_count]
 * from a raw SLAB that holds only rows y in [y_origin, y_origin+y_count) and scan slices
 * z in [z_origin, z_origin+z_count) of the full (Z,Y,X) stack.  Geometry (Z, Y, Xp, the
 * affine row) always refers to the FULL stack, so every voxel is bit-identical to the
 * un-windowed call.  Used by the host pipeline (tilt slabs, halo-free), by tilt/column
 * sharding across GPUs, and by the scan-axis split with halo.
 *   d_raw  element (z,y,x) at d_raw[(z-z_origin)*raw_stride_z + (y-y_origin)*raw_stride_y + x]
 *   d_out  element (p,o1,o2) at d_out[(p-p_begin)*out_stride_p + o1*out_stride_1 + (o2-c_begin)]
 * The slab must contain every row/slice the window needs; otherwise SHRIMPY_EINVAL.
 * shrimpy_deskew_window_needs() reports that range.
 */
typedef struct shrimpy_window {
    int32_t p_begin, p_count; /* tilt blocks (output axis 0)       */
    int32_t c_begin, c_count; /* output columns (output axis 2)    */
    int32_t y_origin, y_count; /* raw tilt rows present in d_raw   */
    int32_t z_origin, z_count; /* raw scan slices present in d_raw */
} shrimpy_window;

int shrimpy_deskew_window_needs(int Z, int Y, int n_avg, double m00, double m02, double shift, int p_begin,
                                int p_count, int c_begin, int c_count, int32_t y_range[2], int32_t z_range[2]);

int shrimpy_deskew_window_device(const void *d_raw, int raw_dtype, float *d_out, int Z, int Y, int X, int Xp,
                                 int n_avg, double m00, double m02, double shift, float cval,
                                 int64_t raw_stride_z, int64_t raw_stride_y, int64_t out_stride_p,
                                 int64_t out_stride_1, const shrimpy_window *window, int kernel, void *stream);

/*
 * Bright-field flat-field correction (replaces _LabelfreePreprocessor._flat_field_BF,
 * shrimpy/preprocessing.py:385-404: static_pattern = volume.quantile(0.5, dim=0);
 * volume / static_pattern * static_pattern.mean()), expressed as a per-pixel scale field
 *     scale[y, x] = mean(pattern) / pattern[y, x],   pattern = median over the scan axis
 * so that it can be fused into the deskew (the step that follows it at preprocessing.py:320-327).
 *   shrimpy_flatfield_pattern_device  per-pixel median over Z (numpy.median semantics), (Y, X) float32
 *   shrimpy_flatfield_scale_device    scale field from the pattern; d_scratch = one double of device scratch
 *   shrimpy_flatfield_apply_device    stand-alone correction: out = raw * scale (float32)
 *   shrimpy_deskew_flatfield_device   deskew with the scale field applied inside the kernel; the scale field
 *                                     always refers to the FULL (Y, X) frame, also for windowed calls
 */
int shrimpy_flatfield_pattern_device(const void *d_raw, int raw_dtype, float *d_pattern, int Z, int Y, int X,
                                     int64_t raw_stride_z, int64_t raw_stride_y, void *stream);
int shrimpy_flatfield_scale_device(const float *d_pattern, int64_t count, float *d_scale, double *d_scratch,
                                   void *stream);
int shrimpy_flatfield_apply_device(const void *d_raw, int raw_dtype, const float *d_scale, float *d_out, int Z, int Y,
                                   int X, void *stream);
int shrimpy_deskew_flatfield_device(const void *d_raw, int raw_dtype, const float *d_scale, float *d_out, int Z, int Y,
                                    int X, int Xp, int n_avg, double m00, double m02, double shift, float cval,
                                    int64_t raw_stride_z, int64_t raw_stride_y, const shrimpy_window *window,
                                    int kernel, void *stream);

/*
 * Deskew with the value range of the result reduced in the same pass (fused epilogue): d_range2 receives
 * (min, max) over every voxel written, padding included -- what torch.min / torch.max of the deskewed volume
 * return in the tracking step that follows the deskew (shrimpy/dynatrack/tracking.py:583-584), so that the
 * histogram pass of shrimpy_hist256_device can start without a pass of its own.  d_scale may be NULL (no
 * flat-field).  Full stack only (no window); asynchronous like every device call.
 */
int shrimpy_deskew_range_device(const void *d_raw, int raw_dtype, const float *d_scale, float *d_out, float *d_range2,
                                int Z, int Y, int X, int Xp, int n_avg, double m00, double m02, double shift,
                                float cval, int64_t raw_stride_z, int64_t raw_stride_y, int kernel, void *stream);

/*
 * Device-resident trilinear resample with a 3x4 (row-major, 12 doubles)
 * output-index -> input-index matrix in ZYX voxel units (the registration
 * resample named by BASELINE.json configs[2]; upstream
 * biahub apply_affine_transform(method="scipy") == scipy.ndimage.affine_transform
 * (order=1, mode="constant")).  NaN inputs are read as 0 when nan_to_zero != 0.
 */
int shrimpy_affine_device(const float *d_in, float *d_out, int iz, int iy, int ix, int oz, int oy, int ox,
                          const double M[12], float cval, int nan_to_zero, void *stream);
/*
 * The same resample from an input whose rows / planes are padded: element (z,y,x) at
 * d_in[z*in_stride_z + y*in_stride_y + x] (strides in elements, multiples of 4 so that rows start on 16-byte
 * boundaries).  The plane-streaming kernels read through a tensor map, which takes the strides and zero-fills
 * beyond the logical ix; this is how inputs whose X is not a multiple of 4 (e.g. a deskewed (100,2048,1279) volume)
 * reach them -- the Python layer pads the rows once.  Returns SHRIMPY_EINVAL ("not eligible") when the matrix needs
 * one of the dense-only kernels; the caller then uses shrimpy_affine_device on the dense array.
 */
int shrimpy_affine_strided_device(const float *d_in, float *d_out, int iz, int iy, int ix, int64_t in_stride_z,
                                  int64_t in_stride_y, int oz, int oy, int ox, const double M[12], float cval,
                                  int nan_to_zero, void *stream);

/*
 * Reductions over a deskewed float32 volume used by the tracking step that consumes it
 * (shrimpy/dynatrack/tracking.py:572-649): value range, torch.histc-style 256-bin histogram for the background
 * percentile, and the sums of the intensity centre of mass with weights w = max(v - background, 0):
 *   d_sums4 = (sum w, sum w*z, sum w*y, sum w*x) in float64.  One streaming pass each.
 */
int shrimpy_minmax_device(const float *d_data, int64_t count, float *d_out2, void *stream);
int shrimpy_hist256_device(const float *d_data, int64_t count, float vmin, float vmax, uint64_t *d_hist, void *stream);
int shrimpy_center_of_mass_device(const float *d_data, int Z, int Y, int X, float background, double *d_sums4,
                                  void *stream);
/* Z max-projection of the background-filtered volume, d_out[y][x] = max_z max(v - background, 0): the image the
 * tracker saves next to the centroid (shrimpy/dynatrack/tracking.py:1447-1455). */
int shrimpy_zmax_projection_device(const float *d_data, int Z, int Y, int X, float background, float *d_out,
                                   void *stream);

/* min over a device array (for cval = min(raw), the scipy-generation default). */
int shrimpy_min_device(const void *d_raw, int raw_dtype, int64_t count, float *d_result, void *stream);

/*
 * Host-buffer deskew (replaces biahub.analysis.deskew.deskew_data as called at
 * scripts/measure_psf.py:239-246: numpy in, numpy out).  The stack is cut into
 * tilt slabs (multiples of n_avg rows, halo-free: SURVEY.md 8e) that are
 * streamed H2D -> kernel -> D2H on three streams with double-buffered device
 * slabs.  h_raw / h_out may be pageable (an ordinary numpy array, as
 * scripts/measure_psf.py:239-246 passes): such a buffer is gathered into /
 * scattered out of a page-locked ring by a few host threads
 * (SHRIMPY_HOST_THREADS, default min(8, cores/2)) while the neighbouring slabs'
 * copies and kernels run, so the GPU only ever sees asynchronous page-locked
 * transfers; pinned buffers are copied directly.
 * Synchronous: returns when h_out is complete, and with nothing in flight
 * whatever it returns.  Calls on one pipeline are serialised by a mutex inside
 * it; distinct pipelines run concurrently.
 */
typedef struct shrimpy_pipeline shrimpy_pipeline;

int shrimpy_pipeline_create(int device, size_t device_bytes_budget, shrimpy_pipeline **out);
void shrimpy_pipeline_destroy(shrimpy_pipeline *p);
int shrimpy_deskew_host(shrimpy_pipeline *p, const void *h_raw, int raw_dtype, float *h_out, int Z, int Y,
                        int X, int Xp, int n_avg, double m00, double m02, double shift, float cval);
/* counters of the last shrimpy_deskew_host call: kernels launched, H2D / D2H bytes */
int shrimpy_pipeline_stats(const shrimpy_pipeline *p, int64_t *launches, int64_t *h2d_bytes, int64_t *d2h_bytes);
/* bytes of the last call that went through the pipeline's own page-locked staging rings (pageable h_raw / h_out) */
int shrimpy_pipeline_staged_bytes(const shrimpy_pipeline *p, int64_t *in_bytes, int64_t *out_bytes);
/* Page-locked host memory of exactly `bytes` bytes (cudaHostAlloc, portable) and its release: what the Python host
 * returns the result of deskew_data in when the caller gives no `out` (scripts/measure_psf.py:239-246 takes what the
 * call returns), so that the device-to-host copies of the pipeline run asynchronously at the PCIe rate. */
int shrimpy_host_alloc(size_t bytes, void **out);
int shrimpy_host_free(void *ptr);

/*
 * Blosc-1 frame codec for the OME-Zarr chunk loader (host memory only; no CUDA call).  The reference acquires with
 * compression="blosc-zstd" inside zarr-v3 shards (shrimpy/mantis/mantis_engine.py:474-481; asserted by
 * shrimpy/tests/test_mantis_integration.py:182-188), so streaming its stores means undoing blosc's framing
 * (16-byte header, block table, per-block split streams, byte/bit shuffle) around zstd / lz4 / zlib streams.
 * The stream codecs are taken from the system's runtime libraries via dlopen.  shrimpy_blosc_decode writes exactly
 * dst_bytes (= the frame's nbytes) into dst, decoding blocks on up to `threads` threads; it may be called
 * concurrently from many threads.  shrimpy_blosc_encode writes one frame (codec: 1 lz4, 3 zlib, 4 zstd; shuffle:
 * 0 none, 1 byte, 2 bit; blocksize 0 = 256 KiB; split != 0 = one stream per byte plane where blosc's rule allows).
 */
int shrimpy_blosc_info(const void *frame, size_t frame_bytes, int64_t *nbytes, int64_t *cbytes, int32_t *blocksize,
                       int32_t *typesize, int32_t *flags);
int shrimpy_blosc_decode(const void *frame, size_t frame_bytes, void *dst, size_t dst_bytes, int threads);
size_t shrimpy_blosc_encode_bound(size_t nbytes, int32_t blocksize, int typesize);
int shrimpy_blosc_encode(const void *data, size_t nbytes, int typesize, int codec, int level, int shuffle,
                         int32_t blocksize, int split, void *frame, size_t frame_capacity, size_t *frame_bytes);

/* Number of kernel launches issued by this library in this process (all entry points). */
int64_t shrimpy_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SHRIMPY_B200_H */
