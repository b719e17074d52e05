/*
 * Plain-C restatement of the deskew / affine-resample oracle.
 *
 * TEST INFRASTRUCTURE ONLY (checker and CPU baseline for tests/, smoke() and
 * bench.py's cpu legs).  The product path never links or loads this file.
 *
 * PARITY UNPINNED: the arithmetic lives in biahub @ b011bca5 (un-vendored; see
 * oracle/deskew_oracle.py for the full statement).  Each function follows the
 * closed form in SURVEY.md appendix C, i.e. what
 *   scipy.ndimage.affine_transform(raw.astype(float32), M, order=1,
 *                                  mode="constant", cval) + edge-padded mean
 * computes for the call sites shrimpy/preprocessing.py:408-413 and
 * scripts/measure_psf.py:239-246:
 *   - coordinates in float64, accumulated as ((shift + o0*M0) + o1*M1) + o2*M2,
 *     no fused multiply-add (compiled with -ffp-contract=off),
 *   - strict inside test 0 <= c <= len-1,
 *   - the lerp evaluated in float64 and rounded once to float32,
 *   - the block mean as a sequential float32 sum divided by n (numpy's
 *     reduction order for a short non-contiguous axis).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off).
 */
#include <float.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>

/*
 * No OpenMP on purpose: a second OpenMP runtime next to torch's is a known
 * crash source (reference tests/conftest.py:11-17).  Both entry points take a
 * [begin, end) range over the outermost output axis so that the Python side
 * can fan slabs out over plain threads (ctypes releases the GIL).
 */

static inline double load_raw(const void *raw, int is_f32, size_t idx) {
    return is_f32 ? (double)((const float *)raw)[idx] : (double)((const uint16_t *)raw)[idx];
}

/*
 * Deskew of one (Z,Y,X) stack -> (ceil(Y/n), X, Xp) float32.
 * m00 = -r*cos(theta), m02 = r, shift = Z_shift, as produced on the host in
 * float64 (oracle/deskew_oracle.py:deskew_affine_matrix).
 */
int oracle_deskew(const void *raw, int is_f32, float *out, int Z, int Y, int X, int Xp, int n,
                  double m00, double m02, double shift, float cval, int p_begin, int p_end) {
    if (Z <= 0 || Y <= 0 || X <= 0 || Xp < 0 || n <= 0) return 1;
    const int Yn = (Y + n - 1) / n;
    const size_t sy = (size_t)X, sz = (size_t)Y * (size_t)X;
    if (p_begin < 0) p_begin = 0;
    if (p_end > Yn) p_end = Yn;
    for (int p = p_begin; p < p_end; ++p) {
        for (int o1 = 0; o1 < X; ++o1) {
            float *row = out + ((size_t)p * X + o1) * (size_t)Xp;
            const int x = X - 1 - o1;
            for (int o2 = 0; o2 < Xp; ++o2) {
                float acc = 0.0f;
                for (int k = 0; k < n; ++k) {
                    int o0 = n * p + k;
                    if (o0 > Y - 1) o0 = Y - 1; /* edge replication of the last plane */
                    const int y = Y - 1 - o0;
                    double z = shift + (double)o0 * m00;
                    z = z + (double)o2 * m02;
                    float v;
                    if (z < 0.0 || z > (double)(Z - 1)) {
                        v = cval;
                    } else {
                        const double fz = floor(z);
                        const double w = z - fz;
                        const int z0 = (int)fz;
                        const int z1 = z0 + 1 > Z - 1 ? Z - 1 : z0 + 1;
                        const double a = load_raw(raw, is_f32, (size_t)z0 * sz + (size_t)y * sy + x);
                        const double b = load_raw(raw, is_f32, (size_t)z1 * sz + (size_t)y * sy + x);
                        v = (float)((1.0 - w) * a + w * b);
                    }
                    acc = (k == 0) ? v : acc + v;
                }
                row[o2] = (n == 1) ? acc : acc / (float)n;
            }
        }
    }
    return 0;
}

/*
 * Trilinear resample with a 3x4 matrix M (row-major, output index -> input
 * index in ZYX voxel units): out[o] = vol(M[:, :3] @ o + M[:, 3]).
 */
int oracle_affine(const float *vol, float *out, int iz, int iy, int ix, int oz, int oy, int ox,
                  const double *M, float cval, int a_begin, int a_end) {
    if (iz <= 0 || iy <= 0 || ix <= 0 || oz < 0 || oy < 0 || ox < 0) return 1;
    const int dims[3] = {iz, iy, ix};
    const size_t strides[3] = {(size_t)iy * ix, (size_t)ix, 1};
    if (a_begin < 0) a_begin = 0;
    if (a_end > oz) a_end = oz;
    for (int a = a_begin; a < a_end; ++a) {
        for (int b = 0; b < oy; ++b) {
            float *row = out + ((size_t)a * oy + b) * (size_t)ox;
            for (int c = 0; c < ox; ++c) {
                int i0[3], i1[3];
                double w[3];
                int inside = 1;
                for (int h = 0; h < 3; ++h) {
                    double t = M[4 * h + 3] + (double)a * M[4 * h + 0];
                    t = t + (double)b * M[4 * h + 1];
                    t = t + (double)c * M[4 * h + 2];
                    if (t < 0.0 || t > (double)(dims[h] - 1)) {
                        inside = 0;
                        break;
                    }
                    const double f = floor(t);
                    w[h] = t - f;
                    i0[h] = (int)f;
                    i1[h] = i0[h] + 1 > dims[h] - 1 ? dims[h] - 1 : i0[h] + 1;
                }
                if (!inside) {
                    row[c] = cval;
                    continue;
                }
                double acc = 0.0;
                for (int dz = 0; dz < 2; ++dz)
                    for (int dy = 0; dy < 2; ++dy)
                        for (int dx = 0; dx < 2; ++dx) {
                            const double wt = (dz ? w[0] : 1.0 - w[0]) * (dy ? w[1] : 1.0 - w[1]) *
                                              (dx ? w[2] : 1.0 - w[2]);
                            const float s = vol[(size_t)(dz ? i1[0] : i0[0]) * strides[0] +
                                                (size_t)(dy ? i1[1] : i0[1]) * strides[1] +
                                                (size_t)(dx ? i1[2] : i0[2])];
                            /* nan_to_num: nan -> 0, +-inf -> +-FLT_MAX */
                            const double sv = isnan(s) ? 0.0 : (isinf(s) ? (s > 0 ? FLT_MAX : -FLT_MAX) : (double)s);
                            acc += wt * sv;
                        }
                row[c] = (float)acc;
            }
        }
    }
    return 0;
}
