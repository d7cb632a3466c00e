// SURVEY.md 8(f) #3 -- the train-split sample pipeline of the reference's loader on the GPU
// (minsu3d/data/dataset/general_dataset.py:80-165, minsu3d/util/transform.py:65-98): affine augmentation, elastic
// distortion (separable 3-tap blur of three noise volumes, trilinear interpolation at every point, in double
// precision like scipy's RegularGridInterpolator), crop test.  The random draws (augmentation matrix, colour jitter,
// noise volumes, crop offsets) are INPUTS, so the result can be compared with the numpy restatement on shared draws.
// HBM-bound elementwise / gather work: one thread per point (or voxel of the noise grid), coalesced.
#include "common.cuh"

namespace b2s {

// xyz' = xyz @ M (float32, like np.matmul on float32 operands); rgb' = rgb + jitter
__global__ void __launch_bounds__(256)
    aug_affine_kernel(const float* __restrict__ xyz, const float* __restrict__ rgb, int64_t n, float m00, float m01,
                      float m02, float m10, float m11, float m12, float m20, float m21, float m22, float j0, float j1,
                      float j2, float* __restrict__ out_xyz, float* __restrict__ out_rgb) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
  // separate multiplies and adds in the order of a row-times-matrix dot product (no FMA contraction)
  out_xyz[3 * i] = __fadd_rn(__fadd_rn(__fmul_rn(x, m00), __fmul_rn(y, m10)), __fmul_rn(z, m20));
  out_xyz[3 * i + 1] = __fadd_rn(__fadd_rn(__fmul_rn(x, m01), __fmul_rn(y, m11)), __fmul_rn(z, m21));
  out_xyz[3 * i + 2] = __fadd_rn(__fadd_rn(__fmul_rn(x, m02), __fmul_rn(y, m12)), __fmul_rn(z, m22));
  if (rgb != nullptr) {
    out_rgb[3 * i] = rgb[3 * i] + j0;
    out_rgb[3 * i + 1] = rgb[3 * i + 1] + j1;
    out_rgb[3 * i + 2] = rgb[3 * i + 2] + j2;
  }
}

// one pass of scipy.ndimage.convolve(n, ones(3)/3 along `axis`, mode="constant", cval=0) over [3][b0][b1][b2]
__global__ void __launch_bounds__(256)
    box3_axis_kernel(const float* __restrict__ in, float* __restrict__ out, int b0, int b1, int b2, int axis,
                     int64_t total) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int i2 = (int)(e % b2);
  const int i1 = (int)((e / b2) % b1);
  const int i0 = (int)((e / ((int64_t)b2 * b1)) % b0);
  const int64_t stride = axis == 0 ? (int64_t)b1 * b2 : (axis == 1 ? b2 : 1);
  const int pos = axis == 0 ? i0 : (axis == 1 ? i1 : i2);
  const int len = axis == 0 ? b0 : (axis == 1 ? b1 : b2);
  const float w = 1.0f / 3.0f;
  // scipy accumulates the (flipped) kernel taps in double; the taps are equal, so the order is pos+1, pos, pos-1
  double acc = 0.0;
  if (pos + 1 < len) acc += (double)in[e + stride] * (double)w;
  acc += (double)in[e] * (double)w;
  if (pos - 1 >= 0) acc += (double)in[e - stride] * (double)w;
  out[e] = (float)acc;
}

// x += mag * trilinear(noise_c, x) for c = 0, 1, 2; grid axes ax_d = linspace(-(b_d - 1) gran, (b_d - 1) gran, b_d);
// points outside the grid get 0 (bounds_error = 0, fill_value = 0).  Double precision, evaluation order of scipy's
// linear RegularGridInterpolator (product of the per-axis weights, corners in lexicographic order).
__global__ void __launch_bounds__(256)
    elastic_apply_kernel(double* __restrict__ x, const float* __restrict__ noise, int64_t n, int b0, int b1, int b2,
                         double gran, double mag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double p[3] = {x[3 * i], x[3 * i + 1], x[3 * i + 2]};
  const int b[3] = {b0, b1, b2};
  int lo[3];
  double t[3];
  bool inside = true;
  for (int d = 0; d < 3; ++d) {
    const double a0 = -(double)(b[d] - 1) * gran, step = 2.0 * gran;  // linspace step = 2 (b-1) gran / (b-1)
    const double a1 = (double)(b[d] - 1) * gran;
    if (!(p[d] >= a0 && p[d] <= a1)) inside = false;
    int k = (int)floor((p[d] - a0) / step);
    k = max(0, min(k, b[d] - 2));
    lo[d] = k;
    const double g0 = a0 + step * k;
    t[d] = (p[d] - g0) / step;
  }
  double add[3] = {0.0, 0.0, 0.0};
  if (inside) {
    const int64_t vol = (int64_t)b0 * b1 * b2;
    for (int c = 0; c < 3; ++c) {
      const float* nz = noise + c * vol;
      double acc = 0.0;
      for (int c0 = 0; c0 < 2; ++c0)
        for (int c1 = 0; c1 < 2; ++c1)
          for (int c2 = 0; c2 < 2; ++c2) {
            const double w = (c0 ? t[0] : 1.0 - t[0]) * (c1 ? t[1] : 1.0 - t[1]) * (c2 ? t[2] : 1.0 - t[2]);
            acc += (double)nz[((int64_t)(lo[0] + c0) * b1 + (lo[1] + c1)) * b2 + (lo[2] + c2)] * w;
          }
      add[c] = acc;
    }
  }
  x[3 * i] = p[0] + add[0] * mag;
  x[3 * i + 1] = p[1] + add[1] * mag;
  x[3 * i + 2] = p[2] + add[2] * mag;
}

// transform.py:96: valid = (pc + offset).min(1) >= 0 and all(pc + offset < range); out = pc + offset
__global__ void __launch_bounds__(256)
    crop_test_kernel(const double* __restrict__ pc, int64_t n, double o0, double o1, double o2, double r0, double r1,
                     double r2, double* __restrict__ out, uint8_t* __restrict__ valid, int32_t* __restrict__ count) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool ok = false;
  if (i < n) {
    const double a = pc[3 * i] + o0, b = pc[3 * i + 1] + o1, c = pc[3 * i + 2] + o2;
    out[3 * i] = a;
    out[3 * i + 1] = b;
    out[3 * i + 2] = c;
    ok = fmin(a, fmin(b, c)) >= 0.0 && a < r0 && b < r1 && c < r2;
    valid[i] = ok ? 1 : 0;
  }
  const unsigned ballot = __ballot_sync(0xffffffffu, ok);
  if ((threadIdx.x & 31) == 0 && ballot) atomicAdd(count, __popc(ballot));  // warp-aggregated
}

}  // namespace b2s

using namespace b2s;

extern "C" {

int b2s_aug_affine(const float* xyz, const float* rgb, int64_t n, const float* m9_host, const float* jitter3_host,
                   float* out_xyz, float* out_rgb, b2s_stream_t stream) {
  if (n < 0 || !xyz || !out_xyz || !m9_host || (rgb && (!out_rgb || !jitter3_host))) {
    set_error("aug_affine: invalid argument");
    return B2S_E_INVALID;
  }
  if (n == 0) return B2S_OK;
  const float* m = m9_host;
  const float j0 = rgb ? jitter3_host[0] : 0.f, j1 = rgb ? jitter3_host[1] : 0.f, j2 = rgb ? jitter3_host[2] : 0.f;
  aug_affine_kernel<<<(unsigned)cdiv(n, 256), 256, 0, stream>>>(xyz, rgb, n, m[0], m[1], m[2], m[3], m[4], m[5], m[6],
                                                                m[7], m[8], j0, j1, j2, out_xyz, out_rgb);
  return check_launch("aug_affine");
}

// noise [3][b0][b1][b2] float32 is blurred in place (six separable passes, tmp = same size scratch)
int b2s_elastic_blur(float* noise, float* tmp, int32_t b0, int32_t b1, int32_t b2, b2s_stream_t stream) {
  if (!noise || !tmp || b0 < 1 || b1 < 1 || b2 < 1) {
    set_error("elastic_blur: invalid argument");
    return B2S_E_INVALID;
  }
  const int64_t total = (int64_t)3 * b0 * b1 * b2;
  float* a = noise;
  float* b = tmp;
  for (int pass = 0; pass < 6; ++pass) {
    box3_axis_kernel<<<(unsigned)cdiv(total, 256), 256, 0, stream>>>(a, b, b0, b1, b2, pass % 3, total);
    float* t = a;
    a = b;
    b = t;
  }
  return check_launch("elastic_blur");  // six passes: the result is back in `noise`
}

int b2s_elastic_apply(double* x, const float* noise, int64_t n, int32_t b0, int32_t b1, int32_t b2, double gran,
                      double mag, b2s_stream_t stream) {
  if (n < 0 || !x || !noise || b0 < 2 || b1 < 2 || b2 < 2) {
    set_error("elastic_apply: invalid argument");
    return B2S_E_INVALID;
  }
  if (n == 0) return B2S_OK;
  elastic_apply_kernel<<<(unsigned)cdiv(n, 256), 256, 0, stream>>>(x, noise, n, b0, b1, b2, gran, mag);
  return check_launch("elastic_apply");
}

int b2s_crop_test(const double* pc, int64_t n, const double* offset3_host, const double* range3_host, double* out,
                  uint8_t* valid, int32_t* d_count, b2s_stream_t stream) {
  if (n < 0 || !pc || !offset3_host || !range3_host || !out || !valid || !d_count) {
    set_error("crop_test: invalid argument");
    return B2S_E_INVALID;
  }
  cudaMemsetAsync(d_count, 0, 4, stream);
  if (n == 0) return check_launch("crop_test(empty)");
  crop_test_kernel<<<(unsigned)cdiv(n, 256), 256, 0, stream>>>(pc, n, offset3_host[0], offset3_host[1], offset3_host[2],
                                                               range3_host[0], range3_host[1], range3_host[2], out, valid,
                                                               d_count);
  return check_launch("crop_test");
}

}  // extern "C"
