// T5: fused BatchNorm(+ReLU) over sparse-tensor feature matrices [N, C].
//
// MinkowskiBatchNorm is nn.BatchNorm1d on .F and MinkowskiReLU is nn.ReLU on .F
// (minsu3d/model/module/common.py:13-14,35-39).  The pair is pure
// bandwidth: statistics = one read, apply(+ReLU) = one read + one write, backward = two reads
// of (x, y, dy) and one write.  Column sums are accumulated per thread in fp32 over short row
// runs, combined across the block and across blocks in double in a fixed order, so the
// statistics are deterministic and accurate for N ~ 1e6.
#include "common.cuh"

namespace b2s {

constexpr int BN_THREADS = 256;
constexpr int BN_SLAB = 1024;  // rows per CTA

// partial[blk][2][c] (double): column sums of (p, q) where
//   mode 0: p = x,            q = x*x
//   mode 1: p = dy*mask(y),   q = dy*mask(y) * (x-mean)*rstd
__global__ void __launch_bounds__(BN_THREADS)
    bn_colsum_kernel(const float* __restrict__ x, const float* __restrict__ y,
                     const float* __restrict__ dy, const float* __restrict__ mean,
                     const float* __restrict__ rstd, int64_t n, int c, int mode, int relu,
                     double* __restrict__ partial) {
  extern __shared__ double s_acc[];  // [rpp][2][c]
  const int rpp = BN_THREADS / c > 0 ? BN_THREADS / c : 1;  // row phases per pass
  const int ch0 = threadIdx.x % c, ph = threadIdx.x / c;
  const int64_t r0 = (int64_t)blockIdx.x * BN_SLAB;
  const int64_t r1 = min(n, r0 + BN_SLAB);
  // channels handled by this thread: ch0, ch0 + BN_THREADS (only when c > BN_THREADS)
  for (int ch = ch0; ch < c; ch += BN_THREADS) {
    double dp = 0.0, dq = 0.0;
    if (ph < rpp) {
      float m = 0.f, rs = 1.f;
      if (mode == 1) {
        m = mean[ch];
        rs = rstd[ch];
      }
      float p = 0.f, q = 0.f;
      int cnt = 0;
      for (int64_t r = r0 + ph; r < r1; r += rpp) {
        int64_t i = r * c + ch;
        if (mode == 0) {
          float v = __ldg(x + i);
          p += v;
          q = fmaf(v, v, q);
        } else {
          float g = __ldg(dy + i);
          if (relu && !(__ldg(y + i) > 0.f)) g = 0.f;
          p += g;
          q = fmaf(g, (__ldg(x + i) - m) * rs, q);
        }
        if (++cnt == 64) {  // flush fp32 run into double
          dp += p; dq += q; p = 0.f; q = 0.f; cnt = 0;
        }
      }
      dp += p;
      dq += q;
      s_acc[(ph * 2 + 0) * c + ch] = dp;
      s_acc[(ph * 2 + 1) * c + ch] = dq;
    }
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += BN_THREADS) {
    double sp = 0.0, sq = 0.0;
    for (int h = 0; h < rpp; ++h) {
      sp += s_acc[(h * 2 + 0) * c + ch];
      sq += s_acc[(h * 2 + 1) * c + ch];
    }
    partial[((int64_t)blockIdx.x * 2 + 0) * c + ch] = sp;
    partial[((int64_t)blockIdx.x * 2 + 1) * c + ch] = sq;
  }
}

__global__ void __launch_bounds__(BN_THREADS)
    bn_finish_stats_kernel(const double* __restrict__ partial, int nblk, int64_t n, int c,
                           float* __restrict__ mean, float* __restrict__ var) {
  int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  double sp = 0.0, sq = 0.0;
  for (int b = 0; b < nblk; ++b) {
    sp += partial[((int64_t)b * 2 + 0) * c + ch];
    sq += partial[((int64_t)b * 2 + 1) * c + ch];
  }
  double m = sp / (double)n;
  double v = sq / (double)n - m * m;
  mean[ch] = (float)m;
  var[ch] = (float)(v > 0.0 ? v : 0.0);
}

__global__ void __launch_bounds__(BN_THREADS)
    bn_finish_grad_kernel(const double* __restrict__ partial, int nblk, int c,
                          float* __restrict__ dgamma, float* __restrict__ dbeta) {
  int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  double sp = 0.0, sq = 0.0;
  for (int b = 0; b < nblk; ++b) {
    sp += partial[((int64_t)b * 2 + 0) * c + ch];
    sq += partial[((int64_t)b * 2 + 1) * c + ch];
  }
  dbeta[ch] = (float)sp;
  dgamma[ch] = (float)sq;
}

__global__ void __launch_bounds__(256)
    bn_apply_kernel(const float4* __restrict__ x, int64_t total4, int c4, const float* __restrict__ mean,
                    const float* __restrict__ rstd, const float* __restrict__ gamma,
                    const float* __restrict__ beta, int relu, float4* __restrict__ y) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total4) return;
  int ch = (int)(t % c4) * 4;
  float4 v = __ldg(x + t);
  float o[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float sc = gamma[ch + j] * rstd[ch + j];
    float r = fmaf(o[j] - mean[ch + j], sc, beta[ch + j]);
    o[j] = (relu && !(r > 0.f)) ? 0.f : r;
  }
  y[t] = make_float4(o[0], o[1], o[2], o[3]);
}

__global__ void __launch_bounds__(256)
    bn_dx_kernel(const float4* __restrict__ x, const float4* __restrict__ y, const float4* __restrict__ dy,
                 int64_t total4, int c4, float inv_n, const float* __restrict__ mean,
                 const float* __restrict__ rstd, const float* __restrict__ gamma,
                 const float* __restrict__ dgamma, const float* __restrict__ dbeta, int relu,
                 int training, float4* __restrict__ dx) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total4) return;
  int ch = (int)(t % c4) * 4;
  float4 xv = __ldg(x + t), gv = __ldg(dy + t);
  float xs[4] = {xv.x, xv.y, xv.z, xv.w}, gs[4] = {gv.x, gv.y, gv.z, gv.w};
  if (relu) {
    float4 yv = __ldg(y + t);
    if (!(yv.x > 0.f)) gs[0] = 0.f;
    if (!(yv.y > 0.f)) gs[1] = 0.f;
    if (!(yv.z > 0.f)) gs[2] = 0.f;
    if (!(yv.w > 0.f)) gs[3] = 0.f;
  }
  float o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float sc = gamma[ch + j] * rstd[ch + j];
    if (training) {
      float xhat = (xs[j] - mean[ch + j]) * rstd[ch + j];
      o[j] = sc * (gs[j] - dbeta[ch + j] * inv_n - xhat * dgamma[ch + j] * inv_n);
    } else {
      o[j] = sc * gs[j];
    }
  }
  dx[t] = make_float4(o[0], o[1], o[2], o[3]);
}

}  // namespace b2s

using namespace b2s;

extern "C" {

size_t b2s_bn_ws_bytes(int64_t n, int32_t c) {
  int64_t nblk = cdiv(n > 0 ? n : 1, BN_SLAB);
  return align_up((size_t)nblk * 2 * c * 8) + 1024;
}

static int bn_check(int64_t n, int32_t c) {
  if (n < 0 || c < 4 || (c & 3) != 0 || c > 1024) {
    set_error("batchnorm: channel count must be a multiple of 4 in [4, 1024]");
    return B2S_E_INVALID;
  }
  return B2S_OK;
}

static size_t bn_smem(int c) {
  int rpp = BN_THREADS / c > 0 ? BN_THREADS / c : 1;
  return (size_t)rpp * 2 * c * 8;
}

int b2s_bn_stats(const float* x, int64_t n, int32_t c, float eps, float* mean, float* var_biased,
                 void* ws, size_t ws_bytes, b2s_stream_t stream) {
  (void)eps;
  int rc = bn_check(n, c);
  if (rc) return rc;
  if (n == 0) return B2S_OK;
  int nblk = (int)cdiv(n, BN_SLAB);
  if (ws_bytes < (size_t)nblk * 2 * c * 8) {
    set_error("bn_stats: workspace too small");
    return B2S_E_WORKSPACE;
  }
  double* partial = (double*)ws;
  bn_colsum_kernel<<<nblk, BN_THREADS, bn_smem(c), stream>>>(x, nullptr, nullptr, nullptr, nullptr, n, c, 0, 0, partial);
  bn_finish_stats_kernel<<<(unsigned)cdiv(c, BN_THREADS), BN_THREADS, 0, stream>>>(partial, nblk, n, c, mean, var_biased);
  return check_launch("bn_stats");
}

int b2s_bn_apply(const float* x, int64_t n, int32_t c, const float* mean, const float* rstd,
                 const float* gamma, const float* beta, int32_t relu, float* y, b2s_stream_t stream) {
  int rc = bn_check(n, c);
  if (rc) return rc;
  if (n == 0) return B2S_OK;
  int64_t total4 = n * (c / 4);
  bn_apply_kernel<<<(unsigned)cdiv(total4, 256), 256, 0, stream>>>((const float4*)x, total4, c / 4, mean, rstd,
                                                                   gamma, beta, relu, (float4*)y);
  return check_launch("bn_apply");
}

int b2s_bn_backward(const float* x, const float* y, const float* dy, int64_t n, int32_t c,
                    const float* mean, const float* rstd, const float* gamma, int32_t relu,
                    int32_t training, float* dx, float* dgamma, float* dbeta, void* ws,
                    size_t ws_bytes, b2s_stream_t stream) {
  int rc = bn_check(n, c);
  if (rc) return rc;
  if (n == 0) {
    cudaMemsetAsync(dgamma, 0, (size_t)c * 4, stream);
    cudaMemsetAsync(dbeta, 0, (size_t)c * 4, stream);
    return check_launch("bn_backward(empty)");
  }
  int nblk = (int)cdiv(n, BN_SLAB);
  if (ws_bytes < (size_t)nblk * 2 * c * 8) {
    set_error("bn_backward: workspace too small");
    return B2S_E_WORKSPACE;
  }
  double* partial = (double*)ws;
  bn_colsum_kernel<<<nblk, BN_THREADS, bn_smem(c), stream>>>(x, y, dy, mean, rstd, n, c, 1, relu, partial);
  bn_finish_grad_kernel<<<(unsigned)cdiv(c, BN_THREADS), BN_THREADS, 0, stream>>>(partial, nblk, c, dgamma, dbeta);
  int64_t total4 = n * (c / 4);
  bn_dx_kernel<<<(unsigned)cdiv(total4, 256), 256, 0, stream>>>(
      (const float4*)x, (const float4*)y, (const float4*)dy, total4, c / 4, 1.0f / (float)n, mean, rstd,
      gamma, dgamma, dbeta, relu, training, (float4*)dx);
  return check_launch("bn_backward");
}

}  // extern "C"
