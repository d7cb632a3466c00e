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
constexpr int BN_SLAB = 512;  // rows per CTA

// partial[blk][2][c] (double): column sums of (p, q) where
//   mode 0: p = x,            q = x*x
//   mode 1: p = dy*mask(y),   q = dy*mask(y) * (x-mean)*rstd
// Thread t owns the float4 column (t % c4) and the row phase (t / c4); four independent 16-byte
// loads are in flight per thread and array, fp32 partials are flushed into double every 32 rows.
struct BnFinish {
  // mode 0 (statistics): mean / var / rstd (+ running stats); mode 1 (gradients): dgamma / dbeta
  float eps, momentum;
  float *running_mean, *running_var, *out_a, *out_b, *out_c;  // stats: mean, var, rstd ; grads: dbeta, dgamma, -
  int32_t* counter;  // zero between launches: the last block to finish does the second stage and resets it
};

__global__ void __launch_bounds__(BN_THREADS)
    bn_colsum_kernel(const float* __restrict__ x, const float* __restrict__ y,
                     const float* __restrict__ dy, const float* __restrict__ mean,
                     const float* __restrict__ rstd, int64_t n, int c, int mode, int relu,
                     double* __restrict__ partial, BnFinish fin) {
  extern __shared__ double s_acc[];  // [rpp][2][c]
  const int c4 = c >> 2;
  const int rpp = BN_THREADS / c4 > 0 ? BN_THREADS / c4 : 1;  // row phases per pass
  const int64_t r0 = (int64_t)blockIdx.x * BN_SLAB;
  const int64_t r1 = min(n, r0 + BN_SLAB);
  for (int vc = threadIdx.x % c4, ph = threadIdx.x / c4; vc < c4 && ph < rpp; vc += BN_THREADS) {
    double dp[4] = {0, 0, 0, 0}, dq[4] = {0, 0, 0, 0};
    float p[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
    float4 m4 = make_float4(0, 0, 0, 0), rs4 = make_float4(1, 1, 1, 1);
    if (mode == 1) {
      m4 = __ldg((const float4*)mean + vc);
      rs4 = __ldg((const float4*)rstd + vc);
    }
    const float m[4] = {m4.x, m4.y, m4.z, m4.w}, rs[4] = {rs4.x, rs4.y, rs4.z, rs4.w};
    int cnt = 0;
    for (int64_t r = r0 + ph; r < r1; r += 4 * (int64_t)rpp) {
      float4 xv[4], gv[4], yv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        int64_t rr = r + (int64_t)u * rpp;
        bool ok = rr < r1;
        int64_t i = rr * c4 + vc;
        xv[u] = ok ? __ldg((const float4*)x + i) : make_float4(0, 0, 0, 0);
        if (mode == 1) {
          gv[u] = ok ? __ldg((const float4*)dy + i) : make_float4(0, 0, 0, 0);
          yv[u] = (ok && relu) ? __ldg((const float4*)y + i) : make_float4(1, 1, 1, 1);
          if (!ok) xv[u] = m4;  // (x - mean) = 0 for padding rows
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float xs[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w};
        if (mode == 0) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            p[j] += xs[j];
            q[j] = fmaf(xs[j], xs[j], q[j]);
          }
        } else {
          float gs[4] = {gv[u].x, gv[u].y, gv[u].z, gv[u].w};
          float ys[4] = {yv[u].x, yv[u].y, yv[u].z, yv[u].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float g = (relu && !(ys[j] > 0.f)) ? 0.f : gs[j];
            p[j] += g;
            q[j] = fmaf(g, (xs[j] - m[j]) * rs[j], q[j]);
          }
        }
      }
      if (++cnt == 8) {  // 32 rows accumulated in fp32
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          dp[j] += p[j]; dq[j] += q[j]; p[j] = 0.f; q[j] = 0.f;
        }
        cnt = 0;
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      s_acc[(ph * 2 + 0) * c + vc * 4 + j] = dp[j] + p[j];
      s_acc[(ph * 2 + 1) * c + vc * 4 + j] = dq[j] + q[j];
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
  // ---- second stage by the last block to finish (fixed summation order -> deterministic) ----
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(fin.counter, 1) == (int)gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int nblk = gridDim.x;
  // thread (g, ch): g-th slice of the blocks for channel ch; slices combined through shared memory
  const int groups = BN_THREADS / c > 0 ? BN_THREADS / c : 1;  // >= 1 row groups when c <= 256
  for (int ch0 = 0; ch0 < c; ch0 += BN_THREADS) {
    const int ch = ch0 + threadIdx.x % (c < BN_THREADS ? c : BN_THREADS);
    const int g = threadIdx.x / (c < BN_THREADS ? c : BN_THREADS);
    double a = 0.0, b = 0.0;
    if (ch < c && g < groups)
      for (int blk = g; blk < nblk; blk += groups) {
        a += __ldcg(partial + ((int64_t)blk * 2 + 0) * c + ch);
        b += __ldcg(partial + ((int64_t)blk * 2 + 1) * c + ch);
      }
    __syncthreads();
    if (ch < c && g < groups) {
      s_acc[(g * 2 + 0) * c + ch] = a;
      s_acc[(g * 2 + 1) * c + ch] = b;
    }
    __syncthreads();
    if (g == 0 && ch < c) {
      double sp = 0.0, sq = 0.0;
      for (int h = 0; h < groups; ++h) {
        sp += s_acc[(h * 2 + 0) * c + ch];
        sq += s_acc[(h * 2 + 1) * c + ch];
      }
      if (mode == 0) {
        double m = sp / (double)n;
        double v = sq / (double)n - m * m;
        v = v > 0.0 ? v : 0.0;
        fin.out_a[ch] = (float)m;
        if (fin.out_b) fin.out_b[ch] = (float)v;
        if (fin.out_c) fin.out_c[ch] = (float)(1.0 / sqrt(v + (double)fin.eps));
        if (fin.running_mean) {
          double unbiased = n > 1 ? v * ((double)n / (double)(n - 1)) : v;
          double mo = (double)fin.momentum;
          fin.running_mean[ch] = (float)((1.0 - mo) * (double)fin.running_mean[ch] + mo * m);
          fin.running_var[ch] = (float)((1.0 - mo) * (double)fin.running_var[ch] + mo * unbiased);
        }
      } else {
        fin.out_a[ch] = (float)sp;  // dbeta
        fin.out_b[ch] = (float)sq;  // dgamma
      }
    }
  }
  if (threadIdx.x == 0) *fin.counter = 0;
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
  int c4 = c / 4;
  int rpp = BN_THREADS / c4 > 0 ? BN_THREADS / c4 : 1;  // first stage: row phases
  return (size_t)rpp * 2 * c * 8;                        // (second stage needs <= BN_THREADS/c groups: smaller)
}

int b2s_bn_stats(const float* x, int64_t n, int32_t c, float eps, float momentum, float* running_mean,
                 float* running_var, float* mean, float* var_biased, float* rstd, int32_t* counter, void* ws,
                 size_t ws_bytes, b2s_stream_t stream) {
  int rc = bn_check(n, c);
  if (rc) return rc;
  if (n == 0) return B2S_OK;
  int nblk = (int)cdiv(n, BN_SLAB);
  if (ws_bytes < (size_t)nblk * 2 * c * 8) {
    set_error("bn_stats: workspace too small");
    return B2S_E_WORKSPACE;
  }
  double* partial = (double*)ws;
  BnFinish fin{eps, momentum, running_mean, running_var, mean, var_biased, rstd, counter};
  bn_colsum_kernel<<<nblk, BN_THREADS, bn_smem(c), stream>>>(x, nullptr, nullptr, nullptr, nullptr, n, c, 0, 0, partial, fin);
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
                    int32_t training, float* dx, float* dgamma, float* dbeta, int32_t* counter, void* ws,
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
  BnFinish fin{0.f, 0.f, nullptr, nullptr, dbeta, dgamma, nullptr, counter};
  bn_colsum_kernel<<<nblk, BN_THREADS, bn_smem(c), stream>>>(x, y, dy, mean, rstd, n, c, 1, relu, partial, fin);
  int64_t total4 = n * (c / 4);
  bn_dx_kernel<<<(unsigned)cdiv(total4, 256), 256, 0, stream>>>(
      (const float4*)x, (const float4*)y, (const float4*)dy, total4, c / 4, 1.0f / (float)n, mean, rstd,
      gamma, dgamma, dbeta, relu, training, (float4*)dx);
  return check_launch("bn_backward");
}

}  // extern "C"
