// T5: fused BatchNorm(+ReLU) over sparse-tensor feature matrices [N, C].
//
// MinkowskiBatchNorm is nn.BatchNorm1d on .F and MinkowskiReLU is nn.ReLU on .F
// (minsu3d/model/module/common.py:13-14,35-39).  The pair is pure
// bandwidth: statistics = one read, apply(+ReLU) = one read + one write, backward = two reads
// of (x, y, dy) and one write.  Column sums are accumulated per thread in fp32 over short row
// runs, combined across the block and across blocks in double in a fixed order, so the
// statistics are deterministic and accurate for N ~ 1e6.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace b2s {

constexpr int BN_THREADS = 256;

// partial[blk][2][c] (double): column sums of (p, q) where
//   MODE 0: p = x,            q = x*x
//   MODE 1: p = dy*mask(y),   q = dy*mask(y) * (x-mean)*rstd
// Thread t owns the float4 column (t % c4) and the row phase (t / c4); U independent 16-byte loads per array are in
// flight per thread (8 for the statistics, 4 for the gradients), fp32 partials are flushed into double every 32
// rows.  The kernel is latency-bound for the small tensors of the deep U-Net levels, so the launch geometry keeps
// every thread at a handful of load batches: CTA b owns `chunk` consecutive rows (a multiple of the rows one batch
// covers) and the grid is capped by bn_grid() so that the second stage (one CTA) stays short too.
struct BnFinish {
  // MODE 0 (statistics): mean / var / rstd (+ running stats); MODE 1 (gradients): dgamma / dbeta
  float eps, momentum;
  float *running_mean, *running_var, *out_a, *out_b, *out_c;  // stats: mean, var, rstd ; grads: dbeta, dgamma, -
  int32_t* counter;  // zero between launches: the last block to finish does the second stage and resets it
};

template <int MODE>
__global__ void __launch_bounds__(BN_THREADS)
    bn_colsum_kernel(const float* __restrict__ x, const float* __restrict__ y,
                     const float* __restrict__ dy, const float* __restrict__ mean,
                     const float* __restrict__ rstd, int64_t n, int c, int chunk, int relu,
                     double* __restrict__ partial, BnFinish fin) {
  extern __shared__ double s_acc[];  // [rpp][2][c]
  constexpr int U = MODE == 0 ? 8 : 4;
  const int c4 = c >> 2;
  const int rpp = BN_THREADS / c4 > 0 ? BN_THREADS / c4 : 1;  // row phases per pass
  const int64_t r0 = (int64_t)blockIdx.x * chunk;
  const int64_t r1 = min(n, r0 + chunk);
  for (int vc = threadIdx.x % c4, ph = threadIdx.x / c4; vc < c4 && ph < rpp; vc += BN_THREADS) {
    double dp[4] = {0, 0, 0, 0}, dq[4] = {0, 0, 0, 0};
    float p[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
    float4 m4 = make_float4(0, 0, 0, 0), rs4 = make_float4(1, 1, 1, 1);
    if (MODE == 1) {
      m4 = __ldg((const float4*)mean + vc);
      rs4 = __ldg((const float4*)rstd + vc);
    }
    const float m[4] = {m4.x, m4.y, m4.z, m4.w}, rs[4] = {rs4.x, rs4.y, rs4.z, rs4.w};
    int cnt = 0;
    for (int64_t r = r0 + ph; r < r1; r += U * (int64_t)rpp) {
      float4 xv[U], gv[MODE == 1 ? U : 1], yv[MODE == 1 ? U : 1];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        int64_t rr = r + (int64_t)u * rpp;
        bool ok = rr < r1;
        int64_t i = rr * c4 + vc;
        xv[u] = ok ? __ldg((const float4*)x + i) : make_float4(0, 0, 0, 0);
        if (MODE == 1) {
          gv[u] = ok ? __ldg((const float4*)dy + i) : make_float4(0, 0, 0, 0);
          yv[u] = (ok && relu) ? __ldg((const float4*)y + i) : make_float4(1, 1, 1, 1);
          if (!ok) xv[u] = m4;  // (x - mean) = 0 for padding rows
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float xs[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w};
        if (MODE == 0) {
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
      cnt += U;
      if (cnt == 32) {  // 32 rows accumulated in fp32
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
    if (ch < c && g < groups) {
      // sixteen independent loads per array in flight (a one-by-one loop exposes an L2 round trip per block);
      // the summation order stays fixed
      int blk = g;
      for (; blk + 15 * groups < nblk; blk += 16 * groups) {
        double va[16], vb[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          va[u] = __ldcg(partial + ((int64_t)(blk + u * groups) * 2 + 0) * c + ch);
          vb[u] = __ldcg(partial + ((int64_t)(blk + u * groups) * 2 + 1) * c + ch);
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          a += va[u];
          b += vb[u];
        }
      }
      {
        double va[16], vb[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int bb = blk + u * groups;
          va[u] = bb < nblk ? __ldcg(partial + ((int64_t)bb * 2 + 0) * c + ch) : 0.0;
          vb[u] = bb < nblk ? __ldcg(partial + ((int64_t)bb * 2 + 1) * c + ch) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          a += va[u];
          b += vb[u];
        }
      }
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
      if (MODE == 0) {
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
                 int training, const float4* __restrict__ add_src, float4* __restrict__ dx) {
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
  if (add_src != nullptr) {  // gradient arriving over the residual shortcut (common.py:48), folded into this pass
    const float4 av = __ldg(add_src + t);
    o[0] += av.x;
    o[1] += av.y;
    o[2] += av.z;
    o[3] += av.w;
  }
  dx[t] = make_float4(o[0], o[1], o[2], o[3]);
}


// ------------------------------------------------------------------------------------------------------------------
// One-launch forms (cooperative grid, co-resident CTAs):
//   MODE 0:  batch statistics -> device-wide barrier -> y = relu(bn(x))          (was bn_colsum<0> + bn_apply)
//   MODE 1:  dgamma / dbeta   -> device-wide barrier -> dx (+ shortcut gradient)  (was bn_colsum<1> + bn_dx)
// The second kernel of each pair could not start before the first had drained and its one-CTA second stage had run;
// here every CTA sums the per-CTA partials itself after the barrier (same fixed order in every CTA -> the same bits
// everywhere, deterministic) and then normalises the rows it has just read, which are still in L1 / L2.  160 fewer
// launches per PointGroup step.  Arithmetic identical to the two-kernel path (tests compare them with torch.equal).
// ------------------------------------------------------------------------------------------------------------------
struct BnFusedArgs {
  const float *x, *y_in, *dy, *mean_in, *rstd_in, *gamma, *beta, *add_src;
  float *out;                      // MODE 0: y ; MODE 1: dx
  float *stat_a, *stat_b;          // MODE 0: mean, rstd ; MODE 1: dbeta, dgamma
  float *running_mean, *running_var;
  double* partial;
  int32_t* counter;                // one int, zero between launches
  int64_t n;
  int c, chunk, relu, training;
  float eps, momentum;
};

// one-shot device-wide barrier on a single int that is zero before and after: arrivals count in the low half, departures
// in the high half, the last CTA to leave resets the word
__device__ __forceinline__ void bn_grid_barrier(int32_t* counter, int nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1);
    while ((*((volatile int32_t*)counter) & 0xFFFF) < nblocks) {
    }
    __threadfence();
    const int old = atomicAdd(counter, 1 << 16);
    if ((old >> 16) == nblocks - 1) *((volatile int32_t*)counter) = 0;
  }
  __syncthreads();
}

template <int MODE>
__global__ void __launch_bounds__(BN_THREADS) bn_fused_kernel(const BnFusedArgs a) {
  extern __shared__ double s_acc[];  // [rpp][2][c] doubles, then [2][c] floats of the final statistics behind them
  constexpr int U = MODE == 0 ? 8 : 4;
  const int c = a.c, c4 = c >> 2;
  const int rpp = BN_THREADS / c4 > 0 ? BN_THREADS / c4 : 1;
  float* s_stat = (float*)(s_acc + (size_t)rpp * 2 * c);
  const int64_t r0 = (int64_t)blockIdx.x * a.chunk;
  const int64_t r1 = min(a.n, r0 + a.chunk);
  // ---- phase 1: this CTA's column sums (same code path and summation order as bn_colsum_kernel) ----------------
  for (int vc = threadIdx.x % c4, ph = threadIdx.x / c4; vc < c4 && ph < rpp; vc += BN_THREADS) {
    double dp[4] = {0, 0, 0, 0}, dq[4] = {0, 0, 0, 0};
    float p[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
    float4 m4 = make_float4(0, 0, 0, 0), rs4 = make_float4(1, 1, 1, 1);
    if (MODE == 1) {
      m4 = __ldg((const float4*)a.mean_in + vc);
      rs4 = __ldg((const float4*)a.rstd_in + vc);
    }
    const float m[4] = {m4.x, m4.y, m4.z, m4.w}, rs[4] = {rs4.x, rs4.y, rs4.z, rs4.w};
    int cnt = 0;
    for (int64_t r = r0 + ph; r < r1; r += U * (int64_t)rpp) {
      float4 xv[U], gv[MODE == 1 ? U : 1], yv[MODE == 1 ? U : 1];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        int64_t rr = r + (int64_t)u * rpp;
        bool ok = rr < r1;
        int64_t i = rr * c4 + vc;
        xv[u] = ok ? __ldg((const float4*)a.x + i) : make_float4(0, 0, 0, 0);
        if (MODE == 1) {
          gv[u] = ok ? __ldg((const float4*)a.dy + i) : make_float4(0, 0, 0, 0);
          yv[u] = (ok && a.relu) ? __ldg((const float4*)a.y_in + i) : make_float4(1, 1, 1, 1);
          if (!ok) xv[u] = m4;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float xs[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w};
        if (MODE == 0) {
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
            float g = (a.relu && !(ys[j] > 0.f)) ? 0.f : gs[j];
            p[j] += g;
            q[j] = fmaf(g, (xs[j] - m[j]) * rs[j], q[j]);
          }
        }
      }
      cnt += U;
      if (cnt == 32) {
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
    a.partial[((int64_t)blockIdx.x * 2 + 0) * c + ch] = sp;
    a.partial[((int64_t)blockIdx.x * 2 + 1) * c + ch] = sq;
  }
  bn_grid_barrier(a.counter, gridDim.x);
  // ---- every CTA: totals over the CTAs in a fixed order (thread (g, ch): g-th slice of the CTAs) ----------------
  const int nblk = gridDim.x;
  const int groups = BN_THREADS / c > 0 ? BN_THREADS / c : 1;
  for (int ch0 = 0; ch0 < c; ch0 += BN_THREADS) {
    const int ch = ch0 + threadIdx.x % (c < BN_THREADS ? c : BN_THREADS);
    const int g = threadIdx.x / (c < BN_THREADS ? c : BN_THREADS);
    double sa = 0.0, sb = 0.0;
    if (ch < c && g < groups) {
      int blk = g;
      for (; blk + 15 * groups < nblk; blk += 16 * groups) {
        double va[16], vb[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          va[u] = __ldcg(a.partial + ((int64_t)(blk + u * groups) * 2 + 0) * c + ch);
          vb[u] = __ldcg(a.partial + ((int64_t)(blk + u * groups) * 2 + 1) * c + ch);
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          sa += va[u];
          sb += vb[u];
        }
      }
      {
        double va[16], vb[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int bb = blk + u * groups;
          va[u] = bb < nblk ? __ldcg(a.partial + ((int64_t)bb * 2 + 0) * c + ch) : 0.0;
          vb[u] = bb < nblk ? __ldcg(a.partial + ((int64_t)bb * 2 + 1) * c + ch) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          sa += va[u];
          sb += vb[u];
        }
      }
    }
    __syncthreads();
    if (ch < c && g < groups) {
      s_acc[(g * 2 + 0) * c + ch] = sa;
      s_acc[(g * 2 + 1) * c + ch] = sb;
    }
    __syncthreads();
    if (g == 0 && ch < c) {
      double sp = 0.0, sq = 0.0;
      for (int h = 0; h < groups; ++h) {
        sp += s_acc[(h * 2 + 0) * c + ch];
        sq += s_acc[(h * 2 + 1) * c + ch];
      }
      if (MODE == 0) {
        double mm = sp / (double)a.n;
        double v = sq / (double)a.n - mm * mm;
        v = v > 0.0 ? v : 0.0;
        const float mean_f = (float)mm, rstd_f = (float)(1.0 / sqrt(v + (double)a.eps));
        s_stat[ch] = mean_f;
        s_stat[c + ch] = rstd_f;
        if (blockIdx.x == 0) {
          a.stat_a[ch] = mean_f;
          a.stat_b[ch] = rstd_f;
          if (a.running_mean) {
            double unbiased = a.n > 1 ? v * ((double)a.n / (double)(a.n - 1)) : v;
            double mo = (double)a.momentum;
            a.running_mean[ch] = (float)((1.0 - mo) * (double)a.running_mean[ch] + mo * mm);
            a.running_var[ch] = (float)((1.0 - mo) * (double)a.running_var[ch] + mo * unbiased);
          }
        }
      } else {
        s_stat[ch] = (float)sp;      // dbeta
        s_stat[c + ch] = (float)sq;  // dgamma
        if (blockIdx.x == 0) {
          a.stat_a[ch] = (float)sp;
          a.stat_b[ch] = (float)sq;
        }
      }
    }
  }
  __syncthreads();
  // ---- phase 2: normalise / differentiate this CTA's rows ---------------------------------------------------------
  const int64_t e0 = r0 * c4, e1 = r1 * c4;
  const float inv_n = 1.0f / (float)a.n;
#pragma unroll 4
  for (int64_t t = e0 + threadIdx.x; t < e1; t += BN_THREADS) {
    const int ch = (int)(t % c4) * 4;
    const float4 xv = __ldg((const float4*)a.x + t);
    float o[4];
    if (MODE == 0) {
      o[0] = xv.x, o[1] = xv.y, o[2] = xv.z, o[3] = xv.w;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float sc = a.gamma[ch + j] * s_stat[c + ch + j];
        float r = fmaf(o[j] - s_stat[ch + j], sc, a.beta[ch + j]);
        o[j] = (a.relu && !(r > 0.f)) ? 0.f : r;
      }
    } else {
      const float4 gv = __ldg((const float4*)a.dy + t);
      float xs[4] = {xv.x, xv.y, xv.z, xv.w}, gs[4] = {gv.x, gv.y, gv.z, gv.w};
      if (a.relu) {
        const float4 yv = __ldg((const float4*)a.y_in + t);
        if (!(yv.x > 0.f)) gs[0] = 0.f;
        if (!(yv.y > 0.f)) gs[1] = 0.f;
        if (!(yv.z > 0.f)) gs[2] = 0.f;
        if (!(yv.w > 0.f)) gs[3] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float sc = a.gamma[ch + j] * a.rstd_in[ch + j];
        if (a.training) {
          float xhat = (xs[j] - a.mean_in[ch + j]) * a.rstd_in[ch + j];
          o[j] = sc * (gs[j] - s_stat[ch + j] * inv_n - xhat * s_stat[c + ch + j] * inv_n);
        } else {
          o[j] = sc * gs[j];
        }
      }
      if (a.add_src != nullptr) {
        const float4 av = __ldg((const float4*)a.add_src + t);
        o[0] += av.x;
        o[1] += av.y;
        o[2] += av.z;
        o[3] += av.w;
      }
    }
    ((float4*)a.out)[t] = make_float4(o[0], o[1], o[2], o[3]);
  }
}

}  // namespace b2s

using namespace b2s;

extern "C" {

// launch geometry shared by the column-sum kernel and the one-launch kernels: as many equal CTAs as the cap allows -- a
// whole number per SM (rounding the chunk to whole load batches left 212 CTAs on 148 SMs for the level-0 tensor).  Both
// paths use the same chunks, so their partial sums and therefore their results are bit-identical.
static void bn_grid_fused(int64_t n, int c, int cap, int* chunk, int* nblk) {
  const int c4 = c / 4;
  const int rpp = BN_THREADS / c4 > 0 ? BN_THREADS / c4 : 1;
  // (<= 296 CTAs: b2s_bn_ws_bytes sizes the partial sums for that many; wide rows: fewer partials to sum per CTA)
  const int gmax = std::min(std::min(cap, 296), c <= 32 ? 2 * sm_count() : (c <= 64 ? sm_count() : sm_count() / 2));
  if (n < 1) n = 1;
  int64_t g = std::max<int64_t>(1, std::min<int64_t>(gmax, n / (2 * (int64_t)rpp)));
  const int64_t ch = cdiv(n, g);
  *chunk = (int)ch;
  *nblk = (int)cdiv(n, ch);
}
static void bn_grid(int64_t n, int c, int unroll, int* chunk, int* nblk) {
  (void)unroll;
  bn_grid_fused(n, c, 296, chunk, nblk);
}

size_t b2s_bn_ws_bytes(int64_t n, int32_t c) {
  (void)n;
  return align_up((size_t)296 * 2 * (c > 0 ? c : 4) * 8) + 1024;
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
  int chunk, nblk;
  bn_grid(n, c, 8, &chunk, &nblk);
  if (ws_bytes < (size_t)nblk * 2 * c * 8) {
    set_error("bn_stats: workspace too small");
    return B2S_E_WORKSPACE;
  }
  double* partial = (double*)ws;
  BnFinish fin{eps, momentum, running_mean, running_var, mean, var_biased, rstd, counter};
  bn_colsum_kernel<0><<<nblk, BN_THREADS, bn_smem(c), stream>>>(x, nullptr, nullptr, nullptr, nullptr, n, c, chunk, 0,
                                                               partial, fin);
  return check_launch("bn_stats");
}

}  // extern "C"

// one cooperative launch per BN forward / backward (B2S_BN_FUSED=0: the two-kernel path)
static bool bn_fused_on() {
  static const int on = [] {
    const char* e = getenv("B2S_BN_FUSED");
    return (e == nullptr || atoi(e) != 0) ? 1 : 0;
  }();
  return on != 0;
}

static size_t bn_fused_smem(int c) { return bn_smem(c) + (size_t)2 * c * 4; }


// CTAs of bn_fused_kernel<MODE> that are resident at once on the current device with this much shared memory
template <int MODE>
static int bn_fused_capacity(int c) {
  static int cached[B2S_MAX_DEVICES][2] = {};
  const int dev = current_device();
  const int slot = c <= 256 ? 0 : 1;  // shared memory grows with c above 256 channels only
  if (cached[dev][slot] == 0) {
    int per_sm = 0;
    const size_t smem = bn_fused_smem(slot == 0 ? 256 : 1024);
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(bn_fused_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bn_fused_kernel<MODE>, BN_THREADS, smem);
    cached[dev][slot] = std::max(1, per_sm) * sm_count();
  }
  return cached[dev][slot];
}

template <int MODE>
static int bn_fused_launch(BnFusedArgs a, int nblk, cudaStream_t stream) {
  void* args[] = {(void*)&a};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)bn_fused_kernel<MODE>, dim3(nblk), dim3(BN_THREADS), args,
                                              bn_fused_smem(a.c), stream);
  if (e != cudaSuccess) {
    set_error(cudaGetErrorString(e));
    cudaGetLastError();
    return B2S_E_LAUNCH;
  }
  return B2S_OK;
}

extern "C" {

int b2s_bn_forward(const float* x, int64_t n, int32_t c, float eps, float momentum, float* running_mean,
                   float* running_var, const float* gamma, const float* beta, int32_t relu, float* y, float* mean,
                   float* rstd, int32_t* counter, void* ws, size_t ws_bytes, b2s_stream_t stream) {
  if (bn_fused_on() && n > 0 && bn_check(n, c) == B2S_OK) {
    int chunk, nblk;
    bn_grid_fused(n, c, bn_fused_capacity<0>(c), &chunk, &nblk);
    if (nblk <= bn_fused_capacity<0>(c) && ws_bytes >= (size_t)nblk * 2 * c * 8) {
      BnFusedArgs a{x, nullptr, nullptr, nullptr, nullptr, gamma, beta, nullptr, y, mean, rstd, running_mean, running_var,
                    (double*)ws, counter, n, c, chunk, relu, 1, eps, momentum};
      return bn_fused_launch<0>(a, nblk, stream);
    }
  }
  int rc = b2s_bn_stats(x, n, c, eps, momentum, running_mean, running_var, mean, nullptr, rstd, counter, ws, ws_bytes,
                        stream);
  if (rc) return rc;
  return b2s_bn_apply(x, n, c, mean, rstd, gamma, beta, relu, y, stream);
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
  return b2s_bn_backward_add(x, y, dy, nullptr, n, c, mean, rstd, gamma, relu, training, dx, dgamma, dbeta, counter, ws,
                             ws_bytes, stream);
}

int b2s_bn_backward_add(const float* x, const float* y, const float* dy, const float* add_src, int64_t n, int32_t c,
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
  int chunk, nblk;
  bn_grid(n, c, 4, &chunk, &nblk);
  if (ws_bytes < (size_t)nblk * 2 * c * 8) {
    set_error("bn_backward: workspace too small");
    return B2S_E_WORKSPACE;
  }
  double* partial = (double*)ws;
  if (bn_fused_on()) {
    int fchunk, fnblk;
    bn_grid_fused(n, c, bn_fused_capacity<1>(c), &fchunk, &fnblk);
    chunk = fchunk;
    nblk = fnblk;
  }
  if (bn_fused_on() && nblk <= bn_fused_capacity<1>(c) && ws_bytes >= (size_t)nblk * 2 * c * 8) {
    BnFusedArgs a{x, y, dy, mean, rstd, gamma, nullptr, add_src, dx, dbeta, dgamma, nullptr, nullptr,
                  partial, counter, n, c, chunk, relu, training, 0.f, 0.f};
    return bn_fused_launch<1>(a, nblk, stream);
  }
  BnFinish fin{0.f, 0.f, nullptr, nullptr, dbeta, dgamma, nullptr, counter};
  bn_colsum_kernel<1><<<nblk, BN_THREADS, bn_smem(c), stream>>>(x, y, dy, mean, rstd, n, c, chunk, relu, partial, fin);
  int64_t total4 = n * (c / 4);
  bn_dx_kernel<<<(unsigned)cdiv(total4, 256), 256, 0, stream>>>(
      (const float4*)x, (const float4*)y, (const float4*)dy, total4, c / 4, 1.0f / (float)n, mean, rstd,
      gamma, dgamma, dbeta, relu, training, (const float4*)add_src, (float4*)dx);
  return check_launch("bn_backward");
}

}  // extern "C"
