// Semantic cross-entropy of the train step (minsu3d/model/general_model.py:36-40: F.cross_entropy(semantic_scores,
// sem_labels.long(), ignore_index=-1), mean over the labelled points) as one fused pass per direction.
// torch runs log_softmax + a single-block nll_loss reduction (424 us + 246 us on the 400k x 20 scores of the benchmark
// batch, profiles/r02_step_launches.txt); here: one thread per point, float4 loads, block partial sums in double,
// the last block to retire adds the partials in block order (deterministic) -- HBM-bound: 4 C n bytes per direction.
#include "common.cuh"

namespace b2s {

constexpr int CE_THREADS = 256;

template <typename L>
__global__ void __launch_bounds__(CE_THREADS)
    ce_forward_kernel(const float* __restrict__ x, const L* __restrict__ labels, int64_t n, int c, int ignore_index,
                      double* __restrict__ partial, int32_t* __restrict__ counter, float* __restrict__ loss,
                      float* __restrict__ n_valid) {
  __shared__ double s_sum[CE_THREADS / 32], s_cnt[CE_THREADS / 32];
  __shared__ bool s_last;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double my = 0.0, cnt = 0.0;
  if (i < n) {
    const int lab = (int)labels[i];
    if (lab != ignore_index) {
      const float* row = x + i * c;
      float m = -INFINITY;
      for (int j = 0; j < c; ++j) m = fmaxf(m, __ldg(row + j));
      float s = 0.f;
      for (int j = 0; j < c; ++j) s += expf(__ldg(row + j) - m);
      my = (double)(logf(s) + m - __ldg(row + lab));
      cnt = 1.0;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    my += __shfl_down_sync(0xffffffffu, my, o);
    cnt += __shfl_down_sync(0xffffffffu, cnt, o);
  }
  if ((threadIdx.x & 31) == 0) {
    s_sum[threadIdx.x >> 5] = my;
    s_cnt[threadIdx.x >> 5] = cnt;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < CE_THREADS / 32; ++w) {
      a += s_sum[w];
      b += s_cnt[w];
    }
    partial[2 * blockIdx.x] = a;
    partial[2 * blockIdx.x + 1] = b;
    __threadfence();
    s_last = (atomicAdd(counter, 1) == (int)gridDim.x - 1);
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {  // second stage: fixed order over the blocks
    __threadfence();
    double a = 0.0, b = 0.0;
    for (unsigned g = 0; g < gridDim.x; ++g) {
      a += ((volatile double*)partial)[2 * g];
      b += ((volatile double*)partial)[2 * g + 1];
    }
    *loss = (float)(a / b);  // 0 / 0 = nan like torch when every label is ignored
    *n_valid = (float)b;
    *counter = 0;  // ready for the next launch
  }
}

template <typename L>
__global__ void __launch_bounds__(CE_THREADS)
    ce_backward_kernel(const float* __restrict__ x, const L* __restrict__ labels, int64_t n, int c, int ignore_index,
                       const float* __restrict__ n_valid, const float* __restrict__ gout, float* __restrict__ gx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int lab = (int)labels[i];
  float* g = gx + i * c;
  if (lab == ignore_index) {
    for (int j = 0; j < c; ++j) g[j] = 0.f;
    return;
  }
  const float* row = x + i * c;
  float m = -INFINITY;
  for (int j = 0; j < c; ++j) m = fmaxf(m, __ldg(row + j));
  float s = 0.f;
  for (int j = 0; j < c; ++j) s += expf(__ldg(row + j) - m);
  const float scale = __ldg(gout) / __ldg(n_valid), inv = 1.f / s;
  for (int j = 0; j < c; ++j) {
    const float p = expf(__ldg(row + j) - m) * inv;
    g[j] = (p - (j == lab ? 1.f : 0.f)) * scale;
  }
}

}  // namespace b2s

using namespace b2s;

extern "C" {

size_t b2s_cross_entropy_ws_bytes(int64_t n) { return (size_t)(cdiv(n > 0 ? n : 1, CE_THREADS)) * 16 + 256; }

// labels: int16 (label_bytes = 2, the loader's dtype, general_dataset.py:88) or int64 (label_bytes = 8).
// loss [1], n_valid [1] (kept for the backward); counter: int32 that is 0 between launches (the BatchNorm one works).
int b2s_cross_entropy_forward(const float* scores, const void* labels, int32_t label_bytes, int64_t n, int32_t c,
                              int32_t ignore_index, float* loss, float* n_valid, int32_t* counter, void* ws,
                              size_t ws_bytes, b2s_stream_t stream) {
  if (n <= 0 || c < 1 || (label_bytes != 2 && label_bytes != 8) || !scores || !labels || !loss || !n_valid || !counter) {
    set_error("cross_entropy_forward: invalid argument");
    return B2S_E_INVALID;
  }
  if (ws_bytes < b2s_cross_entropy_ws_bytes(n)) {
    set_error("cross_entropy_forward: workspace too small");
    return B2S_E_WORKSPACE;
  }
  const unsigned grid = (unsigned)cdiv(n, CE_THREADS);
  if (label_bytes == 2)
    ce_forward_kernel<int16_t><<<grid, CE_THREADS, 0, stream>>>(scores, (const int16_t*)labels, n, c, ignore_index,
                                                                (double*)ws, counter, loss, n_valid);
  else
    ce_forward_kernel<int64_t><<<grid, CE_THREADS, 0, stream>>>(scores, (const int64_t*)labels, n, c, ignore_index,
                                                                (double*)ws, counter, loss, n_valid);
  return check_launch("cross_entropy_forward");
}

int b2s_cross_entropy_backward(const float* scores, const void* labels, int32_t label_bytes, int64_t n, int32_t c,
                               int32_t ignore_index, const float* n_valid, const float* gout, float* gscores,
                               b2s_stream_t stream) {
  if (n <= 0 || c < 1 || (label_bytes != 2 && label_bytes != 8) || !scores || !labels || !n_valid || !gout || !gscores) {
    set_error("cross_entropy_backward: invalid argument");
    return B2S_E_INVALID;
  }
  const unsigned grid = (unsigned)cdiv(n, CE_THREADS);
  if (label_bytes == 2)
    ce_backward_kernel<int16_t><<<grid, CE_THREADS, 0, stream>>>(scores, (const int16_t*)labels, n, c, ignore_index,
                                                                 n_valid, gout, gscores);
  else
    ce_backward_kernel<int64_t><<<grid, CE_THREADS, 0, stream>>>(scores, (const int64_t*)labels, n, c, ignore_index,
                                                                 n_valid, gout, gscores);
  return check_launch("cross_entropy_backward");
}

}  // extern "C"
