// S1-S3, I1-I2, V2: segmented reductions, proposal x instance IoU, mask labels, row gather/scatter.
//
// Reference kernels: minsu3d/common_ops/src/sec_mean/sec_mean.cu:12-85,
// roipool/roipool.cu:12-120, get_iou/get_iou.cu:12-37,
// cal_iou_and_masklabel/cal_iou_and_masklabel.cu:14-140.
// The reference runs one thread per (segment, channel) with a serial scan over the segment
// (C=3 -> 3 active threads per block).  Here the loads are block-parallel and coalesced; the
// sums that the reference defines sequentially (sec_mean's sum(x/count), global_avg_pool's
// sum(x)/n) keep their sequential fp32 addition order, so results are bit-identical.
#include <algorithm>

#include "common.cuh"

namespace b2s {

constexpr int SEG_THREADS = 128;
constexpr int SEG_ROWS = 64;  // rows staged per chunk for the ordered sums

// ordered sum: mode 0 = sum(x / count) (sec_mean), mode 1 = sum(x) / n (global_avg_pool)
__global__ void __launch_bounds__(SEG_THREADS)
    seg_ordered_sum_kernel(const float* __restrict__ inp, const int32_t* __restrict__ offsets,
                           float* __restrict__ out, int n_seg, int c, int mode) {
  extern __shared__ float s_buf[];  // SEG_ROWS * c
  for (int seg = blockIdx.x; seg < n_seg; seg += gridDim.x) {
    int b = offsets[seg], e = offsets[seg + 1];
    float count = (float)(e - b);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};  // channels tid, tid+128, ... (c <= 512)
    for (int r0 = b; r0 < e; r0 += SEG_ROWS) {
      int rows = min(SEG_ROWS, e - r0);
      const float* src = inp + (int64_t)r0 * c;
      for (int i = threadIdx.x; i < rows * c; i += SEG_THREADS) {
        float v = __ldg(src + i);
        s_buf[i] = (mode == 0) ? __fdiv_rn(v, count) : v;
      }
      __syncthreads();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        int ch = threadIdx.x + q * SEG_THREADS;
        if (ch < c) {
          float a = acc[q];
          for (int r = 0; r < rows; ++r) a = __fadd_rn(a, s_buf[r * c + ch]);
          acc[q] = a;
        }
      }
      __syncthreads();
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int ch = threadIdx.x + q * SEG_THREADS;
      if (ch < c) out[(int64_t)seg * c + ch] = (mode == 0) ? acc[q] : __fdiv_rn(acc[q], count);
    }
  }
}

// max / min with first-index-wins argmax: block per segment.  Thread t owns channel (t % Cw) and row phase
// (t / Cw) with Cw = min(C, 128): for the 3-channel coordinate reductions (sec_min / sec_max) 42 row phases work
// in parallel instead of 3 active threads per block in the reference (sec_mean.cu:38-85).  op: 0 = max, 1 = min
__global__ void __launch_bounds__(SEG_THREADS)
    seg_minmax_kernel(const float* __restrict__ inp, const int32_t* __restrict__ offsets,
                      float* __restrict__ out, int32_t* __restrict__ argidx, int n_seg, int c, int op) {
  extern __shared__ unsigned char s_raw[];
  const int cw = c < SEG_THREADS ? c : SEG_THREADS;
  const int phases = SEG_THREADS / cw;                 // >= 1
  float* s_val = (float*)s_raw;                        // [phases][c]
  int32_t* s_idx = (int32_t*)(s_val + phases * c);     // [phases][c]
  const int ch0 = threadIdx.x % cw, ph = threadIdx.x / cw;
  const float init = (op == 0) ? -INFINITY : INFINITY;  // (float)(-1e50), sec_mean.cu:44,70
  for (int seg = blockIdx.x; seg < n_seg; seg += gridDim.x) {
    const int b = offsets[seg], e = offsets[seg + 1];
    if (ph < phases) {
      for (int ch = ch0; ch < c; ch += cw) {
        float best = init;
        int bi = -1;
        // eight independent loads in flight per thread (the compare chain alone exposes one L2 round trip per row)
        int r = b + ph;
        for (; r + 7 * phases < e; r += 8 * phases) {
          float v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = __ldg(inp + (int64_t)(r + u * phases) * c + ch);
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const bool better = (op == 0) ? (v[u] > best) : (v[u] < best);
            if (better) {
              best = v[u];
              bi = r + u * phases;
            }
          }
        }
        for (; r < e; r += phases) {
          float v = __ldg(inp + (int64_t)r * c + ch);
          bool better = (op == 0) ? (v > best) : (v < best);
          if (better) {
            best = v;
            bi = r;
          }
        }
        s_val[ph * c + ch] = best;
        s_idx[ph * c + ch] = bi;
      }
    }
    __syncthreads();
    for (int ch = threadIdx.x; ch < c; ch += SEG_THREADS) {
      float best = init;
      int bi = -1;
      for (int w = 0; w < phases; ++w) {
        float v = s_val[w * c + ch];
        int i = s_idx[w * c + ch];
        if (i < 0) continue;
        bool better = (op == 0) ? (v > best) : (v < best);
        if (bi < 0 || better || (v == best && i < bi)) {
          best = v;
          bi = i;
        }
      }
      out[(int64_t)seg * c + ch] = best;
      if (argidx) argidx[(int64_t)seg * c + ch] = bi;
    }
    __syncthreads();
  }
}

// Long segments (proposals of ~10k points): grid = (segment, part), every CTA scans one contiguous part of its segment
// with the kernel above's thread layout and writes a partial (max, first arg-max) per channel; the combine kernel
// merges the parts in order (ties keep the lower row: same result as one sequential scan, roipool.cu:22-30).
__global__ void __launch_bounds__(SEG_THREADS)
    seg_max_parts_kernel(const float* __restrict__ inp, const int32_t* __restrict__ offsets, float* __restrict__ pval,
                         int32_t* __restrict__ pidx, int n_seg, int c, int parts) {
  extern __shared__ unsigned char s_raw[];
  const int cw = c < SEG_THREADS ? c : SEG_THREADS;
  const int phases = SEG_THREADS / cw;
  float* s_val = (float*)s_raw;
  int32_t* s_idx = (int32_t*)(s_val + phases * c);
  const int ch0 = threadIdx.x % cw, ph = threadIdx.x / cw;
  const int seg = blockIdx.x, part = blockIdx.y;
  const int b0 = offsets[seg], e0 = offsets[seg + 1];
  const int per = (e0 - b0 + parts - 1) / parts;
  const int b = b0 + part * per, e = min(e0, b + per);
  if (ph < phases) {
    for (int ch = ch0; ch < c; ch += cw) {
      float best = -INFINITY;
      int bi = -1;
      int r = b + ph;
      for (; r + 7 * phases < e; r += 8 * phases) {  // eight loads in flight (see seg_minmax_kernel)
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldg(inp + (int64_t)(r + u * phases) * c + ch);
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (v[u] > best) {
            best = v[u];
            bi = r + u * phases;
          }
      }
      for (; r < e; r += phases) {
        const float v = __ldg(inp + (int64_t)r * c + ch);
        if (v > best) {
          best = v;
          bi = r;
        }
      }
      s_val[ph * c + ch] = best;
      s_idx[ph * c + ch] = bi;
    }
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += SEG_THREADS) {
    float best = -INFINITY;
    int bi = -1;
    for (int w = 0; w < phases; ++w) {
      const float v = s_val[w * c + ch];
      const int i = s_idx[w * c + ch];
      if (i < 0) continue;
      if (bi < 0 || v > best || (v == best && i < bi)) {
        best = v;
        bi = i;
      }
    }
    pval[((int64_t)seg * parts + part) * c + ch] = best;
    pidx[((int64_t)seg * parts + part) * c + ch] = bi;
  }
}

__global__ void __launch_bounds__(256)
    seg_max_combine_kernel(const float* __restrict__ pval, const int32_t* __restrict__ pidx, float* __restrict__ out,
                           int32_t* __restrict__ argidx, int64_t total, int c, int parts) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int64_t seg = t / c;
  const int ch = (int)(t - seg * c);
  float best = -INFINITY;
  int bi = -1;
  for (int p = 0; p < parts; ++p) {  // parts are in row order: a strict > keeps the first maximum
    const float v = pval[(seg * parts + p) * c + ch];
    const int i = pidx[(seg * parts + p) * c + ch];
    if (i >= 0 && (bi < 0 || v > best)) {
      best = v;
      bi = i;
    }
  }
  out[t] = best;
  if (argidx) argidx[t] = bi;
}

__global__ void __launch_bounds__(256)
    roipool_bp_kernel(float* __restrict__ d_feats, const int32_t* __restrict__ maxidx,
                      const float* __restrict__ d_out, int64_t total, int c) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  int ch = (int)(t % c);
  int a = maxidx[t];
  if (a >= 0) atomicAdd(d_feats + (int64_t)a * c + ch, d_out[t]);  // segments are disjoint: no contention
}

__global__ void __launch_bounds__(256)
    gap_bp_kernel(float* __restrict__ d_feats, const int32_t* __restrict__ offsets,
                  const float* __restrict__ d_out, int n_seg, int c) {
  for (int seg = blockIdx.x; seg < n_seg; seg += gridDim.x) {
    int b = offsets[seg], e = offsets[seg + 1];
    float n = (float)(e - b);
    int64_t total = (int64_t)(e - b) * c;
    for (int64_t i = threadIdx.x; i < total; i += blockDim.x) {
      int ch = (int)(i % c);
      d_feats[(int64_t)b * c + i] += __fdiv_rn(d_out[(int64_t)seg * c + ch], n);
    }
  }
}

// ---- IoU: one block per proposal, shared-memory histogram over instances --------------------
constexpr int IOU_CHUNK = 8192;

__global__ void __launch_bounds__(256)
    iou_kernel(const int32_t* __restrict__ pidx, const int32_t* __restrict__ poff,
               const int16_t* __restrict__ inst_labels, const int32_t* __restrict__ inst_num,
               const float* __restrict__ mask_scores, float* __restrict__ iou, int n_inst, int n_prop) {
  __shared__ int s_hist[IOU_CHUNK];
  __shared__ int s_total;
  for (int p = blockIdx.x; p < n_prop; p += gridDim.x) {
    int b = poff[p], e = poff[p + 1];
    for (int i0 = 0; i0 < n_inst; i0 += IOU_CHUNK) {
      int ni = min(IOU_CHUNK, n_inst - i0);
      for (int i = threadIdx.x; i < ni; i += blockDim.x) s_hist[i] = 0;
      if (threadIdx.x == 0) s_total = 0;
      __syncthreads();
      int mytot = 0;
      for (int i = b + threadIdx.x; i < e; i += blockDim.x) {
        if (mask_scores && !(mask_scores[i] > 0.5f)) continue;
        ++mytot;
        int lab = (int)inst_labels[pidx[i]] - i0;
        if (lab >= 0 && lab < ni) atomicAdd(&s_hist[lab], 1);
      }
      if (mytot) atomicAdd(&s_total, mytot);
      __syncthreads();
      int ptotal = s_total;
      for (int i = threadIdx.x; i < ni; i += blockDim.x) {
        int inter = s_hist[i];
        int itotal = inst_num[i0 + i];
        // (float)inter / ((float)(p + i - inter) + 1e-5): double division, get_iou.cu:26
        double den = (double)(float)(ptotal + itotal - inter) + 1e-5;
        iou[(int64_t)p * n_inst + i0 + i] = (float)((double)(float)inter / den);
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(256)
    mask_label_kernel(const int32_t* __restrict__ pidx, const int32_t* __restrict__ poff,
                      const int16_t* __restrict__ inst_labels, const int16_t* __restrict__ inst_cls,
                      const float* __restrict__ iou, int n_inst, int n_prop, int ignored, float thr,
                      uint8_t* __restrict__ mask_label, uint8_t* __restrict__ mask_label_mask) {
  __shared__ float s_v[256];
  __shared__ int s_i[256];
  for (int p = blockIdx.x; p < n_prop; p += gridDim.x) {
    // first index of the maximum over non-ignored instances with iou > 0 (else (0, 0))
    float best = 0.f;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < n_inst; i += blockDim.x) {
      float v = iou[(int64_t)p * n_inst + i];
      if (v > best && (int)inst_cls[i] != ignored) {
        best = v;
        bi = i;
      }
    }
    s_v[threadIdx.x] = best;
    s_i[threadIdx.x] = bi;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
      if (threadIdx.x < s) {
        float v = s_v[threadIdx.x + s];
        int i = s_i[threadIdx.x + s];
        if (v > s_v[threadIdx.x] || (v == s_v[threadIdx.x] && i < s_i[threadIdx.x])) {
          s_v[threadIdx.x] = v;
          s_i[threadIdx.x] = i;
        }
      }
      __syncthreads();
    }
    float max_iou = s_v[0];
    int max_ind = (s_i[0] == 0x7fffffff) ? 0 : s_i[0];
    __syncthreads();
    if (max_iou >= thr) {
      int b = poff[p], e = poff[p + 1];
      for (int i = b + threadIdx.x; i < e; i += blockDim.x) {
        if ((int)inst_labels[pidx[i]] == max_ind) mask_label[i] = 1;
        mask_label_mask[i] = 1;
      }
    }
  }
}

// ---- devoxelise gather / scatter-add ---------------------------------------------------------
__global__ void __launch_bounds__(256)
    gather_rows_kernel(const float* __restrict__ feat, const int64_t* __restrict__ idx, int64_t total,
                       int c4, float4* __restrict__ out) {
  // four independent (index -> row piece) chains per thread: one chain per thread keeps 32 KB per SM in flight, which
  // two dependent DRAM latencies turn into ~3 TB/s (profiles/r02 roofline table: 0.44 of HBM)
  const int64_t t0 = (int64_t)blockIdx.x * (blockDim.x * 4) + threadIdx.x;
  int64_t src[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int64_t t = t0 + (int64_t)u * blockDim.x;
    const int64_t r = t / c4;
    src[u] = t < total ? __ldg(idx + r) * c4 + (t - r * c4) : -1;
  }
  float4 v[4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
    if (src[u] >= 0) v[u] = __ldg((const float4*)feat + src[u]);
#pragma unroll
  for (int u = 0; u < 4; ++u)
    if (src[u] >= 0) out[t0 + (int64_t)u * blockDim.x] = v[u];
}
__global__ void __launch_bounds__(256)
    gather_rows_scalar_kernel(const float* __restrict__ feat, const int64_t* __restrict__ idx,
                              int64_t total, int c, float* __restrict__ out) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  int64_t r = t / c;
  int q = (int)(t - r * c);
  out[t] = __ldg(feat + idx[r] * c + q);
}
__global__ void __launch_bounds__(256)
    scatter_add_rows_kernel(const float* __restrict__ grad, const int64_t* __restrict__ idx,
                            int64_t total, int c4, float* __restrict__ gfeat) {
  const int64_t t0 = (int64_t)blockIdx.x * (blockDim.x * 4) + threadIdx.x;
  int64_t dsti[4];
  float4 v[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int64_t t = t0 + (int64_t)u * blockDim.x;
    const int64_t r = t / c4;
    dsti[u] = t < total ? __ldg(idx + r) * c4 + (t - r * c4) : -1;
    if (t < total) v[u] = __ldg((const float4*)grad + t);
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    if (dsti[u] < 0) continue;
    float* dst = gfeat + dsti[u] * 4;
    // 16-byte vector reduction (sm_90+): one L2 atomic per 4 channels
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v[u].x), "f"(v[u].y), "f"(v[u].z),
                 "f"(v[u].w)
                 : "memory");
  }
}
__global__ void __launch_bounds__(256)
    scatter_add_rows_scalar_kernel(const float* __restrict__ grad, const int64_t* __restrict__ idx,
                                   int64_t total, int c, float* __restrict__ gfeat) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  int64_t r = t / c;
  int q = (int)(t - r * c);
  atomicAdd(gfeat + idx[r] * c + q, grad[t]);
}

static int seg_grid(int n_seg) { return n_seg < 8 * B2S_SM_COUNT ? n_seg : 8 * B2S_SM_COUNT; }


// ------------------------------------------------------------------------------------------
// clusters_voxelization (general_model.py:152-193) as two kernels: every float operation of the reference's
// torch expression sequence is reproduced with its own rounding (no FMA contraction), so the integer voxel
// coordinates are bit-identical:
//   mean_c  = sec_mean(x)                      (sequential sum of x / count, sec_mean.cu:12-27)
//   lo, hi  = min / max over the cluster of (x - mean_c)
//   scale_c = min(1 / max_d((hi_d - lo_d) / S) - 0.01, scale)
//   off_c,d = -(lo * scale) + max(S - (hi*scale - lo*scale) - 0.001, 0) * r0_d + min(S - (...) + 0.001, 0) * r1_d
//   voxel   = int((x - mean_c) * scale_c + off_c)            (truncation, .int())
// ------------------------------------------------------------------------------------------
constexpr int CV_THREADS = 128;

template <typename IdxT>
__global__ void __launch_bounds__(CV_THREADS)
    cv_cluster_kernel(const IdxT* __restrict__ clusters_idx, const int32_t* __restrict__ offsets, int n_cluster,
                      const float* __restrict__ coords, float scale, float shape, const float* __restrict__ rnd,
                      float* __restrict__ params /* [nC][8]: mean xyz, scale, offset xyz, - */) {
  __shared__ float s_x[CV_THREADS * 3];
  __shared__ float s_red[2][CV_THREADS / 32][3];
  const int c = blockIdx.x;
  if (c >= n_cluster) return;
  const int b = offsets[c], e = offsets[c + 1];
  const float count = (float)(e - b);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // (1) mean: points staged in shared memory CV_THREADS at a time, three threads add sequentially
  float acc = 0.f;
  for (int r0 = b; r0 < e; r0 += CV_THREADS) {
    const int rows = min(CV_THREADS, e - r0);
    if (tid < rows) {
      const int64_t pt = (int64_t)clusters_idx[2 * (int64_t)(r0 + tid) + 1];
#pragma unroll
      for (int d = 0; d < 3; ++d) s_x[tid * 3 + d] = __fdiv_rn(__ldg(coords + pt * 3 + d), count);
    }
    __syncthreads();
    if (tid < 3)
      for (int r = 0; r < rows; ++r) acc = __fadd_rn(acc, s_x[r * 3 + tid]);
    __syncthreads();
  }
  __shared__ float s_mean[3];
  if (tid < 3) s_mean[tid] = acc;
  __syncthreads();
  const float m0 = s_mean[0], m1 = s_mean[1], m2 = s_mean[2];
  // (2) min / max of the centred coordinates (order-independent)
  float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
  for (int r = b + tid; r < e; r += CV_THREADS) {
    const int64_t pt = (int64_t)clusters_idx[2 * (int64_t)r + 1];
    const float v[3] = {__fsub_rn(__ldg(coords + pt * 3), m0), __fsub_rn(__ldg(coords + pt * 3 + 1), m1),
                        __fsub_rn(__ldg(coords + pt * 3 + 2), m2)};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      lo[d] = fminf(lo[d], v[d]);
      hi[d] = fmaxf(hi[d], v[d]);
    }
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
      hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
    }
    if (lane == 0) {
      s_red[0][warp][d] = lo[d];
      s_red[1][warp][d] = hi[d];
    }
  }
  __syncthreads();
  if (tid == 0) {
    float L[3], H[3];
    for (int d = 0; d < 3; ++d) {
      L[d] = s_red[0][0][d];
      H[d] = s_red[1][0][d];
      for (int w = 1; w < CV_THREADS / 32; ++w) {
        L[d] = fminf(L[d], s_red[0][w][d]);
        H[d] = fmaxf(H[d], s_red[1][w][d]);
      }
    }
    // (3) scale and offset, one rounding per reference operation
    // tensor / python-scalar on CUDA multiplies by the fp32 reciprocal of the scalar (ATen BinaryDivTrueKernel.cu)
    const float inv_shape = __fdiv_rn(1.0f, shape);
    float ext = __fmul_rn(__fsub_rn(H[0], L[0]), inv_shape);
    ext = fmaxf(ext, __fmul_rn(__fsub_rn(H[1], L[1]), inv_shape));
    ext = fmaxf(ext, __fmul_rn(__fsub_rn(H[2], L[2]), inv_shape));
    float sc = __fsub_rn(__fdiv_rn(1.0f, ext), 0.01f);
    sc = fminf(sc, scale);  // torch.clamp(max=scale)
    float* out = params + (int64_t)c * 8;
    out[0] = m0;
    out[1] = m1;
    out[2] = m2;
    out[3] = sc;
    for (int d = 0; d < 3; ++d) {
      const float mn = __fmul_rn(L[d], sc), mx = __fmul_rn(H[d], sc);
      const float rg = __fsub_rn(mx, mn);
      const float room = __fsub_rn(shape, rg);
      float off = __fadd_rn(-mn, __fmul_rn(fmaxf(__fsub_rn(room, 0.001f), 0.f), rnd[d]));
      off = __fadd_rn(off, __fmul_rn(fminf(__fadd_rn(room, 0.001f), 0.f), rnd[3 + d]));
      out[4 + d] = off;
    }
    out[7] = 0.f;
  }
}

template <typename IdxT>
__global__ void __launch_bounds__(256)
    cv_point_kernel(const IdxT* __restrict__ clusters_idx, int64_t total, const float* __restrict__ coords,
                    const float* __restrict__ params, int32_t* __restrict__ out_xyz) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int64_t c = (int64_t)clusters_idx[2 * i], pt = (int64_t)clusters_idx[2 * i + 1];
  const float4 p0 = __ldg((const float4*)(params + c * 8)), p1 = __ldg((const float4*)(params + c * 8 + 4));
  const float mean[3] = {p0.x, p0.y, p0.z}, off[3] = {p1.x, p1.y, p1.z};
  int v[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const float x = __fsub_rn(__ldg(coords + pt * 3 + d), mean[d]);
    v[d] = (int)__fadd_rn(__fmul_rn(x, p0.w), off[d]);  // float -> int32 conversion truncates like .int()
  }
  ((int4*)out_xyz)[i] = make_int4((int)c, v[0], v[1], v[2]);
}

}  // namespace b2s

using namespace b2s;

extern "C" {

static int ordered_sum(const float* inp, const int32_t* offsets, float* out, int32_t n_seg, int32_t c,
                       int mode, b2s_stream_t stream) {
  if (n_seg < 0 || c < 1 || c > 4 * SEG_THREADS) {
    set_error("segmented sum: invalid n_seg / channel count (c <= 512)");
    return B2S_E_INVALID;
  }
  if (n_seg == 0) return B2S_OK;
  size_t smem = (size_t)SEG_ROWS * c * 4;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(seg_ordered_sum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  seg_ordered_sum_kernel<<<seg_grid(n_seg), SEG_THREADS, smem, stream>>>(inp, offsets, out, n_seg, c, mode);
  return check_launch("seg_ordered_sum");
}

static int minmax(const float* inp, const int32_t* offsets, float* out, int32_t* arg, int32_t n_seg,
                  int32_t c, int op, b2s_stream_t stream) {
  if (n_seg < 0 || c < 1 || c > 2048) {
    set_error("segmented min/max: invalid argument");
    return B2S_E_INVALID;
  }
  if (n_seg == 0) return B2S_OK;
  const int cw = c < SEG_THREADS ? c : SEG_THREADS;
  size_t smem = (size_t)2 * (SEG_THREADS / cw) * c * 4;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(seg_minmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  seg_minmax_kernel<<<seg_grid(n_seg), SEG_THREADS, smem, stream>>>(inp, offsets, out, arg, n_seg, c, op);
  return check_launch("seg_minmax");
}

int b2s_sec_mean(const float* inp, const int32_t* offsets, float* out, int32_t n_seg, int32_t c, b2s_stream_t s) {
  return ordered_sum(inp, offsets, out, n_seg, c, 0, s);
}
int b2s_sec_min(const float* inp, const int32_t* offsets, float* out, int32_t n_seg, int32_t c, b2s_stream_t s) {
  return minmax(inp, offsets, out, nullptr, n_seg, c, 1, s);
}
int b2s_sec_max(const float* inp, const int32_t* offsets, float* out, int32_t n_seg, int32_t c, b2s_stream_t s) {
  return minmax(inp, offsets, out, nullptr, n_seg, c, 0, s);
}
int b2s_roipool_fp(const float* feats, const int32_t* offsets, float* out, int32_t* maxidx, int32_t n_seg,
                   int32_t c, b2s_stream_t s) {
  return minmax(feats, offsets, out, maxidx, n_seg, c, 0, s);
}
size_t b2s_roipool_ws_bytes(int32_t n_seg, int32_t c) { return (size_t)(n_seg > 0 ? n_seg : 1) * 32 * c * 8 + 256; }

// roipool forward for long segments: n_rows = rows of feats (picks the number of parts per segment); ws from
// b2s_roipool_ws_bytes.  Same result as b2s_roipool_fp.
int b2s_roipool_fp_ws(const float* feats, const int32_t* offsets, float* out, int32_t* maxidx, int32_t n_seg,
                      int32_t c, int64_t n_rows, void* ws, size_t ws_bytes, b2s_stream_t stream) {
  if (n_seg < 0 || c < 1 || c > 2048) {
    set_error("roipool_fp_ws: invalid argument");
    return B2S_E_INVALID;
  }
  if (n_seg == 0) return B2S_OK;
  const int64_t avg = n_rows / n_seg;
  int parts = (int)std::min<int64_t>(32, std::max<int64_t>(1, avg / 1024));
  // enough CTAs to fill the GPU, no more parts than that needs
  while (parts > 1 && (int64_t)n_seg * (parts / 2) >= 4 * (int64_t)sm_count()) parts /= 2;
  if (parts <= 1 || ws == nullptr || ws_bytes < b2s_roipool_ws_bytes(n_seg, c))
    return minmax(feats, offsets, out, maxidx, n_seg, c, 0, stream);
  float* pval = (float*)ws;
  int32_t* pidx = (int32_t*)((char*)ws + (size_t)n_seg * parts * c * 4);
  const int cw = c < SEG_THREADS ? c : SEG_THREADS;
  const size_t smem = (size_t)2 * (SEG_THREADS / cw) * c * 4;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(seg_max_parts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  seg_max_parts_kernel<<<dim3((unsigned)n_seg, (unsigned)parts), SEG_THREADS, smem, stream>>>(feats, offsets, pval, pidx,
                                                                                              n_seg, c, parts);
  const int64_t total = (int64_t)n_seg * c;
  seg_max_combine_kernel<<<(unsigned)cdiv(total, 256), 256, 0, stream>>>(pval, pidx, out, maxidx, total, c, parts);
  return check_launch("roipool_fp_ws");
}

int b2s_roipool_bp(float* d_feats, const int32_t* offsets, const int32_t* maxidx, const float* d_out,
                   int32_t n_seg, int32_t c, b2s_stream_t stream) {
  (void)offsets;
  int64_t total = (int64_t)n_seg * c;
  if (total <= 0) return B2S_OK;
  roipool_bp_kernel<<<(unsigned)cdiv(total, 256), 256, 0, stream>>>(d_feats, maxidx, d_out, total, c);
  return check_launch("roipool_bp");
}
int b2s_global_avg_pool_fp(const float* feats, const int32_t* offsets, float* out, int32_t n_seg, int32_t c,
                           b2s_stream_t s) {
  return ordered_sum(feats, offsets, out, n_seg, c, 1, s);
}
int b2s_global_avg_pool_bp(float* d_feats, const int32_t* offsets, const float* d_out, int32_t n_seg,
                           int32_t c, b2s_stream_t stream) {
  if (n_seg <= 0) return B2S_OK;
  gap_bp_kernel<<<seg_grid(n_seg), 256, 0, stream>>>(d_feats, offsets, d_out, n_seg, c);
  return check_launch("global_avg_pool_bp");
}

int b2s_get_iou(const int32_t* proposals_idx, const int32_t* proposals_offset,
                const int16_t* instance_labels, const int32_t* instance_pointnum,
                const float* mask_scores, float* proposals_iou, int32_t n_instance,
                int32_t n_proposal, b2s_stream_t stream) {
  if (n_instance <= 0 || n_proposal <= 0) return B2S_OK;
  iou_kernel<<<seg_grid(n_proposal), 256, 0, stream>>>(proposals_idx, proposals_offset, instance_labels,
                                                       instance_pointnum, mask_scores, proposals_iou,
                                                       n_instance, n_proposal);
  return check_launch("get_iou");
}

int b2s_get_mask_label(const int32_t* proposals_idx, const int32_t* proposals_offset,
                       const int16_t* instance_labels, const int16_t* instance_cls,
                       const float* proposals_iou, int32_t n_instance, int32_t n_proposal,
                       int32_t ignored_label, float iou_thr, uint8_t* mask_label,
                       uint8_t* mask_label_mask, b2s_stream_t stream) {
  if (n_proposal <= 0) return B2S_OK;
  mask_label_kernel<<<seg_grid(n_proposal), 256, 0, stream>>>(
      proposals_idx, proposals_offset, instance_labels, instance_cls, proposals_iou, n_instance,
      n_proposal, ignored_label, iou_thr, mask_label, mask_label_mask);
  return check_launch("get_mask_label");
}

int b2s_gather_rows(const float* feat, const int64_t* idx, int64_t n, int32_t c, float* out,
                    b2s_stream_t stream) {
  if (n <= 0) return B2S_OK;
  if ((c & 3) == 0) {
    int64_t total = n * (c / 4);
    gather_rows_kernel<<<(unsigned)cdiv(total, 1024), 256, 0, stream>>>(feat, idx, total, c / 4, (float4*)out);
  } else {
    int64_t total = n * c;
    gather_rows_scalar_kernel<<<(unsigned)cdiv(total, 256), 256, 0, stream>>>(feat, idx, total, c, out);
  }
  return check_launch("gather_rows");
}

int b2s_scatter_add_rows(const float* grad, const int64_t* idx, int64_t n, int32_t c, float* gfeat,
                         b2s_stream_t stream) {
  if (n <= 0) return B2S_OK;
  if ((c & 3) == 0) {
    int64_t total = n * (c / 4);
    scatter_add_rows_kernel<<<(unsigned)cdiv(total, 1024), 256, 0, stream>>>(grad, idx, total, c / 4, gfeat);
  } else {
    int64_t total = n * c;
    scatter_add_rows_scalar_kernel<<<(unsigned)cdiv(total, 256), 256, 0, stream>>>(grad, idx, total, c, gfeat);
  }
  return check_launch("scatter_add_rows");
}

int b2s_clusters_voxelize(const void* clusters_idx, int32_t idx_is_int64, const int32_t* clusters_offset,
                          int64_t sum_npoint, int32_t n_cluster, const float* coords, float scale,
                          int32_t spatial_shape, const float* rand6, int32_t* out_xyz, float* cluster_params,
                          b2s_stream_t stream) {
  if (sum_npoint < 0 || n_cluster < 0 || spatial_shape < 1) {
    set_error("clusters_voxelize: invalid argument");
    return B2S_E_INVALID;
  }
  if (sum_npoint == 0 || n_cluster == 0) return B2S_OK;
  const float shape = (float)spatial_shape;
  if (idx_is_int64) {
    cv_cluster_kernel<int64_t><<<n_cluster, CV_THREADS, 0, stream>>>((const int64_t*)clusters_idx, clusters_offset,
                                                                    n_cluster, coords, scale, shape, rand6, cluster_params);
    cv_point_kernel<int64_t><<<(unsigned)cdiv(sum_npoint, 256), 256, 0, stream>>>(
        (const int64_t*)clusters_idx, sum_npoint, coords, cluster_params, out_xyz);
  } else {
    cv_cluster_kernel<int32_t><<<n_cluster, CV_THREADS, 0, stream>>>((const int32_t*)clusters_idx, clusters_offset,
                                                                    n_cluster, coords, scale, shape, rand6, cluster_params);
    cv_point_kernel<int32_t><<<(unsigned)cdiv(sum_npoint, 256), 256, 0, stream>>>(
        (const int32_t*)clusters_idx, sum_npoint, coords, cluster_params, out_xyz);
  }
  return check_launch("clusters_voxelize");
}

}  // extern "C"
