// S1-S3, I1-I2, V2: segmented reductions, proposal x instance IoU, mask labels, row gather/scatter.
//
// Reference kernels: minsu3d/common_ops/src/sec_mean/sec_mean.cu:12-85,
// roipool/roipool.cu:12-120, get_iou/get_iou.cu:12-37,
// cal_iou_and_masklabel/cal_iou_and_masklabel.cu:14-140.
// The reference runs one thread per (segment, channel) with a serial scan over the segment
// (C=3 -> 3 active threads per block).  Here the loads are block-parallel and coalesced; the
// sums that the reference defines sequentially (sec_mean's sum(x/count), global_avg_pool's
// sum(x)/n) keep their sequential fp32 addition order, so results are bit-identical.
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
        for (int r = b + ph; r < e; r += phases) {
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
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  int64_t r = t / c4;
  int q = (int)(t - r * c4);
  out[t] = __ldg((const float4*)feat + idx[r] * c4 + q);
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
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  int64_t r = t / c4;
  int q = (int)(t - r * c4);
  float4 v = __ldg((const float4*)grad + t);
  float* dst = gfeat + (idx[r] * c4 + q) * 4;
  // 16-byte vector reduction (sm_90+): one L2 atomic per 4 channels
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
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
    gather_rows_kernel<<<(unsigned)cdiv(total, 256), 256, 0, stream>>>(feat, idx, total, c / 4, (float4*)out);
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
    scatter_add_rows_kernel<<<(unsigned)cdiv(total, 256), 256, 0, stream>>>(grad, idx, total, c / 4, gfeat);
  } else {
    int64_t total = n * c;
    scatter_add_rows_scalar_kernel<<<(unsigned)cdiv(total, 256), 256, 0, stream>>>(grad, idx, total, c, gfeat);
  }
  return check_launch("scatter_add_rows");
}

}  // extern "C"
