// SURVEY 8(f) #2 -- instance post-processing on the device.
//
// Reference: PointGroup._get_pred_instances / _get_nms_instances (minsu3d/model/pointgroup.py:197-265) and
// HAIS._get_pred_instances (minsu3d/model/hais.py:210-247) build a dense bool mask [nProposal, N] on the CPU,
// multiply it with its transpose (torch.mm) for the intersections and run a numpy NMS.  Here the proposals stay in
// their (proposal, point) pair-list form:
//   1. keys (point << 32 | proposal) of the valid pairs are radix-sorted, so the proposals of one point are adjacent
//      and duplicate pairs collapse (the dense mask is a set);
//   2. npoint[p] = number of distinct points of p (mask.sum(1));
//   3. every pair of proposals sharing a point gets +1 in inter[a, b] -- exactly (mask @ mask.T)[a, b], an integer;
//   4. iou = inter / (n_a + n_b - inter) in fp32 with the reference's operation order (pointgroup.py:236-241);
//   5. greedy NMS over the proposals in descending score order (pointgroup.py:209-218), one CTA.
// Integer results (npoint, inter, picks) are bit-exact; iou is the same fp32 expression.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace b2s {

constexpr uint64_t PP_INVALID = ~0ull;

__global__ void __launch_bounds__(256)
    pp_key_kernel(const int32_t* __restrict__ prop_idx, const uint8_t* __restrict__ valid, int64_t S,
                  uint64_t* __restrict__ keys) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S) return;
  const int32_t p = prop_idx[2 * i], pt = prop_idx[2 * i + 1];
  const bool ok = (valid == nullptr || valid[i] != 0) && p >= 0 && pt >= 0;
  keys[i] = ok ? (((uint64_t)(uint32_t)pt << 32) | (uint32_t)p) : PP_INVALID;
}

__global__ void __launch_bounds__(256)
    pp_npoint_kernel(const uint64_t* __restrict__ keys, int64_t S, int32_t* __restrict__ npoint) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S) return;
  const uint64_t k = keys[i];
  if (k == PP_INVALID || (i > 0 && keys[i - 1] == k)) return;  // invalid or duplicate pair
  atomicAdd(npoint + (uint32_t)k, 1);
}

// one thread per distinct (point, proposal) pair: +1 for itself and for every later proposal of the same point
__global__ void __launch_bounds__(256)
    pp_inter_kernel(const uint64_t* __restrict__ keys, int64_t S, const int32_t* __restrict__ remap, int P_kept,
                    int32_t* __restrict__ inter) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S) return;
  const uint64_t k = keys[i];
  if (k == PP_INVALID || (i > 0 && keys[i - 1] == k)) return;
  const int a = remap[(uint32_t)k];
  if (a < 0) return;
  atomicAdd(inter + (int64_t)a * P_kept + a, 1);
  const uint64_t point = k >> 32;
  uint64_t prev = k;
  for (int64_t j = i + 1; j < S; ++j) {
    const uint64_t kj = keys[j];
    if (kj == PP_INVALID || (kj >> 32) != point) break;
    if (kj == prev) continue;
    prev = kj;
    const int b = remap[(uint32_t)kj];
    if (b < 0) continue;
    atomicAdd(inter + (int64_t)a * P_kept + b, 1);
    atomicAdd(inter + (int64_t)b * P_kept + a, 1);
  }
}

__global__ void __launch_bounds__(256)
    pp_iou_kernel(const int32_t* __restrict__ inter, int P, float* __restrict__ iou) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)P * P) return;
  const int a = (int)(e / P), b = (int)(e - (int64_t)a * P);
  const float i = (float)inter[e];
  const float na = (float)inter[(int64_t)a * P + a], nb = (float)inter[(int64_t)b * P + b];
  iou[e] = __fdiv_rn(i, __fsub_rn(__fadd_rn(na, nb), i));  // intersection / (h + v - intersection)
}

// greedy NMS in the given order; suppressed[] lives in global scratch (P bytes)
__global__ void __launch_bounds__(1024)
    pp_nms_kernel(const float* __restrict__ iou, const int32_t* __restrict__ order, int P, float thr,
                  uint8_t* __restrict__ suppressed, int32_t* __restrict__ pick, int32_t* __restrict__ d_count) {
  __shared__ int s_npick;
  if (threadIdx.x == 0) s_npick = 0;
  for (int j = threadIdx.x; j < P; j += blockDim.x) suppressed[j] = 0;
  __syncthreads();
  for (int i = 0; i < P; ++i) {
    const int cur = order[i];
    if (suppressed[cur]) continue;  // uniform: written before the last barrier
    if (threadIdx.x == 0) pick[s_npick++] = cur;
    const float* row = iou + (int64_t)cur * P;
    for (int jj = i + 1 + threadIdx.x; jj < P; jj += blockDim.x) {
      const int o = order[jj];
      if (row[o] > thr) suppressed[o] = 1;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) d_count[0] = s_npick;
}

static size_t pp_sort_tmp(int64_t S) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortKeys((void*)nullptr, bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr, (int)S, 0, 64);
  return bytes;
}

}  // namespace b2s

using namespace b2s;

extern "C" {

size_t b2s_proposal_sort_ws_bytes(int64_t S) {
  if (S < 1) S = 1;
  return align_up((size_t)S * 8) + align_up(pp_sort_tmp(S)) + 1024;
}

int b2s_proposal_sort(const int32_t* prop_idx, const uint8_t* valid, int64_t S, uint64_t* keys_sorted, void* ws,
                      size_t ws_bytes, b2s_stream_t stream) {
  if (S < 0 || S > 0x7FFFFFFF || (S > 0 && (!prop_idx || !keys_sorted))) {
    set_error("proposal_sort: invalid argument");
    return B2S_E_INVALID;
  }
  if (S == 0) return B2S_OK;
  Workspace w(ws, ws_bytes);
  uint64_t* keys = w.take<uint64_t>(S);
  const size_t tmp_bytes = pp_sort_tmp(S);
  char* tmp = w.take<char>(tmp_bytes);
  if (!keys || !tmp) {
    set_error("proposal_sort: workspace too small");
    return B2S_E_WORKSPACE;
  }
  pp_key_kernel<<<(unsigned)cdiv(S, 256), 256, 0, stream>>>(prop_idx, valid, S, keys);
  size_t tb = tmp_bytes;
  cudaError_t e = cub::DeviceRadixSort::SortKeys(tmp, tb, keys, keys_sorted, (int)S, 0, 64, stream);
  if (e != cudaSuccess) {
    set_error("proposal_sort: radix sort failed");
    return B2S_E_LAUNCH;
  }
  return check_launch("proposal_sort");
}

int b2s_proposal_npoint(const uint64_t* keys_sorted, int64_t S, int32_t n_proposal, int32_t* npoint,
                        b2s_stream_t stream) {
  if (S < 0 || n_proposal < 0 || (n_proposal > 0 && !npoint)) {
    set_error("proposal_npoint: invalid argument");
    return B2S_E_INVALID;
  }
  if (n_proposal > 0) cudaMemsetAsync(npoint, 0, (size_t)n_proposal * 4, stream);
  if (S > 0) pp_npoint_kernel<<<(unsigned)cdiv(S, 256), 256, 0, stream>>>(keys_sorted, S, npoint);
  return check_launch("proposal_npoint");
}

int b2s_proposal_iou(const uint64_t* keys_sorted, int64_t S, const int32_t* remap, int32_t n_kept, int32_t* inter,
                     float* iou, b2s_stream_t stream) {
  if (S < 0 || n_kept < 0 || n_kept > 46340 || (n_kept > 0 && (!remap || !inter || !iou))) {
    set_error("proposal_iou: invalid argument");
    return B2S_E_INVALID;
  }
  if (n_kept == 0) return B2S_OK;
  const int64_t pp = (int64_t)n_kept * n_kept;
  cudaMemsetAsync(inter, 0, (size_t)pp * 4, stream);
  if (S > 0) pp_inter_kernel<<<(unsigned)cdiv(S, 256), 256, 0, stream>>>(keys_sorted, S, remap, n_kept, inter);
  pp_iou_kernel<<<(unsigned)cdiv(pp, 256), 256, 0, stream>>>(inter, n_kept, iou);
  return check_launch("proposal_iou");
}

int b2s_nms(const float* iou, const int32_t* order, int32_t n, float threshold, int32_t* pick, int32_t* d_count,
            void* ws, size_t ws_bytes, b2s_stream_t stream) {
  if (n < 0 || !d_count || (n > 0 && (!iou || !order || !pick))) {
    set_error("nms: invalid argument");
    return B2S_E_INVALID;
  }
  if (n == 0) {
    cudaMemsetAsync(d_count, 0, 4, stream);
    return check_launch("nms(empty)");
  }
  if (ws_bytes < (size_t)n) {
    set_error("nms: workspace too small");
    return B2S_E_WORKSPACE;
  }
  pp_nms_kernel<<<1, 1024, 0, stream>>>(iou, order, n, threshold, (uint8_t*)ws, pick, d_count);
  return check_launch("nms");
}

}  // extern "C"
