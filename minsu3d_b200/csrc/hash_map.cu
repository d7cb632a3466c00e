// T1 / T2 / V1: coordinate hash, first-occurrence unique, kernel maps, canonical pair lists.
//
// Semantics restated from the MinkowskiEngine call sites in the reference
// (minsu3d/model/module/common.py:12-93, backbone.py:38, general_model.py:187-191,
//  data/dataset/general_dataset.py:159-163) and SURVEY.md appendix A items 1-6.
// Design: 64-bit packed keys in an open-addressing table (linear probing, load <= 0.5).
// All integer results are independent of thread scheduling: the winner among duplicate
// coordinates is chosen with atomicMin(row) (first occurrence), positions come from scans.
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace b2s {

static thread_local char g_err[256] = "";
void set_error(const char* msg) { snprintf(g_err, sizeof(g_err), "%s", msg); }

size_t scan_ws_bytes(int64_t n) {
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum((void*)nullptr, bytes, (const int32_t*)nullptr, (int32_t*)nullptr,
                                (int)(n > 0 ? n : 1));
  return align_up(bytes) + 256;
}

int exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, void* ws, size_t ws_bytes,
                       cudaStream_t stream) {
  if (n <= 0) return B2S_OK;
  size_t need = 0;
  cub::DeviceScan::ExclusiveSum((void*)nullptr, need, in, out, (int)n);
  if (need > ws_bytes) {
    set_error("scan workspace too small");
    return B2S_E_WORKSPACE;
  }
  cudaError_t e = cub::DeviceScan::ExclusiveSum(ws, need, in, out, (int)n, stream);
  if (e != cudaSuccess) {
    set_error(cudaGetErrorString(e));
    return B2S_E_LAUNCH;
  }
  return B2S_OK;
}

// ------------------------------------------------------------------------------------------
// insert: one thread per row.  vals[slot] = min row index holding that key.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) insert_kernel(const int4* __restrict__ coords, int n, int quant,
                                                     uint64_t* __restrict__ keys,
                                                     int32_t* __restrict__ vals, uint64_t mask,
                                                     int32_t* __restrict__ slot_of_row,
                                                     int32_t* __restrict__ d_count,
                                                     const int32_t* __restrict__ d_n) {
  int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n) return;
  if (d_n && row >= *d_n) {  // n is only an upper bound: the real row count lives on the device
    slot_of_row[row] = -1;
    return;
  }
  int4 c = coords[row];  // (b, x, y, z)
  if (quant > 1) {
    c.y = floor_div(c.y, quant) * quant;
    c.z = floor_div(c.z, quant) * quant;
    c.w = floor_div(c.w, quant) * quant;
  }
  if (!coord_in_range(c.x, c.y, c.z, c.w)) {
    d_count[1] = 1;
    slot_of_row[row] = -1;
    return;
  }
  uint64_t key = pack_coord(c.x, c.y, c.z, c.w);
  uint64_t slot = hash64(key) & mask;
  while (true) {
    unsigned long long prev =
        atomicCAS((unsigned long long*)(keys + slot), (unsigned long long)B2S_EMPTY_KEY,
                  (unsigned long long)key);
    if (prev == B2S_EMPTY_KEY || prev == key) {
      atomicMin(vals + slot, row);
      slot_of_row[row] = (int32_t)slot;
      return;
    }
    slot = (slot + 1) & mask;
  }
}

__global__ void __launch_bounds__(256) flag_kernel(int n, const int32_t* __restrict__ vals,
                                                   const int32_t* __restrict__ slot_of_row,
                                                   int32_t* __restrict__ flag) {
  int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n) return;
  int s = slot_of_row[row];
  flag[row] = (s >= 0 && vals[s] == row) ? 1 : 0;
}

__global__ void __launch_bounds__(256)
    emit_unique_kernel(const int4* __restrict__ coords, int n, int quant,
                       const int32_t* __restrict__ vals, const int32_t* __restrict__ slot_of_row,
                       const int32_t* __restrict__ rank, int32_t* __restrict__ unique_idx,
                       int32_t* __restrict__ inverse, int4* __restrict__ out_coords,
                       int32_t* __restrict__ d_count, const int32_t* __restrict__ d_n) {
  int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n) return;
  if (d_n && row >= *d_n) {
    if (row == n - 1) d_count[0] = rank[row];
    return;
  }
  int s = slot_of_row[row];
  int w = s >= 0 ? vals[s] : row;
  int r = rank[w];
  inverse[row] = s >= 0 ? r : -1;
  bool win = (s >= 0 && w == row);
  if (win) {
    int4 c = coords[row];
    if (quant > 1) {
      c.y = floor_div(c.y, quant) * quant;
      c.z = floor_div(c.z, quant) * quant;
      c.w = floor_div(c.w, quant) * quant;
    }
    unique_idx[r] = row;
    out_coords[r] = c;
  }
  if (row == n - 1) d_count[0] = rank[row] + (win ? 1 : 0);
}

// after emit: table value := unique row (so lookups return rows of the new map)
__global__ void __launch_bounds__(256) remap_vals_kernel(int n, int32_t* __restrict__ vals,
                                                         const int32_t* __restrict__ slot_of_row,
                                                         const int32_t* __restrict__ flag,
                                                         const int32_t* __restrict__ rank) {
  int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n) return;
  if (flag[row]) vals[slot_of_row[row]] = rank[row];
}

// ------------------------------------------------------------------------------------------
// kernel map: one thread per (output row, kernel offset); coalesced nbr writes.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    kernel_map_kernel(const int4* __restrict__ out_coords, int64_t total, int K, int ksize, int dil,
                      const uint64_t* __restrict__ keys, const int32_t* __restrict__ vals,
                      uint64_t mask, int32_t* __restrict__ nbr, uint32_t* __restrict__ tile_mask) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = t < total;
  int o = 0, kidx = 0, r = -1;
  if (valid) {
    o = (int)(t / K);
    kidx = (int)(t - (int64_t)o * K);
    int ix = kidx % ksize;
    int iy = (kidx / ksize) % ksize;
    int iz = kidx / (ksize * ksize);
    int lo = (ksize & 1) ? (ksize - 1) / 2 : 0;
    int4 c = __ldg(out_coords + o);
    int x = c.y + (ix - lo) * dil, y = c.z + (iy - lo) * dil, z = c.w + (iz - lo) * dil;
    if (coord_in_range(c.x, x, y, z)) r = hash_lookup(keys, vals, mask, pack_coord(c.x, x, y, z));
    nbr[t] = r;
  }
  if (tile_mask != nullptr) {
    // active-offset mask per 128-row tile (bit k set iff some row of the tile has a neighbour at offset k): a
    // warp covers 32 consecutive (row, offset) entries, i.e. at most two tiles
    const int tile = o >> 7;
    const int tile0 = __shfl_sync(0xffffffffu, tile, 0);
    const uint32_t bit = (valid && r >= 0) ? (1u << kidx) : 0u;
    const uint32_t m0 = __reduce_or_sync(0xffffffffu, tile == tile0 ? bit : 0u);
    const uint32_t m1 = __reduce_or_sync(0xffffffffu, tile != tile0 ? bit : 0u);
    if ((threadIdx.x & 31) == 0) {
      if (m0) atomicOr(tile_mask + tile0, m0);
      if (m1) atomicOr(tile_mask + tile0 + 1, m1);
    }
  }
}

// ------------------------------------------------------------------------------------------
// pair lists sorted by (kidx, out row): block-wise counts laid out [K][nblk] so that ONE
// exclusive scan yields every (kidx, block) base in canonical order.
// ------------------------------------------------------------------------------------------
#define PAIR_ROWS 256

__global__ void __launch_bounds__(256)
    pair_count_kernel(const int32_t* __restrict__ nbr, int n, int K, int nblk,
                      int32_t* __restrict__ block_counts) {
  extern __shared__ int32_t s_cnt[];  // K
  for (int k = threadIdx.x; k < K; k += blockDim.x) s_cnt[k] = 0;
  __syncthreads();
  int row0 = blockIdx.x * PAIR_ROWS;
  int rows = min(PAIR_ROWS, n - row0);
  const int32_t* p = nbr + (int64_t)row0 * K;
  int total = rows * K;
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    if (p[e] >= 0) atomicAdd(&s_cnt[e % K], 1);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x) block_counts[(int64_t)k * nblk + blockIdx.x] = s_cnt[k];
}

__global__ void __launch_bounds__(256)
    pair_fill_kernel(const int32_t* __restrict__ nbr, int n, int K, int nblk,
                     const int32_t* __restrict__ block_base, const int32_t* __restrict__ block_counts,
                     int64_t capacity, int32_t* __restrict__ pair_in, int32_t* __restrict__ pair_out,
                     int32_t* __restrict__ k_offsets, int32_t* __restrict__ d_count) {
  extern __shared__ int32_t s_tile[];  // PAIR_ROWS * K
  int row0 = blockIdx.x * PAIR_ROWS;
  int rows = min(PAIR_ROWS, n - row0);
  const int32_t* p = nbr + (int64_t)row0 * K;
  int total = rows * K;
  for (int e = threadIdx.x; e < total; e += blockDim.x) s_tile[e] = p[e];
  __syncthreads();
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  for (int k = warp; k < K; k += nwarp) {
    int64_t base = block_base[(int64_t)k * nblk + blockIdx.x];
    for (int r0 = 0; r0 < rows; r0 += 32) {
      int r = r0 + lane;
      int v = (r < rows) ? s_tile[r * K + k] : -1;
      unsigned m = __ballot_sync(0xffffffffu, v >= 0);
      if (v >= 0) {
        int64_t pos = base + __popc(m & ((1u << lane) - 1));
        if (pos < capacity) {
          pair_in[pos] = v;
          pair_out[pos] = row0 + r;
        }
      }
      base += __popc(m);
    }
  }
  if (blockIdx.x == 0) {
    for (int k = threadIdx.x; k < K; k += blockDim.x) k_offsets[k] = block_base[(int64_t)k * nblk];
    if (threadIdx.x == 0) {
      int64_t last = (int64_t)K * nblk - 1;
      int tot = block_base[last] + block_counts[last];
      k_offsets[K] = tot;
      d_count[0] = tot;
    }
  }
}

}  // namespace b2s

using namespace b2s;

extern "C" {

int b2s_version(void) { return 100; }
const char* b2s_last_error_string(void) { return g_err; }

int64_t b2s_hash_capacity(int64_t n) {
  int64_t cap = 1024;
  while (cap < 2 * n) cap <<= 1;
  return cap;
}

size_t b2s_coord_unique_ws_bytes(int64_t n) {
  if (n < 1) n = 1;
  return 3 * align_up((size_t)n * 4) + scan_ws_bytes(n) + 1024;
}

int b2s_coord_unique(const int32_t* coords, int64_t n, int32_t quant, uint64_t* table_keys,
                     int32_t* table_vals, int64_t cap, int32_t* unique_idx, int32_t* inverse,
                     int32_t* out_coords, int32_t* d_count, const int32_t* d_n, void* ws, size_t ws_bytes,
                     b2s_stream_t stream) {
  if (n < 0 || quant < 1 || cap < 2 * n || (cap & (cap - 1)) != 0 || n > 0x7fffffff / 32) {
    set_error("coord_unique: invalid n/quant/capacity");
    return B2S_E_INVALID;
  }
  cudaMemsetAsync(table_keys, 0xFF, (size_t)cap * 8, stream);
  cudaMemsetAsync(table_vals, 0x7F, (size_t)cap * 4, stream);
  cudaMemsetAsync(d_count, 0, 8, stream);
  if (n == 0) return check_launch("coord_unique(empty)");
  Workspace w(ws, ws_bytes);
  int32_t* slot_of_row = w.take<int32_t>(n);
  int32_t* flag = w.take<int32_t>(n);
  int32_t* rank = w.take<int32_t>(n);
  size_t sbytes = scan_ws_bytes(n);
  char* sws = w.take<char>(sbytes);
  if (!slot_of_row || !flag || !rank || !sws) {
    set_error("coord_unique: workspace too small");
    return B2S_E_WORKSPACE;
  }
  int grid = (int)cdiv(n, 256);
  insert_kernel<<<grid, 256, 0, stream>>>((const int4*)coords, (int)n, quant, table_keys, table_vals,
                                          (uint64_t)cap - 1, slot_of_row, d_count, d_n);
  flag_kernel<<<grid, 256, 0, stream>>>((int)n, table_vals, slot_of_row, flag);
  int rc = exclusive_scan_i32(flag, rank, n, sws, sbytes, stream);
  if (rc) return rc;
  emit_unique_kernel<<<grid, 256, 0, stream>>>((const int4*)coords, (int)n, quant, table_vals,
                                               slot_of_row, rank, unique_idx, inverse,
                                               (int4*)out_coords, d_count, d_n);
  remap_vals_kernel<<<grid, 256, 0, stream>>>((int)n, table_vals, slot_of_row, flag, rank);
  return check_launch("coord_unique");
}

int b2s_kernel_map(const int32_t* out_coords, int64_t n_out, int32_t ksize, int32_t dil,
                   const uint64_t* table_keys, const int32_t* table_vals, int64_t cap, int32_t* nbr,
                   uint32_t* tile_mask, b2s_stream_t stream) {
  if (n_out < 0 || ksize < 1 || ksize > 7 || dil < 1 || (cap & (cap - 1)) != 0) {
    set_error("kernel_map: invalid argument");
    return B2S_E_INVALID;
  }
  if (n_out == 0) return B2S_OK;
  int K = ksize * ksize * ksize;
  int64_t total = n_out * K;
  if (tile_mask != nullptr) {
    if (K > 32) {
      set_error("kernel_map: tile masks need ksize^3 <= 32");
      return B2S_E_INVALID;
    }
    cudaMemsetAsync(tile_mask, 0, (size_t)cdiv(n_out, (int64_t)128) * 4, stream);
  }
  kernel_map_kernel<<<(unsigned)cdiv(total, 256), 256, 0, stream>>>(
      (const int4*)out_coords, total, K, ksize, dil, table_keys, table_vals, (uint64_t)cap - 1, nbr, tile_mask);
  return check_launch("kernel_map");
}

size_t b2s_pairs_ws_bytes(int64_t n_out, int32_t K) {
  int64_t nblk = cdiv(n_out > 0 ? n_out : 1, PAIR_ROWS);
  int64_t cnt = nblk * K;
  return 2 * align_up((size_t)cnt * 4) + scan_ws_bytes(cnt) + 1024;
}

int b2s_pairs_from_nbr(const int32_t* nbr, int64_t n_out, int32_t K, int64_t pair_capacity,
                       int32_t* pair_in, int32_t* pair_out, int32_t* k_offsets, int32_t* d_count,
                       void* ws, size_t ws_bytes, b2s_stream_t stream) {
  if (n_out < 0 || K < 1 || K > 125) {
    set_error("pairs_from_nbr: invalid argument");
    return B2S_E_INVALID;
  }
  if (n_out == 0) {
    cudaMemsetAsync(k_offsets, 0, (size_t)(K + 1) * 4, stream);
    cudaMemsetAsync(d_count, 0, 4, stream);
    return check_launch("pairs_from_nbr(empty)");
  }
  int nblk = (int)cdiv(n_out, PAIR_ROWS);
  int64_t cnt = (int64_t)nblk * K;
  Workspace w(ws, ws_bytes);
  int32_t* bc = w.take<int32_t>(cnt);
  int32_t* bb = w.take<int32_t>(cnt);
  size_t sbytes = scan_ws_bytes(cnt);
  char* sws = w.take<char>(sbytes);
  if (!bc || !bb || !sws) {
    set_error("pairs_from_nbr: workspace too small");
    return B2S_E_WORKSPACE;
  }
  pair_count_kernel<<<nblk, 256, K * 4, stream>>>(nbr, (int)n_out, K, nblk, bc);
  int rc = exclusive_scan_i32(bc, bb, cnt, sws, sbytes, stream);
  if (rc) return rc;
  size_t smem = (size_t)PAIR_ROWS * K * 4;
  if (smem > 48 * 1024) {
    cudaFuncSetAttribute(pair_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
  pair_fill_kernel<<<nblk, 256, smem, stream>>>(nbr, (int)n_out, K, nblk, bb, bc, pair_capacity,
                                                pair_in, pair_out, k_offsets, d_count);
  return check_launch("pairs_from_nbr");
}

}  // extern "C"
