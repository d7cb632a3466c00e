// T2 extension: mask-sorted tile order for the output-stationary 3^3 convolution.
//
// The tcgen05 kernel (conv_tc.cu) works on tiles of 128 output rows and skips a kernel offset only when NO
// row of the tile has a neighbour there.  Surface voxels have ~9 of 27 neighbours, but in first-occurrence
// row order (the order MinkowskiEngine defines) a tile mixes floors, walls and edges, so
// every tile sees all 27 offsets (measured on the benchmark batch: 26.96 active offsets per tile).  Grouping
// rows with similar neighbourhoods fixes that without touching the result: rows are sorted by
//     key = [which of the six faces (-x,+x,-y,+y,-z,+z) of the 3^3 stencil hold any neighbour] : [27-bit mask]
// and the convolution runs over the permuted table, storing row t of a tile at out[row_perm[t]].  On the
// benchmark batch this leaves 12.7 active offsets per tile (tools/tile_mask_probe.py), i.e. 2.1x fewer
// gather slabs and MMAs.  The output is the same sum in the same k order (skipped terms are exact zeros).
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace b2s {

constexpr int TO_ROWS = 256;  // rows per CTA in the key kernel

// bit sets of the 3^3 stencil faces, kidx = ix + 3*iy + 9*iz (x fastest, appendix A.4)
__host__ __device__ constexpr uint32_t face_bits(int axis, int side) {
  uint32_t m = 0;
  for (int k = 0; k < 27; ++k) {
    int i = axis == 0 ? k % 3 : (axis == 1 ? (k / 3) % 3 : k / 9);
    if (i == side) m |= 1u << k;
  }
  return m;
}

__global__ void __launch_bounds__(TO_ROWS)
    row_key_kernel(const int32_t* __restrict__ nbr, int n, uint32_t* __restrict__ keys, int32_t* __restrict__ iota) {
  // stage 256 rows x 27 entries coalesced, then one thread per row (stride 27 words: conflict-free)
  __shared__ int32_t s[TO_ROWS * 27];
  const int row0 = blockIdx.x * TO_ROWS;
  const int rows = min(TO_ROWS, n - row0);
  const int32_t* p = nbr + (int64_t)row0 * 27;
  for (int e = threadIdx.x; e < rows * 27; e += TO_ROWS) s[e] = __ldg(p + e);
  __syncthreads();
  const int r = threadIdx.x;
  if (r >= rows) return;
  uint32_t m = 0;
#pragma unroll
  for (int k = 0; k < 27; ++k) m |= (s[r * 27 + k] >= 0) ? (1u << k) : 0u;
  constexpr uint32_t XM = face_bits(0, 0), XP = face_bits(0, 2), YM = face_bits(1, 0), YP = face_bits(1, 2),
                     ZM = face_bits(2, 0), ZP = face_bits(2, 2);
  const uint32_t six = ((m & XM) ? 1u : 0u) | ((m & XP) ? 2u : 0u) | ((m & YM) ? 4u : 0u) | ((m & YP) ? 8u : 0u) |
                       ((m & ZM) ? 16u : 0u) | ((m & ZP) ? 32u : 0u);
  // the centre bit (kidx 13) is set for every row of a stride-1 map: dropping it makes the key 32 bits
  const uint32_t m26 = (m & 0x1FFFu) | ((m >> 14) << 13);
  keys[row0 + r] = (six << 26) | m26;
  iota[row0 + r] = row0 + r;
}

__global__ void __launch_bounds__(256)
    permute_nbr_kernel(const int32_t* __restrict__ nbr, const int32_t* __restrict__ perm, int64_t total, int K,
                       int32_t* __restrict__ nbr_sorted, uint32_t* __restrict__ tile_mask) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = e < total;
  int t = 0, k = 0, v = -1;
  if (valid) {
    t = (int)(e / K);
    k = (int)(e - (int64_t)t * K);
    v = __ldg(nbr + (int64_t)__ldg(perm + t) * K + k);
    nbr_sorted[e] = v;
  }
  // a warp covers 32 consecutive (row, offset) entries, i.e. at most two 128-row tiles
  const int tile = t >> 7;
  const int tile0 = __shfl_sync(0xffffffffu, tile, 0);
  const uint32_t bit = (valid && v >= 0) ? (1u << k) : 0u;
  const uint32_t m0 = __reduce_or_sync(0xffffffffu, tile == tile0 ? bit : 0u);
  const uint32_t m1 = __reduce_or_sync(0xffffffffu, tile != tile0 ? bit : 0u);
  if ((threadIdx.x & 31) == 0) {
    if (m0) atomicOr(tile_mask + tile0, m0);
    if (m1) atomicOr(tile_mask + tile0 + 1, m1);
  }
}

static size_t sort_tmp_bytes(int64_t n) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs((void*)nullptr, bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, (int)n, 0, 32);
  return bytes;
}

}  // namespace b2s

using namespace b2s;

extern "C" {

size_t b2s_tile_order_ws_bytes(int64_t n_out) {
  if (n_out < 1) n_out = 1;
  return 3 * align_up((size_t)n_out * 4) + align_up(sort_tmp_bytes(n_out)) + 1024;
}

int b2s_tile_order(const int32_t* nbr, int64_t n_out, int32_t K, int32_t* row_perm, int32_t* nbr_sorted,
                   uint32_t* tile_mask, void* ws, size_t ws_bytes, b2s_stream_t stream) {
  if (n_out < 0 || K != 27 || n_out > 0x7FFFFFFF / 32 || !nbr || !row_perm || !nbr_sorted || !tile_mask) {
    set_error("tile_order: invalid argument (needs the 3^3 neighbour table, K = 27)");
    return B2S_E_INVALID;
  }
  if (n_out == 0) return B2S_OK;
  Workspace w(ws, ws_bytes);
  uint32_t* keys = w.take<uint32_t>(n_out);
  uint32_t* skeys = w.take<uint32_t>(n_out);
  int32_t* iota = w.take<int32_t>(n_out);
  size_t tmp_bytes = sort_tmp_bytes(n_out);
  char* tmp = w.take<char>(tmp_bytes);
  if (!keys || !skeys || !iota || !tmp) {
    set_error("tile_order: workspace too small");
    return B2S_E_WORKSPACE;
  }
  row_key_kernel<<<(unsigned)cdiv(n_out, TO_ROWS), TO_ROWS, 0, stream>>>(nbr, (int)n_out, keys, iota);
  // stable: rows with equal keys stay in ascending (first-occurrence) order
  cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, skeys, iota, row_perm, (int)n_out, 0, 32, stream);
  if (e != cudaSuccess) {
    set_error("tile_order: radix sort failed");
    return B2S_E_LAUNCH;
  }
  cudaMemsetAsync(tile_mask, 0, (size_t)cdiv(n_out, (int64_t)128) * 4, stream);
  const int64_t total = n_out * K;
  permute_nbr_kernel<<<(unsigned)cdiv(total, 256), 256, 0, stream>>>(nbr, row_perm, total, K, nbr_sorted, tile_mask);
  return check_launch("tile_order");
}

}  // extern "C"
