// C1: ballquery_batch_p on a uniform cell grid (cell edge = radius * (1 + 1e-5)).
//
// Reference: minsu3d/common_ops/src/bfs_cluster/bfs_cluster.cu:15-60 scans every point of the
// scene for every query (O(n * n_scene)).  Here points are bucketed into cells by a stable radix
// sort (so each cell's list is ascending in the original index), every query visits its 27 cells,
// and the per-query neighbour list is emitted in ascending index order -- the order the
// reference's ascending scan produces -- including the "lowest 1000" truncation
// (bfs_cluster.cu:38-43).  The distance predicate reproduces the reference SASS:
//     d2 = fma(dz, dz, fma(dx, dx, dy*dy));  keep iff d2 < radius*radius   (fp32)
// start = exclusive scan of len, so the CSR layout is deterministic (the reference's start is
// atomicAdd-arrival order, bfs_cluster.cu:48).
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace b2s {

constexpr int BQ_CAP = 1000;   // reference cap, bfs_cluster.cu:20,38
constexpr int BQ_STEP = 16;    // per-cell window per merge round
constexpr int BQ_CAND = 27 * BQ_STEP;
constexpr int BQ_SORT = 512;   // >= BQ_CAND, power of two
constexpr int BQ_WARPS = 8;
constexpr int BQ_DENSE = 48;   // lists at least this long are emitted by the bitmap kernel instead of the merge ...
constexpr int BQ_BM_WORDS = 2048;  // ... when their candidates span at most 65536 point indices (8 KB bitmap per warp)

__device__ __forceinline__ uint64_t cell_key(int b, int cx, int cy, int cz) {
  const int bias = 1 << 17;
  return ((uint64_t)(uint32_t)(b & 0x3FF) << 54) | ((uint64_t)(uint32_t)((cx + bias) & 0x3FFFF) << 36) |
         ((uint64_t)(uint32_t)((cy + bias) & 0x3FFFF) << 18) | (uint64_t)(uint32_t)((cz + bias) & 0x3FFFF);
}
__device__ __forceinline__ int cell_of(float v, double inv_cell) {
  double q = floor((double)v * inv_cell);
  q = fmin(fmax(q, -131000.0), 131000.0);
  return (int)q;
}

__global__ void __launch_bounds__(256)
    bq_keys_kernel(const float* __restrict__ xyz, const uint8_t* __restrict__ batch_idxs, int n,
                   double inv_cell, uint64_t* __restrict__ keys, int32_t* __restrict__ iota) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
  keys[i] = cell_key(batch_idxs[i], cell_of(x, inv_cell), cell_of(y, inv_cell), cell_of(z, inv_cell));
  iota[i] = i;
}

// sorted position i: build the packed candidate record (x,y,z,idx) and the cell table
__global__ void __launch_bounds__(256)
    bq_cells_kernel(const float* __restrict__ xyz, const uint64_t* __restrict__ skeys,
                    const int32_t* __restrict__ order, int n, float4* __restrict__ rec,
                    uint64_t* __restrict__ tkeys, int32_t* __restrict__ tstart,
                    int32_t* __restrict__ tend, uint64_t mask) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int p = order[i];
  rec[i] = make_float4(xyz[3 * p], xyz[3 * p + 1], xyz[3 * p + 2], __int_as_float(p));
  uint64_t key = skeys[i];
  bool first = (i == 0) || (skeys[i - 1] != key);
  bool last = (i == n - 1) || (skeys[i + 1] != key);
  if (!first && !last) return;
  // cooperative insert: whoever arrives first claims the slot, both write their field
  uint64_t slot = hash64(key) & mask;
  while (true) {
    unsigned long long prev = atomicCAS((unsigned long long*)(tkeys + slot),
                                        (unsigned long long)B2S_EMPTY_KEY, (unsigned long long)key);
    if (prev == B2S_EMPTY_KEY || prev == key) break;
    slot = (slot + 1) & mask;
  }
  if (first) tstart[slot] = i;
  if (last) tend[slot] = i + 1;
}

struct CellRange {
  int begin, end;
};
__device__ __forceinline__ CellRange find_cell(const uint64_t* __restrict__ tkeys,
                                               const int32_t* __restrict__ tstart,
                                               const int32_t* __restrict__ tend, uint64_t mask,
                                               uint64_t key) {
  uint64_t slot = hash64(key) & mask;
  CellRange r{0, 0};
  while (true) {
    uint64_t k = __ldg(tkeys + slot);
    if (k == key) {
      r.begin = __ldg(tstart + slot);
      r.end = __ldg(tend + slot);
      return r;
    }
    if (k == B2S_EMPTY_KEY) return r;
    slot = (slot + 1) & mask;
  }
}

__device__ __forceinline__ bool in_ball(float ox, float oy, float oz, float4 c, float r2) {
  float dx = ox - c.x, dy = oy - c.y, dz = oz - c.z;
  float d2 = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
  return d2 < r2;
}

// index span of a query's candidates in 32-bit words (cell lists are ascending: first / last entries bound it)
__device__ __forceinline__ int cand_span(const float4* __restrict__ rec, CellRange cr, int* lo_word_out) {
  int lo = 0x7fffffff, hi = -1;
  if (cr.begin < cr.end) {
    lo = __float_as_int(__ldg(&rec[cr.begin].w));
    hi = __float_as_int(__ldg(&rec[cr.end - 1].w));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  *lo_word_out = lo >> 5;
  return (hi >> 5) - (lo >> 5) + 1;
}

// ---- count: one warp per query, lanes 0..26 own one cell each ------------------------------
__global__ void __launch_bounds__(256)
    bq_count_kernel(const float* __restrict__ xyz, const uint8_t* __restrict__ batch_idxs, int n,
                    double inv_cell, float r2, const float4* __restrict__ rec,
                    const uint64_t* __restrict__ tkeys, const int32_t* __restrict__ tstart,
                    const int32_t* __restrict__ tend, uint64_t mask, int32_t* __restrict__ len) {
  int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (q >= n) return;
  float ox = xyz[3 * q], oy = xyz[3 * q + 1], oz = xyz[3 * q + 2];
  int b = batch_idxs[q];
  int cx = cell_of(ox, inv_cell), cy = cell_of(oy, inv_cell), cz = cell_of(oz, inv_cell);
  CellRange cr{0, 0};
  if (lane < 27) cr = find_cell(tkeys, tstart, tend, mask,
                                cell_key(b, cx + lane % 3 - 1, cy + (lane / 3) % 3 - 1, cz + lane / 9 - 1));
  int cnt = 0;
  for (int j = 0; j < 27; ++j) {
    int begin = __shfl_sync(0xffffffffu, cr.begin, j), end = __shfl_sync(0xffffffffu, cr.end, j);
    for (int i = begin + lane; i < end; i += 32) cnt += in_ball(ox, oy, oz, __ldg(rec + i), r2) ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) len[q] = min(cnt, BQ_CAP);
}

__global__ void __launch_bounds__(256)
    bq_startlen_kernel(const int32_t* __restrict__ len, const int32_t* __restrict__ start, int n,
                       int32_t* __restrict__ start_len, int32_t* __restrict__ d_count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  start_len[2 * i] = start[i];
  start_len[2 * i + 1] = len[i];
  if (i == n - 1) d_count[0] = start[i] + len[i];
}

// ---- fill: windowed 27-way merge, each round sorted in shared memory -------------------------
__global__ void __launch_bounds__(BQ_WARPS * 32)
    bq_fill_kernel(const float* __restrict__ xyz, const uint8_t* __restrict__ batch_idxs, int n,
                   double inv_cell, float r2, const float4* __restrict__ rec,
                   const uint64_t* __restrict__ tkeys, const int32_t* __restrict__ tstart,
                   const int32_t* __restrict__ tend, uint64_t mask,
                   const int32_t* __restrict__ start_len, int32_t* __restrict__ idx, int dense_min) {
  __shared__ int32_t s_cand[BQ_WARPS][BQ_CAND];
  __shared__ int32_t s_hit[BQ_WARPS][BQ_SORT];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int q = blockIdx.x * BQ_WARPS + warp;
  if (q >= n) return;
  const int need = start_len[2 * q + 1];
  if (need == 0) return;
  int32_t* out = idx + start_len[2 * q];
  float ox = xyz[3 * q], oy = xyz[3 * q + 1], oz = xyz[3 * q + 2];
  int b = batch_idxs[q];
  int cx = cell_of(ox, inv_cell), cy = cell_of(oy, inv_cell), cz = cell_of(oz, inv_cell);
  CellRange cr{0, 0};
  if (lane < 27) cr = find_cell(tkeys, tstart, tend, mask,
                                cell_key(b, cx + lane % 3 - 1, cy + (lane / 3) % 3 - 1, cz + lane / 9 - 1));
  if (need >= dense_min) {  // long list: the bitmap kernel emits it unless its candidates span too many indices
    int lw;
    if (cand_span(rec, cr, &lw) <= BQ_BM_WORDS) return;
  }
  int cur = cr.begin;
  const int end = cr.end;
  int32_t* cand = s_cand[warp];
  int32_t* hit = s_hit[warp];
  int written = 0;
  while (written < need) {
    // adaptive window: the BQ_CAND candidate slots of a round are shared by the cells that still have entries
    // (dense blobs of shifted coordinates sit in 1-8 cells: 54-432 entries per cell and round instead of 16)
    const unsigned alive = __ballot_sync(0xffffffffu, cur < end);
    if (alive == 0) break;
    const int step = BQ_CAND / __popc(alive);
    // window bound: only cells that extend past their window constrain the threshold
    int wend = min(cur + step, end);
    int hi = 0x7fffffff;
    if (wend < end) hi = __float_as_int(__ldg(&rec[wend - 1].w));
    int T = hi;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) T = min(T, __shfl_xor_sync(0xffffffffu, T, o));
    // entries of my window with index <= T: a prefix (lists are ascending) -> binary search
    int lo_i = cur, hi_i = wend;
    while (lo_i < hi_i) {
      int mid = (lo_i + hi_i) >> 1;
      if (__float_as_int(__ldg(&rec[mid].w)) <= T) lo_i = mid + 1; else hi_i = mid;
    }
    int take = lo_i - cur;
    // exclusive prefix over lanes
    int off = take;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int v = __shfl_up_sync(0xffffffffu, off, o);
      if (lane >= o) off += v;
    }
    int total = __shfl_sync(0xffffffffu, off, 31);
    off -= take;
    // candidate positions, written cooperatively cell by cell (a single dense cell may contribute the whole round)
    unsigned todo = __ballot_sync(0xffffffffu, take > 0);
    while (todo) {
      const int j = __ffs(todo) - 1;
      todo &= todo - 1;
      const int tj = __shfl_sync(0xffffffffu, take, j), oj = __shfl_sync(0xffffffffu, off, j),
                cj = __shfl_sync(0xffffffffu, cur, j);
      for (int e = lane; e < tj; e += 32) cand[oj + e] = cj + e;
    }
    cur += take;
    __syncwarp();
    if (total == 0) break;  // nothing left anywhere (cannot happen while written < need)
    // distance test + compaction
    int nhit = 0;
    for (int c0 = 0; c0 < total; c0 += 32) {
      int c = c0 + lane;
      bool ok = false;
      int pidx = 0;
      if (c < total) {
        float4 r = __ldg(rec + cand[c]);
        ok = in_ball(ox, oy, oz, r, r2);
        pidx = __float_as_int(r.w);
      }
      unsigned m = __ballot_sync(0xffffffffu, ok);
      if (ok) hit[nhit + __popc(m & ((1u << lane) - 1))] = pidx;
      nhit += __popc(m);
    }
    // bitonic sort of hit[0..S)
    int S = 32;
    while (S < nhit) S <<= 1;
    for (int i = nhit + lane; i < S; i += 32) hit[i] = 0x7fffffff;
    __syncwarp();
    for (int k = 2; k <= S; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = lane; i < S; i += 32) {
          int p = i ^ j;
          if (p > i) {
            int a = hit[i], bb = hit[p];
            bool up = ((i & k) == 0);
            if ((a > bb) == up) {
              hit[i] = bb;
              hit[p] = a;
            }
          }
        }
        __syncwarp();
      }
    }
    int w = min(nhit, need - written);
    for (int i = lane; i < w; i += 32) out[written + i] = hit[i];
    written += w;
    __syncwarp();
  }
}

// ---- fill, dense lists: bitmap over the point-index range of the candidates -------------------
// The output must be ascending in the point index and keep the LOWEST 1000 (bfs_cluster.cu:36-44).  For long
// lists (shifted coordinates: hundreds of neighbours per point) sorting rounds dominate the merge above; here
// every in-radius candidate sets one bit of a per-warp bitmap over [lowest, highest] candidate index (the cell
// lists are ascending, so the bounds are their first / last entries) and the set bits are enumerated in order.
__global__ void bq_fill_bitmap_kernel(const float* __restrict__ xyz, const uint8_t* __restrict__ batch_idxs, int n,
                                      double inv_cell, float r2, const float4* __restrict__ rec,
                                      const uint64_t* __restrict__ tkeys, const int32_t* __restrict__ tstart,
                                      const int32_t* __restrict__ tend, uint64_t mask,
                                      const int32_t* __restrict__ start_len, int32_t* __restrict__ idx,
                                      int dense_min) {
  extern __shared__ uint32_t s_bits[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * (blockDim.x >> 5) + warp;
  if (q >= n) return;
  const int need = start_len[2 * q + 1];
  if (need < dense_min) return;
  uint32_t* bm = s_bits + (size_t)warp * BQ_BM_WORDS;
  int32_t* out = idx + start_len[2 * q];
  const float ox = xyz[3 * q], oy = xyz[3 * q + 1], oz = xyz[3 * q + 2];
  const int b = batch_idxs[q];
  const int cx = cell_of(ox, inv_cell), cy = cell_of(oy, inv_cell), cz = cell_of(oz, inv_cell);
  CellRange cr{0, 0};
  if (lane < 27) cr = find_cell(tkeys, tstart, tend, mask,
                                cell_key(b, cx + lane % 3 - 1, cy + (lane / 3) % 3 - 1, cz + lane / 9 - 1));
  int lo_word;
  const int nwords = cand_span(rec, cr, &lo_word);
  if (nwords > BQ_BM_WORDS) return;  // handled by the merge kernel
  for (int w = lane; w < nwords; w += 32) bm[w] = 0u;
  __syncwarp();
  for (int j = 0; j < 27; ++j) {
    const int begin = __shfl_sync(0xffffffffu, cr.begin, j), end = __shfl_sync(0xffffffffu, cr.end, j);
    for (int i = begin + lane; i < end; i += 32) {
      const float4 r = __ldg(rec + i);
      if (in_ball(ox, oy, oz, r, r2)) {
        const int p = __float_as_int(r.w);
        atomicOr(bm + ((p >> 5) - lo_word), 1u << (p & 31));
      }
    }
  }
  __syncwarp();
  int written = 0;
  for (int w0 = 0; w0 < nwords && written < need; w0 += 32) {
    uint32_t word = (w0 + lane < nwords) ? bm[w0 + lane] : 0u;
    const int cnt = __popc(word);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    int pos = written + incl - cnt;
    const int base_idx = (lo_word + w0 + lane) << 5;
    while (word != 0u && pos < need) {
      const int bit = __ffs(word) - 1;
      word &= word - 1;
      out[pos++] = base_idx + bit;
    }
    written += total;
  }
}

__global__ void bq_zero_kernel(int n, int32_t* __restrict__ start_len, int32_t* __restrict__ d_count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 2 * n) start_len[i] = 0;
  if (i == 0) d_count[0] = 0;
}

struct BqWs {
  uint64_t *keys, *skeys, *tkeys;
  int32_t *iota, *order, *tstart, *tend, *len, *start;
  float4* rec;
  char* tmp;
  size_t tmp_bytes;
  int64_t cap;
};

static size_t bq_tmp_bytes(int64_t n) {
  size_t sort_bytes = 0;
  cub::DeviceRadixSort::SortPairs((void*)nullptr, sort_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, (int)(n > 0 ? n : 1));
  return std::max(align_up(sort_bytes), scan_ws_bytes(n)) + 256;
}

static bool bq_carve(void* ws, size_t ws_bytes, int64_t n, BqWs& w) {
  Workspace a(ws, ws_bytes);
  w.cap = b2s_hash_capacity(n);
  w.keys = a.take<uint64_t>(n);
  w.skeys = a.take<uint64_t>(n);
  w.tkeys = a.take<uint64_t>(w.cap);
  w.iota = a.take<int32_t>(n);
  w.order = a.take<int32_t>(n);
  w.tstart = a.take<int32_t>(w.cap);
  w.tend = a.take<int32_t>(w.cap);
  w.len = a.take<int32_t>(n);
  w.start = a.take<int32_t>(n);
  w.rec = a.take<float4>(n);
  w.tmp_bytes = bq_tmp_bytes(n);
  w.tmp = a.take<char>(w.tmp_bytes);
  return w.tmp != nullptr;
}

}  // namespace b2s

using namespace b2s;

extern "C" {

size_t b2s_ballquery_ws_bytes(int64_t n) {
  if (n < 1) n = 1;
  int64_t cap = b2s_hash_capacity(n);
  return align_up(n * 8) * 2 + align_up(cap * 8) + align_up(n * 4) * 4 + align_up(cap * 4) * 2 +
         align_up(n * 16) + align_up(bq_tmp_bytes(n)) + 4096;
}

int b2s_ballquery_count(const float* xyz, const uint8_t* batch_idxs, const int32_t* batch_offsets,
                        int64_t n, int32_t n_batch, float radius, int32_t* start_len,
                        int32_t* d_count, void* ws, size_t ws_bytes, b2s_stream_t stream) {
  (void)batch_offsets;
  (void)n_batch;
  if (n < 0 || n > (1 << 30)) {
    set_error("ballquery: invalid n");
    return B2S_E_INVALID;
  }
  if (n == 0) {
    cudaMemsetAsync(d_count, 0, 4, stream);
    return check_launch("ballquery_count(empty)");
  }
  if (!(radius > 0.f)) {  // d2 < r*r never holds
    bq_zero_kernel<<<(unsigned)cdiv(2 * n, 256), 256, 0, stream>>>((int)n, start_len, d_count);
    return check_launch("ballquery_count(r<=0)");
  }
  BqWs w;
  if (!bq_carve(ws, ws_bytes, n, w)) {
    set_error("ballquery: workspace too small");
    return B2S_E_WORKSPACE;
  }
  const double inv_cell = 1.0 / ((double)radius * (1.0 + 1e-5));
  const float r2 = radius * radius;
  int grid = (int)cdiv(n, 256);
  bq_keys_kernel<<<grid, 256, 0, stream>>>(xyz, batch_idxs, (int)n, inv_cell, w.keys, w.iota);
  size_t sort_bytes = w.tmp_bytes;
  cudaError_t e = cub::DeviceRadixSort::SortPairs(w.tmp, sort_bytes, w.keys, w.skeys, w.iota, w.order,
                                                  (int)n, 0, 64, stream);
  if (e != cudaSuccess) {
    set_error(cudaGetErrorString(e));
    return B2S_E_LAUNCH;
  }
  cudaMemsetAsync(w.tkeys, 0xFF, (size_t)w.cap * 8, stream);
  bq_cells_kernel<<<grid, 256, 0, stream>>>(xyz, w.skeys, w.order, (int)n, w.rec, w.tkeys, w.tstart,
                                            w.tend, (uint64_t)w.cap - 1);
  bq_count_kernel<<<(unsigned)cdiv(n * 32, 256), 256, 0, stream>>>(
      xyz, batch_idxs, (int)n, inv_cell, r2, w.rec, w.tkeys, w.tstart, w.tend, (uint64_t)w.cap - 1, w.len);
  int rc = exclusive_scan_i32(w.len, w.start, n, w.tmp, w.tmp_bytes, stream);
  if (rc) return rc;
  bq_startlen_kernel<<<grid, 256, 0, stream>>>(w.len, w.start, (int)n, start_len, d_count);
  return check_launch("ballquery_count");
}

int b2s_ballquery_fill(const float* xyz, const uint8_t* batch_idxs, const int32_t* batch_offsets,
                       int64_t n, int32_t n_batch, float radius, const int32_t* start_len,
                       int32_t* idx, void* ws, size_t ws_bytes, b2s_stream_t stream) {
  (void)batch_offsets;
  (void)n_batch;
  if (n <= 0 || !(radius > 0.f)) return B2S_OK;
  BqWs w;
  if (!bq_carve(ws, ws_bytes, n, w)) {
    set_error("ballquery: workspace too small");
    return B2S_E_WORKSPACE;
  }
  const double inv_cell = 1.0 / ((double)radius * (1.0 + 1e-5));
  const float r2 = radius * radius;
  // long lists whose candidates span <= 65536 point indices are emitted by the bitmap kernel, the rest by the merge
  const int dense_min = BQ_DENSE;
  bq_fill_kernel<<<(unsigned)cdiv(n, BQ_WARPS), BQ_WARPS * 32, 0, stream>>>(
      xyz, batch_idxs, (int)n, inv_cell, r2, w.rec, w.tkeys, w.tstart, w.tend, (uint64_t)w.cap - 1,
      start_len, idx, dense_min);
  {
    const size_t smem = (size_t)BQ_WARPS * BQ_BM_WORDS * 4;
    static bool configured = false;
    if (!configured) {
      cudaFuncSetAttribute(bq_fill_bitmap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      configured = true;
    }
    bq_fill_bitmap_kernel<<<(unsigned)cdiv(n, BQ_WARPS), BQ_WARPS * 32, smem, stream>>>(
        xyz, batch_idxs, (int)n, inv_cell, r2, w.rec, w.tkeys, w.tstart, w.tend, (uint64_t)w.cap - 1, start_len,
        idx, dense_min);
  }
  return check_launch("ballquery_fill");
}

}  // extern "C"
