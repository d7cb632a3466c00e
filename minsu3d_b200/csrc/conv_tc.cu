// T3 / T4 on the 5th-generation tensor cores: output-stationary implicit GEMM with tcgen05.mma
// (kind::tf32), accumulators in TMEM, weights staged by the TMA engine (cp.async.bulk), gathered
// activations staged by producer warps, mbarrier pipelines between the three roles.
//
//   out[o, :] = sum_k in[nbr[o,k], :] @ W[k]        (appendix A.7; common.py:12,37,40,69,77)
//
// GEMM view per CTA: D[128 x Cout] (TMEM, fp32) += A[128 x 16] (TMEM) * B[16 x Cout] (smem) for every
// "slab" = (active kernel offset k, 16-channel chunk c).  A slabs are gathered rows (zero rows where the
// neighbour is missing); offsets with no neighbour in the whole tile are skipped.  The gathered A operand
// is written straight from registers into TENSOR MEMORY (tcgen05.st) and consumed by the TS form of
// tcgen05.mma: the first version staged A in shared memory and was bound by shared-memory bandwidth
// (ncu: LSU + tensor-core smem wavefronts at 97 %, profiles/r01_conv_tc_smem_bound.txt) because the
// 3xTF32 split re-reads every operand three times; with A in TMEM only the small weight tile is read
// from shared memory.
//
// Precision: the reference computes in fp32.  NSPLIT = 3 runs the 3xTF32 split (x = hi + lo with
// hi = x truncated to TF32; D += Ahi*Bhi + Ahi*Blo + Alo*Bhi), which is fp32-class (~1e-6 relative);
// NSPLIT = 1 is plain TF32 (~5e-4 relative) for callers that accept it.
//
// Shared-memory layout of the B (weight) operand: K-major, SWIZZLE_64B (a slab row is 16 fp32 = 64 bytes):
//   byte(row r, k j) = r*64 + (((j>>2) ^ ((r>>1)&3)) << 4) + (j&3)*4        [Swizzle<2,4,3>]
// Weights are pre-packed by pack_weights_kernel into exactly this byte image per slab, so one 1-D bulk
// copy (UBLKCP) per slab moves them, no tensor map needed.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace b2s {

// Round-2 experiment switches (compile time, `B2S_NVCC_EXTRA="-DB2S_TC_MIN_CTAS=4 -DB2S_TC_IDX_AHEAD=1" python
// minsu3d_b200/csrc/build.py --force`); the defaults reproduce the measured round-1 kernel exactly.
//   B2S_TC_MIN_CTAS : CTAs per SM the register allocation aims for (4 -> 80 registers, 64-80 B of spills; combine with
//                     B2S_TC_SB=16 at run time so that four weight rings fit the shared memory)
//   B2S_TC_IDX_AHEAD: producers read the NEXT offset's gather index from shared memory one slab ahead of its use
//                     (profiles/r01_conv_tc_sorted_stalls.txt: 11 % of the samples wait on that LDS)
#ifndef B2S_TC_MIN_CTAS
#define B2S_TC_MIN_CTAS 3
#endif
#ifndef B2S_TC_IDX_AHEAD
#define B2S_TC_IDX_AHEAD 0
#endif

// ---------------------------------------------------------------------------------------------
// weight packing: W[K][c_in][c_out] (or its transpose view) -> per-slab SW64 images, hi then lo
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    pack_weights_kernel(const float* __restrict__ W, float* __restrict__ Bp, int K, int c_in, int c_out,
                        int w_transposed, int64_t total) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  if (gridDim.y == 2) {
    // both orientations in one launch (b2s_conv_pack): y = 0 the forward image pair, y = 1 the transposed pair for the
    // data gradient, which is a c_out -> c_in product: roles of the two channel counts swap
    Bp += (int64_t)blockIdx.y * 2 * total;
    if (blockIdx.y == 1) {
      w_transposed = 1;
      const int t = c_in;
      c_in = c_out;
      c_out = t;
    }
  }
  const int per_slab = c_out * 16;
  int64_t slab = e / per_slab;
  int within = (int)(e - slab * per_slab);
  int n = within >> 4, j = within & 15;
  int nc = c_in >> 4;
  int k = (int)(slab / nc), c = (int)(slab - (int64_t)k * nc);
  int ci = c * 16 + j;
  const float* Wk = W + (int64_t)k * c_in * c_out;
  float v = w_transposed ? Wk[(int64_t)n * c_in + ci] : Wk[(int64_t)ci * c_out + n];
  float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
  float lo = v - hi;
  int64_t off = slab * per_slab + (sw64_offset(n, j) >> 2);
  Bp[off] = hi;
  Bp[total + off] = lo;
}

// what a missing neighbour reads (keeps the gather loop branch-free); 256 floats = widest supported row
__device__ float g_zero_row[256];

struct TcArgs {
  const float* A;
  const float* Bp;       // packed weights: [hi image | lo image], each K*(c_in/16) slabs of c_out*16 floats
  const int32_t* idx;    // TABLE: nbr [n_out, K] (or NULL = identity, K == 1); PAIRS: src
  const int32_t* dst;    // PAIRS: destination rows
  const int32_t* k_offsets;
  const uint32_t* tile_mask;  // TABLE, optional: active-offset mask per tile (b2s_kernel_map / b2s_tile_order)
  const int32_t* out_rows;    // TABLE, optional: tile row t is stored at out[out_rows[t]] (b2s_tile_order)
  float* out;
  int64_t n_out;
  int64_t bp_half;       // floats in one image
  int K, c_in, c_out, k_reversed, stages, tmem_cols, a_col0;
  int sb;  // weight-slab ring depth (shared memory), decoupled from the A stages (tensor memory)
};

template <bool PAIRS, int NSPLIT>
__global__ void __launch_bounds__(TC_THREADS, B2S_TC_MIN_CTAS) conv_tc_kernel(const TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = a.stages, SB = a.sb;
  constexpr int NB = NSPLIT == 3 ? 2 : 1;        // weight images per slab (hi, lo)
  constexpr int A_COLS = NSPLIT == 3 ? 32 : 16;  // tensor-memory columns of one A stage (hi, lo)
  const int b_slab = a.c_out * 64;               // bytes of one weight slab image
  const int b_slot = NB * b_slab;                // bytes of one ring slot
  // shared memory: [SB weight-slab slots][barriers][tmem ptr][mask][index tile]
  // tensor memory: [accumulator: c_out columns][S stages x A_COLS columns of the gathered A operand]
  // The weight ring is as deep as ~56 KB allows and has its own barriers: the TMA bulk copies have ~1-2 us
  // latency, so they must run many slabs ahead of the MMA (for C = 16 the whole layer fits: every weight byte
  // is fetched once per CTA).
  uint64_t* bars = (uint64_t*)(sm + (size_t)SB * b_slot);
  const uint32_t bar0 = base + (uint32_t)SB * b_slot;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (S + s); };
  auto bfull_bar = [&](int j) { return bar0 + 8u * (2 * S + j); };
  auto bempty_bar = [&](int j) { return bar0 + 8u * (2 * S + SB + j); };
  const uint32_t tmem_full_bar = bar0 + 8u * (2 * S + 2 * SB);
  uint32_t* s_tmem = (uint32_t*)(bars + 2 * S + 2 * SB + 1);
  uint32_t* s_mask = s_tmem + 1;
  // 16-byte aligned (base is 1024-aligned); plain pointer arithmetic on the shared array keeps the address space
  int32_t* s_idx = (int32_t*)(sm + (((size_t)SB * b_slot + 8u * (2 * S + 2 * SB + 1) + 16u + 15u) & ~(size_t)15));

  // ---- tile -> rows ------------------------------------------------------------------------
  int64_t row0 = 0;
  int rows = 0, k_single = -1, p0 = 0;
  const int K = a.K;
  if (!PAIRS) {
    row0 = (int64_t)blockIdx.x * TC_BM;
    rows = (int)min((int64_t)TC_BM, a.n_out - row0);
  } else {
    int chunk = blockIdx.x, kk, begin = 0, end = 0;
    bool found = false;
    for (kk = 0; kk < K; ++kk) {
      begin = a.k_offsets[kk];
      end = a.k_offsets[kk + 1];
      int nch = (end - begin + TC_BM - 1) / TC_BM;
      if (chunk < nch) {
        found = true;
        break;
      }
      chunk -= nch;
    }
    if (!found) return;
    k_single = kk;
    p0 = begin + chunk * TC_BM;
    rows = min(TC_BM, end - p0);
  }

  // ---- prologue ----------------------------------------------------------------------------
  {
    // one barrier per thread (a serial loop over ~60 barriers by one thread cost ~1.8k cycles per tile)
    const int nbar = 2 * S + 2 * SB + 1;
    if (tid < nbar) mbar_init(bar0 + 8u * tid, (tid < S) ? 4u : 1u);  // full[s]: one arrival per producer warp
    if (tid == 0) *s_mask = 0;
    if (tid < nbar) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) tmem_alloc(smem_u32(s_tmem), (uint32_t)a.tmem_cols);
  // Stage the tile's gather indices for all 128 rows (-1 = no row: the producers then read the zero row, so
  // their loop has no row/tile-edge branches) and build the active-offset mask from the same registers
  // (or take it from b2s_kernel_map's per-tile masks).
  uint32_t pre_mask = 0;
  if (!PAIRS) {
    if (a.idx != nullptr) {
      const int32_t* p = a.idx + row0 * K;
      const int total = rows * K;
      const bool scan = (a.tile_mask == nullptr);
      if (!scan) pre_mask = __ldg(a.tile_mask + blockIdx.x);
      uint32_t m = 0;
      if ((((uintptr_t)p) & 15) == 0) {
        const int n4 = total >> 2;
        constexpr int U = 5;  // 128 * 27 / 4 = 864 int4 <= 192 * 5; all loads issued before the first store
        int4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          int e4 = tid + u * TC_THREADS;
          if (e4 < n4) v[u] = __ldg((const int4*)p + e4);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          int e4 = tid + u * TC_THREADS;
          if (e4 < n4) {
            ((int4*)s_idx)[e4] = v[u];
            if (scan) {
              int k0 = (e4 * 4) % K;
              const int vv[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                m |= (vv[i] >= 0) ? (1u << k0) : 0u;
                k0 = (k0 + 1 == K) ? 0 : k0 + 1;
              }
            }
          }
        }
        for (int e4 = tid + U * TC_THREADS; e4 < n4; e4 += TC_THREADS) {  // K > 27 only
          int4 w = __ldg((const int4*)p + e4);
          ((int4*)s_idx)[e4] = w;
          const int vv[4] = {w.x, w.y, w.z, w.w};
          for (int i = 0; i < 4; ++i) m |= (vv[i] >= 0) ? (1u << ((e4 * 4 + i) % K)) : 0u;
        }
        for (int e = (n4 << 2) + tid; e < total; e += TC_THREADS) {
          int g = __ldg(p + e);
          s_idx[e] = g;
          m |= (g >= 0) ? (1u << (e % K)) : 0u;
        }
      } else {
        for (int e = tid; e < total; e += TC_THREADS) {
          int g = __ldg(p + e);
          s_idx[e] = g;
          m |= (g >= 0) ? (1u << (e % K)) : 0u;
        }
      }
      for (int e = total + tid; e < TC_BM * K; e += TC_THREADS) s_idx[e] = -1;  // rows past the end of the table
      if (scan) {
        m = __reduce_or_sync(0xffffffffu, m);
        if (lane == 0 && m) atomicOr(s_mask, m);
      }
    } else {
      if (tid < TC_BM) s_idx[tid] = (tid < rows) ? (int)(row0 + tid) : -1;  // identity map (1x1 conv), K == 1
    }
  } else {
    for (int e = tid; e < TC_BM; e += TC_THREADS) {
      s_idx[e] = (e < rows) ? a.idx[p0 + e] : -1;
      s_idx[TC_BM + e] = (e < rows) ? a.dst[p0 + e] : -1;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t kmask;
  if (PAIRS) kmask = 1u << k_single;
  else if (a.idx == nullptr) kmask = 1u;
  else kmask = (a.tile_mask != nullptr) ? pre_mask : *s_mask;
  const uint32_t tmem_base = *s_tmem;
  const int nc = a.c_in >> 4;
  const int T = __popc(kmask) * nc;  // slabs of this tile

  if (warp < 4) {
    // =========================== gather producers (thread = tile row = TMEM lane) ============
    // Three register slots rotate (slab t+2 is loaded while slab t is written to tensor memory): two slabs of
    // gathered data (2 x 64 B per thread) are in flight.  The loop body is branch-free apart from the uniform
    // trip-count tests: a warp issues in order, so every divergent branch / dependent latency in the per-slab
    // chain directly lengthens the slab period.
    const int r = tid;
    // mask-sorted tiles (tile_order.cu): the destination row comes from the permutation; loaded here so that
    // its latency is hidden behind the main loop
    int64_t orow = -1;
    int orow_perm = -1;
    if (!PAIRS && a.out_rows != nullptr && r < rows) orow_perm = __ldg(a.out_rows + row0 + r);
    const float* __restrict__ Ag = a.A;
    const int c_in = a.c_in;
    const int32_t* my_idx = s_idx + (PAIRS ? r : r * K);
    uint32_t km = kmask;
    int lk = 0, lc = 0;  // kernel offset / channel chunk of the NEXT slab to load
    float ra[16], rb[16], rc[16];
#if B2S_TC_IDX_AHEAD
    int g_cur = -1;
    int g_ahead = my_idx[PAIRS ? 0 : max(__ffs(km) - 1, 0)];  // index of the first offset
#endif
    auto load_next = [&](float (&dst)[16]) {
      const bool adv = (lc == 0);
      const int kn = __ffs(km) - 1;
      lk = adv ? kn : lk;
      km = adv ? (km & (km - 1)) : km;
#if B2S_TC_IDX_AHEAD
      g_cur = adv ? g_ahead : g_cur;
      const int g = g_cur;
      // the index of the offset after this one, read one slab ahead of its use
      g_ahead = my_idx[PAIRS ? 0 : max(__ffs(km) - 1, 0)];
#else
      const int g = my_idx[PAIRS ? 0 : lk];
#endif
      const float* p = (g >= 0) ? Ag + ((int64_t)g * c_in + lc * 16) : g_zero_row;
      // two 256-bit loads (LDG.E.256, sm_100) per slab and thread
      asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=f"(dst[0]), "=f"(dst[1]), "=f"(dst[2]), "=f"(dst[3]), "=f"(dst[4]), "=f"(dst[5]), "=f"(dst[6]),
                     "=f"(dst[7])
                   : "l"(p));
      asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=f"(dst[8]), "=f"(dst[9]), "=f"(dst[10]), "=f"(dst[11]), "=f"(dst[12]), "=f"(dst[13]),
                     "=f"(dst[14]), "=f"(dst[15])
                   : "l"(p + 8));
      lc = (lc + 1 == nc) ? 0 : lc + 1;
    };
    const uint32_t a_lane = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)a.a_col0;
    int st_s = 0;        // stage of the next slab to store
    uint32_t st_ph = 0;  // its phase bit
    auto store_slab = [&](const float (&src)[16]) {
      mbar_wait(empty_bar(st_s), st_ph ^ 1u);
      tc_fence_after();
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (NSPLIT == 3) {
          const uint32_t h = __float_as_uint(src[j]) & 0xFFFFE000u;
          hi[j] = h;
          lo[j] = __float_as_uint(src[j] - __uint_as_float(h));
        } else {
          hi[j] = __float_as_uint(src[j]);
        }
      }
      const uint32_t col = a_lane + (uint32_t)(st_s * A_COLS);
      tmem_st16(col, hi);
      if (NSPLIT == 3) tmem_st16(col + 16, lo);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar(st_s));  // 4 arrivals per slab instead of 128 serialised ones
      const bool wrap = (st_s + 1 == S);
      st_s = wrap ? 0 : st_s + 1;
      st_ph ^= wrap ? 1u : 0u;
    };
    if (T > 0) load_next(ra);
    if (T > 1) load_next(rb);
    for (int t = 0; t < T; t += 3) {
      if (t + 2 < T) load_next(rc);
      store_slab(ra);
      if (t + 3 < T) load_next(ra);
      if (t + 1 < T) store_slab(rb);
      if (t + 4 < T) load_next(rb);
      if (t + 2 < T) store_slab(rc);
    }
    // =========================== epilogue (thread = TMEM lane = tile row) ====================
    if (r < rows) orow = PAIRS ? (int64_t)s_idx[TC_BM + r] : (orow_perm >= 0 ? (int64_t)orow_perm : row0 + r);
    if (T > 0) {
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
    }
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    float* const out_base = a.out;
    for (int col = 0; col < a.c_out; col += 16) {
      uint32_t v[16];
      if (T > 0) {
        tmem_ld16(taddr + (uint32_t)col, v);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0u;
      }
      if (orow >= 0) {
        float4* o = (float4*)(out_base + orow * a.c_out + col);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          o[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                             __uint_as_float(v[4 * q + 3]));
      }
    }
    tc_fence_before();
  } else if (warp == 4) {
    // =========================== MMA issuer ===================================================
    // the whole warp runs the loop (uniform control flow); one elected lane issues
    if (T > 0) {
      const uint32_t idesc = make_idesc_tf32(a.c_out);
      const bool recycle_b = T > SB;
      int s = 0, j = 0;
      uint32_t ph = 0, bph = 0;
      for (int t = 0; t < T; ++t) {
        mbar_wait(bfull_bar(j), bph);
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t b_hi = base + (uint32_t)j * b_slot;
          const uint64_t db_hi = make_desc_sw64(b_hi);
          const uint32_t ta_hi = tmem_base + (uint32_t)a.a_col0 + (uint32_t)(s * A_COLS);
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {  // two K = 8 slices: +8 TMEM columns (A), +32 bytes = +2 encoded (B)
            const uint32_t acc = (t > 0 || ks > 0) ? 1u : 0u;
            if (NSPLIT == 3) {
              const uint64_t db_lo = make_desc_sw64(b_hi + b_slab);
              umma_tf32_ts(tmem_base, ta_hi + 16 + 8 * ks, db_hi + 2 * ks, idesc, acc);  // lo * hi
              umma_tf32_ts(tmem_base, ta_hi + 8 * ks, db_lo + 2 * ks, idesc, 1u);         // hi * lo
              umma_tf32_ts(tmem_base, ta_hi + 8 * ks, db_hi + 2 * ks, idesc, 1u);         // hi * hi
            } else {
              umma_tf32_ts(tmem_base, ta_hi + 8 * ks, db_hi + 2 * ks, idesc, acc);
            }
          }
          umma_commit(empty_bar(s));                  // frees the A stage once the MMAs above have read it
          if (recycle_b) umma_commit(bempty_bar(j));  // ... and the weight slot (only when the layer does not fit)
          if (t == T - 1) umma_commit(tmem_full_bar);
        }
        __syncwarp();
        if (++s == S) {
          s = 0;
          ph ^= 1u;
        }
        if (++j == SB) {
          j = 0;
          bph ^= 1u;
        }
      }
    }
  } else {
    // =========================== weight loader (TMA engine) ===================================
    if (T > 0) {
      uint32_t km = kmask;
      int k = -1, c = nc, j = 0;
      uint32_t bph = 0;
      const bool recycle_b = T > SB;
      for (int t = 0; t < T; ++t) {
        if (c == nc) {
          k = __ffs(km) - 1;
          km &= km - 1;
          c = 0;
        }
        const int kw = a.k_reversed ? (K - 1 - k) : k;
        if (recycle_b) mbar_wait(bempty_bar(j), bph ^ 1u);
        if (elect_one()) {
          const uint32_t b_hi = base + (uint32_t)j * b_slot;
          const float* src = a.Bp + ((int64_t)kw * nc + c) * (int64_t)(a.c_out * 16);
          mbar_arrive_expect_tx(bfull_bar(j), (uint32_t)b_slot);
          bulk_g2s(b_hi, src, (uint32_t)b_slab, bfull_bar(j));
          if (NSPLIT == 3) bulk_g2s(b_hi + b_slab, src + a.bp_half, (uint32_t)b_slab, bfull_bar(j));
        }
        __syncwarp();
        ++c;
        if (++j == SB) {
          j = 0;
          bph ^= 1u;
        }
      }
    }
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
  }
}

bool conv_tc_supported(int K, int c_in, int c_out) {
  return K >= 1 && K <= 32 && c_in >= 16 && (c_in % 16) == 0 && c_out >= 16 && (c_out % 16) == 0 && c_out <= 256;
}

size_t conv_tcp_split_ws_bytes();

size_t conv_tc_ws_bytes(int K, int c_in, int c_out) {
  return align_up((size_t)K * c_in * c_out * 4 * 2) + 256 + conv_tcp_split_ws_bytes();
}

template <bool PAIRS, int NSPLIT>
static int launch_tc(TcArgs a, int64_t grid_x, cudaStream_t stream) {
  auto bucket = [](int c) { int b = 32; while (b < c) b <<= 1; return b; };
  const int b_slot = (NSPLIT == 3 ? 2 : 1) * a.c_out * 64;
  const int a_cols = NSPLIT == 3 ? 32 : 16;
  // A stages: as many as the tensor-memory bucket of (accumulator + 2 stages) holds, at most 6
  int stages = 6;
  while (stages > 2 && bucket(a.c_out + stages * a_cols) > bucket(a.c_out + 2 * a_cols)) --stages;
  a.stages = stages;
  a.a_col0 = a.c_out;
  a.tmem_cols = bucket(a.c_out + stages * a_cols);
  // weight ring: every slab of the layer if that fits in 56 KB, else as many slots as 56 KB holds (>= 2)
  const int total_slabs = a.K * (a.c_in / 16);
  a.sb = std::max(2, std::min(std::min(total_slabs, 64), (56 * 1024) / b_slot));
  if (const char* e = getenv("B2S_TC_SB")) a.sb = std::max(2, std::min(a.sb, atoi(e)));
  if (const char* e = getenv("B2S_TC_STAGES")) {
    stages = std::max(2, std::min(stages, atoi(e)));
    a.stages = stages;
    a.tmem_cols = bucket(a.c_out + stages * a_cols);
  }
  const size_t fixed = 1024 /*align slack*/ + 8 * (2 * 8 + 2 * 64 + 1) + 16 + (size_t)(PAIRS ? 2 * TC_BM : TC_BM * a.K) * 4 + 64;
  size_t smem = (size_t)a.sb * b_slot + fixed;
  auto kern = conv_tc_kernel<PAIRS, NSPLIT>;
  static size_t configured[B2S_MAX_DEVICES] = {0};  // the attribute is per device (and per kernel instantiation)
  const int dev = current_device();
  if (smem > configured[dev]) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured[dev] = smem;
  }
  kern<<<(unsigned)grid_x, TC_THREADS, smem, stream>>>(a);
  return check_launch("conv_tc");
}

// W[K][c_in][c_out] -> [fwd hi | fwd lo | transposed hi | transposed lo], 4 * K * c_in * c_out floats, one launch.
// The transposed pair is the operand of the data gradient (a c_out -> c_in product with w_transposed = 1).
int conv_tc_pack_both(const float* W, float* Bp, int K, int c_in, int c_out, cudaStream_t stream) {
  const int64_t half = (int64_t)K * c_in * c_out;
  if (half == 0) return B2S_OK;
  pack_weights_kernel<<<dim3((unsigned)cdiv(half, 256), 2), 256, 0, stream>>>(W, Bp, K, c_in, c_out, 0, half);
  return check_launch("conv_pack");
}

// Every convolution kernel of a model in ONE launch (the trainer calls this right after optimizer.step()).
// desc[i] = {W pointer, Wp pointer, K, c_in, c_out, first element index} as six int64; element = one (k, c_in, c_out)
// entry of a layer; each thread writes the four images of its element (forward hi/lo, transposed hi/lo).
__global__ void __launch_bounds__(256)
    pack_weights_multi_kernel(const int64_t* __restrict__ desc, int n_desc, int64_t total) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  int lo = 0, hi = n_desc - 1;  // last descriptor whose start <= e
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(desc + mid * 6 + 5) <= e) lo = mid;
    else hi = mid - 1;
  }
  const int64_t* d = desc + lo * 6;
  const float* W = (const float*)__ldg(d + 0);
  float* Bp = (float*)__ldg(d + 1);
  const int c_in = (int)__ldg(d + 3), c_out = (int)__ldg(d + 4);
  const int64_t half = __ldg(d + 2) * c_in * c_out;
  const int64_t i = e - __ldg(d + 5);  // flat index into W[k][ci][co]
  const int co = (int)(i % c_out);
  const int ci = (int)((i / c_out) % c_in);
  const int64_t k = i / ((int64_t)c_out * c_in);
  const float v = W[i];
  const float vh = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
  const float vl = v - vh;
  // forward image: product c_in -> c_out, slab (k, ci / 16), row n = co, column j = ci % 16
  {
    const int64_t slab = k * (c_in >> 4) + (ci >> 4);
    const int64_t off = slab * (c_out * 16) + (sw64_offset(co, ci & 15) >> 2);
    Bp[off] = vh;
    Bp[half + off] = vl;
  }
  // transposed image: product c_out -> c_in, slab (k, co / 16), row n = ci, column j = co % 16
  {
    const int64_t slab = k * (c_out >> 4) + (co >> 4);
    const int64_t off = slab * (c_in * 16) + (sw64_offset(ci, co & 15) >> 2);
    Bp[2 * half + off] = vh;
    Bp[3 * half + off] = vl;
  }
}

int conv_tc_pack_multi(const int64_t* desc, int n_desc, int64_t total, cudaStream_t stream) {
  if (n_desc <= 0 || total <= 0) return B2S_OK;
  pack_weights_multi_kernel<<<(unsigned)cdiv(total, 256), 256, 0, stream>>>(desc, n_desc, total);
  return check_launch("conv_pack_multi");
}

// the image pair one product needs: packed by the caller (Wp, both orientations) or packed here into ws
const float* conv_tc_weights(const float* W, const float* Wp, int K, int c_in, int c_out, int wT, void* ws,
                             cudaStream_t stream) {
  const int64_t half = (int64_t)K * c_in * c_out;
  if (Wp != nullptr) return Wp + (wT ? 2 * half : 0);
  float* Bp = (float*)ws;
  pack_weights_kernel<<<(unsigned)cdiv(half, 256), 256, 0, stream>>>(W, Bp, K, c_in, c_out, wT, half);
  return Bp;
}

// ws: conv_tc_ws_bytes(K, c_in, c_out) bytes (packed weights when Wp == NULL, split scratch)
int conv_tc(const float* A, const float* W, const float* Wp, const int32_t* idx, const int32_t* dst, const int32_t* k_offsets,
            const uint32_t* tile_mask, const int32_t* out_rows, float* out, int64_t n_out, int64_t max_pairs, int K, int c_in, int c_out, int wT, int krev, int nsplit,
            bool pairs, void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (ws_bytes < conv_tc_ws_bytes(K, c_in, c_out)) {
    set_error("conv_tc: workspace too small");
    return B2S_E_WORKSPACE;
  }
  if ((!pairs && n_out == 0) || (pairs && max_pairs == 0)) return B2S_OK;
  int64_t half = (int64_t)K * c_in * c_out;
  const float* Bp = conv_tc_weights(W, Wp, K, c_in, c_out, wT, ws, stream);
  TcArgs a;
  a.A = A;
  a.Bp = Bp;
  a.idx = idx;
  a.dst = dst;
  a.k_offsets = k_offsets;
  a.tile_mask = (!pairs && K <= 32) ? tile_mask : nullptr;
  a.out_rows = pairs ? nullptr : out_rows;
  a.out = out;
  a.n_out = n_out;
  a.bp_half = half;
  a.K = K;
  a.c_in = c_in;
  a.c_out = c_out;
  a.k_reversed = krev;
  a.stages = a.tmem_cols = a.a_col0 = a.sb = 0;
  if (!pairs) {
    int64_t gx = cdiv(n_out, TC_BM);
    return nsplit == 3 ? launch_tc<false, 3>(a, gx, stream) : launch_tc<false, 1>(a, gx, stream);
  }
  int64_t gx = cdiv(max_pairs, TC_BM) + K;
  return nsplit == 3 ? launch_tc<true, 3>(a, gx, stream) : launch_tc<true, 1>(a, gx, stream);
}

}  // namespace b2s
