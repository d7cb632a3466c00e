// T3 on the 5th-generation tensor cores, PERSISTENT table-mode kernel (round 2).
//
//   out[o, :] = sum_k in[nbr[o,k], :] @ W[k]   (+ shortcut[o, :])          (appendix A.7; common.py:12,37,40,48)
//
// Same GEMM view and operand formats as conv_tc.cu (one tile = 128 output rows, accumulator D[128 x Cout] fp32 in
// TENSOR MEMORY, gathered A slabs written by producer warps with tcgen05.st and consumed by the TS form of
// tcgen05.mma.kind::tf32, weight slabs moved by the TMA engine with cp.async.bulk into a SWIZZLE_64B ring), but the
// per-tile fixed cost of the round-1 kernel -- launch of the CTA, ~60 mbarrier inits, the TMEM allocation, staging the
// tile's 13.8 KB of gather indices, the first weight slabs and the epilogue: 8.2 us of an 11.7 us 16-channel tile
// (profiles/r01_conv_tc_sorted_tiles.txt) -- is paid once per CTA or hidden:
//
//   * CTAs are persistent: each walks work items i = blockIdx.x, blockIdx.x + gridDim.x, ... ; barriers and the TMEM
//     allocation live for the whole kernel, pipeline state (stage / slot / phase) carries over from item to item.
//   * Weights stay RESIDENT in shared memory when the whole layer fits (16 -> 16: 27 slabs x 2 KB): one bulk copy per
//     slab per CTA instead of one per slab per tile; otherwise the ring streams them as before.
//   * The next items' gather indices are prefetched by the TMA engine (one 13.8 KB bulk copy) into a double buffer,
//     requested by producer warp 0 the moment a buffer is free.
//   * Gather staging.  Default: three register slots per producer thread (two slabs of LDG.256 in flight), the next
//     item's first loads issued before the accumulator is drained.  The round-2 ncu capture of this kernel on the
//     level-0 map (profiles/r02_conv_tcp_ncu.txt) shows a latency-bound gather -- 41 % of the warp samples wait on the
//     first use of a gathered register, 0.43 eligible warps per scheduler, issue slots 32 %, L2 throughput 33 % of its
//     peak -- and two remedies were built and MEASURED SLOWER, so they are not the default:
//       - ASYNC (B2S_TC_ASYNC=1, kept as an experiment switch): every producer thread copies its neighbour rows with
//         cp.async.cg into a thread-private, XOR-swizzled row of a 5-deep shared-memory ring (five slabs in flight, no
//         registers, cursor runs ahead across item boundaries) and reads it back with conflict-free LDS.128.  cp.async
//         moves at most 16 bytes per lane: 2.8x the L1TEX wavefronts of LDG.256 (6.6 M vs 2.4 M per launch), L1TEX
//         65 % busy -- 16->16 level 0: 64 us vs 50 us, 32->32 level 1: 44 us vs 37 us.
//       - lane pairs that fetch the two halves of a row in one instruction and swap them with shuffles: wavefronts
//         halve (1.28 M) but the 8 SHFL + 24 SEL per slab cost more than they save: 56-60 us vs 50 us.
//   * Two accumulators in TMEM: the MMA warp starts the next item while the producer warps drain the previous
//     accumulator (tcgen05.ld -> global); the producers issue the next item's first gathers BEFORE that epilogue.
//   * The residual add of the MinkUNet block (common.py:48 `x += shortcut`) is folded into the epilogue.
//   * A work item is (tile, part): maps with few tiles (U-Net levels 3-6: 1-80 tiles that each walk 100-190 slabs)
//     split the K offsets into `splits` fixed ranges that run on different SMs and write partial sums, added in a
//     fixed order by tcp_split_reduce_kernel (deterministic, independent of the tile masks).  Round 1 ran those
//     levels on a handful of SMs.
//
// Roles (192 threads): warps 0-3 gather producers + epilogue (thread = tile row = TMEM lane), warp 4 MMA issuer,
// warp 5 TMA loader (weights + index tiles).
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace b2s {

__device__ float g_zero_row_p[256];  // register mode: what a missing neighbour reads (keeps the gather loop branch-free)

constexpr int TCP_MAX_SB = 64;
constexpr int TCP_MAX_S = 6;
constexpr int TCP_DEPTH = 5;            // ASYNC: slabs of cp.async copies in flight per thread
constexpr int TCP_GSTAGE = TC_BM * 64;  // bytes of one gather stage (128 rows x 16 channels)

struct TcpArgs {
  const float* A;
  const float* Bp;            // packed weights [hi image | lo image] (pack_weights_kernel)
  const int32_t* idx;         // nbr [n_out, K] (row order or mask-sorted), or NULL = identity (K == 1)
  const uint32_t* tile_mask;  // active-offset mask per 128-row tile, or NULL = all K offsets
  const int32_t* out_rows;    // optional: tile row t is stored at out[out_rows[t]] (b2s_tile_order)
  const float* add_src;       // optional residual, same rows as out (splits == 1 only; else added by the reduce)
  float* out;                 // splits == 1: final output; else partial sums [splits][n_out][c_out]
  int64_t n_out;
  int64_t bp_half;            // floats in one weight image
  int K, c_in, c_out, k_reversed;
  int stages, sb, acc_bufs, tmem_cols, a_col0;
  int resident;               // every slab of the layer has its own slot: loaded once per CTA
  int splits;
  int n_tiles, n_items;
  int idx_bulk;               // nbr base is 16-byte aligned: index tiles are moved with cp.async.bulk
  // offset split with the reduction folded in: the warp that delivers the LAST part of its 32 rows of a tile (arrival
  // counter per (tile, warp), self-cleaning) adds the parts in ascending order (+ residual) and stores the final rows
  float* final_out;           // NULL: partial sums only, tcp_split_reduce_kernel finishes
  const float* final_add;
  int* split_done;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int NSPLIT, bool ASYNC>
__device__ __forceinline__ void conv_tcp_body(const TcpArgs& a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = a.stages, SB = a.sb, K = a.K;
  constexpr int NB = NSPLIT == 3 ? 2 : 1;
  constexpr int A_COLS = NSPLIT == 3 ? 32 : 16;
  const int b_slab = a.c_out * 64;
  const int b_slot = NB * b_slab;
  const int nc = a.c_in >> 4;
  // shared memory: [SB weight slots][barriers][tmem ptr][index tile 0][index tile 1][ASYNC: gather ring]
  const uint32_t bar0 = base + (uint32_t)SB * b_slot;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (S + s); };
  auto bfull_bar = [&](int j) { return bar0 + 8u * (2 * S + j); };
  auto bempty_bar = [&](int j) { return bar0 + 8u * (2 * S + SB + j); };
  const int misc0 = 2 * S + 2 * SB;
  auto tfull_bar = [&](int b) { return bar0 + 8u * (misc0 + b); };       // accumulator b complete (MMA commit)
  auto aempty_bar = [&](int b) { return bar0 + 8u * (misc0 + 2 + b); };  // accumulator b drained (4 warp arrivals)
  auto ifull_bar = [&](int b) { return bar0 + 8u * (misc0 + 4 + b); };   // index tile b landed
  auto iempty_bar = [&](int b) { return bar0 + 8u * (misc0 + 6 + b); };  // index tile b no longer read (4 arrivals)
  const int nbar = misc0 + 8;
  uint32_t* s_tmem = (uint32_t*)(sm + (size_t)SB * b_slot + 8u * nbar);
  const size_t idx_off = ((size_t)SB * b_slot + 8u * nbar + 16u + 15u) & ~(size_t)15;
  const int idx_tile_ints = TC_BM * K;
  int32_t* s_idx0 = (int32_t*)(sm + idx_off);
  const uint32_t s_idx0_u32 = base + (uint32_t)idx_off;
  const bool has_idx = a.idx != nullptr;
  const size_t gat_off = (idx_off + (has_idx ? (size_t)2 * idx_tile_ints * 4 : 0) + 127) & ~(size_t)127;

  // ---- one-time prologue ---------------------------------------------------------------------
  if (tid < nbar) {
    uint32_t cnt = 1u;
    if (tid < S) cnt = 4u;                                           // full[s]: one arrival per producer warp
    else if (tid >= misc0 + 2 && tid < misc0 + 4) cnt = 4u;          // aempty
    else if (tid >= misc0 + 6) cnt = 4u;                             // iempty
    mbar_init(bar0 + 8u * tid, cnt);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) tmem_alloc(smem_u32(s_tmem), (uint32_t)a.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const uint32_t all_mask = has_idx ? (K >= 32 ? 0xffffffffu : ((1u << K) - 1u)) : 1u;
  const int G = (int)gridDim.x;

  // mask of the offsets work item `item` covers: the tile's active offsets, restricted to this part's range
  auto item_mask = [&](int item) -> uint32_t {
    const int tile = a.splits > 1 ? item / a.splits : item;
    uint32_t kmask = (has_idx && a.tile_mask != nullptr) ? __ldg(a.tile_mask + tile) & all_mask : all_mask;
    if (a.splits > 1) {
      // part p owns the FIXED offset range [p * per, (p + 1) * per): the grouping of the partial sums does not depend
      // on the tile's mask, so a row's result is the same whatever tile (row order, mask-sorted order) it sits in
      const int part = item - tile * a.splits;
      const int per = (K + a.splits - 1) / a.splits;
      const int k0 = part * per, k1 = min(K, k0 + per);
      const uint32_t range = (k1 >= 32 ? 0xffffffffu : ((1u << k1) - 1u)) & ~((1u << k0) - 1u);
      kmask &= (k0 < k1) ? range : 0u;
    }
    return kmask;
  };

  if (warp < 4) {
    // =========================== gather producers + epilogue =================================
    const int r = tid;
    const float* __restrict__ Ag = a.A;
    const int c_in = a.c_in;
    const uint32_t a_lane = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)a.a_col0;
    const uint32_t d_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
    int st_s = 0;
    uint32_t st_ph = 0;

    // ---- index tiles: requested by warp 0 as soon as a buffer is free (local item number j -> buffer j & 1) -------
    auto load_idx = [&](int item, int j) {  // warp 0 only, converged
      const int b = j & 1;
      const int tile = a.splits > 1 ? item / a.splits : item;
      const int64_t row0 = (int64_t)tile * TC_BM;
      const int rows = (int)min((int64_t)TC_BM, a.n_out - row0);
      const int32_t* src = a.idx + row0 * K;
      int32_t* dst = s_idx0 + b * idx_tile_ints;
      const uint32_t bytes = (uint32_t)rows * (uint32_t)K * 4u;
      if (a.idx_bulk && rows == TC_BM) {
        if (elect_one()) {
          mbar_arrive_expect_tx(ifull_bar(b), bytes);
          bulk_g2s(s_idx0_u32 + (uint32_t)(b * idx_tile_ints) * 4u, src, bytes, ifull_bar(b));
        }
        __syncwarp();
      } else {  // last (partial) tile or unaligned table: plain copies, rows past the end hold -1
        const int total = rows * K;
        for (int e = lane; e < total; e += 32) dst[e] = __ldg(src + e);
        for (int e = total + lane; e < idx_tile_ints; e += 32) dst[e] = -1;
        __syncwarp();
        if (lane == 0) mbar_arrive(ifull_bar(b));
      }
    };
    // the gather cursor has read every index of local item j: release the buffer; warp 0 refills it for item j + 2
    auto release_idx = [&](int item, int j) {
      if (!has_idx) return;
      __syncwarp();
      if (lane == 0) mbar_arrive(iempty_bar(j & 1));
      if (warp == 0) {
        const int nxt = item + 2 * G;
        if (nxt < a.n_items) {
          mbar_wait(iempty_bar(j & 1), (uint32_t)(j >> 1) & 1u);  // the other three producer warps are at most one slab behind
          load_idx(nxt, j + 2);
        }
      }
    };
    if (has_idx && warp == 0) {
      if ((int)blockIdx.x < a.n_items) load_idx(blockIdx.x, 0);
      if ((int)blockIdx.x + G < a.n_items) load_idx(blockIdx.x + G, 1);
    }

    // ---- gather cursor: (item, offset, channel chunk) of the next slab to fetch ----------------------------
    int g_item = blockIdx.x, g_it = 0, g_rem = 0;
    bool g_valid = false;
    uint32_t g_km = 0, g_next_mask = (g_item < a.n_items) ? item_mask(g_item) : 0u;
    int g_lk = 0, g_lc = 0, g_ident = -1;
    const int32_t* g_idx = nullptr;
    auto gather_begin = [&](int item) {
      g_item = item;
      const int tile = a.splits > 1 ? item / a.splits : item;
      const int64_t row0 = (int64_t)tile * TC_BM;
      const int rows = (int)min((int64_t)TC_BM, a.n_out - row0);
      g_km = g_next_mask;
      if (item + G < a.n_items) g_next_mask = item_mask(item + G);  // one item ahead: latency hidden behind this item
      g_rem = __popc(g_km) * nc;
      if (has_idx) {
        mbar_wait(ifull_bar(g_it & 1), (uint32_t)(g_it >> 1) & 1u);
        g_idx = s_idx0 + (g_it & 1) * idx_tile_ints + r * K;  // rows past the end of the table hold -1
      } else {
        g_ident = (r < rows) ? (int)(row0 + r) : -1;
      }
      g_lk = 0;
      g_lc = 0;
      ++g_it;
      g_valid = true;
    };
    // moves the cursor to an item that still has slabs (releasing finished items' index tiles); false = no work left
    auto gather_advance = [&]() -> bool {
      while (g_valid && g_rem == 0) {
        release_idx(g_item, g_it - 1);
        if (g_item + G < a.n_items) gather_begin(g_item + G);
        else g_valid = false;
      }
      return g_valid;
    };
    // gather index + source pointer of the cursor's slab, then step the cursor
    auto gather_next = [&](int& g, const float*& p) {
      const bool adv = (g_lc == 0);
      const int kn = __ffs(g_km) - 1;
      g_lk = adv ? kn : g_lk;
      g_km = adv ? (g_km & (g_km - 1)) : g_km;
      g = has_idx ? g_idx[g_lk] : g_ident;
      p = Ag + ((int64_t)max(g, 0) * c_in + g_lc * 16);
      g_lc = (g_lc + 1 == nc) ? 0 : g_lc + 1;
      --g_rem;
    };

    // ---- TMEM store of one gathered slab (thread = tile row = TMEM lane) ---------------------------------
    auto split_store = [&](const float (&src)[16], auto&& between) {
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
      between();  // ASYNC: refill of the ring slot just read (its values are in hi / lo now)
      mbar_wait(empty_bar(st_s), st_ph ^ 1u);
      tc_fence_after();
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

    // ---- epilogue: TMEM -> registers -> (+ residual) -> global ----------------------------------------------
    auto epilogue = [&](int e_T, int e_part, int64_t e_orow, int e_ab, uint32_t e_ph, int e_tile) {
      if (e_T > 0) {
        mbar_wait(tfull_bar(e_ab), e_ph);
        tc_fence_after();
      }
      float* const out_base = a.out + (int64_t)e_part * a.n_out * a.c_out;
      const uint32_t taddr = d_lane + (uint32_t)(e_ab * a.c_out);
      for (int col = 0; col < a.c_out; col += 16) {
        uint32_t v[16];
        if (e_T > 0) {
          tmem_ld16(taddr + (uint32_t)col, v);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = 0u;
        }
        if (e_orow >= 0) {
          float4* o = (float4*)(out_base + e_orow * a.c_out + col);
          if (a.add_src != nullptr) {
            const float4* sp = (const float4*)(a.add_src + e_orow * a.c_out + col);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 sv = __ldg(sp + q);
              o[q] = make_float4(__uint_as_float(v[4 * q]) + sv.x, __uint_as_float(v[4 * q + 1]) + sv.y,
                                 __uint_as_float(v[4 * q + 2]) + sv.z, __uint_as_float(v[4 * q + 3]) + sv.w);
            }
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              o[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                                 __uint_as_float(v[4 * q + 3]));
          }
        }
      }
      if (e_T > 0) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(aempty_bar(e_ab));
      }
      if (a.splits > 1 && a.final_out != nullptr) {
        // same summation order as tcp_split_reduce_kernel (part 0, 1, ..., then the residual): identical bits, whichever
        // part arrives last
        __threadfence();
        __syncwarp();
        int old = 0;
        if (lane == 0) old = atomicAdd(a.split_done + e_tile * 4 + warp, 1);
        old = __shfl_sync(0xffffffffu, old, 0);
        if (old == a.splits - 1) {
          __threadfence();
          if (e_orow >= 0) {
            const int64_t part_stride = a.n_out * a.c_out;
            const float* p0 = a.out + e_orow * a.c_out;
            float4* o = (float4*)(a.final_out + e_orow * a.c_out);
            for (int col = 0; col < a.c_out; col += 4) {
              float4 acc = __ldcg((const float4*)(p0 + col));
              for (int sp = 1; sp < a.splits; ++sp) {
                const float4 v = __ldcg((const float4*)(p0 + (int64_t)sp * part_stride + col));
                acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
              }
              if (a.final_add != nullptr) {
                const float4 v = __ldg((const float4*)(a.final_add + e_orow * a.c_out + col));
                acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
              }
              o[col >> 2] = acc;
            }
          }
          if (lane == 0) a.split_done[e_tile * 4 + warp] = 0;
        }
      }
    };

    // ---- consume cursor: items in order, T slabs each --------------------------------------------------------
    int n_acc = 0;  // items with at least one slab so far (accumulator uses)
    uint32_t c_next_mask = ((int)blockIdx.x < a.n_items) ? item_mask(blockIdx.x) : 0u;
    if ((int)blockIdx.x < a.n_items) gather_begin(blockIdx.x);

    if constexpr (ASYNC) {
      // ring slot of slab number q (counted over the CTA's whole slab stream) = q % TCP_DEPTH; row r of a slot is
      // thread-private: 64 bytes, its four 16-byte chunks XOR-swizzled with (r >> 1) & 3 (conflict-free LDS.128)
      const uint32_t my_row = base + (uint32_t)gat_off + (uint32_t)r * 64u;
      const uint32_t sw = (uint32_t)((r >> 1) & 3);
      auto issue = [&](int slot) {  // one commit group per call, empty when the stream has ended
        if (gather_advance()) {
          int g;
          const float* p;
          gather_next(g, p);
          const uint32_t nb = (g >= 0) ? 16u : 0u;  // missing neighbour: zero-fill, nothing is read
          const uint32_t dst = my_row + (uint32_t)slot * TCP_GSTAGE;
#pragma unroll
          for (uint32_t q = 0; q < 4; ++q) cp_async16(dst + ((q ^ sw) << 4), p + 4 * q, nb);
        }
        cp_async_commit();
      };
#pragma unroll
      for (int d = 0; d < TCP_DEPTH; ++d) issue(d);
      int slot = 0;
      for (int item = blockIdx.x; item < a.n_items; item += G) {
        const uint32_t kmask = c_next_mask;
        if (item + G < a.n_items) c_next_mask = item_mask(item + G);
        const int T = __popc(kmask) * nc;
        const int tile = a.splits > 1 ? item / a.splits : item;
        const int64_t row0 = (int64_t)tile * TC_BM;
        const int rows = (int)min((int64_t)TC_BM, a.n_out - row0);
        const int orow_perm = (a.out_rows != nullptr && r < rows) ? __ldg(a.out_rows + row0 + r) : -1;
        for (int t = 0; t < T; ++t) {
          cp_async_wait<TCP_DEPTH - 1>();  // the oldest group (this slab) has landed
          float src[16];
          const uint32_t row = my_row + (uint32_t)slot * TCP_GSTAGE;
#pragma unroll
          for (uint32_t q = 0; q < 4; ++q)
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                         : "=f"(src[4 * q]), "=f"(src[4 * q + 1]), "=f"(src[4 * q + 2]), "=f"(src[4 * q + 3])
                         : "r"(row + ((q ^ sw) << 4)));
          const int cur = slot;
          split_store(src, [&]() { issue(cur); });
          slot = (slot + 1 == TCP_DEPTH) ? 0 : slot + 1;
        }
        const int e_ab = (T > 0) ? (n_acc % a.acc_bufs) : 0;
        const uint32_t e_ph = (uint32_t)(n_acc / a.acc_bufs) & 1u;
        if (T > 0) ++n_acc;
        const int64_t e_orow = (r < rows) ? (orow_perm >= 0 ? (int64_t)orow_perm : row0 + r) : -1;
        epilogue(T, a.splits > 1 ? item - tile * a.splits : 0, e_orow, e_ab, e_ph, tile);
      }
      cp_async_wait<0>();
    } else {
      // register mode (fallback when the gather ring does not fit): three register slots rotate, two slabs of loads
      // in flight per thread; the next item's first loads are issued before the accumulator is drained
      float ra[16], rb[16], rc[16];
      auto load_next = [&](float (&dst)[16]) {
        int g;
        const float* p;
        gather_next(g, p);
        p = (g >= 0) ? p : g_zero_row_p;
        asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(dst[0]), "=f"(dst[1]), "=f"(dst[2]), "=f"(dst[3]), "=f"(dst[4]), "=f"(dst[5]), "=f"(dst[6]),
                       "=f"(dst[7])
                     : "l"(p));
        asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(dst[8]), "=f"(dst[9]), "=f"(dst[10]), "=f"(dst[11]), "=f"(dst[12]), "=f"(dst[13]),
                       "=f"(dst[14]), "=f"(dst[15])
                     : "l"(p + 8));
      };
      auto nothing = []() {};
      int T = 0;
      for (int item = blockIdx.x; item < a.n_items;) {
        if (item == (int)blockIdx.x) {  // first item: nothing is in flight yet
          T = g_rem;
          if (T > 0) load_next(ra);
          if (T > 1) load_next(rb);
        }
        const int tile = a.splits > 1 ? item / a.splits : item;
        const int64_t row0 = (int64_t)tile * TC_BM;
        const int rows = (int)min((int64_t)TC_BM, a.n_out - row0);
        const int orow_perm = (a.out_rows != nullptr && r < rows) ? __ldg(a.out_rows + row0 + r) : -1;
        for (int t = 0; t < T; t += 3) {
          if (t + 2 < T) load_next(rc);
          split_store(ra, nothing);
          if (t + 3 < T) load_next(ra);
          if (t + 1 < T) split_store(rb, nothing);
          if (t + 4 < T) load_next(rb);
          if (t + 2 < T) split_store(rc, nothing);
        }
        const int e_T = T;
        const int e_ab = (e_T > 0) ? (n_acc % a.acc_bufs) : 0;
        const uint32_t e_ph = (uint32_t)(n_acc / a.acc_bufs) & 1u;
        if (e_T > 0) ++n_acc;
        const int64_t e_orow = (r < rows) ? (orow_perm >= 0 ? (int64_t)orow_perm : row0 + r) : -1;
        const int e_part = a.splits > 1 ? item - tile * a.splits : 0;
        const int e_tile = tile;
        // every index of this item has been read: release its tile, then start the next item's gathers before
        // draining the accumulator
        release_idx(item, g_it - 1);
        item += G;
        if (item < a.n_items) {
          gather_begin(item);
          T = g_rem;
          if (T > 0) load_next(ra);
          if (T > 1) load_next(rb);
        }
        epilogue(e_T, e_part, e_orow, e_ab, e_ph, e_tile);
      }
    }
    tc_fence_before();
  } else if (warp == 4) {
    // =========================== MMA issuer ===================================================
    const uint32_t idesc = make_idesc_tf32(a.c_out);
    int s = 0, j = 0, n_acc = 0;
    uint32_t ph = 0, bph = 0;
    uint32_t next_mask = ((int)blockIdx.x < a.n_items) ? item_mask(blockIdx.x) : 0u;
    for (int item = blockIdx.x; item < a.n_items; item += G) {
      uint32_t km = next_mask;
      if (item + G < a.n_items) next_mask = item_mask(item + G);
      const int T = __popc(km) * nc;
      if (T == 0) continue;
      const int ab = n_acc % a.acc_bufs;
      const int use = n_acc / a.acc_bufs;
      ++n_acc;
      if (use > 0) {  // the accumulator still holds an earlier item until the producers have drained it
        mbar_wait(aempty_bar(ab), (uint32_t)(use - 1) & 1u);
        tc_fence_after();
      }
      const uint32_t d_addr = tmem_base + (uint32_t)(ab * a.c_out);
      int k = -1, c = nc;
      for (int t = 0; t < T; ++t) {
        if (c == nc) {
          k = __ffs(km) - 1;
          km &= km - 1;
          c = 0;
        }
        const int slot = a.resident ? (k * nc + c) : j;
        mbar_wait(bfull_bar(slot), a.resident ? 0u : bph);
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t b_hi = base + (uint32_t)slot * b_slot;
          const uint64_t db_hi = make_desc_sw64(b_hi);
          const uint32_t ta_hi = tmem_base + (uint32_t)a.a_col0 + (uint32_t)(s * A_COLS);
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint32_t acc = (t > 0 || ks > 0) ? 1u : 0u;
            if (NSPLIT == 3) {
              const uint64_t db_lo = make_desc_sw64(b_hi + b_slab);
              umma_tf32_ts(d_addr, ta_hi + 16 + 8 * ks, db_hi + 2 * ks, idesc, acc);  // lo * hi
              umma_tf32_ts(d_addr, ta_hi + 8 * ks, db_lo + 2 * ks, idesc, 1u);         // hi * lo
              umma_tf32_ts(d_addr, ta_hi + 8 * ks, db_hi + 2 * ks, idesc, 1u);         // hi * hi
            } else {
              umma_tf32_ts(d_addr, ta_hi + 8 * ks, db_hi + 2 * ks, idesc, acc);
            }
          }
          umma_commit(empty_bar(s));
          if (!a.resident) umma_commit(bempty_bar(j));
          if (t == T - 1) umma_commit(tfull_bar(ab));
        }
        __syncwarp();
        ++c;
        if (++s == S) {
          s = 0;
          ph ^= 1u;
        }
        if (!a.resident && ++j == SB) {
          j = 0;
          bph ^= 1u;
        }
      }
    }
  } else {
    // =========================== TMA loader: weight slabs =====================================
    auto load_weight = [&](int slot, int kw, int c) {
      const uint32_t b_hi = base + (uint32_t)slot * b_slot;
      const float* src = a.Bp + ((int64_t)kw * nc + c) * (int64_t)(a.c_out * 16);
      mbar_arrive_expect_tx(bfull_bar(slot), (uint32_t)b_slot);
      bulk_g2s(b_hi, src, (uint32_t)b_slab, bfull_bar(slot));
      if (NSPLIT == 3) bulk_g2s(b_hi + b_slab, src + a.bp_half, (uint32_t)b_slab, bfull_bar(slot));
    };
    const int first = blockIdx.x;
    if (a.resident) {
      if (first < a.n_items) {
        // every slab of the layer once per CTA, slot = (tile offset k, chunk c); the source honours k_reversed
        for (int k = 0; k < K; ++k) {
          const int kw = a.k_reversed ? (K - 1 - k) : k;
          for (int c = 0; c < nc; ++c) {
            if (elect_one()) load_weight(k * nc + c, kw, c);
            __syncwarp();
          }
        }
        // a CTA must not exit with bulk copies in flight: slabs its items never touch are awaited here
        for (int slot = 0; slot < K * nc; ++slot) mbar_wait(bfull_bar(slot), 0u);
      }
    } else {
      int j = 0;
      uint32_t bph = 0;
      long long issued = 0;  // the first SB slabs need no bempty wait
      uint32_t next_mask = (first < a.n_items) ? item_mask(first) : 0u;
      for (int item = first; item < a.n_items; item += G) {
        uint32_t km = next_mask;
        if (item + G < a.n_items) next_mask = item_mask(item + G);
        const int T = __popc(km) * nc;
        int k = -1, c = nc;
        for (int t = 0; t < T; ++t) {
          if (c == nc) {
            k = __ffs(km) - 1;
            km &= km - 1;
            c = 0;
          }
          const int kw = a.k_reversed ? (K - 1 - k) : k;
          if (issued >= SB) mbar_wait(bempty_bar(j), bph ^ 1u);
          if (elect_one()) load_weight(j, kw, c);
          __syncwarp();
          ++issued;
          ++c;
          if (++j == SB) {
            j = 0;
            bph ^= 1u;
          }
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

// entry points: 2 CTAs/SM (up to 168 registers, the default) and 3 CTAs/SM (112 registers, 8 bytes of spills;
// B2S_TC_CTAS=3, narrow layers only).  The 3-CTA variant was MEASURED SLOWER on the level-0 map -- 16->16 cold 87.7 us vs
// 68.9 us, step 27.6 vs 26.0 ms (profiles/r02_conv_tcp_variants.txt) -- with three stages, a streamed instead of
// resident weight set and 50 % more gather streams contending for the same L1TEX; it stays as an experiment switch.
template <int NSPLIT, bool ASYNC>
__global__ void __launch_bounds__(TC_THREADS, 2) conv_tcp_kernel(const TcpArgs a) {
  conv_tcp_body<NSPLIT, ASYNC>(a);
}
template <int NSPLIT>
__global__ void __maxnreg__(112) conv_tcp3_kernel(const TcpArgs a) {
  conv_tcp_body<NSPLIT, false>(a);
}

// partial sums of the offset split -> output, fixed order (deterministic), optional residual
__global__ void __launch_bounds__(256)
    tcp_split_reduce_kernel(const float4* __restrict__ partial, const float4* __restrict__ add_src,
                            float4* __restrict__ out, int64_t total4, int splits) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  float4 acc = __ldg(partial + i);
  for (int s = 1; s < splits; ++s) {
    const float4 v = __ldg(partial + (int64_t)s * total4 + i);
    acc.x += v.x;
    acc.y += v.y;
    acc.z += v.z;
    acc.w += v.w;
  }
  if (add_src != nullptr) {
    const float4 v = __ldg(add_src + i);
    acc.x += v.x;
    acc.y += v.y;
    acc.z += v.z;
    acc.w += v.w;
  }
  out[i] = acc;
}

constexpr size_t TCP_SPLIT_WS = (size_t)16 << 20;  // scratch for the partial outputs of the offset split

size_t conv_tcp_split_ws_bytes() { return TCP_SPLIT_WS + 256; }

static int tcp_env(const char* name, int dflt) {
  const char* e = getenv(name);
  return e != nullptr ? atoi(e) : dflt;
}

template <int NSPLIT>
static int launch_tcp(TcpArgs a, void* split_ws, float* final_out, const float* add_src, cudaStream_t stream) {
  const int a_cols = NSPLIT == 3 ? 32 : 16;
  const int b_slot = (NSPLIT == 3 ? 2 : 1) * a.c_out * 64;
  // ---- tensor memory: [acc_bufs accumulators of c_out columns][S stages of a_cols columns] -------------------
  // 512 columns per SM: <= 128 per CTA lets three CTAs share an SM, <= 256 two, more only one
  auto fits = [&](int ab, int S, int cap) { return ab * a.c_out + S * a_cols <= cap; };
  int cap = 256, ab = 2;
  const bool allow3 = tcp_env("B2S_TC_CTAS", 2) >= 3;  // measured slower (see conv_tcp3_kernel): off by default
  if (allow3 && fits(2, 3, 128)) {
    cap = 128;
    ab = 2;
  } else if (allow3 && fits(1, 3, 128)) {
    cap = 128;
    ab = 1;
  } else if (fits(2, 4, 256)) {
    ab = 2;
  } else if (fits(1, 3, 256)) {
    ab = 1;
  } else {
    cap = 512;
    ab = fits(2, 4, 512) ? 2 : 1;
  }
  int S = std::min(TCP_MAX_S, (cap - ab * a.c_out) / a_cols);
  S = std::max(2, std::min(S, tcp_env("B2S_TC_STAGES", TCP_MAX_S)));
  int cols = ab * a.c_out + S * a_cols;
  int bucket = 32;
  while (bucket < cols) bucket <<= 1;
  a.acc_bufs = ab;
  a.stages = S;
  a.a_col0 = ab * a.c_out;
  a.tmem_cols = bucket;
  const int ctas_per_sm = bucket > 256 ? 1 : (bucket > 128 ? 2 : 3);
  // ---- shared memory plan: [weights][barriers][2 index tiles][ASYNC: gather ring of TCP_DEPTH x 8 KB] ------------
  // budget per CTA: 112 KB when two CTAs share an SM, 220 KB when the TMEM allocation allows only one
  const int total_slabs = a.K * (a.c_in / 16);
  const size_t budget = (size_t)(ctas_per_sm == 3 ? 74 : (ctas_per_sm == 2 ? 112 : 220)) * 1024;
  const size_t idx_bytes = a.idx != nullptr ? (size_t)2 * TC_BM * a.K * 4 : 0;
  const size_t fixed = 1024 /*align slack*/ + 8 * (size_t)(2 * TCP_MAX_S + 8) + 32 + idx_bytes + 64 + 128;
  auto plan = [&](size_t extra, bool* resident, int* sb) {  // weights in what is left after `extra`; false = does not fit
    if (fixed + extra + 2 * (size_t)b_slot + 64 > budget) return false;
    const size_t room = budget - fixed - extra;
    // every slab costs its slot + two barriers
    const int cap_slots = (int)std::min<size_t>(TCP_MAX_SB, room / ((size_t)b_slot + 16));
    *resident = total_slabs <= cap_slots;
    *sb = *resident ? total_slabs : std::min(total_slabs, cap_slots);
    return *resident || *sb >= 2;
  };
  bool resident = false, async = tcp_env("B2S_TC_ASYNC", 0) != 0;
  int sb = 2;
  if (async) {
    async = ctas_per_sm == 2 && plan((size_t)TCP_DEPTH * TCP_GSTAGE, &resident, &sb) && (resident || sb >= 4);
  }
  if (!async && !plan(0, &resident, &sb)) {
    set_error("conv_tcp: shared memory plan failed");
    return B2S_E_INVALID;
  }
  if (!resident) sb = std::max(2, std::min(sb, tcp_env("B2S_TC_SB", TCP_MAX_SB)));
  a.resident = resident ? 1 : 0;
  a.sb = sb;
  // ---- work items -------------------------------------------------------------------------------------
  a.n_tiles = (int)cdiv(a.n_out, TC_BM);
  const int slots = sm_count() * ctas_per_sm;
  int splits = 1;
  if (a.idx != nullptr && a.K >= 8 && total_slabs >= 24 && split_ws != nullptr && tcp_env("B2S_TC_SPLIT", 1) != 0) {
    const int64_t by_fill = slots / std::max(1, a.n_tiles);  // parts that still fit in one wave
    const int64_t by_mem = (int64_t)(TCP_SPLIT_WS / ((size_t)a.n_out * a.c_out * 4 + 1));
    const int64_t s = std::min<int64_t>(std::min<int64_t>(by_fill, by_mem), std::min(9, a.K / 3));
    if (s >= 2) {
      const int per = (a.K + (int)s - 1) / (int)s;  // offsets per part (fixed ranges, see item_mask)
      splits = (a.K + per - 1) / per;               // no empty trailing part
    }
  }
  a.splits = splits;
  a.n_items = a.n_tiles * splits;
  a.idx_bulk = (a.idx != nullptr && (((uintptr_t)a.idx) & 15) == 0) ? 1 : 0;
  a.final_out = nullptr;
  a.final_add = nullptr;
  a.split_done = nullptr;
  if (splits > 1) {
    a.out = (float*)split_ws;
    a.add_src = nullptr;
    // B2S_TC_SPLIT_FOLD=1 folds the reduction into the kernel (last-arriving warp sums the parts): 72 launches less per
    // PointGroup step but MEASURED SLOWER -- step 23.4 vs 22.1 ms: one warp adding 32 rows x splits parts serially at the
    // end of a latency-bound kernel costs more than the 3 us reduce kernel that spreads the same reads over the GPU.
    // Default: separate tcp_split_reduce_kernel launch.
    if (tcp_env("B2S_TC_SPLIT_FOLD", 0) != 0 && (size_t)a.n_tiles * 4 <= B2S_COUNTERS) {
      a.split_done = zeroed_counters(stream);
      if (a.split_done != nullptr) {
        a.final_out = final_out;
        a.final_add = add_src;
      }
    }
  } else {
    a.out = final_out;
    a.add_src = add_src;
  }
  const int nbar = 2 * S + 2 * a.sb + 8;
  size_t smem = 1024 /*align slack*/ + (size_t)a.sb * b_slot + 8 * (size_t)nbar + 32 + idx_bytes + 64 + 128 +
                (async ? (size_t)TCP_DEPTH * TCP_GSTAGE : 0);
  if (ctas_per_sm == 1) smem = std::max(smem, (size_t)120 * 1024);  // 512 TMEM columns: keep a second CTA off the SM
  auto kern = ctas_per_sm == 3 ? conv_tcp3_kernel<NSPLIT>
                               : (async ? conv_tcp_kernel<NSPLIT, true> : conv_tcp_kernel<NSPLIT, false>);
  static size_t configured[B2S_MAX_DEVICES][3] = {};
  const int dev = current_device();
  const int which = ctas_per_sm == 3 ? 2 : (async ? 1 : 0);
  if (smem > configured[dev][which]) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured[dev][which] = smem;
  }
  const int grid = std::min(a.n_items, slots);
  kern<<<grid, TC_THREADS, smem, stream>>>(a);
  int rc = check_launch("conv_tcp");
  if (rc || splits == 1 || a.final_out != nullptr) return rc;
  const int64_t total4 = a.n_out * (a.c_out / 4);
  tcp_split_reduce_kernel<<<(unsigned)cdiv(total4, 256), 256, 0, stream>>>((const float4*)split_ws, (const float4*)add_src,
                                                                           (float4*)final_out, total4, splits);
  return check_launch("conv_tcp split reduce");
}

// Persistent table-mode convolution.  Bp: packed weights (both images); split_ws: conv_tcp_split_ws_bytes() of scratch
// (or NULL: never split).  add_src (optional): residual added to the output rows.
int conv_tcp(const float* A, const float* Bp, const int32_t* idx, const uint32_t* tile_mask, const int32_t* out_rows,
             const float* add_src, float* out, int64_t n_out, int K, int c_in, int c_out, int krev, int nsplit,
             void* split_ws, cudaStream_t stream) {
  if (n_out == 0) return B2S_OK;
  TcpArgs a{};
  a.A = A;
  a.Bp = Bp;
  a.idx = idx;
  a.tile_mask = (K <= 32) ? tile_mask : nullptr;
  a.out_rows = out_rows;
  a.n_out = n_out;
  a.bp_half = (int64_t)K * c_in * c_out;
  a.K = K;
  a.c_in = c_in;
  a.c_out = c_out;
  a.k_reversed = krev;
  return nsplit == 3 ? launch_tcp<3>(a, split_ws, out, add_src, stream) : launch_tcp<1>(a, split_ws, out, add_src, stream);
}

}  // namespace b2s
