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

namespace b2s {

constexpr int TC_BM = 128;          // rows of one output tile = TMEM lanes
constexpr int TC_NPW = 16;          // gather producer warps: 4 groups x 4 lane quarters
constexpr int TC_G = TC_NPW / 4;    // producer groups; group g owns the slabs with (global slab index) % G == g
constexpr int TC_W_MMA = 16;        // MMA issuer warp (also owns the tensor-memory allocation)
constexpr int TC_W_TILE = 17;       // tile loader: neighbour-table tile, active-offset mask, slab list
constexpr int TC_W_WGT = 18;        // weight loader (TMA bulk copies)
constexpr int TC_W_EPI = 20;        // 4 epilogue warps (20..23 -> lane quarters 0..3)
constexpr int TC_THREADS = 24 * 32;
constexpr int TC_MAX_NS = 16;       // A stages (tensor memory)
constexpr int TC_MAX_SB = 32;       // weight ring slots (ring mode)
constexpr int TC_NBUF = 4;          // tile index buffers (bulk copies run two tiles ahead)
constexpr int TC_MAX_T = 512;       // slabs per tile: K <= 32 offsets x c_in / 16 <= 16 chunks

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
// one try_wait attempt; the result is consumed later so that the ~90-cycle barrier probe overlaps other work
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// one lane of a converged warp (ELECT): keeps the surrounding control flow warp-uniform, so descriptors and
// barrier addresses stay in uniform registers (a lane == 0 branch makes ptxas wrap every tcgen05/UBLKCP
// instruction in a R2UR waterfall loop: ~40 instructions and ~80 cycles per MMA, measured)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u)
      : "memory");
}
// TS form: A operand from tensor memory (lane = tile row, 8 consecutive 32-bit columns per K = 8 slice)
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_64B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// start>>4 [0,14) | LBO>>4 [16,30) (ignored for swizzled K-major) | SBO>>4 [32,46) = 512 B between
// 8-row groups | version=1 [46,48) | layout_type=4 (SWIZZLE_64B) [61,64)
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)4 << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6)=1, a=TF32 [7,10)=2, b=TF32 [10,13)=2,
// a/b K-major, N>>3 [17,23), M>>4 [24,29)
__host__ __device__ __forceinline__ uint32_t make_idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}
__host__ __device__ __forceinline__ int sw64_offset(int row, int j) {  // byte offset inside a slab
  return row * 64 + ((((j >> 2) ^ ((row >> 1) & 3))) << 4) + (j & 3) * 4;
}

// ---------------------------------------------------------------------------------------------
// weight packing: W[K][c_in][c_out] (or its transpose view) -> per slab (kernel offset k, 16-channel chunk c):
// [hi image][lo image (nb == 2 only)], each the SW64 byte image of a (c_out rows x 16 k) K-major tile
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    pack_weights_kernel(const float* __restrict__ W, float* __restrict__ Bp, int K, int c_in, int c_out,
                        int w_transposed, int nb, int64_t total) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int per_slab = c_out * 16;
  int64_t slab = e / per_slab;
  int within = (int)(e - slab * per_slab);
  int n = within >> 4, j = within & 15;
  int nc = c_in >> 4;
  int k = (int)(slab / nc), c = (int)(slab - (int64_t)k * nc);
  int ci = c * 16 + j;
  const float* Wk = W + (int64_t)k * c_in * c_out;
  float v = w_transposed ? Wk[(int64_t)n * c_in + ci] : Wk[(int64_t)ci * c_out + n];
  int64_t off = slab * per_slab * nb + (sw64_offset(n, j) >> 2);
  if (nb == 2) {
    float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    Bp[off] = hi;
    Bp[off + per_slab] = v - hi;
  } else {
    Bp[off] = v;
  }
}

// what a missing neighbour reads (keeps the gather loop branch-free); 256 floats = widest supported row
__device__ float g_zero_row[256];

struct TcArgs {
  const float* A;
  const float* Bp;       // packed weights, one slot of nb * c_out * 16 floats per slab (k, c)
  const int32_t* idx;    // TABLE: nbr [n_out, K] (or NULL = identity, K == 1); PAIRS: src
  const int32_t* dst;    // PAIRS: destination rows
  const int32_t* k_offsets;
  const uint32_t* tile_mask;  // TABLE, optional: active-offset mask per tile (b2s_kernel_map)
  float* out;
  int64_t n_out;
  int n_tiles;           // TABLE: exact; PAIRS: upper bound (the tile loader detects the end)
  int K, c_in, c_out, k_reversed;
  int ns;                // A stages in tensor memory
  int nacc;              // accumulators in tensor memory (2 = epilogue overlaps the next tile)
  int a_col0;            // first tensor-memory column of the A stages
  int sb;                // weight slots in shared memory (resident: all K * c_in/16 slabs)
  int resident;          // 1: every weight slab is loaded once per CTA and stays
  int idx_bulk;          // 1: neighbour-table tiles are 16-byte aligned -> one bulk copy per tile
  int groups;            // producer groups in use (<= TC_G); env B2S_TC_GROUPS, experiments
  int flags;             // env B2S_TC_DEBUG: 8 = skip the tcgen05.mma issue (ablation)
};

struct TcMeta {  // per tile, written by the tile loader
  int T;         // slabs (active offsets x channel chunks); -1 = no more tiles (PAIRS)
  int rows;      // valid rows of the tile
  int base;      // TABLE: first output row; PAIRS: first pair
  int pad;       // active-offset mask (bit k)
};

// Persistent, warp-specialised implicit GEMM: one CTA per SM walks tiles blockIdx.x, +gridDim.x, ...
//   producers (16 warps) : gather 64 B per row and slab into registers, split hi/lo, tcgen05.st into an A stage
//   MMA warp             : tcgen05.mma (TS form) per slab into one of two accumulators, commits free the stages
//   tile loader          : next tile's neighbour-table slice (one bulk copy), active-offset mask, slab list
//   weight loader        : weight slabs by bulk copy; resident for small layers, a ring otherwise
//   epilogue (4 warps)   : tcgen05.ld of the finished accumulator -> global, overlapped with the next tile
// Why this shape (profiles/r01_conv_tc_*.txt): one warp issues in order, so the per-slab chain (index -> pointer
// -> load -> split -> tcgen05.st -> wait -> arrive, ~400-800 cycles) only overlaps across warps; the previous
// one-CTA-per-tile kernel had 3 producer warps per scheduler and paid prologue + drain per tile.
template <bool PAIRS, int NSPLIT, bool RESIDENT>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int NB = NSPLIT == 3 ? 2 : 1;  // weight images per slab (hi, lo)
  constexpr int A_COLS = 16 * NB;          // tensor-memory columns of one A stage (hi, lo)
  const int K = a.K, nc = a.c_in >> 4, NS = a.ns, SB = a.sb;
  const uint32_t b_slab = (uint32_t)a.c_out * 64u;  // bytes of one weight image
  const uint32_t b_slot = NB * b_slab;              // bytes of one weight slot
  const int idx_words = PAIRS ? 2 * TC_BM : TC_BM * K;
  // shared memory: [weights][index tile x4][slab list x4][meta x4][barriers][tmem ptr]
  const uint32_t off_idx = (uint32_t)SB * b_slot;
  const uint32_t off_list = off_idx + (((uint32_t)TC_NBUF * idx_words * 4u + 15u) & ~15u);
  const uint32_t off_meta = off_list + (uint32_t)TC_NBUF * TC_MAX_T * 2u;
  const uint32_t off_bar = off_meta + (uint32_t)TC_NBUF * (uint32_t)sizeof(TcMeta);
  int32_t* s_idx = (int32_t*)(sm + off_idx);
  uint16_t* s_list = (uint16_t*)(sm + off_list);
  TcMeta* s_meta = (TcMeta*)(sm + off_meta);
  const uint32_t bar0 = base + off_bar;
  // barrier indices
  constexpr int B_FULL = 0, B_EMPTY = B_FULL + TC_MAX_NS, B_WFULL = B_EMPTY + TC_MAX_NS, B_WEMPTY = B_WFULL + TC_MAX_SB,
                B_TCOPY = B_WEMPTY + TC_MAX_SB, B_TREADY = B_TCOPY + TC_NBUF, B_TEMPTY = B_TREADY + TC_NBUF,
                B_AFULL = B_TEMPTY + TC_NBUF,
                B_AEMPTY = B_AFULL + 2, B_COUNT = B_AEMPTY + 2;
  auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  uint32_t* s_tmem = (uint32_t*)(sm + off_bar + 8u * B_COUNT);
  const int G = a.groups;
  const int n_consumers = 4 * G + 1 + 4 + (RESIDENT ? 0 : 1);  // warps that read a tile's index/list/meta

  if (tid < B_COUNT) {
    uint32_t cnt = 1;
    if (tid < B_EMPTY) cnt = 4u * (uint32_t)G;  // full[s]: every producer warp fills its part of the fat stage
    else if (tid >= B_TEMPTY && tid < B_AFULL) cnt = (uint32_t)n_consumers;
    else if (tid >= B_AEMPTY) cnt = 4;  // epilogue warps
    mbar_init(bar(tid), cnt);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == TC_W_MMA) tmem_alloc(smem_u32(s_tmem), 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const int n_tiles = a.n_tiles;
  const int nacc_shift = a.nacc - 1;  // nacc is 1 or 2

  if (warp < 4 * G) {
    // =========================== gather producers ============================================
    // thread = tile row = TMEM lane.  A *fat stage* holds G consecutive slabs (one per producer group), so the
    // MMA warp does one barrier handshake and one burst of MMAs per G slabs; a tile's slab count is padded to
    // a multiple of G with null slabs (arrive only).  One slab is loaded ahead (also across tiles) while the
    // previous one is split and written to tensor memory.
    const int g = warp >> 2, q = warp & 3;
    const int r = q * 32 + lane;
    const float* __restrict__ Ag = a.A;
    const int c_in = a.c_in;
    // cursor over the owned slabs
    int it = -1, T = 0, Tp = 0, n_rel = g, tloc = 0;
    bool valid = false;
    const int32_t* my_idx = nullptr;
    const uint16_t* list = nullptr;
    auto next_tile = [&]() -> bool {
      if (it >= 0) {  // done reading the previous tile's indices
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_TEMPTY + (it & (TC_NBUF - 1))));
      }
      ++it;
      if (blockIdx.x + (int64_t)it * gridDim.x >= n_tiles) return false;
      const int buf = it & (TC_NBUF - 1);
      mbar_wait(bar(B_TREADY + buf), (uint32_t)(it >> 2) & 1u);
      T = s_meta[buf].T;
      if (T < 0) return false;
      Tp = (T + G - 1) / G * G;
      my_idx = s_idx + buf * idx_words + (PAIRS ? r : r * K);
      list = s_list + buf * TC_MAX_T;
      return true;
    };
    auto advance = [&]() {
      while (n_rel >= Tp) {
        n_rel -= Tp;
        if (!next_tile()) {
          valid = false;
          Tp = 0x40000000;  // stay invalid
          n_rel = 0;
          return;
        }
      }
      valid = true;
      tloc = n_rel;
      n_rel += G;
    };
    auto issue = [&](float (&dst)[16]) {
      const uint32_t e = list[tloc];
      const int gi = my_idx[PAIRS ? 0 : (int)(e & 255u)];
      const float* p = (gi >= 0) ? Ag + ((int64_t)gi * c_in + (int)(e >> 8) * 16) : g_zero_row;
      asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=f"(dst[0]), "=f"(dst[1]), "=f"(dst[2]), "=f"(dst[3]), "=f"(dst[4]), "=f"(dst[5]), "=f"(dst[6]),
                     "=f"(dst[7])
                   : "l"(p));
      asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=f"(dst[8]), "=f"(dst[9]), "=f"(dst[10]), "=f"(dst[11]), "=f"(dst[12]), "=f"(dst[13]),
                     "=f"(dst[14]), "=f"(dst[15])
                   : "l"(p + 8));
    };
    const uint32_t a_lane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a.a_col0 + g * A_COLS);
    const int fat_cols = G * A_COLS;
    int st_s = 0;        // fat stage of the next owned slab
    uint32_t st_ph = 0;  // its phase bit
    auto store = [&](const float (&src)[16], bool real) {
      mbar_wait(bar(B_EMPTY + st_s), st_ph ^ 1u);
      if (real) {
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
        const uint32_t col = a_lane + (uint32_t)(st_s * fat_cols);
        tmem_st16(col, hi);
        if (NSPLIT == 3) tmem_st16(col + 16, lo);
        tmem_wait_st();
        tc_fence_before();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_FULL + st_s));
      const bool wrap = (st_s + 1 == NS);
      st_s = wrap ? 0 : st_s + 1;
      st_ph ^= wrap ? 1u : 0u;
    };
    float ra[16], rb[16];
    advance();
    if (valid) {
      bool real_a = tloc < T, real_b = false;
      if (real_a) issue(ra);
      while (true) {
        advance();
        const bool vb = valid;
        real_b = vb && tloc < T;
        if (real_b) issue(rb);
        store(ra, real_a);
        if (!vb) break;
        advance();
        const bool va = valid;
        real_a = va && tloc < T;
        if (real_a) issue(ra);
        store(rb, real_b);
        if (!va) break;
      }
    }
  } else if (warp == TC_W_MMA) {
    // =========================== MMA issuer ===================================================
    // The whole warp runs the loop with warp-uniform values only (kernel arguments, tile meta read through
    // one broadcast): everything stays in uniform registers and one elected lane issues.  Per fat stage: one
    // barrier handshake, then up to G slabs x 2 K-slices x (1 | 3) tcgen05.mma back to back, one commit.
    // Weight slots are derived arithmetically from the tile's offset mask (no shared-memory lookups): a
    // per-slab handshake / lookup chain kept this warp 100 % busy at 600+ cycles per slab.
    const uint32_t idesc = make_idesc_tf32(a.c_out);
    const uint64_t desc_hi = make_desc_sw64(0u);  // everything but the start-address field
    const uint32_t fat_cols = (uint32_t)(G * A_COLS);
    const uint32_t wbase16 = (base >> 4) & 0x3FFFu;
    const uint32_t slot16 = b_slot >> 4, slab16 = b_slab >> 4;
    const uint32_t a_base = tmem_base + (uint32_t)a.a_col0;
    const int krev = a.k_reversed;
    int s = 0, j = 0;
    uint32_t ph = 0, jph = 0;
    if (RESIDENT) mbar_wait(bar(B_WFULL), 0);
    bool ready = false;  // result of the early probe of full[s]
    for (int it = 0;; ++it) {
      if (blockIdx.x + (int64_t)it * gridDim.x >= n_tiles) break;
      const int buf = it & (TC_NBUF - 1);
      mbar_wait(bar(B_TREADY + buf), (uint32_t)(it >> 2) & 1u);
      const int T = __shfl_sync(0xffffffffu, s_meta[buf].T, 0);
      if (T < 0) break;
      uint32_t km = (uint32_t)__shfl_sync(0xffffffffu, s_meta[buf].pad, 0);  // active-offset mask of the tile
      const int acc = it & nacc_shift;  // nacc 1 -> 0; nacc 2 -> it & 1
      const int use = it >> nacc_shift;
      if (use >= 1) mbar_wait(bar(B_AEMPTY + acc), (uint32_t)(use - 1) & 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * a.c_out);
      if (T == 0 && lane == 0) mbar_arrive(bar(B_AFULL + acc));
      int k = 0, c = nc;  // (kernel offset, channel chunk) of the next slab
      for (int t0 = 0; t0 < T; t0 += G) {
        const int nv = min(G, T - t0);
        if (!RESIDENT) {
          int jj = j;
          uint32_t pp = jph;
          for (int u = 0; u < nv; ++u) {
            mbar_wait(bar(B_WFULL + jj), pp);
            if (++jj == SB) {
              jj = 0;
              pp ^= 1u;
            }
          }
        }
        if (!ready) mbar_wait(bar(B_FULL + s), ph);
        {  // probe the next stage now: the probe latency hides behind the MMA issue below
          const bool wrap = (s + 1 == NS);
          ready = mbar_try(bar(B_FULL + (wrap ? 0 : s + 1)), wrap ? (ph ^ 1u) : ph);
        }
        tc_fence_after();
        // weight descriptors of the stage's slabs (uniform arithmetic)
        uint32_t dlo[TC_G];
#pragma unroll
        for (int u = 0; u < TC_G; ++u) {
          if (u < nv) {
            if (c == nc) {
              k = __ffs(km) - 1;
              km &= km - 1;
              c = 0;
            }
            if (RESIDENT) {
              dlo[u] = wbase16 + (uint32_t)((krev ? (K - 1 - k) : k) * nc + c) * slot16;
            } else {
              int jj = j + u;
              jj -= (jj >= SB) ? SB : 0;
              dlo[u] = wbase16 + (uint32_t)jj * slot16;
            }
            ++c;
          } else {
            dlo[u] = 0;
          }
        }
        if (elect_one()) {
          const uint32_t ta0 = a_base + (uint32_t)s * fat_cols;
#pragma unroll
          for (int u = 0; u < TC_G; ++u) {
            if (u < nv) {
              const uint64_t db_hi = desc_hi | (uint64_t)dlo[u];
              const uint64_t db_lo = db_hi + (uint64_t)slab16;
              const uint32_t ta_hi = ta0 + (uint32_t)(u * A_COLS);
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) {  // two K = 8 slices: +8 TMEM columns (A), +32 bytes = +2 encoded (B)
                const uint32_t accf = (t0 > 0 || u > 0 || ks > 0) ? 1u : 0u;
                if (NSPLIT == 3) {
                  umma_tf32_ts(d_tmem, ta_hi + 16 + 8 * ks, db_hi + 2 * ks, idesc, accf);  // lo * hi
                  umma_tf32_ts(d_tmem, ta_hi + 8 * ks, db_lo + 2 * ks, idesc, 1u);         // hi * lo
                  umma_tf32_ts(d_tmem, ta_hi + 8 * ks, db_hi + 2 * ks, idesc, 1u);         // hi * hi
                } else {
                  umma_tf32_ts(d_tmem, ta_hi + 8 * ks, db_hi + 2 * ks, idesc, accf);
                }
              }
              if (!RESIDENT) {  // the weight slot is free once the MMAs above have read it
                int jj = j + u;
                jj -= (jj >= SB) ? SB : 0;
                umma_commit(bar(B_WEMPTY + jj));
              }
            }
          }
          umma_commit(bar(B_EMPTY + s));                       // frees the fat A stage
          if (t0 + G >= T) umma_commit(bar(B_AFULL + acc));    // accumulator complete
        }
        __syncwarp();
        if (++s == NS) {
          s = 0;
          ph ^= 1u;
        }
        if (!RESIDENT) {
          j += nv;
          if (j >= SB) {
            j -= SB;
            jph ^= 1u;
          }
        }
      }
      if (lane == 0) mbar_arrive(bar(B_TEMPTY + buf));
    }
  } else if (warp == TC_W_TILE) {
    // =========================== tile loader ==================================================
    // the bulk copy of a tile's neighbour-table slice is issued two tiles ahead of its use
    uint32_t cp_ph = 0;  // bit b = phase of B_TCOPY + b
    const bool can_bulk = !PAIRS && a.idx != nullptr && a.idx_bulk;
    auto issue_copy = [&](int j) {
      const int64_t tile = blockIdx.x + (int64_t)j * gridDim.x;
      if (tile >= n_tiles) return;
      const int b = j & (TC_NBUF - 1);
      if (j >= TC_NBUF) mbar_wait(bar(B_TEMPTY + b), (uint32_t)((j >> 2) - 1) & 1u);
      if (can_bulk && (tile + 1) * TC_BM <= a.n_out) {
        if (elect_one()) {
          mbar_arrive_expect_tx(bar(B_TCOPY + b), (uint32_t)(TC_BM * K * 4));
          bulk_g2s(smem_u32(s_idx + b * idx_words), a.idx + tile * TC_BM * K, (uint32_t)(TC_BM * K * 4),
                   bar(B_TCOPY + b));
        }
        __syncwarp();
      }
    };
    issue_copy(0);
    issue_copy(1);
    for (int it = 0;; ++it) {
      const int64_t tile = blockIdx.x + (int64_t)it * gridDim.x;
      if (tile >= n_tiles) break;
      issue_copy(it + 2);
      const int buf = it & (TC_NBUF - 1);
      int32_t* sidx = s_idx + buf * idx_words;
      uint16_t* list = s_list + buf * TC_MAX_T;
      int T = 0, rows = 0, base_row = 0;
      uint32_t kmask = 0;
      if (!PAIRS) {
        const int64_t row0 = tile * TC_BM;
        rows = (int)min((int64_t)TC_BM, a.n_out - row0);
        base_row = (int)row0;
        if (a.idx == nullptr) {  // identity map (1x1 conv), K == 1
          for (int e = lane; e < TC_BM; e += 32) sidx[e] = (e < rows) ? (int)(row0 + e) : -1;
          kmask = 1u;
        } else {
          uint32_t m = 0;
          if (a.tile_mask != nullptr) m = __ldg(a.tile_mask + tile);
          if (can_bulk && rows == TC_BM) {
            mbar_wait(bar(B_TCOPY + buf), (cp_ph >> buf) & 1u);
            cp_ph ^= 1u << buf;
            if (a.tile_mask == nullptr) {  // no precomputed mask: scan the tile in shared memory
              const int n4 = TC_BM * K / 4;
              for (int e4 = lane; e4 < n4; e4 += 32) {
                const int4 v = ((const int4*)sidx)[e4];
                int k0 = (e4 * 4) % K;
                const int vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  m |= (vv[i] >= 0) ? (1u << k0) : 0u;
                  k0 = (k0 + 1 == K) ? 0 : k0 + 1;
                }
              }
            }
          } else {
            const int32_t* p = a.idx + row0 * K;
            const int total = rows * K;
            for (int e = lane; e < TC_BM * K; e += 32) {
              const int gi = (e < total) ? __ldg(p + e) : -1;
              sidx[e] = gi;
              m |= (gi >= 0) ? (1u << (e % K)) : 0u;
            }
          }
          kmask = __reduce_or_sync(0xffffffffu, m);
        }
        T = __popc(kmask) * nc;
      } else {
        // pair tile -> (kernel offset, chunk): warp scan over the per-offset chunk counts
        int begin = 0, end = 0, nch = 0;
        if (lane < K) {
          begin = a.k_offsets[lane];
          end = a.k_offsets[lane + 1];
          nch = (end - begin + TC_BM - 1) / TC_BM;
        }
        int incl = nch;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int o = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += o;
        }
        const int excl = incl - nch;
        const bool hit = (lane < K) && (int64_t)excl <= tile && tile < (int64_t)incl;
        const uint32_t bal = __ballot_sync(0xffffffffu, hit);
        if (bal == 0u) {
          T = -1;
        } else {
          const int src = __ffs(bal) - 1;
          const int p0 = __shfl_sync(0xffffffffu, begin + ((int)tile - excl) * TC_BM, src);
          const int e_end = __shfl_sync(0xffffffffu, end, src);
          rows = min(TC_BM, e_end - p0);
          base_row = p0;
          for (int e = lane; e < TC_BM; e += 32) {
            sidx[e] = (e < rows) ? a.idx[p0 + e] : -1;
            sidx[TC_BM + e] = (e < rows) ? a.dst[p0 + e] : -1;
          }
          kmask = 1u << src;
          T = nc;
        }
      }
      for (int t = lane; t < T; t += 32) {
        const int ki = t / nc, c = t - ki * nc;
        const int k = __fns(kmask, 0, ki + 1);
        list[t] = (uint16_t)(k | (c << 8));
      }
      if (lane == 0) {
        s_meta[buf].T = T;
        s_meta[buf].rows = rows;
        s_meta[buf].base = base_row;
        s_meta[buf].pad = (int)kmask;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_TREADY + buf));
      if (T < 0) break;
    }
  } else if (warp == TC_W_WGT) {
    // =========================== weight loader (TMA bulk copies) ==============================
    if (RESIDENT) {
      if (elect_one()) {
        const uint32_t bytes = (uint32_t)SB * b_slot;
        mbar_arrive_expect_tx(bar(B_WFULL), bytes);
        bulk_g2s(base, a.Bp, bytes, bar(B_WFULL));
      }
      __syncwarp();
    } else {
      int j = 0;
      uint32_t jph = 0;
      for (int it = 0;; ++it) {
        if (blockIdx.x + (int64_t)it * gridDim.x >= n_tiles) break;
        const int buf = it & (TC_NBUF - 1);
        mbar_wait(bar(B_TREADY + buf), (uint32_t)(it >> 2) & 1u);
        const int T = s_meta[buf].T;
        if (T < 0) break;
        const uint16_t* list = s_list + buf * TC_MAX_T;
        for (int t = 0; t < T; ++t) {
          const uint32_t e = list[t];
          const int k = (int)(e & 255u);
          const int widx = (a.k_reversed ? (K - 1 - k) : k) * nc + (int)(e >> 8);
          mbar_wait(bar(B_WEMPTY + j), jph ^ 1u);
          if (elect_one()) {
            mbar_arrive_expect_tx(bar(B_WFULL + j), b_slot);
            bulk_g2s(base + (uint32_t)j * b_slot, a.Bp + (int64_t)widx * (b_slot / 4), b_slot, bar(B_WFULL + j));
          }
          __syncwarp();
          if (++j == SB) {
            j = 0;
            jph ^= 1u;
          }
        }
        if (lane == 0) mbar_arrive(bar(B_TEMPTY + buf));
      }
    }
  } else if (warp >= TC_W_EPI) {
    // =========================== epilogue (thread = TMEM lane = tile row) =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;
    for (int it = 0;; ++it) {
      if (blockIdx.x + (int64_t)it * gridDim.x >= n_tiles) break;
      const int buf = it & (TC_NBUF - 1);
      mbar_wait(bar(B_TREADY + buf), (uint32_t)(it >> 2) & 1u);
      const int T = s_meta[buf].T;
      if (T < 0) break;
      const int rows = s_meta[buf].rows;
      int64_t orow = -1;
      if (r < rows) orow = PAIRS ? (int64_t)s_idx[buf * idx_words + TC_BM + r] : (int64_t)s_meta[buf].base + r;
      const int acc = it & nacc_shift;
      const int use = it >> nacc_shift;
      mbar_wait(bar(B_AFULL + acc), (uint32_t)use & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * a.c_out);
      for (int col = 0; col < a.c_out; col += 16) {
        uint32_t v[16];
        if (T > 0) {
          tmem_ld16(taddr + (uint32_t)col, v);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = 0u;
        }
        if (orow >= 0) {
          float4* o = (float4*)(a.out + orow * a.c_out + col);
#pragma unroll
          for (int qq = 0; qq < 4; ++qq)
            o[qq] = make_float4(__uint_as_float(v[4 * qq]), __uint_as_float(v[4 * qq + 1]),
                                __uint_as_float(v[4 * qq + 2]), __uint_as_float(v[4 * qq + 3]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar(B_AEMPTY + acc));
        mbar_arrive(bar(B_TEMPTY + buf));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == TC_W_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

bool conv_tc_supported(int K, int c_in, int c_out) {
  return K >= 1 && K <= 32 && c_in >= 16 && (c_in % 16) == 0 && c_in <= 256 && c_out >= 16 && (c_out % 16) == 0 &&
         c_out <= 256;
}

size_t conv_tc_ws_bytes(int K, int c_in, int c_out) { return align_up((size_t)K * c_in * c_out * 4 * 2) + 256; }

static int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <bool PAIRS, int NSPLIT, bool RESIDENT>
static int launch_tc2(TcArgs a, size_t smem, cudaStream_t stream) {
  auto kern = conv_tc_kernel<PAIRS, NSPLIT, RESIDENT>;
  static size_t configured = 0;
  if (smem > configured) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = smem;
  }
  const int grid = std::max(1, std::min(a.n_tiles, sm_count()));
  kern<<<grid, TC_THREADS, smem, stream>>>(a);
  return check_launch("conv_tc");
}

template <bool PAIRS, int NSPLIT>
static int launch_tc(TcArgs a, cudaStream_t stream) {
  constexpr int NB = NSPLIT == 3 ? 2 : 1;
  constexpr int A_COLS = 16 * NB;
  const int b_slot = NB * a.c_out * 64;
  const int total_slabs = a.K * (a.c_in / 16);
  // tensor memory (all 512 columns, one CTA per SM): two accumulators when that leaves >= 2 fat A stages
  a.groups = TC_G;
  a.flags = 0;
  if (const char* e = getenv("B2S_TC_GROUPS")) a.groups = std::max(1, std::min(TC_G, atoi(e)));
  if (const char* e = getenv("B2S_TC_DEBUG")) a.flags = atoi(e);
  const int fat_cols = a.groups * A_COLS;
  a.nacc = 2;
  a.ns = (512 - 2 * a.c_out) / fat_cols;
  if (a.ns < 2) {
    a.nacc = 1;
    a.ns = (512 - a.c_out) / fat_cols;
  }
  a.ns = std::min(a.ns, TC_MAX_NS);
  if (const char* e = getenv("B2S_TC_NS")) a.ns = std::max(2, std::min(a.ns, atoi(e)));
  a.a_col0 = a.nacc * a.c_out;
  // weights: resident when the whole layer fits in 128 KB, else a ring of up to 32 slots in <= 128 KB
  const int w_cap = 128 * 1024;
  a.resident = (int64_t)total_slabs * b_slot <= w_cap ? 1 : 0;
  a.sb = a.resident ? total_slabs : std::max(2, std::min(TC_MAX_SB, w_cap / b_slot));
  const int idx_words = PAIRS ? 2 * TC_BM : TC_BM * a.K;
  size_t smem = 1024 /*align slack*/ + (size_t)a.sb * b_slot + ((TC_NBUF * (size_t)idx_words * 4 + 15) & ~(size_t)15) +
                TC_NBUF * TC_MAX_T * 2 + TC_NBUF * sizeof(TcMeta) +
                8 * (2 * TC_MAX_NS + 2 * TC_MAX_SB + 3 * TC_NBUF + 4) + 16;
  return a.resident ? launch_tc2<PAIRS, NSPLIT, true>(a, smem, stream) : launch_tc2<PAIRS, NSPLIT, false>(a, smem, stream);
}

// ws: packed weights, conv_tc_ws_bytes(K, c_in, c_out) bytes
int conv_tc(const float* A, const float* W, const int32_t* idx, const int32_t* dst, const int32_t* k_offsets,
            const uint32_t* tile_mask, float* out, int64_t n_out, int64_t max_pairs, int K, int c_in, int c_out, int wT, int krev, int nsplit,
            bool pairs, void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (ws_bytes < conv_tc_ws_bytes(K, c_in, c_out)) {
    set_error("conv_tc: workspace too small");
    return B2S_E_WORKSPACE;
  }
  if ((!pairs && n_out == 0) || (pairs && max_pairs == 0)) return B2S_OK;
  const int64_t total = (int64_t)K * c_in * c_out;
  float* Bp = (float*)ws;
  pack_weights_kernel<<<(unsigned)cdiv(total, 256), 256, 0, stream>>>(W, Bp, K, c_in, c_out, wT, nsplit == 3 ? 2 : 1,
                                                                      total);
  TcArgs a;
  a.A = A;
  a.Bp = Bp;
  a.idx = idx;
  a.dst = dst;
  a.k_offsets = k_offsets;
  a.tile_mask = (!pairs && K <= 32) ? tile_mask : nullptr;
  a.out = out;
  a.n_out = n_out;
  a.K = K;
  a.c_in = c_in;
  a.c_out = c_out;
  a.k_reversed = krev;
  a.ns = a.nacc = a.a_col0 = a.sb = a.resident = 0;
  a.idx_bulk = (!pairs && idx != nullptr && (((uintptr_t)idx) & 15) == 0) ? 1 : 0;
  const int64_t tiles = pairs ? cdiv(max_pairs, (int64_t)TC_BM) + K : cdiv(n_out, (int64_t)TC_BM);
  if (tiles > 0x7fffffff) {
    set_error("conv_tc: too many tiles");
    return B2S_E_INVALID;
  }
  a.n_tiles = (int)tiles;
  if (!pairs) return nsplit == 3 ? launch_tc<false, 3>(a, stream) : launch_tc<false, 1>(a, stream);
  return nsplit == 3 ? launch_tc<true, 3>(a, stream) : launch_tc<true, 1>(a, stream);
}

}  // namespace b2s
