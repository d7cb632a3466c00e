// T3 weight gradient on the tensor cores:  gW[k][a][g] = sum over the pairs (i, o) of offset k of A[i][a] * G[o][g]
// (gkernel[k] = in[I_k]^T @ gout[O_k]).
//
// Why warp-level mma.sync (m16n8k8, TF32) and not tcgen05 here.  The contraction runs over PAIRS, so the pair index
// is the K dimension of the MMA and both operands arrive "transposed" (a gathered row holds one K index and all M or
// N indices).  tcgen05 would need (i) the A operand in shared memory (TMEM operands cannot be transposed:
// cute/arch/mma_sm100_umma.hpp "A from TMEM can't be transposed"), i.e. the SS form whose 3xTF32 variant re-reads
// every operand three times from shared memory -- the round-1 forward kernel was bound by exactly that traffic
// (profiles/r01_conv_tc_smem_bound.txt) -- and (ii) M = 64/128 with N = c_g: for the narrow layers that carry most
// of the time (c = 16, 32) that is 64 x 16 x 8 = 16 kFLOP per instruction, > 300 single-thread-issued MMAs per
// 128-row tile.  With mma.sync the operands are built in REGISTERS straight from the staged rows (conflict-free
// LDS.32 fragments, hi/lo split in registers), M = 16 fits c_a = 16 exactly, and the fp32 FMA kernel's 62 M warp
// instructions per level-0 launch (38 % of them FMAs, FMA pipe 31 % busy, smem-latency bound: ncu,
// profiles/r02_wgrad_ncu.txt) shrink ~6x.  3xTF32 (hi*hi + lo*hi + hi*lo, fp32 accumulate) keeps fp32-class accuracy.
//
// Structure (same chunk walk as conv_wgrad_kernel in conv_simt.cu): persistent CTAs walk contiguous runs of
// WG_PC-pair chunks of the canonical pair lists (sorted by offset, then output row); the gathered A / G rows of the
// next chunk are prefetched into registers while the current chunk is multiplied out of shared memory; a CTA owns a
// BA x BG channel tile; its 8 warps split the tile (WM x WN warps) and the chunk's pairs (KG = 8 / (WM * WN) groups);
// partial sums are combined through shared memory when the offset changes and leave the CTA as one atomicAdd per
// tile element.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace b2s {

constexpr int WM_THREADS = 256;

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// BA x BG: channel tile of the CTA; MT x NT: 16 x 8 MMA tiles per warp; PC: pairs per chunk (multiple of 8 * KG)
template <int BA, int BG, int MT, int NT, int PC>
__global__ void __launch_bounds__(WM_THREADS)
    conv_wgrad_mma_kernel(const float* __restrict__ A, const float* __restrict__ G, const int32_t* __restrict__ src,
                          const int32_t* __restrict__ dst, const int32_t* __restrict__ k_offsets,
                          float* __restrict__ gW, int K, int c_a, int c_g) {
  constexpr int WM = BA / (16 * MT), WN = BG / (8 * NT);  // warps along the tile
  constexpr int KG = (WM_THREADS / 32) / (WM * WN);        // pair-splitting warp groups
  static_assert(WM * WN * KG == WM_THREADS / 32 && KG >= 1, "bad warp tiling");
  static_assert(PC % (8 * KG) == 0, "chunk must split into whole k-steps per group");
  constexpr int LDA = BA + 8, LDG = BG + 8;  // row stride = 8 mod 32 floats: the fragment loads hit 32 distinct banks
  extern __shared__ __align__(16) float s_dyn[];
  float* As = s_dyn;               // [PC][LDA]
  float* Gs = s_dyn + PC * LDA;    // [PC][LDG]
  __shared__ int s_cum[130];       // cumulative chunk counts per k (K <= 125)
  __shared__ int s_koff[130];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gid = lane >> 2, tq = lane & 3;  // mma fragment coordinates
  const int kg = warp / (WM * WN), wt = warp % (WM * WN);
  const int wm = wt / WN, wn = wt % WN;
  const int a0 = blockIdx.y * BA, g0 = blockIdx.z * BG;
  const int am0 = wm * 16 * MT, gn0 = wn * 8 * NT;  // this warp's tile origin inside the CTA tile

  for (int k = tid; k <= K; k += WM_THREADS) s_koff[k] = __ldg(k_offsets + k);
  __syncthreads();
  if (tid == 0) {
    int cum = 0;
    for (int k = 0; k < K; ++k) {
      s_cum[k] = cum;
      cum += (s_koff[k + 1] - s_koff[k] + PC - 1) / PC;
    }
    s_cum[K] = cum;
  }
  __syncthreads();
  const int total_chunks = s_cum[K];
  const int per = (total_chunks + gridDim.x - 1) / gridDim.x;
  const int ch_begin = blockIdx.x * per;
  const int ch_end = min(total_chunks, ch_begin + per);

  float acc[MT][NT][4];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[i][j][q] = 0.f;
  // M / N tiles of this warp that lie inside the real channel counts (c = 48 in a 64-wide tile skips a quarter)
  bool m_ok[MT], n_ok[NT];
#pragma unroll
  for (int i = 0; i < MT; ++i) m_ok[i] = a0 + am0 + 16 * i < c_a;
#pragma unroll
  for (int j = 0; j < NT; ++j) n_ok[j] = g0 + gn0 + 8 * j < c_g;

  // flush: the KG groups hold partial sums of the same tile; combined through shared memory (the staging buffers are
  // free at that point), one atomicAdd per tile element leaves the CTA
  auto flush = [&](int k) {
    float* base = gW + (int64_t)k * c_a * c_g;
    float* red = s_dyn;  // KG * BA * BG floats <= staging size (checked on the host)
    __syncthreads();
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const int r = am0 + 16 * i + gid, c = gn0 + 8 * j + 2 * tq;
        float* p = red + (size_t)kg * BA * BG;
        p[r * BG + c] = acc[i][j][0];
        p[r * BG + c + 1] = acc[i][j][1];
        p[(r + 8) * BG + c] = acc[i][j][2];
        p[(r + 8) * BG + c + 1] = acc[i][j][3];
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[i][j][q] = 0.f;
      }
    __syncthreads();
    for (int e = tid; e < BA * BG; e += WM_THREADS) {
      float v = 0.f;
#pragma unroll
      for (int q = 0; q < KG; ++q) v += red[q * BA * BG + e];
      const int a = a0 + e / BG, g = g0 + e % BG;
      if (a < c_a && g < c_g && v != 0.f) atomicAdd(base + (int64_t)a * c_g + g, v);
    }
    __syncthreads();
  };

  // software pipeline over the chunks: indices, then rows of chunk ch + 1 into registers while chunk ch is multiplied
  constexpr int NA4 = (PC * (BA / 4) + WM_THREADS - 1) / WM_THREADS;
  constexpr int NG4 = (PC * (BG / 4) + WM_THREADS - 1) / WM_THREADS;
  float4 ra[NA4], rg[NG4];
  int k = 0, cur_k = -1, pk = 0, pnp = 0;
  auto prefetch = [&](int ch) {
    while (ch >= s_cum[k + 1]) ++k;
    pk = k;
    const int p0 = s_koff[k] + (ch - s_cum[k]) * PC;
    pnp = min(PC, s_koff[k + 1] - p0);
    int ia[NA4], ig[NG4];
#pragma unroll
    for (int u = 0; u < NA4; ++u) {
      const int e = tid + u * WM_THREADS, p = e / (BA / 4), c = (e - p * (BA / 4)) * 4;
      ia[u] = (e < PC * (BA / 4) && p < pnp && a0 + c < c_a) ? __ldg(src + p0 + p) : -1;
    }
#pragma unroll
    for (int u = 0; u < NG4; ++u) {
      const int e = tid + u * WM_THREADS, p = e / (BG / 4), c = (e - p * (BG / 4)) * 4;
      ig[u] = (e < PC * (BG / 4) && p < pnp && g0 + c < c_g) ? __ldg(dst + p0 + p) : -1;
    }
#pragma unroll
    for (int u = 0; u < NA4; ++u) {
      const int e = tid + u * WM_THREADS, p = e / (BA / 4), c = (e - p * (BA / 4)) * 4;
      ra[u] = ia[u] >= 0 ? __ldg((const float4*)(A + (int64_t)ia[u] * c_a + a0 + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < NG4; ++u) {
      const int e = tid + u * WM_THREADS, p = e / (BG / 4), c = (e - p * (BG / 4)) * 4;
      rg[u] = ig[u] >= 0 ? __ldg((const float4*)(G + (int64_t)ig[u] * c_g + g0 + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  if (ch_begin < ch_end) prefetch(ch_begin);
  for (int ch = ch_begin; ch < ch_end; ++ch) {
    if (pk != cur_k) {
      if (cur_k >= 0) flush(cur_k);
      cur_k = pk;
    }
    const int np = pnp;
#pragma unroll
    for (int u = 0; u < NA4; ++u) {
      const int e = tid + u * WM_THREADS, p = e / (BA / 4), c = (e - p * (BA / 4)) * 4;
      if (e < PC * (BA / 4)) *(float4*)&As[p * LDA + c] = ra[u];
    }
#pragma unroll
    for (int u = 0; u < NG4; ++u) {
      const int e = tid + u * WM_THREADS, p = e / (BG / 4), c = (e - p * (BG / 4)) * 4;
      if (e < PC * (BG / 4)) *(float4*)&Gs[p * LDG + c] = rg[u];
    }
    __syncthreads();
    if (ch + 1 < ch_end) prefetch(ch + 1);
    // ---- multiply: k-steps of 8 pairs, interleaved over the KG groups; rows past np are zero ------------------
    const int ksteps = (np + 7) >> 3;
    for (int ks = kg; ks < ksteps; ks += KG) {
      const float* ap = As + (ks * 8 + tq) * LDA + am0 + gid;
      const float* gp = Gs + (ks * 8 + tq) * LDG + gn0 + gid;
      uint32_t bh[NT][2], bl[NT][2];
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const float b0 = gp[8 * j], b1 = gp[4 * LDG + 8 * j];
        bh[j][0] = __float_as_uint(b0) & 0xFFFFE000u;
        bh[j][1] = __float_as_uint(b1) & 0xFFFFE000u;
        bl[j][0] = __float_as_uint(b0 - __uint_as_float(bh[j][0]));
        bl[j][1] = __float_as_uint(b1 - __uint_as_float(bh[j][1]));
      }
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        if (!m_ok[i]) continue;
        const float x0 = ap[16 * i], x1 = ap[16 * i + 8], x2 = ap[4 * LDA + 16 * i], x3 = ap[4 * LDA + 16 * i + 8];
        uint32_t ah[4], al[4];
        ah[0] = __float_as_uint(x0) & 0xFFFFE000u;
        ah[1] = __float_as_uint(x1) & 0xFFFFE000u;
        ah[2] = __float_as_uint(x2) & 0xFFFFE000u;
        ah[3] = __float_as_uint(x3) & 0xFFFFE000u;
        al[0] = __float_as_uint(x0 - __uint_as_float(ah[0]));
        al[1] = __float_as_uint(x1 - __uint_as_float(ah[1]));
        al[2] = __float_as_uint(x2 - __uint_as_float(ah[2]));
        al[3] = __float_as_uint(x3 - __uint_as_float(ah[3]));
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          if (!n_ok[j]) continue;
          mma_tf32(acc[i][j], al, bh[j]);  // small terms first
          mma_tf32(acc[i][j], ah, bl[j]);
          mma_tf32(acc[i][j], ah, bh[j]);
        }
      }
    }
    __syncthreads();
  }
  if (cur_k >= 0) flush(cur_k);
}

template <int BA, int BG, int MT, int NT, int PC>
static int launch_wgrad_mma(const float* A, const float* G, const int32_t* src, const int32_t* dst,
                            const int32_t* k_offsets, float* gW, int K, int c_a, int c_g, int64_t max_pairs,
                            cudaStream_t stream) {
  constexpr int KG = 8 / ((BA / (16 * MT)) * (BG / (8 * NT)));
  constexpr size_t stage = (size_t)PC * (BA + 8 + BG + 8) * 4;
  static_assert((size_t)KG * BA * BG * 4 <= stage, "flush scratch must fit the staging buffers");
  auto kern = conv_wgrad_mma_kernel<BA, BG, MT, NT, PC>;
  static bool configured[B2S_MAX_DEVICES] = {};
  const int dev = current_device();
  if (stage + 2048 > 48 * 1024 && !configured[dev]) {  // + static shared memory of the kernel
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage);
    configured[dev] = true;
  }
  const int64_t chunks = cdiv(max_pairs, PC) + K;
  const int ty = (int)cdiv(c_a, BA), tz = (int)cdiv(c_g, BG);
  const int gx = (int)std::min<int64_t>(chunks, std::max(1, 4 * sm_count() / (ty * tz)));
  kern<<<dim3(gx, ty, tz), WM_THREADS, stage, stream>>>(A, G, src, dst, k_offsets, gW, K, c_a, c_g);
  return check_launch("conv_wgrad_mma");
}

bool conv_wgrad_mma_supported(int c_a, int c_g) { return (c_a % 4) == 0 && (c_g % 4) == 0 && c_a >= 16 && c_g >= 16; }

int conv_wgrad_mma(const float* A, const float* G, const int32_t* src, const int32_t* dst, const int32_t* k_offsets,
                   float* gW, int K, int c_a, int c_g, int64_t max_pairs, cudaStream_t stream) {
  cudaMemsetAsync(gW, 0, (size_t)K * c_a * c_g * 4, stream);
  if (max_pairs == 0) return check_launch("conv_wgrad(empty)");
  if (c_a <= 16 && c_g <= 16)
    return launch_wgrad_mma<16, 16, 1, 2, 256>(A, G, src, dst, k_offsets, gW, K, c_a, c_g, max_pairs, stream);
  if (c_a <= 32 && c_g <= 32)
    return launch_wgrad_mma<32, 32, 2, 2, 128>(A, G, src, dst, k_offsets, gW, K, c_a, c_g, max_pairs, stream);
  return launch_wgrad_mma<64, 64, 2, 4, 64>(A, G, src, dst, k_offsets, gW, K, c_a, c_g, max_pairs, stream);
}

}  // namespace b2s
