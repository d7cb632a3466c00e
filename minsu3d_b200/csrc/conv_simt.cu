// T3 / T4: sparse convolution, fp32 FMA (SIMT) implicit-GEMM path.
//
// Restates MinkowskiConvolution / MinkowskiConvolutionTranspose forward/backward as used by
// minsu3d (minsu3d/model/module/common.py:12,31,37,40,69,77; backbone.py:14; SURVEY.md
// appendix A items 7-8):  out[o] = sum_k in[nbr[o][k]] @ W[k], k ascending.
//
// One output-stationary kernel serves every data-path product:
//   TABLE mode  : tile = BM consecutive output rows, loops over the K offsets of the neighbour
//                 table (conv fwd, conv dgrad with reversed/transposed weights, strided conv fwd,
//                 transposed-conv dgrad, 1x1 conv with nbr == NULL)
//   PAIRS mode  : tile = BM pairs of ONE offset, gather src -> matmul -> store at dst
//                 (transposed-conv fwd, strided-conv dgrad)
// and one pair-list kernel computes the weight gradient.  Accumulation is fp32 FMA in a fixed
// order (k ascending, channel ascending): results are run-to-run deterministic.
#include <algorithm>

#include "common.cuh"

namespace b2s {

constexpr int IG_BM = 128;      // output rows (or pairs) per CTA
constexpr int IG_BK = 16;       // reduction channels per shared-memory stage
constexpr int IG_THREADS = 128; // 16 row groups x 8 column groups
constexpr int IG_TM = 8;

template <int BN, bool PAIRS>
__global__ void __launch_bounds__(IG_THREADS)
    conv_igemm_kernel(const float* __restrict__ A, const float* __restrict__ W,
                      const int32_t* __restrict__ nbr_or_src, const int32_t* __restrict__ dst,
                      const int32_t* __restrict__ k_offsets, float* __restrict__ out, int64_t n_out,
                      int K, int c_in, int c_out, int w_transposed, int k_reversed) {
  constexpr int TN = BN / 8;
  constexpr int LDA = IG_BM + 4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* As = (float*)smem_raw;                 // [IG_BK][LDA]
  float* Ws = As + IG_BK * LDA;                 // [IG_BK][BN]
  int32_t* s_idx = (int32_t*)(Ws + IG_BK * BN); // TABLE: [IG_BM*K] ; PAIRS: src[IG_BM], dst[IG_BM]

  const int tid = threadIdx.x;
  const int tx = tid & 7, ty = tid >> 3;
  const int col0 = blockIdx.y * BN;

  int64_t row0 = 0;
  int rows = 0;
  int k_begin = 0, k_end = K;
  if (!PAIRS) {
    row0 = (int64_t)blockIdx.x * IG_BM;
    rows = (int)min((int64_t)IG_BM, n_out - row0);
    if (nbr_or_src != nullptr) {
      const int32_t* p = nbr_or_src + row0 * K;
      for (int e = tid; e < rows * K; e += IG_THREADS) s_idx[e] = p[e];
    }
  } else {
    // chunk id -> (k, range)
    int chunk = blockIdx.x;
    int kk = 0;
    int begin = 0, end = 0;
    bool found = false;
    for (kk = 0; kk < K; ++kk) {
      begin = k_offsets[kk];
      end = k_offsets[kk + 1];
      int nch = (end - begin + IG_BM - 1) / IG_BM;
      if (chunk < nch) {
        found = true;
        break;
      }
      chunk -= nch;
    }
    if (!found) return;
    k_begin = kk;
    k_end = kk + 1;
    int p0 = begin + chunk * IG_BM;
    rows = min(IG_BM, end - p0);
    for (int e = tid; e < rows; e += IG_THREADS) {
      s_idx[e] = nbr_or_src[p0 + e];
      s_idx[IG_BM + e] = dst[p0 + e];
    }
  }
  __syncthreads();

  float acc[IG_TM][TN];
#pragma unroll
  for (int i = 0; i < IG_TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const bool vec_a = ((c_in & 3) == 0);
  for (int k = k_begin; k < k_end; ++k) {
    // gather row of this thread for offset k
    int g = -1;
    if (tid < rows) {
      if (PAIRS) g = s_idx[tid];
      else g = (nbr_or_src == nullptr) ? (int)(row0 + tid) : s_idx[tid * K + k];
    }
    if (!__syncthreads_or(g >= 0)) continue;  // whole tile has no input at this offset
    const int kw = k_reversed ? (K - 1 - k) : k;
    const float* Wk = W + (int64_t)kw * c_in * c_out;
    const float* arow = A + (int64_t)(g >= 0 ? g : 0) * c_in;

    for (int c0 = 0; c0 < c_in; c0 += IG_BK) {
      // ---- stage A tile: thread t owns tile row t (transposed store, conflict-free) ----
      if (vec_a) {
#pragma unroll
        for (int q = 0; q < IG_BK / 4; ++q) {
          int c = c0 + q * 4;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (g >= 0 && c < c_in) v = __ldg((const float4*)(arow + c));
          As[(q * 4 + 0) * LDA + tid] = v.x;
          As[(q * 4 + 1) * LDA + tid] = v.y;
          As[(q * 4 + 2) * LDA + tid] = v.z;
          As[(q * 4 + 3) * LDA + tid] = v.w;
        }
      } else {
#pragma unroll
        for (int q = 0; q < IG_BK; ++q) {
          int c = c0 + q;
          As[q * LDA + tid] = (g >= 0 && c < c_in) ? __ldg(arow + c) : 0.f;
        }
      }
      // ---- stage W tile [IG_BK][BN] ----
      if (!w_transposed) {
        for (int e = tid; e < IG_BK * BN; e += IG_THREADS) {
          int ci = e / BN, co = e - ci * BN;
          int c = c0 + ci, n = col0 + co;
          Ws[e] = (c < c_in && n < c_out) ? __ldg(Wk + (int64_t)c * c_out + n) : 0.f;
        }
      } else {
        for (int e = tid; e < IG_BK * BN; e += IG_THREADS) {
          int co = e / IG_BK, ci = e - co * IG_BK;
          int c = c0 + ci, n = col0 + co;
          Ws[ci * BN + co] = (c < c_in && n < c_out) ? __ldg(Wk + (int64_t)n * c_in + c) : 0.f;
        }
      }
      __syncthreads();
      // ---- register-tiled FMA ----
#pragma unroll
      for (int ci = 0; ci < IG_BK; ++ci) {
        float a[IG_TM], b[TN];
        const float4* ap = (const float4*)(As + ci * LDA + ty * IG_TM);
        float4 a0 = ap[0], a1 = ap[1];
        a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
        a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = Ws[ci * BN + tx * TN + j];
#pragma unroll
        for (int i = 0; i < IG_TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  // ---- epilogue ----
#pragma unroll
  for (int i = 0; i < IG_TM; ++i) {
    int r = ty * IG_TM + i;
    if (r >= rows) continue;
    int64_t orow = PAIRS ? (int64_t)s_idx[IG_BM + r] : row0 + r;
    float* op = out + orow * c_out + col0 + tx * TN;
#pragma unroll
    for (int j = 0; j < TN; ++j)
      if (col0 + tx * TN + j < c_out) op[j] = acc[i][j];
  }
}

// ------------------------------------------------------------------------------------------
// weight gradient: gW[k][a][g] = sum over pairs of offset k of A[src][a] * G[dst][g]
// persistent CTAs walk contiguous runs of 64-pair chunks; BA x BG channel tile per CTA.
// ------------------------------------------------------------------------------------------
constexpr int WG_THREADS = 256;

template <int BA, int BG, int TA, int TG, int WG_PC>
__global__ void __launch_bounds__(WG_THREADS)
    conv_wgrad_kernel(const float* __restrict__ A, const float* __restrict__ G,
                      const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                      const int32_t* __restrict__ k_offsets, float* __restrict__ gW, int K, int c_a,
                      int c_g) {
  constexpr int NT = (BA / TA) * (BG / TG);  // threads covering the tile once
  constexpr int NG = WG_THREADS / NT;        // pair-splitting groups
  static_assert(NG >= 1 && NG * NT == WG_THREADS, "bad wgrad tiling");
  __shared__ __align__(16) float s_stage[WG_PC * (BA + 4) + WG_PC * (BG + 4)];
  float (*As)[BA + 4] = (float (*)[BA + 4])s_stage;
  float (*Gs)[BG + 4] = (float (*)[BG + 4])(s_stage + WG_PC * (BA + 4));
  __shared__ int s_cum[130];  // cumulative chunk counts per k (K <= 125)
  __shared__ int s_koff[130];

  const int tid = threadIdx.x;
  const int grp = tid / NT, t = tid % NT;
  const int tg = t % (BG / TG), ta = t / (BG / TG);
  const int a0 = blockIdx.y * BA, g0 = blockIdx.z * BG;

  // offsets first, all threads in parallel; the serial scan then runs out of shared memory (thread 0 walking K + 1
  // dependent global loads cost ~8 us per launch, a third of the small-map launches' time)
  for (int k = tid; k <= K; k += WG_THREADS) s_koff[k] = __ldg(k_offsets + k);
  __syncthreads();
  if (tid == 0) {
    int cum = 0;
    for (int k = 0; k < K; ++k) {
      s_cum[k] = cum;
      cum += (s_koff[k + 1] - s_koff[k] + WG_PC - 1) / WG_PC;
    }
    s_cum[K] = cum;
  }
  __syncthreads();
  const int total_chunks = s_cum[K];
  const int per = (total_chunks + gridDim.x - 1) / gridDim.x;
  const int ch_begin = blockIdx.x * per;
  const int ch_end = min(total_chunks, ch_begin + per);

  float acc[TA][TG];
#pragma unroll
  for (int i = 0; i < TA; ++i)
#pragma unroll
    for (int j = 0; j < TG; ++j) acc[i][j] = 0.f;

  // flush: the NG pair-splitting groups hold partial sums of the same BA x BG tile; they are combined through
  // shared memory (reusing the staging buffers) so that one atomicAdd per tile element leaves the CTA instead of NG
  auto flush = [&](int k) {
    float* base = gW + (int64_t)k * c_a * c_g;
    if (NG > 1 && NG * BA * BG <= WG_PC * (BA + 4) + WG_PC * (BG + 4)) {
      float* red = &As[0][0];  // As and Gs are adjacent static arrays; NG*BA*BG floats fit (checked above)
      __syncthreads();
#pragma unroll
      for (int i = 0; i < TA; ++i)
#pragma unroll
        for (int j = 0; j < TG; ++j) {
          red[(grp * BA + ta * TA + i) * BG + tg * TG + j] = acc[i][j];
          acc[i][j] = 0.f;
        }
      __syncthreads();
      for (int e = tid; e < BA * BG; e += WG_THREADS) {
        float v = 0.f;
#pragma unroll
        for (int q = 0; q < NG; ++q) v += red[q * BA * BG + e];
        const int a = a0 + e / BG, g = g0 + e % BG;
        if (a < c_a && g < c_g && v != 0.f) atomicAdd(base + (int64_t)a * c_g + g, v);
      }
      __syncthreads();
      return;
    }
#pragma unroll
    for (int i = 0; i < TA; ++i) {
      int a = a0 + ta * TA + i;
#pragma unroll
      for (int j = 0; j < TG; ++j) {
        int g = g0 + tg * TG + j;
        if (a < c_a && g < c_g && acc[i][j] != 0.f) atomicAdd(base + (int64_t)a * c_g + g, acc[i][j]);
        acc[i][j] = 0.f;
      }
    }
  };

  int k = 0;
  int cur_k = -1;
  const bool vec = ((c_a | c_g) & 3) == 0;
  // Software pipeline over the chunks (vector path): the gathered rows of chunk ch+1 are loaded into registers --
  // all indices first, then all rows, so one L2 round trip each -- while chunk ch is multiplied out of shared memory.
  // (The unpipelined loop spent 62 % of its time in four exposed index -> row load chains per chunk.)
  constexpr int NA4 = (WG_PC * (BA / 4) + WG_THREADS - 1) / WG_THREADS;
  constexpr int NG4 = (WG_PC * (BG / 4) + WG_THREADS - 1) / WG_THREADS;
  float4 ra[NA4], rg[NG4];
  int pk = 0, pnp = 0;  // offset and pair count of the prefetched chunk
  auto prefetch = [&](int ch) {
    while (ch >= s_cum[k + 1]) ++k;
    pk = k;
    const int p0 = s_koff[k] + (ch - s_cum[k]) * WG_PC;
    pnp = min(WG_PC, s_koff[k + 1] - p0);
    int ia[NA4], ig[NG4];
#pragma unroll
    for (int u = 0; u < NA4; ++u) {
      const int e = tid + u * WG_THREADS, p = e / (BA / 4), c = (e - p * (BA / 4)) * 4;
      ia[u] = (e < WG_PC * (BA / 4) && p < pnp && a0 + c < c_a) ? __ldg(src + p0 + p) : -1;
    }
#pragma unroll
    for (int u = 0; u < NG4; ++u) {
      const int e = tid + u * WG_THREADS, p = e / (BG / 4), c = (e - p * (BG / 4)) * 4;
      ig[u] = (e < WG_PC * (BG / 4) && p < pnp && g0 + c < c_g) ? __ldg(dst + p0 + p) : -1;
    }
#pragma unroll
    for (int u = 0; u < NA4; ++u) {
      const int e = tid + u * WG_THREADS, p = e / (BA / 4), c = (e - p * (BA / 4)) * 4;
      ra[u] = ia[u] >= 0 ? __ldg((const float4*)(A + (int64_t)ia[u] * c_a + a0 + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < NG4; ++u) {
      const int e = tid + u * WG_THREADS, p = e / (BG / 4), c = (e - p * (BG / 4)) * 4;
      rg[u] = ig[u] >= 0 ? __ldg((const float4*)(G + (int64_t)ig[u] * c_g + g0 + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  if (vec && ch_begin < ch_end) prefetch(ch_begin);
  for (int ch = ch_begin; ch < ch_end; ++ch) {
    int np;
    if (vec) {
      if (pk != cur_k) {
        if (cur_k >= 0) flush(cur_k);
        cur_k = pk;
      }
      np = pnp;
#pragma unroll
      for (int u = 0; u < NA4; ++u) {
        const int e = tid + u * WG_THREADS, p = e / (BA / 4), c = (e - p * (BA / 4)) * 4;
        if (e < WG_PC * (BA / 4)) *(float4*)&As[p][c] = ra[u];
      }
#pragma unroll
      for (int u = 0; u < NG4; ++u) {
        const int e = tid + u * WG_THREADS, p = e / (BG / 4), c = (e - p * (BG / 4)) * 4;
        if (e < WG_PC * (BG / 4)) *(float4*)&Gs[p][c] = rg[u];
      }
      __syncthreads();
      if (ch + 1 < ch_end) prefetch(ch + 1);
    } else {
      while (ch >= s_cum[k + 1]) ++k;
      if (k != cur_k) {
        if (cur_k >= 0) flush(cur_k);
        cur_k = k;
      }
      int p0 = s_koff[k] + (ch - s_cum[k]) * WG_PC;
      np = min(WG_PC, s_koff[k + 1] - p0);
      for (int e = tid; e < WG_PC * BA; e += WG_THREADS) {
        int p = e / BA, c = e - p * BA;
        float v = 0.f;
        if (p < np && a0 + c < c_a) v = __ldg(A + (int64_t)__ldg(src + p0 + p) * c_a + a0 + c);
        As[p][c] = v;
      }
      for (int e = tid; e < WG_PC * BG; e += WG_THREADS) {
        int p = e / BG, c = e - p * BG;
        float v = 0.f;
        if (p < np && g0 + c < c_g) v = __ldg(G + (int64_t)__ldg(dst + p0 + p) * c_g + g0 + c);
        Gs[p][c] = v;
      }
      __syncthreads();
    }
    for (int p = grp; p < np; p += NG) {
      float a[TA], g[TG];
      if (TA == 4 && TG == 4) {  // rows are 16-byte aligned (row stride BA + 4 floats): one LDS.128 per operand
        const float4 av = *(const float4*)&As[p][ta * 4], gv = *(const float4*)&Gs[p][tg * 4];
        a[0] = av.x; a[1] = av.y; a[2] = av.z; a[3] = av.w;
        g[0] = gv.x; g[1] = gv.y; g[2] = gv.z; g[3] = gv.w;
      } else {
#pragma unroll
        for (int i = 0; i < TA; ++i) a[i] = As[p][ta * TA + i];
#pragma unroll
        for (int j = 0; j < TG; ++j) g[j] = Gs[p][tg * TG + j];
      }
#pragma unroll
      for (int i = 0; i < TA; ++i)
#pragma unroll
        for (int j = 0; j < TG; ++j) acc[i][j] = fmaf(a[i], g[j], acc[i][j]);
    }
    __syncthreads();
  }
  if (cur_k >= 0) flush(cur_k);
}

template <int BN, bool PAIRS>
static int launch_igemm(const float* A, const float* W, const int32_t* idx, const int32_t* dst,
                        const int32_t* k_offsets, float* out, int64_t n_out, int64_t grid_x, int K,
                        int c_in, int c_out, int wT, int krev, cudaStream_t stream) {
  size_t smem = (size_t)(IG_BK * (IG_BM + 4) + IG_BK * BN) * 4 +
                (size_t)(PAIRS ? 2 * IG_BM : IG_BM * K) * 4;
  auto kern = conv_igemm_kernel<BN, PAIRS>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid((unsigned)grid_x, (unsigned)cdiv(c_out, BN));
  kern<<<grid, IG_THREADS, smem, stream>>>(A, W, idx, dst, k_offsets, out, n_out, K, c_in, c_out, wT, krev);
  return check_launch("conv_igemm");
}

int conv_table_simt(const float* A, const float* W, const int32_t* nbr, float* out, int64_t n_out,
                    int K, int c_in, int c_out, int wT, int krev, cudaStream_t stream) {
  if (n_out == 0) return B2S_OK;
  int64_t gx = cdiv(n_out, IG_BM);
  if (c_out <= 16) return launch_igemm<16, false>(A, W, nbr, nullptr, nullptr, out, n_out, gx, K, c_in, c_out, wT, krev, stream);
  if (c_out <= 32) return launch_igemm<32, false>(A, W, nbr, nullptr, nullptr, out, n_out, gx, K, c_in, c_out, wT, krev, stream);
  return launch_igemm<64, false>(A, W, nbr, nullptr, nullptr, out, n_out, gx, K, c_in, c_out, wT, krev, stream);
}

int conv_pairs_simt(const float* A, const float* W, const int32_t* src, const int32_t* dst,
                    const int32_t* k_offsets, float* out, int K, int c_in, int c_out, int wT,
                    int64_t max_pairs, cudaStream_t stream) {
  if (max_pairs == 0) return B2S_OK;
  int64_t gx = cdiv(max_pairs, IG_BM) + K;
  if (c_out <= 16) return launch_igemm<16, true>(A, W, src, dst, k_offsets, out, 0, gx, K, c_in, c_out, wT, 0, stream);
  if (c_out <= 32) return launch_igemm<32, true>(A, W, src, dst, k_offsets, out, 0, gx, K, c_in, c_out, wT, 0, stream);
  return launch_igemm<64, true>(A, W, src, dst, k_offsets, out, 0, gx, K, c_in, c_out, wT, 0, stream);
}

int conv_wgrad_simt(const float* A, const float* G, const int32_t* src, const int32_t* dst,
                    const int32_t* k_offsets, float* gW, int K, int c_a, int c_g, int64_t max_pairs,
                    cudaStream_t stream) {
  cudaMemsetAsync(gW, 0, (size_t)K * c_a * c_g * 4, stream);
  if (max_pairs == 0) return check_launch("conv_wgrad(empty)");
  int64_t chunks = cdiv(max_pairs, 64) + K;
  if (c_a <= 16 && c_g <= 16) {
    int gx = (int)std::min<int64_t>(chunks, 4 * B2S_SM_COUNT);
    conv_wgrad_kernel<16, 16, 4, 4, 256><<<dim3(gx, 1, 1), WG_THREADS, 0, stream>>>(A, G, src, dst, k_offsets, gW, K, c_a, c_g);
  } else if (c_a <= 32 && c_g <= 32) {
    int gx = (int)std::min<int64_t>(chunks, 4 * B2S_SM_COUNT);
    conv_wgrad_kernel<32, 32, 4, 4, 128><<<dim3(gx, 1, 1), WG_THREADS, 0, stream>>>(A, G, src, dst, k_offsets, gW, K, c_a, c_g);
  } else {
    int ty = (int)cdiv(c_a, 64), tz = (int)cdiv(c_g, 64);
    int gx = (int)std::min<int64_t>(chunks, std::max(1, 4 * B2S_SM_COUNT / (ty * tz)));
    conv_wgrad_kernel<64, 64, 4, 4, 64><<<dim3(gx, ty, tz), WG_THREADS, 0, stream>>>(A, G, src, dst, k_offsets, gW, K, c_a, c_g);
  }
  return check_launch("conv_wgrad");
}

}  // namespace b2s
