// T3 for the NARROW layers (c_in, c_out in {16, 32}: U-Net levels 0-1, ScoreNet): warp-stream table convolution
//
//   out[o, :] = sum_k in[nbr[o,k], :] @ W[k]   (+ shortcut[o, :])          (appendix A.7; common.py:12,37,40,48)
//
// conv_tcp.cu (tcgen05, 128-row tiles, TMEM accumulators) is latency-bound on these layers: 12 warps per SM, a
// ~100-instruction dependent chain per (tile, offset) slab in the producer warps, an MMA of N = 16 that the tensor
// pipe finishes in 8 cycles -- profiles/r02_conv_tcp_ncu.txt: issue slots 32 %, 0.43 eligible warps per scheduler,
// tensor pipe 10 %.  The weight gradient showed what these shapes want instead (wgrad_det.cu: 117 -> 60 us): no shared
// staging of activations, no block barriers, many independent warps.  Here:
//   * a WARP owns 16 output rows at a time (M = 16 of mma.sync.m16n8k8.tf32) and walks only the offsets that are
//     active within those 16 rows (finer than the 128-row tile masks: fewer empty slabs);
//   * the A fragment is loaded STRAIGHT from the gathered rows: lane (gid, tq) reads c_in/4 consecutive channels of
//     rows gid and gid + 8 (one or two 16-byte loads per row; the K index of the MMA is a permutation of the
//     channels that both operands share);
//   * weights live in shared memory in B-fragment order for that permutation (16-byte conflict-free reads), written
//     once per CTA from the fp32 kernel -- transposed / offset-reversed on the fly for the data gradient -- and split
//     into TF32 hi / lo in registers (3xTF32: lo*hi + hi*lo + hi*hi, fp32 accumulate);
//   * the 16 x K index block of the next group is prefetched into registers, the rows of the next active offset are
//     in flight while the current one is multiplied;
//   * output columns are permuted so that a lane holds c_out/4 consecutive columns: 16-byte stores, residual add
//     folded in.
// Same sums in the same k order as conv_tcp / the oracle (zeros for missing neighbours); fp32-class accuracy.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace b2s {

constexpr int CW_MAXK = 32;

__device__ __forceinline__ void cw_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct CwArgs {
  const float* A;
  const float* W;             // fp32 kernel [K][c_in][c_out], or [K][c_out][c_in] when w_transposed
  const int32_t* nbr;         // [n_out, K] (row order or mask-sorted)
  const int32_t* out_rows;    // optional: table row t is stored at out[out_rows[t]]
  const float* add_src;       // optional residual, same rows as out
  float* out;
  int64_t n_out;
  int K, w_transposed, k_reversed;
};

// THREADS: 256 (two CTAs per SM) or, when the weights leave room for one CTA only (32 -> 32: 111 KB), 512
template <int CIN, int COUT, int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == 256 ? 2 : 1) conv_ws_kernel(const CwArgs a) {
  constexpr int CW_THREADS = THREADS, CW_WARPS = THREADS / 32;
  constexpr int S = CIN / 8, NTL = COUT / 8, XV = CIN / 4;  // k-steps, n-tiles, channels per lane and row
  constexpr int D = CIN == 16 ? 3 : 2;                       // ring slots of gathered rows
  constexpr int WK = S * NTL * 2 * 32;                     // floats of one offset's B fragments
  extern __shared__ __align__(16) float smem[];
  float* sW = smem;                                        // [K][S][NTL*2/4][32 lanes][4]
  int32_t* sIdx = (int32_t*)(smem + (size_t)a.K * WK) + (threadIdx.x >> 5) * (16 * CW_MAXK);
  const int K = a.K;
  const int lane = threadIdx.x & 31, gid = lane >> 2, tq = lane & 3;

  // ---- weights -> B-fragment order -----------------------------------------------------------------------------------
  // element (k, s, j, r, lane): Wk[channel][column], channel = XV*tq + 2s + r (K index tq + 4r of k-step s),
  // column = (gid >> 1) * 2*NTL + 2j + (gid & 1) (column gid of n-tile j); stored as 16-byte groups q = (2j + r) / 4.
  // The source is read in memory order (coalesced, four independent loads per thread in flight).
  {
    constexpr int Q = NTL * 2 / 4;
    const int total = K * CIN * COUT;
#pragma unroll 4
    for (int i = threadIdx.x; i < total; i += CW_THREADS) {
      const float w = __ldg(a.W + i);
      const int kk = i / (CIN * COUT), rem = i - kk * (CIN * COUT);
      const int ch = a.w_transposed ? rem % CIN : rem / COUT;
      const int col = a.w_transposed ? rem / CIN : rem % COUT;
      const int k = a.k_reversed ? K - 1 - kk : kk;
      const int lt = ch / XV, cr = ch % XV, s2 = cr >> 1, r = cr & 1;
      const int lg = (col / (2 * NTL)) * 2 + (col & 1), j = (col % (2 * NTL)) >> 1;
      const int jr = 2 * j + r;
      sW[((((k * S + s2) * Q + (jr >> 2)) * 32) + lg * 4 + lt) * 4 + (jr & 3)] = w;
    }
  }
  __syncthreads();

  const int64_t n_groups = (a.n_out + 15) >> 4;
  const int64_t gw = (int64_t)blockIdx.x * CW_WARPS + (threadIdx.x >> 5), GW = (int64_t)gridDim.x * CW_WARPS;
  const int blk = 16 * K;                       // ints in one group's index block
  constexpr int NI = (16 * CW_MAXK + 31) / 32;  // registers that hold a prefetched block
  int pre[NI];
  auto fetch_idx = [&](int64_t g) {
    const int64_t base = g * blk, lim = a.n_out * K;
#pragma unroll
    for (int u = 0; u < NI; ++u) {
      const int e = lane + 32 * u;
      pre[u] = (g < n_groups && e < blk && base + e < lim) ? __ldg(a.nbr + base + e) : -1;
    }
  };
  fetch_idx(gw);
  const int kstep32 = 32 % K, k_lane0 = lane % K;  // offset of element lane + 32u: advances by 32 mod K per register
  for (int64_t g = gw; g < n_groups; g += GW) {
    // ---- this group's indices into the warp's shared block; active offsets of the 16 rows ---------------------------
    uint32_t mask = 0;
    __syncwarp();
    {
      int ku = k_lane0;
#pragma unroll
      for (int u = 0; u < NI; ++u) {
        const int e = lane + 32 * u;
        if (e < blk) {
          sIdx[e] = pre[u];
          if (pre[u] >= 0) mask |= 1u << ku;
        }
        ku += kstep32;
        if (ku >= K) ku -= K;
      }
    }
    mask = __reduce_or_sync(0xffffffffu, mask);
    __syncwarp();
    fetch_idx(g + GW);

    float acc[NTL][4];
#pragma unroll
    for (int j = 0; j < NTL; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;

    float4 xa[D][2][XV / 4];  // [ring slot][row gid / gid+8][16-byte piece]
    int kq[D];                // offset held by each ring slot (-1: none)
    auto next_k = [&]() {
      const int k = mask ? __ffs(mask) - 1 : -1;
      mask &= mask - 1;
      return k;
    };
    auto issue = [&](int k, int slot) {
      const int i0 = sIdx[gid * K + k], i1 = sIdx[(gid + 8) * K + k];
#pragma unroll
      for (int v = 0; v < XV / 4; ++v) {
        xa[slot][0][v] = i0 >= 0 ? __ldg((const float4*)(a.A + (int64_t)i0 * CIN + XV * tq) + v) : make_float4(0.f, 0.f, 0.f, 0.f);
        xa[slot][1][v] = i1 >= 0 ? __ldg((const float4*)(a.A + (int64_t)i1 * CIN + XV * tq) + v) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto compute = [&](int k, int slot) {
      const float* x0 = (const float*)xa[slot][0];
      const float* x1 = (const float*)xa[slot][1];
      const float4* wk = (const float4*)(sW + (size_t)k * WK) + lane;
#pragma unroll
      for (int s = 0; s < S; ++s) {
        uint32_t ah[4], al[4];
        const float av[4] = {x0[2 * s], x1[2 * s], x0[2 * s + 1], x1[2 * s + 1]};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          ah[e] = __float_as_uint(av[e]) & 0xFFFFE000u;
          al[e] = __float_as_uint(av[e] - __uint_as_float(ah[e]));
        }
        float bv[NTL * 2];
#pragma unroll
        for (int q = 0; q < NTL * 2 / 4; ++q) {
          const float4 t = wk[(s * (NTL * 2 / 4) + q) * 32];
          bv[4 * q] = t.x, bv[4 * q + 1] = t.y, bv[4 * q + 2] = t.z, bv[4 * q + 3] = t.w;
        }
#pragma unroll
        for (int j = 0; j < NTL; ++j) {
          const uint32_t bh0 = __float_as_uint(bv[2 * j]) & 0xFFFFE000u, bh1 = __float_as_uint(bv[2 * j + 1]) & 0xFFFFE000u;
          const uint32_t bl0 = __float_as_uint(bv[2 * j] - __uint_as_float(bh0));
          const uint32_t bl1 = __float_as_uint(bv[2 * j + 1] - __uint_as_float(bh1));
          cw_mma(acc[j], al, bh0, bh1);  // small terms first
          cw_mma(acc[j], ah, bl0, bl1);
          cw_mma(acc[j], ah, bh0, bh1);
        }
      }
    };
    // ---- active offsets in ascending k through a ring of D register slots: D - 1 offsets' rows in flight --------------
#pragma unroll
    for (int d = 0; d < D; ++d) {
      kq[d] = next_k();
      if (kq[d] >= 0) issue(kq[d], d);
    }
    bool more = kq[0] >= 0;
    while (more) {
#pragma unroll
      for (int d = 0; d < D; ++d) {
        if (kq[d] >= 0) {
          compute(kq[d], d);
          kq[d] = next_k();
          if (kq[d] >= 0) issue(kq[d], d);
        }
      }
      more = kq[0] >= 0;
    }
    // ---- epilogue: lane tq holds columns tq*2*NTL .. +2*NTL-1 of rows gid and gid + 8 ------------------------------
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t t = g * 16 + gid + 8 * h;
      if (t < a.n_out) {
        const int64_t o = a.out_rows != nullptr ? (int64_t)__ldg(a.out_rows + t) : t;
        float4* op = (float4*)(a.out + o * COUT + tq * 2 * NTL);
#pragma unroll
        for (int v = 0; v < NTL / 2; ++v) {
          float4 r = make_float4(acc[2 * v][2 * h], acc[2 * v][2 * h + 1], acc[2 * v + 1][2 * h], acc[2 * v + 1][2 * h + 1]);
          if (a.add_src != nullptr) {
            const float4 sv = __ldg((const float4*)(a.add_src + o * COUT + tq * 2 * NTL) + v);
            r.x += sv.x, r.y += sv.y, r.z += sv.z, r.w += sv.w;
          }
          op[v] = r;
        }
      }
    }
  }
}

bool conv_ws_supported(int K, int c_in, int c_out) {
  return K >= 1 && K <= CW_MAXK && (c_in == 16 || c_in == 32) && (c_out == 16 || c_out == 32);
}

template <int CIN, int COUT>
static int cw_launch(const CwArgs& a, cudaStream_t stream) {
  constexpr size_t w_max = (size_t)CW_MAXK * CIN * COUT * 4;
  constexpr int THREADS = (2 * (w_max + 8 * 16 * CW_MAXK * 4) <= 200 * 1024) ? 256 : 512;
  constexpr int WARPS = THREADS / 32;
  const size_t smem = (size_t)a.K * CIN * COUT * 4 + (size_t)WARPS * 16 * CW_MAXK * 4;
  auto kern = conv_ws_kernel<CIN, COUT, THREADS>;
  static bool configured[B2S_MAX_DEVICES] = {};
  const int dev = current_device();
  if (!configured[dev]) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(w_max + WARPS * 16 * CW_MAXK * 4));
    configured[dev] = true;
  }
  const int64_t n_groups = (a.n_out + 15) / 16;
  const int per_sm = THREADS == 256 ? 2 : 1;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(cdiv(n_groups, WARPS), (int64_t)per_sm * sm_count()));
  kern<<<grid, THREADS, smem, stream>>>(a);
  return check_launch("conv_ws");
}

int conv_ws(const float* A, const float* W, const int32_t* nbr, const int32_t* out_rows, const float* add_src, float* out,
            int64_t n_out, int K, int c_in, int c_out, int wT, int krev, cudaStream_t stream) {
  if (n_out == 0) return B2S_OK;
  const CwArgs a{A, W, nbr, out_rows, add_src, out, n_out, K, wT, krev};
  if (c_in == 16 && c_out == 16) return cw_launch<16, 16>(a, stream);
  if (c_in == 32 && c_out == 16) return cw_launch<32, 16>(a, stream);
  if (c_in == 16 && c_out == 32) return cw_launch<16, 32>(a, stream);
  if (c_in == 32 && c_out == 32) return cw_launch<32, 32>(a, stream);
  set_error("conv_ws: unsupported channel counts");
  return B2S_E_INVALID;
}

}  // namespace b2s
