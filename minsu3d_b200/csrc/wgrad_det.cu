// T3 weight gradient, deterministic warp-stream form:
//     gW[k][a][g] = sum over the pairs (i, o) of offset k of A[i][a] * G[o][g]      
//
// What bounded the two earlier kernels (conv_simt.cu: fp32 FMA, wgrad_mma.cu: mma.sync out of shared memory) was not
// the arithmetic but the staging pipeline around it: gather -> registers -> shared memory -> __syncthreads ->
// fragment loads -> __syncthreads per 64-256 pairs, 8 warps in lock step, and an atomicAdd flush per offset change
// (7 M atomics per launch on the small wide maps, run-to-run different sums).  Moving the multiply to tensor cores
// changed 116 us to 114 us on the level-0 map (profiles/r02_wgrad.txt).
//
// This kernel has no shared-memory staging, no block barrier in its main loop and no atomics:
//   * a CTA owns a contiguous run of CH-pair chunks of ONE kernel offset and one (a, g) channel tile; its 8 warps take
//     the chunks round-robin and run independently of each other;
//   * the MMA's K dimension is the pair index, and K indices as well as the M / N channel indices may be permuted
//     freely as long as both operands agree, so the m16n8k8 fragments are loaded STRAIGHT from the gathered rows:
//     lane (gid, tq) reads 2*MT consecutive A channels and NT consecutive G channels of pairs tq and tq + 4 of a
//     k-step as 8 / 16-byte vector loads (a warp-wide load touches 4 rows x 32..128 contiguous bytes), and channel
//     2*MT*gid + 2*i + h plays the role of row gid + 8*h of M tile i (column NT*gid + j that of column gid of N
//     tile j);
//   * pair indices are read coalesced (lane L holds pairs L, L+32, .. of the chunk) and handed out by shuffles; the
//     loads of the next batch of k-steps are in flight while the current one is multiplied (register double buffer),
//     across chunk boundaries as well;
//   * 3xTF32 in registers (hi*hi + lo*hi + hi*lo, small terms first, fp32 accumulate): fp32-class accuracy;
//   * the 8 warps' tiles are summed through shared memory in a fixed tree and every CTA leaves ONE partial tile in the
//     workspace, in fragment order (coalesced 128-byte rows); the CTA that finishes an offset last (arrival counter)
//     sums that offset's partials in ascending CTA order -> bit-reproducible weight gradients whichever CTA is last,
//     no float atomics, no second launch.
// Work split: q = chunks per CTA is derived on the device from k_offsets so that sum_k ceil(chunks_k / q) <= B (no
// host read of the pair counts).  Small maps use 32-pair chunks so that
// the few thousand pairs of a coarse U-Net level still spread over all warps of many CTAs.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace b2s {

constexpr int WD_THREADS = 256;
constexpr int WD_WARPS = WD_THREADS / 32;
constexpr int WD_MAXK = 125;

__device__ __forceinline__ void wd_mma(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// The CTA split shared by both kernels (called by all threads of a CTA with >= 128 threads; K <= 125):
// s_base[k] = first CTA of offset k, s_base[K] = CTAs in use, *s_q = chunks per CTA, s_koff = k_offsets.
__device__ __forceinline__ void wd_split(const int32_t* __restrict__ k_offsets, int K, int B, int chunk, int* s_koff,
                                         int* s_base, int* s_q) {
  const int t = threadIdx.x;
  if (t <= K) s_koff[t] = __ldg(k_offsets + t);
  __syncthreads();
  if (t < 32) {
    // lane t owns offsets 4t .. 4t+3
    int ck[4], sum = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int k = 4 * t + r;
      ck[r] = k < K ? (s_koff[k + 1] - s_koff[k] + chunk - 1) / chunk : 0;
      sum += ck[r];
    }
    int total = sum;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
    const int q = max(1, (total + (B - K) - 1) / max(1, B - K));
    int seg[4], mine = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      seg[r] = (ck[r] + q - 1) / q;
      mine += seg[r];
    }
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (t >= o) incl += v;
    }
    int b = incl - mine;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int k = 4 * t + r;
      if (k <= K) s_base[k] = b;
      b += seg[r];
    }
    if (t == 0) *s_q = q;
  }
  __syncthreads();
}

// N consecutive floats (N in {2, 4, 6, 8}) as the widest loads the alignment allows; zeros when row < 0 or past the
// channel count (cols_left = channels from the lane's first column to the end; vectors are all in or all out because
// channel counts are multiples of 16)
template <int N, bool RAG>
__device__ __forceinline__ void wd_load(float (&x)[N], const float* __restrict__ base, int row, int ld, int cols_left) {
  const float* p = base + (int64_t)row * ld;
  if constexpr (RAG) {
    // any channel count (3-channel input convolution, 20-class linear head): scalar loads, guarded one by one
#pragma unroll
    for (int v = 0; v < N; ++v) x[v] = (row >= 0 && v < cols_left) ? __ldg(p + v) : 0.f;
  } else {
    constexpr int V = (N % 4 == 0) ? 4 : 2;
#pragma unroll
    for (int v = 0; v < N / V; ++v) {
      if (row >= 0 && V * v < cols_left) {
        if constexpr (V == 4) {
          const float4 t = __ldg((const float4*)p + v);
          x[4 * v] = t.x, x[4 * v + 1] = t.y, x[4 * v + 2] = t.z, x[4 * v + 3] = t.w;
        } else {
          const float2 t = __ldg((const float2*)p + v);
          x[2 * v] = t.x, x[2 * v + 1] = t.y;
        }
      } else {
#pragma unroll
        for (int e = 0; e < V; ++e) x[V * v + e] = 0.f;
      }
    }
  }
}

// MT x NT m16n8 tiles per warp (channel tile 16*MT x 8*NT); KB k-steps per register buffer; U*32 pairs per chunk
template <int MT, int NT, int KB, int U, bool RAG>
__global__ void __launch_bounds__(WD_THREADS, (MT * NT <= 8) ? 2 : 1)
    wgrad_det_kernel(const float* __restrict__ A, const float* __restrict__ G, const int32_t* __restrict__ src,
                     const int32_t* __restrict__ dst, const int32_t* __restrict__ k_offsets, float* __restrict__ part,
                     int* __restrict__ done, float* __restrict__ gW, int K, int c_a, int c_g, int B) {
  constexpr int NA = 2 * MT, NG = NT, CH = 32 * U, NB = (CH / 8) / KB, TILE = MT * NT * 128;
  static_assert((CH / 8) % KB == 0 && NB % 2 == 0, "batches per chunk must be even (static double buffer)");
  extern __shared__ __align__(16) float s_red[];  // [4][TILE]: tree reduction over the 8 warps
  __shared__ int s_koff[WD_MAXK + 3], s_base[WD_MAXK + 3], s_q;
  wd_split(k_offsets, K, B, CH, s_koff, s_base, &s_q);
  const int cta = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gid = lane >> 2, tq = lane & 3;
  const int tile = blockIdx.y * gridDim.z + blockIdx.z;
  // element (register r, lane l) of a tile in fragment order -> channel pair (a, g); used by the final sums
  auto tile_elem = [&](int r, int l, int& a, int& g) {
    const int ij = r >> 2, e = r & 3, i = ij / NT, j = ij % NT;
    a = blockIdx.y * (16 * MT) + NA * (l >> 2) + 2 * i + (e >> 1);
    g = blockIdx.z * (8 * NT) + NG * (2 * (l & 3) + (e & 1)) + j;
  };
  if (cta == 0) {
    // offsets without a single pair have no CTA: their gradient tile is zero
    for (int kk = 0; kk < K; ++kk) {
      if (s_base[kk + 1] != s_base[kk]) continue;
      for (int x = threadIdx.x; x < TILE; x += WD_THREADS) {
        int a, g;
        tile_elem(x >> 5, x & 31, a, g);
        if (a < c_a && g < c_g) gW[((int64_t)kk * c_a + a) * c_g + g] = 0.f;
      }
    }
  }
  if (cta >= s_base[K]) return;
  int k = 0;
  while (s_base[k + 1] <= cta) ++k;
  const int q = s_q;
  const int pk0 = s_koff[k], pk1 = s_koff[k + 1];
  const int cta_ch0 = (cta - s_base[k]) * q;
  const int ch1 = min(cta_ch0 + q, (pk1 - pk0 + CH - 1) / CH);
  const int ch0 = cta_ch0 + warp;  // this warp: chunks ch0, ch0 + 8, ...

  const int a0 = blockIdx.y * (16 * MT) + NA * gid, g0 = blockIdx.z * (8 * NT) + NG * gid;
  const int a_left = c_a - a0, g_left = c_g - g0;  // <= 0: the lane's columns lie past the channel count (ragged tile)
  const float* Ab = A + a0;
  const float* Gb = G + g0;

  float acc[MT][NT][4];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;

  float xa[2][KB][2][NA], xg[2][KB][2][NG];
  int is_cur[U], id_cur[U], is_nxt[U], id_nxt[U];
  auto load_idx = [&](int ch, int (&is)[U], int (&id)[U]) {
    const int p0 = pk0 + ch * CH;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = p0 + lane + 32 * u;
      const bool ok = ch < ch1 && p < pk1;
      is[u] = ok ? __ldg(src + p) : -1;
      id[u] = ok ? __ldg(dst + p) : -1;
    }
  };
  // batch b of a chunk = k-steps b*KB .. b*KB+KB-1; k-step s holds pairs 8s + tq (fragment K index tq) and 8s + tq + 4
  auto issue = [&](const int (&is)[U], const int (&id)[U], int b, int buf) {
#pragma unroll
    for (int t = 0; t < KB; ++t) {
      const int s = b * KB + t;  // compile-time after unrolling
      const int u = s >> 2, l0 = 8 * (s & 3) + tq;
      const int ra0 = __shfl_sync(0xffffffffu, is[u], l0), ra1 = __shfl_sync(0xffffffffu, is[u], l0 + 4);
      const int rg0 = __shfl_sync(0xffffffffu, id[u], l0), rg1 = __shfl_sync(0xffffffffu, id[u], l0 + 4);
      wd_load<NA, RAG>(xa[buf][t][0], Ab, ra0, c_a, a_left);
      wd_load<NA, RAG>(xa[buf][t][1], Ab, ra1, c_a, a_left);
      wd_load<NG, RAG>(xg[buf][t][0], Gb, rg0, c_g, g_left);
      wd_load<NG, RAG>(xg[buf][t][1], Gb, rg1, c_g, g_left);
    }
  };
  // (a term-major MMA order -- MT*NT independent accumulators between the three terms of one -- was measured slower:
  // it keeps all hi/lo fragments of a k-step live; the warps of a sub-core interleave their chains instead)
  auto compute = [&](int buf) {
#pragma unroll
    for (int t = 0; t < KB; ++t) {
      uint32_t bh[NT][2], bl[NT][2];
#pragma unroll
      for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const float v = xg[buf][t][r][j];
          bh[j][r] = __float_as_uint(v) & 0xFFFFE000u;
          bl[j][r] = __float_as_uint(v - __uint_as_float(bh[j][r]));
        }
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        uint32_t ah[4], al[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          // fragment element e: row gid + 8*(e & 1), K index tq + 4*(e >> 1)
          const float v = xa[buf][t][e >> 1][2 * i + (e & 1)];
          ah[e] = __float_as_uint(v) & 0xFFFFE000u;
          al[e] = __float_as_uint(v - __uint_as_float(ah[e]));
        }
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          wd_mma(acc[i][j], al, bh[j]);  // small terms first
          wd_mma(acc[i][j], ah, bl[j]);
          wd_mma(acc[i][j], ah, bh[j]);
        }
      }
    }
  };

  if (ch0 < ch1) {
    load_idx(ch0, is_cur, id_cur);
    issue(is_cur, id_cur, 0, 0);
    for (int ch = ch0; ch < ch1; ch += WD_WARPS) {
      load_idx(ch + WD_WARPS, is_nxt, id_nxt);  // all -1 past the end
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        if (b + 1 < NB)
          issue(is_cur, id_cur, b + 1, (b + 1) & 1);
        else
          issue(is_nxt, id_nxt, 0, 0);
        compute(b & 1);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) is_cur[u] = is_nxt[u], id_cur[u] = id_nxt[u];
    }
  }
  // fixed-order tree over the warps: (w, w+4), (w, w+2), (0, 1); tiles in fragment order [register][lane]
#pragma unroll
  for (int half = 4; half >= 1; half >>= 1) {
    if (warp >= half && warp < 2 * half) {
      float* o = s_red + (warp - half) * TILE + lane;
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e) o[((i * NT + j) * 4 + e) * 32] = acc[i][j][e];
    }
    __syncthreads();
    if (warp < half) {
      const float* o = s_red + warp * TILE + lane;
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[i][j][e] += o[((i * NT + j) * 4 + e) * 32];
    }
    __syncthreads();
  }
  float* tile_part = part + (int64_t)tile * B * TILE;
  if (warp == 0) {
    float* o = tile_part + (int64_t)cta * TILE + lane;
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) o[((i * NT + j) * 4 + e) * 32] = acc[i][j][e];
    __threadfence();
    __syncwarp();
    if (lane == 0) s_q = atomicAdd(done + tile * K + k, 1);
  }
  __syncthreads();
  // the CTA that finishes an offset last sums its partial tiles in ascending CTA order (fixed order -> the same bits
  // on every run, whichever CTA happens to be last) and stores the gradient tile
  const int b0 = s_base[k], nb = s_base[k + 1] - b0;
  if (s_q != nb - 1) return;
  __threadfence();
  for (int x = threadIdx.x; x < TILE; x += WD_THREADS) {
    const float* p = tile_part + (int64_t)b0 * TILE + x;
    float sum = 0.f;
    int b = 0;
    for (; b + 4 <= nb; b += 4) {
      const float v0 = __ldcg(p + (int64_t)b * TILE), v1 = __ldcg(p + (int64_t)(b + 1) * TILE),
                  v2 = __ldcg(p + (int64_t)(b + 2) * TILE), v3 = __ldcg(p + (int64_t)(b + 3) * TILE);
      sum += v0;
      sum += v1;
      sum += v2;
      sum += v3;
    }
    for (; b < nb; ++b) sum += __ldcg(p + (int64_t)b * TILE);
    int a, g;
    tile_elem(x >> 5, x & 31, a, g);
    if (a < c_a && g < c_g) gW[((int64_t)k * c_a + a) * c_g + g] = sum;
  }
  if (threadIdx.x == 0) done[tile * K + k] = 0;  // leave the counters zero for the next launch
}

struct WdShape {
  int mt, nt, ty, tz, B, u;
  bool rag;
};

static int wd_pick(int units, int maxu) {  // units of 16 channels -> tiles of at most maxu units, as even as possible
  const int nt = (units + maxu - 1) / maxu;
  return (units + nt - 1) / nt;
}

static int wd_env(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

static WdShape wd_shape(int K, int c_a, int c_g, int64_t max_pairs) {
  // Register tile.  Big maps are bound by the gathers: 64 x 64 tiles (every row fetched once per pair).  Below ~15k
  // rows the tensor pipe is the bound (mma.sync TF32: 8 cycles per m16n8k8 and sub-core, 21 cycles latency, measured
  // with tools/experiments/mma_rate.cu) and 254-register tiles leave two warps per sub-core to hide it: 32 x 32
  // tiles, 16 warps per SM, the re-fetched rows come out of L2 (profiles/r02_wgrad_det.txt).
  static const int env_maxu = wd_env("B2S_WGRAD_MAXT", 0);
  const int maxu = env_maxu > 0 ? std::min(4, env_maxu) : (max_pairs >= 400000 ? 4 : 2);
  WdShape s;
  s.rag = (c_a % 16) != 0 || (c_g % 16) != 0;
  const int ua = (c_a + 15) / 16, ug = (c_g + 15) / 16;
  if (s.rag) {  // tiles up to 32 x 32 only (four instantiations)
    s.mt = std::min(2, ua);
    s.nt = 2 * std::min(2, ug);
  } else {
    s.mt = c_a == 48 ? 3 : wd_pick(ua, maxu);
    s.nt = c_g == 48 ? 6 : 2 * wd_pick(ug, maxu);
  }
  s.ty = (ua + s.mt - 1) / s.mt;
  s.tz = (2 * ug + s.nt - 1) / s.nt;
  const int prod = s.mt * s.nt;
  // CTAs over all channel tiles = one resident wave: 2 CTAs per SM up to 128 registers (tiles <= 32 x 32), else 1
  const int cap_total = sm_count() * (prod <= 8 ? 2 : 1);
  const int cap = std::max(1, cap_total / (s.ty * s.tz));
  // 128-pair chunks when every warp of every CTA gets at least one, else 32-pair chunks
  // (tiles up to 32 x 32 only: the unrolled chunk body of a bigger tile would not fit the instruction cache)
  s.u = (prod <= 8 && cdiv(max_pairs, 128) >= (int64_t)WD_WARPS * cap) ? 4 : 1;
  const int64_t chunks_ub = cdiv(max_pairs, 32 * s.u) + K;
  // the device-side split uses at most B CTAs (q = ceil(chunks / (B - K))): B = cap keeps the launch to one wave
  // (when the channel tiles leave fewer than 4 K CTAs per tile the K per-offset remainders no longer fit one wave)
  s.B = (int)std::min<int64_t>(cdiv(chunks_ub, WD_WARPS) + K, cap >= 4 * K ? cap : cap + K);
  return s;
}

// channel counts that are multiples of 16 take the vector-load kernels; anything else up to 1024 channels the
// element-guarded ("ragged") ones
bool conv_wgrad_det_supported(int K, int c_a, int c_g) {
  return K >= 1 && K <= WD_MAXK && c_a >= 1 && c_g >= 1 && c_a <= 1024 && c_g <= 1024;
}

// upper bound over max_pairs (the CTA count is capped by the SM count)
size_t conv_wgrad_det_ws_bytes(int K, int c_a, int c_g) {
  if (!conv_wgrad_det_supported(K, c_a, c_g)) return 0;
  size_t need = 0;
  for (int64_t mp : {(int64_t)1 << 40, (int64_t)399999}) {  // both tilings, each at its largest CTA count
    const WdShape s = wd_shape(K, c_a, c_g, mp);
    need = std::max(need, align_up((size_t)s.ty * s.tz * K * 4) + (size_t)s.ty * s.tz * s.B * s.mt * s.nt * 128 * 4);
  }
  return need + 256;
}

template <int MT, int NT, int KB, int U, bool RAG>
static int wd_launch(const WdShape& s, const float* A, const float* G, const int32_t* src, const int32_t* dst,
                     const int32_t* k_offsets, float* gW, float* part, int* done, int K, int c_a, int c_g,
                     cudaStream_t stream) {
  constexpr size_t smem = (size_t)4 * MT * NT * 128 * 4;
  auto kern = wgrad_det_kernel<MT, NT, KB, U, RAG>;
  static bool configured[B2S_MAX_DEVICES] = {};
  const int dev = current_device();
  if (smem + 2048 > 48 * 1024 && !configured[dev]) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured[dev] = true;
  }
  kern<<<dim3(s.B, s.ty, s.tz), WD_THREADS, smem, stream>>>(A, G, src, dst, k_offsets, part, done, gW, K, c_a, c_g, s.B);
  return check_launch("conv_wgrad_det");
}

template <int MT, int NT, int KB4, int KB1>
static int wd_dispatch(const WdShape& s, const float* A, const float* G, const int32_t* src, const int32_t* dst,
                       const int32_t* k_offsets, float* gW, float* part, int* done, int K, int c_a, int c_g,
                       cudaStream_t stream) {
  if constexpr (MT <= 2 && NT <= 4) {
    if (s.rag) {
      constexpr int KBR = (MT * NT <= 2) ? 2 : 1;  // scalar loads hold more address registers: shorter batches
      if (s.u == 4)
        return wd_launch<MT, NT, KBR, 4, true>(s, A, G, src, dst, k_offsets, gW, part, done, K, c_a, c_g, stream);
      return wd_launch<MT, NT, KBR, 1, true>(s, A, G, src, dst, k_offsets, gW, part, done, K, c_a, c_g, stream);
    }
  }
  if constexpr (MT * NT <= 8) {
    if (s.u == 4)
      return wd_launch<MT, NT, KB4, 4, false>(s, A, G, src, dst, k_offsets, gW, part, done, K, c_a, c_g, stream);
  }
  return wd_launch<MT, NT, KB1, 1, false>(s, A, G, src, dst, k_offsets, gW, part, done, K, c_a, c_g, stream);
}

int conv_wgrad_det(const float* A, const float* G, const int32_t* src, const int32_t* dst, const int32_t* k_offsets,
                   float* gW, int K, int c_a, int c_g, int64_t max_pairs, void* ws, size_t ws_bytes,
                   cudaStream_t stream) {
  if (max_pairs == 0) {
    cudaMemsetAsync(gW, 0, (size_t)K * c_a * c_g * 4, stream);
    return check_launch("conv_wgrad(empty)");
  }
  const WdShape s = wd_shape(K, c_a, c_g, max_pairs);
  const size_t cnt_bytes = align_up((size_t)s.ty * s.tz * K * 4);
  const size_t need = cnt_bytes + (size_t)s.ty * s.tz * s.B * s.mt * s.nt * 128 * 4;
  if (ws == nullptr || ws_bytes < need) {
    set_error("conv_wgrad: workspace too small (b2s_conv_wgrad_ws_bytes)");
    return B2S_E_WORKSPACE;
  }
  // per-(tile, offset) arrival counters.  Every launch leaves them zero (the last CTA of an offset resets its
  // counter), so they live in a small library-owned buffer per (device, stream) that is cleared once when it is
  // created: no memset per call (86 per PointGroup step).  Shapes with more counters than the buffer holds use the
  // head of the caller's workspace, which other products share, and clear it here.
  int* done = (size_t)s.ty * s.tz * K <= B2S_COUNTERS ? zeroed_counters(stream) : nullptr;
  if (done == nullptr) {
    done = (int*)ws;
    cudaMemsetAsync(done, 0, cnt_bytes, stream);
  }
  float* part = (float*)((char*)ws + cnt_bytes);
  // KB: k-steps per register buffer (two buffers), smaller for the big tiles
#define WD_CASE(M, N, B4, B1) \
  if (s.mt == M && s.nt == N) return wd_dispatch<M, N, B4, B1>(s, A, G, src, dst, k_offsets, gW, part, done, K, c_a, c_g, stream);
  WD_CASE(1, 2, 4, 2)
  WD_CASE(1, 4, 2, 2)
  WD_CASE(1, 6, 2, 2)
  WD_CASE(1, 8, 1, 1)
  WD_CASE(2, 2, 2, 2)
  WD_CASE(2, 4, 2, 2)
  WD_CASE(2, 6, 1, 1)
  WD_CASE(2, 8, 1, 1)
  WD_CASE(3, 2, 2, 2)
  WD_CASE(3, 4, 1, 1)
  WD_CASE(3, 6, 1, 1)
  WD_CASE(3, 8, 1, 1)
  WD_CASE(4, 2, 1, 1)
  WD_CASE(4, 4, 1, 1)
  WD_CASE(4, 6, 1, 1)
  WD_CASE(4, 8, 1, 1)
#undef WD_CASE
  set_error("conv_wgrad: no kernel for this channel tiling");
  return B2S_E_INVALID;
}

}  // namespace b2s
