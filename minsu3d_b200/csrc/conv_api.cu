// C-ABI dispatch for the sparse convolution products (b2s.h: T3 / T4).
#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"

namespace b2s {
int* zeroed_counters(cudaStream_t stream) {
  static std::mutex mu;
  static std::map<std::pair<int, cudaStream_t>, int*> bufs;
  const std::pair<int, cudaStream_t> key(current_device(), stream);
  std::lock_guard<std::mutex> lock(mu);
  auto it = bufs.find(key);
  if (it != bufs.end()) return it->second;
  int* p = nullptr;
  if (cudaMalloc(&p, B2S_COUNTERS * sizeof(int)) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  cudaMemsetAsync(p, 0, B2S_COUNTERS * sizeof(int), stream);
  bufs[key] = p;
  return p;
}

int conv_table_simt(const float*, const float*, const int32_t*, float*, int64_t, int, int, int, int, int, cudaStream_t);
int conv_pairs_simt(const float*, const float*, const int32_t*, const int32_t*, const int32_t*, float*, int, int, int, int, int64_t, cudaStream_t);
int conv_wgrad_simt(const float*, const float*, const int32_t*, const int32_t*, const int32_t*, float*, int, int, int, int64_t, cudaStream_t);
// weight gradient on the tensor cores (wgrad_mma.cu: mma.sync m16n8k8 3xTF32)
bool conv_wgrad_mma_supported(int c_a, int c_g);
int conv_wgrad_mma(const float*, const float*, const int32_t*, const int32_t*, const int32_t*, float*, int, int, int, int64_t, cudaStream_t);
// deterministic warp-stream kernel (wgrad_det.cu: mma.sync 3xTF32 straight from the gathered rows, partials + ordered sum)
bool conv_wgrad_det_supported(int K, int c_a, int c_g);
size_t conv_wgrad_det_ws_bytes(int K, int c_a, int c_g);
int conv_wgrad_det(const float*, const float*, const int32_t*, const int32_t*, const int32_t*, float*, int, int, int, int64_t,
                   void*, size_t, cudaStream_t);
// tcgen05 paths (conv_tc.cu: per-tile kernel, PAIRS mode + fallback; conv_tcp.cu: persistent table kernel)
bool conv_tc_supported(int K, int c_in, int c_out);
size_t conv_tc_ws_bytes(int K, int c_in, int c_out);
int conv_tc_pack_both(const float* W, float* Bp, int K, int c_in, int c_out, cudaStream_t stream);
int conv_tc_pack_multi(const int64_t* desc, int n_desc, int64_t total, cudaStream_t stream);
const float* conv_tc_weights(const float* W, const float* Wp, int K, int c_in, int c_out, int wT, void* ws,
                             cudaStream_t stream);
int conv_tc(const float* A, const float* W, const float* Wp, const int32_t* idx, const int32_t* dst, const int32_t* k_offsets,
            const uint32_t* tile_mask, const int32_t* out_rows,
            float* out, int64_t n_out, int64_t max_pairs, int K, int c_in, int c_out, int wT, int krev, int nsplit,
            bool pairs, void* ws, size_t ws_bytes, cudaStream_t stream);
int conv_tcp(const float* A, const float* Bp, const int32_t* idx, const uint32_t* tile_mask, const int32_t* out_rows,
             const float* add_src, float* out, int64_t n_out, int K, int c_in, int c_out, int krev, int nsplit,
             void* split_ws, cudaStream_t stream);

// narrow layers (c_in, c_out in {16, 32}): warp-stream mma.sync kernel (conv_ws.cu), 3xTF32
bool conv_ws_supported(int K, int c_in, int c_out);
int conv_ws(const float* A, const float* W, const int32_t* nbr, const int32_t* out_rows, const float* add_src, float* out,
            int64_t n_out, int K, int c_in, int c_out, int wT, int krev, cudaStream_t stream);

__global__ void __launch_bounds__(256)
    add_rows_kernel(float4* __restrict__ y, const float4* __restrict__ s, int64_t total4) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  float4 a = y[i];
  const float4 b = __ldg(s + i);
  a.x += b.x;
  a.y += b.y;
  a.z += b.z;
  a.w += b.w;
  y[i] = a;
}

// y += s  (residual add for the paths that cannot fold it into their epilogue)
int add_rows(float* y, const float* s, int64_t n, int c, cudaStream_t stream) {
  const int64_t total4 = n * (c / 4);
  if (total4 == 0) return B2S_OK;
  add_rows_kernel<<<(unsigned)cdiv(total4, 256), 256, 0, stream>>>((float4*)y, (const float4*)s, total4);
  return check_launch("add_rows");
}

static bool tc_persistent() {
  static const int on = [] {
    const char* e = getenv("B2S_TC_PERSIST");
    return (e == nullptr || atoi(e) != 0) ? 1 : 0;
  }();
  return on != 0;
}

// table-mode product on the tensor cores: persistent kernel (default) or the round-1 per-tile kernel (B2S_TC_PERSIST=0)
int conv_table_tc(const float* A, const float* W, const float* Wp, const int32_t* nbr, const uint32_t* tile_mask,
                  const int32_t* out_rows, const float* add_src, float* out, int64_t n_out, int K, int c_in, int c_out,
                  int wT, int krev, int nsplit, void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (ws_bytes < conv_tc_ws_bytes(K, c_in, c_out)) {
    set_error("conv_table: workspace too small");
    return B2S_E_WORKSPACE;
  }
  if (n_out == 0) return B2S_OK;
  // warp-stream kernel for the narrow layers: explicit (algo 4) or, for 3xTF32 calls, switched on with B2S_CONV_WS=1
  // (measured: faster than the tcgen05 kernel only for 32 -> 16 products, profiles/r02_conv_ws.txt; default off)
  static const int ws_env = [] {
    const char* e = getenv("B2S_CONV_WS");
    return (e != nullptr && atoi(e) != 0) ? 1 : 0;
  }();
  if ((nsplit == 4 || (nsplit == 3 && ws_env)) && nbr != nullptr && W != nullptr && conv_ws_supported(K, c_in, c_out))
    return conv_ws(A, W, nbr, out_rows, add_src, out, n_out, K, c_in, c_out, wT, krev, stream);
  if (nsplit == 4) nsplit = 3;
  if (!tc_persistent()) {
    int rc = conv_tc(A, W, Wp, nbr, nullptr, nullptr, tile_mask, out_rows, out, n_out, 0, K, c_in, c_out, wT, krev, nsplit,
                     false, ws, ws_bytes, stream);
    if (rc || add_src == nullptr) return rc;
    return add_rows(out, add_src, n_out, c_out, stream);
  }
  const float* Bp = conv_tc_weights(W, Wp, K, c_in, c_out, wT, ws, stream);
  void* split_ws = (char*)ws + align_up((size_t)K * c_in * c_out * 4 * 2) + 256;
  return conv_tcp(A, Bp, nbr, tile_mask, out_rows, add_src, out, n_out, K, c_in, c_out, krev, nsplit, split_ws, stream);
}
}  // namespace b2s

using namespace b2s;

// algo: 0 auto (tcgen05 3xTF32 when the shape allows, else fp32 FMA), 1 fp32 FMA, 2 tcgen05 3xTF32, 3 tcgen05 TF32
static int pick(int algo, int K, int c_in, int c_out, const char* what, int* nsplit) {
  bool ok = conv_tc_supported(K, c_in, c_out);
  if (algo == 0) algo = ok ? 2 : 1;
  if (algo == 4 && !conv_ws_supported(K, c_in, c_out)) ok = false;
  if ((algo == 2 || algo == 3 || algo == 4) && !ok) {
    char buf[160];
    snprintf(buf, sizeof(buf), "%s: shape (K=%d, c_in=%d, c_out=%d) not supported by the tcgen05 path", what, K, c_in, c_out);
    set_error(buf);
    return -1;
  }
  *nsplit = algo == 2 ? 3 : (algo == 4 ? 4 : 1);
  return algo;
}

extern "C" {

size_t b2s_conv_ws_bytes(int32_t K, int32_t c_in, int32_t c_out) {
  return conv_tc_supported(K, c_in, c_out) ? conv_tc_ws_bytes(K, c_in, c_out) : 256;
}

int64_t b2s_conv_packed_floats(int32_t K, int32_t c_in, int32_t c_out) {
  return conv_tc_supported(K, c_in, c_out) && conv_tc_supported(K, c_out, c_in) ? (int64_t)4 * K * c_in * c_out : 0;
}

int b2s_conv_pack(const float* W, float* Wp, int32_t K, int32_t c_in, int32_t c_out, b2s_stream_t stream) {
  if (!W || !Wp || b2s_conv_packed_floats(K, c_in, c_out) == 0) {
    set_error("conv_pack: shape not supported by the tcgen05 path");
    return B2S_E_INVALID;
  }
  return conv_tc_pack_both(W, Wp, K, c_in, c_out, stream);
}

int b2s_conv_pack_multi(const int64_t* desc, int32_t n_desc, int64_t total, b2s_stream_t stream) {
  if (!desc || n_desc < 0 || total < 0) {
    set_error("conv_pack_multi: invalid argument");
    return B2S_E_INVALID;
  }
  return conv_tc_pack_multi(desc, n_desc, total, stream);
}

int b2s_conv_table(const float* A, const float* W, const float* Wp, const int32_t* nbr, const uint32_t* tile_mask,
                   const float* add_src, float* out, int64_t n_out, int32_t K, int32_t c_in, int32_t c_out,
                   int32_t w_transposed, int32_t k_reversed, int32_t algo, void* ws, size_t ws_bytes,
                   b2s_stream_t stream) {
  if (n_out < 0 || K < 1 || K > 125 || c_in < 1 || c_out < 1 || (nbr == nullptr && K != 1)) {
    set_error("conv_table: invalid argument");
    return B2S_E_INVALID;
  }
  int nsplit = 0;
  algo = pick(algo, K, c_in, c_out, "conv_table", &nsplit);
  if (algo < 0) return B2S_E_INVALID;
  if (algo >= 2)
    return conv_table_tc(A, W, Wp, nbr, tile_mask, nullptr, add_src, out, n_out, K, c_in, c_out, w_transposed, k_reversed,
                         nsplit, ws, ws_bytes, stream);
  int rc = conv_table_simt(A, W, nbr, out, n_out, K, c_in, c_out, w_transposed, k_reversed, stream);
  if (rc || add_src == nullptr) return rc;
  if (c_out & 3) {
    set_error("conv_table: the residual add needs c_out % 4 == 0");
    return B2S_E_INVALID;
  }
  return add_rows(out, add_src, n_out, c_out, stream);
}

int b2s_conv_table_rows(const float* A, const float* W, const float* Wp, const int32_t* nbr_sorted,
                        const uint32_t* tile_mask, const int32_t* out_rows, const float* add_src, float* out,
                        int64_t n_out, int32_t K, int32_t c_in, int32_t c_out, int32_t w_transposed, int32_t k_reversed,
                        int32_t algo, void* ws, size_t ws_bytes, b2s_stream_t stream) {
  if (n_out < 0 || K < 1 || K > 32 || c_in < 1 || c_out < 1 || !nbr_sorted || !out_rows) {
    set_error("conv_table_rows: invalid argument");
    return B2S_E_INVALID;
  }
  if (algo == 1 || !conv_tc_supported(K, c_in, c_out)) {
    set_error("conv_table_rows: only the tcgen05 path takes a row permutation (c_in, c_out multiples of 16)");
    return B2S_E_INVALID;
  }
  const int nsplit = (algo == 3) ? 1 : (algo == 4 ? 4 : 3);
  return conv_table_tc(A, W, Wp, nbr_sorted, tile_mask, out_rows, add_src, out, n_out, K, c_in, c_out, w_transposed,
                       k_reversed, nsplit, ws, ws_bytes, stream);
}

int b2s_conv_pairs(const float* A, const float* W, const float* Wp, const int32_t* src, const int32_t* dst,
                   const int32_t* k_offsets, float* out, int32_t K, int32_t c_in, int32_t c_out,
                   int32_t w_transposed, int64_t max_pairs, int32_t algo, void* ws, size_t ws_bytes,
                   b2s_stream_t stream) {
  if (K < 1 || K > 125 || c_in < 1 || c_out < 1 || max_pairs < 0) {
    set_error("conv_pairs: invalid argument");
    return B2S_E_INVALID;
  }
  int nsplit = 0;
  if (algo == 4) algo = 2;  // the warp-stream kernel is table-mode only
  algo = pick(algo, K, c_in, c_out, "conv_pairs", &nsplit);
  if (algo < 0) return B2S_E_INVALID;
  if (algo >= 2)
    return conv_tc(A, W, Wp, src, dst, k_offsets, nullptr, nullptr, out, 0, max_pairs, K, c_in, c_out, w_transposed, 0,
                   nsplit, true, ws, ws_bytes, stream);
  return conv_pairs_simt(A, W, src, dst, k_offsets, out, K, c_in, c_out, w_transposed, max_pairs, stream);
}

int b2s_conv_wgrad(const float* A, const float* G, const int32_t* src, const int32_t* dst,
                   const int32_t* k_offsets, float* gW, int32_t K, int32_t c_a, int32_t c_g,
                   int64_t max_pairs, int32_t algo, b2s_stream_t stream) {
  if (K < 1 || K > 125 || c_a < 1 || c_g < 1 || max_pairs < 0) {
    set_error("conv_wgrad: invalid argument");
    return B2S_E_INVALID;
  }
  // algo 1 = fp32 FMA (strict); otherwise the tensor-core kernel when the channel counts allow (B2S_WGRAD_MMA=0 turns it off)
  static const int mma_on = [] {
    const char* e = getenv("B2S_WGRAD_MMA");
    return (e == nullptr || atoi(e) != 0) ? 1 : 0;
  }();
  if (algo != 1 && mma_on && conv_wgrad_mma_supported(c_a, c_g))
    return conv_wgrad_mma(A, G, src, dst, k_offsets, gW, K, c_a, c_g, max_pairs, stream);
  return conv_wgrad_simt(A, G, src, dst, k_offsets, gW, K, c_a, c_g, max_pairs, stream);
}

size_t b2s_conv_wgrad_ws_bytes(int32_t K, int32_t c_a, int32_t c_g) {
  return std::max<size_t>(conv_wgrad_det_ws_bytes(K, c_a, c_g), 256);
}

int b2s_conv_wgrad_ws(const float* A, const float* G, const int32_t* src, const int32_t* dst,
                      const int32_t* k_offsets, float* gW, int32_t K, int32_t c_a, int32_t c_g, int64_t max_pairs,
                      int32_t algo, void* ws, size_t ws_bytes, b2s_stream_t stream) {
  if (K < 1 || K > 125 || c_a < 1 || c_g < 1 || max_pairs < 0) {
    set_error("conv_wgrad: invalid argument");
    return B2S_E_INVALID;
  }
  // algo 1 = fp32 FMA (strict); otherwise the deterministic tensor-core kernel when the channel counts are multiples
  // of 16 (B2S_WGRAD_DET=0: the round-2 shared-memory kernels with their atomicAdd flush, kept for comparison)
  static const int det_on = [] {
    const char* e = getenv("B2S_WGRAD_DET");
    return (e == nullptr || atoi(e) != 0) ? 1 : 0;
  }();
  if (algo != 1 && det_on && conv_wgrad_det_supported(K, c_a, c_g))
    return conv_wgrad_det(A, G, src, dst, k_offsets, gW, K, c_a, c_g, max_pairs, ws, ws_bytes, stream);
  return b2s_conv_wgrad(A, G, src, dst, k_offsets, gW, K, c_a, c_g, max_pairs, algo, stream);
}

}  // extern "C"
