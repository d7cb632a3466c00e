// C-ABI dispatch for the sparse convolution products (b2s.h: T3 / T4).
#include "common.cuh"

namespace b2s {
int conv_table_simt(const float*, const float*, const int32_t*, float*, int64_t, int, int, int, int, int, cudaStream_t);
int conv_pairs_simt(const float*, const float*, const int32_t*, const int32_t*, const int32_t*, float*, int, int, int, int, int64_t, cudaStream_t);
int conv_wgrad_simt(const float*, const float*, const int32_t*, const int32_t*, const int32_t*, float*, int, int, int, int64_t, cudaStream_t);
// tcgen05 path (conv_tc.cu); returns B2S_E_INVALID when the shape is unsupported
int conv_table_tc(const float*, const float*, const int32_t*, float*, int64_t, int, int, int, int, int, int, cudaStream_t);
bool conv_tc_supported(int K, int c_in, int c_out);
}  // namespace b2s

using namespace b2s;

extern "C" {

int b2s_conv_table(const float* A, const float* W, const int32_t* nbr, float* out, int64_t n_out,
                   int32_t K, int32_t c_in, int32_t c_out, int32_t w_transposed, int32_t k_reversed,
                   int32_t algo, b2s_stream_t stream) {
  if (n_out < 0 || K < 1 || K > 125 || c_in < 1 || c_out < 1 || (nbr == nullptr && K != 1)) {
    set_error("conv_table: invalid argument");
    return B2S_E_INVALID;
  }
  if (algo == 2 || algo == 3) {
    if (!conv_tc_supported(K, c_in, c_out)) {
      set_error("conv_table: shape not supported by the tcgen05 path");
      return B2S_E_INVALID;
    }
    return conv_table_tc(A, W, nbr, out, n_out, K, c_in, c_out, w_transposed, k_reversed, algo == 2 ? 3 : 1, stream);
  }
  return conv_table_simt(A, W, nbr, out, n_out, K, c_in, c_out, w_transposed, k_reversed, stream);
}

int b2s_conv_pairs(const float* A, const float* W, const int32_t* src, const int32_t* dst,
                   const int32_t* k_offsets, float* out, int32_t K, int32_t c_in, int32_t c_out,
                   int32_t w_transposed, int64_t max_pairs, int32_t algo, b2s_stream_t stream) {
  if (K < 1 || K > 125 || c_in < 1 || c_out < 1 || max_pairs < 0) {
    set_error("conv_pairs: invalid argument");
    return B2S_E_INVALID;
  }
  (void)algo;
  return conv_pairs_simt(A, W, src, dst, k_offsets, out, K, c_in, c_out, w_transposed, max_pairs, stream);
}

int b2s_conv_wgrad(const float* A, const float* G, const int32_t* src, const int32_t* dst,
                   const int32_t* k_offsets, float* gW, int32_t K, int32_t c_a, int32_t c_g,
                   int64_t max_pairs, int32_t algo, b2s_stream_t stream) {
  if (K < 1 || K > 125 || c_a < 1 || c_g < 1 || max_pairs < 0) {
    set_error("conv_wgrad: invalid argument");
    return B2S_E_INVALID;
  }
  (void)algo;
  return conv_wgrad_simt(A, G, src, dst, k_offsets, gW, K, c_a, c_g, max_pairs, stream);
}

}  // extern "C"
