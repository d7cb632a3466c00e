// Shared device/host helpers for libb2s (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/b2s.h"

#define B2S_MAX_DEVICES 64

namespace b2s {

void set_error(const char* msg);

// current device ordinal (clamped to the per-device cache size)
inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < B2S_MAX_DEVICES) ? dev : 0;
}

// SM count of the current device, queried once per device (148 on B200); used by the grid heuristics
inline int sm_count() {
  static int cached[B2S_MAX_DEVICES] = {0};
  const int dev = current_device();
  int v = cached[dev];
  if (v == 0) {
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    cached[dev] = v;
  }
  return v;
}
#define B2S_SM_COUNT (::b2s::sm_count())

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    char buf[256];
    snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
    set_error(buf);
    return B2S_E_LAUNCH;
  }
  return B2S_OK;
}

inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// bump allocator over the caller's workspace
struct Workspace {
  char* base;
  size_t size;
  size_t off;
  Workspace(void* p, size_t n) : base((char*)p), size(n), off(0) {}
  template <typename T>
  T* take(size_t count) {
    size_t bytes = align_up(count * sizeof(T));
    if (off + bytes > size) return nullptr;
    T* r = (T*)(base + off);
    off += bytes;
    return r;
  }
};

// ---- packed coordinate key: b:19 | x:15 | y:15 | z:15, spatial bias 2^14 -------------------
#define B2S_EMPTY_KEY 0xFFFFFFFFFFFFFFFFull
#define B2S_COORD_BIAS 16384
#define B2S_COORD_LIM 16384

__host__ __device__ __forceinline__ bool coord_in_range(int b, int x, int y, int z) {
  return (unsigned)b < (1u << 19) && x >= -B2S_COORD_LIM && x < B2S_COORD_LIM &&
         y >= -B2S_COORD_LIM && y < B2S_COORD_LIM && z >= -B2S_COORD_LIM && z < B2S_COORD_LIM;
}
__host__ __device__ __forceinline__ uint64_t pack_coord(int b, int x, int y, int z) {
  return ((uint64_t)(uint32_t)b << 45) | ((uint64_t)(uint32_t)(x + B2S_COORD_BIAS) << 30) |
         ((uint64_t)(uint32_t)(y + B2S_COORD_BIAS) << 15) | (uint64_t)(uint32_t)(z + B2S_COORD_BIAS);
}
__host__ __device__ __forceinline__ uint64_t hash64(uint64_t k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return k;
}
__host__ __device__ __forceinline__ int floor_div(int a, int q) {
  int d = a / q;
  return (a % q != 0 && ((a < 0) != (q < 0))) ? d - 1 : d;
}

__device__ __forceinline__ int hash_lookup(const uint64_t* __restrict__ keys,
                                           const int32_t* __restrict__ vals, uint64_t mask,
                                           uint64_t key) {
  uint64_t slot = hash64(key) & mask;
  while (true) {
    uint64_t k = __ldg(keys + slot);
    if (k == key) return __ldg(vals + slot);
    if (k == B2S_EMPTY_KEY) return -1;
    slot = (slot + 1) & mask;
  }
}

// Library-owned buffer of B2S_COUNTERS int32 per (device, stream), zero when handed out.  Kernels that use it as arrival
// counters reset every word they touched before they finish ("self-cleaning"), so consecutive launches on the stream can
// share it without a memset in between.  nullptr when the allocation fails.
constexpr size_t B2S_COUNTERS = 16384;
int* zeroed_counters(cudaStream_t stream);

// device-wide exclusive scan of int32 (CUB underneath), temp storage from the workspace
size_t scan_ws_bytes(int64_t n);
int exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, void* ws, size_t ws_bytes,
                       cudaStream_t stream);

}  // namespace b2s
