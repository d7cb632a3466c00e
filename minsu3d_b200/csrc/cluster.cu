// C2 / C3 / C4: BFS clustering on the GPU, bit-exact with the reference's sequential CPU BFS.
//
// Reference: minsu3d/common_ops/src/bfs_cluster/bfs_cluster.cpp:28-187 (pg_/sg_ variants) and
// hierarchical_aggregation/hierarchical_aggregation.cpp:8-97, .cu:20-91.
//
// The reference visits seeds in ascending index and grows each cluster by a FIFO BFS over the
// (possibly asymmetric, because of the 1000-neighbour cap) ball-query lists.  Two facts make a
// parallel restatement exact:
//   (1) membership: v belongs to the cluster of m(v) = the lowest index that reaches v along
//       directed edges (proof in DESIGN.md).  m is the fixpoint of  m(w) = min(m(w), m(u)) over
//       edges u->w; pointer jumping m(v) = m(m(v)) is valid because reachability is transitive.
//   (2) order: BFS visit order is level-synchronous; inside a level nodes are ordered by
//       (queue position of the first frontier node listing them, position in that node's list).
// Both run as persistent cooperative kernels with a device-wide barrier: no host round trip
// per level (the reference copies the whole CSR to the host and runs single-threaded).
#include <algorithm>

#include "common.cuh"

namespace b2s {

// ---- device-wide barrier for cooperative (co-resident) grids --------------------------------
__device__ __forceinline__ void grid_sync(unsigned* bar, unsigned nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned gen = *((volatile unsigned*)(bar + 1));
    if (atomicAdd(bar, 1u) == nblocks - 1) {
      *((volatile unsigned*)bar) = 0;
      __threadfence();
      atomicAdd(bar + 1, 1u);
    } else {
      while (*((volatile unsigned*)(bar + 1)) == gen) {
      }
    }
    __threadfence();
  }
  __syncthreads();
}

constexpr int CL_THREADS = 256;

// ------------------------------------------------------------------------------------------
// (1) membership labels
// ------------------------------------------------------------------------------------------
// (1a) Union-find over the SYMMETRIC edges.  Nodes joined by edges present in both directions are
// mutually reachable, hence share m(v); contracting them first (ECL-CC style hooking: one pass over
// the edges, no device-wide barrier, independent of the graph diameter) leaves the directed fixpoint
// iteration below with nothing to do unless ball-query lists were truncated at 1000 neighbours.
// An edge u->w is symmetric iff u is in list(w); the distance predicate is symmetric, so that can
// only fail when list(w) hit the cap; a capped list keeps the lowest indices, so one comparison with
// its last element decides.
constexpr int BQ_LIST_CAP = 1000;

__device__ __forceinline__ int uf_find(int32_t* parent, int v) {
  int p = __ldcg(parent + v);
  while (p != v) {
    int gp = __ldcg(parent + p);
    if (gp != p) parent[v] = gp;  // path halving (values only decrease: benign race)
    v = p;
    p = gp;
  }
  return v;
}

// lastv[w] = largest listed index of a capped list (u is in list(w) iff u <= lastv[w] for u in radius), INT_MAX
// for complete lists: one L2-resident read decides the symmetry of an edge instead of two dependent loads into
// the big neighbour array.
__global__ void __launch_bounds__(256)
    cc_init_kernel(const int32_t* __restrict__ nbr_idx, const int32_t* __restrict__ start_len, int n,
                   int32_t* __restrict__ root, int32_t* __restrict__ m, int32_t* __restrict__ lastv,
                   uint8_t* __restrict__ asym, int* __restrict__ n_capped) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  m[v] = v;
  asym[v] = 0;
  const int s = __ldg(start_len + 2 * v), l = __ldg(start_len + 2 * v + 1);
  lastv[v] = (l >= BQ_LIST_CAP) ? __ldg(nbr_idx + s + l - 1) : 0x7fffffff;
  if (l >= BQ_LIST_CAP) atomicAdd(n_capped, 1);  // rare; 0 = every list is complete, every edge symmetric
  root[v] = v;
}

// ECL-CC style initialisation: every node starts hooked to its lowest-index symmetric neighbour (lists are ascending,
// so that is one of the first entries).  Parents are strictly smaller, so no cycle can form and no atomics are
// needed; in dense blobs most nodes then already hang under the blob's lowest index when the hooking pass runs.
__global__ void __launch_bounds__(256)
    cc_prehook_kernel(const int32_t* __restrict__ nbr_idx, const int32_t* __restrict__ start_len,
                      const int16_t* __restrict__ labels, const int32_t* __restrict__ lastv, int n,
                      int32_t* __restrict__ root) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  const int s = __ldg(start_len + 2 * v), l = min(__ldg(start_len + 2 * v + 1), 4);
  const int lab = labels ? labels[v] : 0;
  for (int e = 0; e < l; ++e) {
    const int w = __ldg(nbr_idx + s + e);
    if (w >= v) break;  // ascending: nothing smaller follows
    if (labels && labels[w] != lab) continue;
    if (v > __ldg(lastv + w)) continue;  // v is not in list(w): directed-only edge
    root[v] = w;
    break;
  }
}

// one warp per node: union with the symmetric neighbours of higher index; nodes that own a directed-only edge
// (u -> w with u not in list(w)) are flagged for the fixpoint below
__global__ void __launch_bounds__(256)
    cc_hook_kernel(const int32_t* __restrict__ nbr_idx, const int32_t* __restrict__ start_len,
                   const int16_t* __restrict__ labels, const int32_t* __restrict__ lastv, int n, int32_t* root,
                   uint8_t* __restrict__ asym, int* __restrict__ n_asym, const int* __restrict__ n_capped) {
  const int u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (u >= n) return;
  const bool any_capped = *n_capped > 0;  // written by cc_init_kernel; uniform
  const int s = __ldg(start_len + 2 * u), l = __ldg(start_len + 2 * u + 1);
  const int lab = labels ? labels[u] : 0;
  bool directed = false;
  int ru_hint = uf_find(root, u);  // root of u as last seen by this lane
  for (int e = lane; e < l; e += 32) {
    // (unrolling this loop for memory-level parallelism was measured slower: prefetched roots go stale and the
    // shortcut below misses more often)
    const int w = __ldg(nbr_idx + s + e);
    // without capped lists every edge is symmetric and the lower endpoint handles it: the edges towards lower indices
    // (half of them) need no further read
    if (!any_capped && w <= u) continue;
    if (labels && labels[w] != lab) continue;
    if (any_capped && u > __ldg(lastv + w)) {  // list(w) is capped and does not reach up to u
      directed = true;
      continue;
    }
    if (w <= u) continue;  // each undirected pair is handled from its lower endpoint
    // dense blobs: after the first few unions nearly every neighbour already hangs under u's root
    if (__ldcg(root + w) == ru_hint) continue;
    int ru = uf_find(root, u), rw = uf_find(root, w);
    ru_hint = min(ru, rw);
    while (ru != rw) {
      const int hi = max(ru, rw), lo = min(ru, rw);
      const int old = atomicCAS(root + hi, hi, lo);
      if (old == hi) break;
      ru = uf_find(root, old);
      rw = lo;
    }
  }
  if (__any_sync(0xffffffffu, directed) && lane == 0) {
    asym[u] = 1;
    atomicAdd(n_asym, 1);
  }
}

__global__ void __launch_bounds__(256) cc_compress_kernel(int n, int32_t* root) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < n) {
    int r = v, p = __ldcg(root + r);
    while (p != r) {
      r = p;
      p = __ldcg(root + r);
    }
    root[v] = r;
  }
}

// (1b) directed fixpoint on the contracted graph.  m[c] (c a union-find root) = lowest index that reaches
// component c; only directed-only edges can lower it, and only flagged nodes own such edges, so a sweep reads
// the lists of the flagged nodes instead of every edge.  Pointer jumping m[c] = m[root[m[c]]] is valid because
// reachability is transitive.  Ends with comp[v] = m[root[v]].  Without truncated lists nothing is flagged and
// the kernel only writes comp.
__global__ void __launch_bounds__(CL_THREADS)
    cc_label_kernel(const int32_t* __restrict__ nbr_idx, const int32_t* __restrict__ start_len,
                    const int16_t* __restrict__ labels, const int32_t* __restrict__ lastv,
                    const int32_t* __restrict__ root, const uint8_t* __restrict__ asym,
                    const int* __restrict__ n_asym, int n, int32_t* m, int32_t* __restrict__ comp, unsigned* bar,
                    int* flags) {
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const int nthreads = gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  const int gwarp = gtid >> 5, nwarps = nthreads >> 5;
  if (*n_asym > 0) {  // uniform across the grid (written by an earlier kernel)
    int it = 0;
    while (true) {
      int* flag = flags + (it & 1);
      for (int u = gwarp; u < n; u += nwarps) {
        if (!asym[u]) continue;
        const int cu = __ldcg(m + root[u]);
        const int s = __ldg(start_len + 2 * u), l = __ldg(start_len + 2 * u + 1);
        const int lab = labels ? labels[u] : 0;
        bool ch = false;
        for (int e = lane; e < l; e += 32) {
          const int w = __ldg(nbr_idx + s + e);
          if (labels && labels[w] != lab) continue;
          if (u <= __ldg(lastv + w)) continue;  // symmetric edge: same component
          const int rw = root[w];
          if (cu < __ldcg(m + rw)) {
            const int old = atomicMin(m + rw, cu);
            ch |= (cu < old);
          }
        }
        if (__any_sync(0xffffffffu, ch) && lane == 0) *flag = 1;
      }
      grid_sync(bar, gridDim.x);
      if (gtid == 0) flags[(it + 1) & 1] = 0;
      for (int v = gtid; v < n; v += nthreads) {
        if (root[v] != v) continue;
        const int c = __ldcg(m + v);
        const int cc = __ldcg(m + root[c]);
        if (cc < c) {
          atomicMin(m + v, cc);
          *flag = 1;
        }
      }
      grid_sync(bar, gridDim.x);
      const int changed = *((volatile int*)flag);
      if (!changed) break;
      ++it;
    }
  }
  for (int v = gtid; v < n; v += nthreads) comp[v] = __ldcg(m + root[v]);
}

// ------------------------------------------------------------------------------------------
// selection
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cl_hist_kernel(const int32_t* __restrict__ comp, int n,
                                                      int32_t* __restrict__ size) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < n) atomicAdd(size + comp[v], 1);
}

__global__ void __launch_bounds__(256)
    cl_keep_kernel(const int32_t* __restrict__ comp, const int16_t* __restrict__ labels, int n,
                   const int32_t* __restrict__ size, int mode, int thr_i, float thr_f,
                   const float* __restrict__ point_num_avg, int group, int32_t* __restrict__ flag,
                   int32_t* __restrict__ ksize) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  bool keep = false;
  int sz = size[v];
  if (comp[v] == v) {
    if (mode == 0) keep = sz >= thr_i;
    else if (mode == 1) keep = (float)sz >= thr_f;  // bfs_cluster.cpp:116-121 (int compared as float)
    else if (mode == 3) keep = (float)sz >= point_num_avg[labels[v]];  // mode 1 with one threshold per class
    else {
      // hierarchical_aggregation.cpp:58-74: double literal x float mean, rounded into float
      float mean = point_num_avg[labels[v]];
      float low = (float)(0.05 * (double)mean);
      float high = (float)(0.3 * (double)mean);
      float fs = (float)sz;
      bool frag = fs < high;
      if (group == 1) keep = frag && fs >= low;  // kept fragments
      else if (group == 2) keep = !frag;         // primary
      else keep = frag;                          // every fragment
    }
  }
  flag[v] = keep ? 1 : 0;
  ksize[v] = keep ? sz : 0;
}

__global__ void __launch_bounds__(256)
    cl_emit_kernel(int n, const int32_t* __restrict__ flag, const int32_t* __restrict__ ksize,
                   const int32_t* __restrict__ cidx, const int32_t* __restrict__ coff,
                   int32_t* __restrict__ cluster_offsets, int32_t* __restrict__ seeds,
                   int32_t* __restrict__ d_count) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  if (flag[v]) {
    seeds[cidx[v]] = v;
    cluster_offsets[cidx[v]] = coff[v];
  }
  if (v == n - 1) {
    int nc = cidx[v] + flag[v], tot = coff[v] + ksize[v];
    cluster_offsets[nc] = tot;
    d_count[0] = nc;
    d_count[1] = tot;
  }
}

// ------------------------------------------------------------------------------------------
// (2) BFS visit order: one CTA per kept cluster, level-synchronous with block-level barriers only
// ------------------------------------------------------------------------------------------
// The BFS order of a cluster is the concatenation of its levels, so the output slice
// cluster_idxs[off .. off+size) doubles as the frontier queue: level L is the contiguous range
// [lvl_begin, lvl_end) of positions, its children are appended behind it.  Inside a level a node's key is
// (queue position of the first frontier node listing it [atomicMin], its slot in that node's list), which is
// exactly the order the reference's FIFO queue produces (bfs_cluster.cpp:38-52).  Surface clusters are
// ~100 levels deep: with one CTA per cluster a level costs four __syncthreads instead of four device-wide
// barriers, and all clusters advance concurrently.
struct OrderArgs {
  const int32_t* nbr_idx;
  const int32_t* start_len;
  const int16_t* labels;
  const int32_t* comp;
  const int32_t* cluster_offsets;
  const int32_t* seeds;
  int32_t* cluster_idxs;
  int32_t *cid, *seedcid, *pos, *parent, *cnt;
  int n, n_cluster;
};

__global__ void __launch_bounds__(256) bfs_prep_kernel(OrderArgs a, int phase) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (phase == 0) {
    if (v < a.n) {
      a.seedcid[v] = -1;
      a.pos[v] = -1;
      a.parent[v] = 0x7fffffff;
    }
  } else if (phase == 1) {
    if (v < a.n_cluster) a.seedcid[a.seeds[v]] = v;
  } else {
    if (v < a.n) a.cid[v] = a.seedcid[a.comp[v]];
  }
}

__global__ void __launch_bounds__(CL_THREADS) bfs_order_kernel(OrderArgs a) {
  __shared__ int s_scan[CL_THREADS / 32];
  __shared__ int s_total;
  const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
  constexpr int NW = CL_THREADS / 32;
  for (int c = blockIdx.x; c < a.n_cluster; c += gridDim.x) {
    const int off = a.cluster_offsets[c];
    const int size = a.cluster_offsets[c + 1] - off;
    const int seed = a.seeds[c];
    int32_t* out = a.cluster_idxs + 2 * (int64_t)off;  // rows (cluster id, point) of this cluster
    if (tid == 0) {
      out[0] = c;
      out[1] = seed;
      a.pos[seed] = off;
    }
    __syncthreads();
    int lvl_begin = 0, lvl_end = 1;
    while (lvl_begin < lvl_end && lvl_end < size) {
      // ---- A: every frontier node claims its unvisited neighbours (lowest queue position wins)
      for (int f = lvl_begin + wib; f < lvl_end; f += NW) {
        const int u = out[2 * f + 1];
        const int s = __ldg(a.start_len + 2 * u), l = __ldg(a.start_len + 2 * u + 1);
        const int lab = a.labels ? a.labels[u] : 0;
        for (int e = lane; e < l; e += 32) {
          const int w = __ldg(a.nbr_idx + s + e);
          if (a.cid[w] != c) continue;
          if (a.labels && a.labels[w] != lab) continue;
          if (__ldcg(a.pos + w) >= 0) continue;
          atomicMin(a.parent + w, f);
        }
      }
      __syncthreads();
      // ---- B: children per frontier node
      for (int f = lvl_begin + wib; f < lvl_end; f += NW) {
        const int u = out[2 * f + 1];
        const int s = __ldg(a.start_len + 2 * u), l = __ldg(a.start_len + 2 * u + 1);
        int cc = 0;
        for (int e0 = 0; e0 < l; e0 += 32) {
          const int e = e0 + lane;
          bool child = false;
          if (e < l) {
            const int w = __ldg(a.nbr_idx + s + e);
            child = (__ldcg(a.parent + w) == f) && (a.cid[w] == c) && (__ldcg(a.pos + w) < 0);
          }
          cc += __popc(__ballot_sync(0xffffffffu, child));
        }
        if (lane == 0) a.cnt[off + f] = cc;
      }
      __syncthreads();
      // ---- scan: exclusive prefix of the children counts over the level (in place)
      int running = 0;
      for (int f0 = lvl_begin; f0 < lvl_end; f0 += CL_THREADS) {
        const int f = f0 + tid;
        const int v = (f < lvl_end) ? a.cnt[off + f] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          int t = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += t;
        }
        if (lane == 31) s_scan[wib] = inc;
        __syncthreads();
        int woff = 0, ctot = 0;
        for (int i = 0; i < NW; ++i) {
          const int t = s_scan[i];
          if (i < wib) woff += t;
          ctot += t;
        }
        if (f < lvl_end) a.cnt[off + f] = running + woff + inc - v;
        running += ctot;
        __syncthreads();
      }
      if (tid == 0) s_total = running;
      __syncthreads();
      const int total = s_total;
      // ---- C: append the children behind the current level, in (parent position, list slot) order
      for (int f = lvl_begin + wib; f < lvl_end; f += NW) {
        const int u = out[2 * f + 1];
        const int s = __ldg(a.start_len + 2 * u), l = __ldg(a.start_len + 2 * u + 1);
        int j = lvl_end + a.cnt[off + f];
        if (a.cnt[off + f] == (f + 1 < lvl_end ? a.cnt[off + f + 1] : total)) continue;  // no children: skip the rescan
        for (int e0 = 0; e0 < l; e0 += 32) {
          const int e = e0 + lane;
          bool child = false;
          int w = 0;
          if (e < l) {
            w = __ldg(a.nbr_idx + s + e);
            child = (__ldcg(a.parent + w) == f) && (a.cid[w] == c) && (__ldcg(a.pos + w) < 0);
          }
          const unsigned m = __ballot_sync(0xffffffffu, child);
          if (child) {
            const int jj = j + __popc(m & ((1u << lane) - 1));
            out[2 * jj] = c;
            out[2 * jj + 1] = w;
          }
          j += __popc(m);
        }
      }
      __syncthreads();
      // positions are published only after the whole level is emitted (phase C tests pos < 0)
      for (int jj = lvl_end + tid; jj < lvl_end + total; jj += CL_THREADS) a.pos[out[2 * jj + 1]] = off + jj;
      __syncthreads();
      lvl_begin = lvl_end;
      lvl_end += total;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// (2b) BFS visit order, device-wide variant for DENSE graphs (shifted coordinates: ~1000 neighbours per
// node, few levels, clusters of thousands of nodes): all kept clusters advance level by level together in a
// persistent cooperative kernel (4 device-wide barriers per level), so one huge cluster is expanded by the
// whole GPU.  b2s_cluster_order uses this variant for everything but small inputs.
// ------------------------------------------------------------------------------------------
struct GridOrderArgs {
  const int32_t* nbr_idx;
  const int32_t* start_len;
  const int16_t* labels;
  const int32_t* comp;
  const int32_t* cluster_offsets;
  const int32_t* seeds;
  int32_t* cluster_idxs;
  int32_t *cid, *seedcid, *parent, *F0, *F1, *cnt, *base, *fstart, *filled, *blocksum;
  unsigned* bar;
  int n, n_cluster;
};

__device__ __forceinline__ int block_sum(int v, int* s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  int t = 0;
  for (int i = 0; i < CL_THREADS / 32; ++i) t += s_red[i];
  return t;
}

__global__ void __launch_bounds__(CL_THREADS) bfs_order_grid_kernel(GridOrderArgs a) {
  __shared__ int s_red[CL_THREADS / 32];
  __shared__ int s_scan[CL_THREADS / 32];
  const int G = gridDim.x;
  const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
  const int gtid = blockIdx.x * blockDim.x + tid, nthreads = G * blockDim.x;
  const int gwarp = gtid >> 5, nwarps = nthreads >> 5;
  const int n = a.n, nC = a.n_cluster;
  // a.cid[v]: cluster id while v is unvisited, -1 outside the kept clusters, -2 - cluster id once visited: one
  // read per edge decides "same cluster and unvisited" (members of a cluster share the seed's label, so the
  // label needs no separate check).  a.parent[v]: queue position of the first frontier node that lists v.
  int32_t* remaining = a.seedcid;        // [nC] nodes of the cluster not yet placed (reuses seedcid after the prologue)
  int32_t* fstart = a.fstart;            // [2][n + 2] per-cluster start of the frontier, double-buffered per level
  int32_t* filled = a.filled;            // [2][n + 2] nodes of the cluster already placed
  const int fstride = n + 2;

  for (int v = gtid; v < n; v += nthreads) {
    a.seedcid[v] = -1;
    a.parent[v] = 0x7fffffff;
  }
  grid_sync(a.bar, G);
  for (int c = gtid; c < nC; c += nthreads) a.seedcid[a.seeds[c]] = c;
  if (gtid == 0) fstart[nC] = nC;
  grid_sync(a.bar, G);
  for (int v = gtid; v < n; v += nthreads) a.cid[v] = a.seedcid[a.comp[v]];
  grid_sync(a.bar, G);
  for (int c = gtid; c < nC; c += nthreads) {
    const int sd = a.seeds[c];
    const int off = a.cluster_offsets[c];
    a.F0[c] = sd;
    remaining[c] = a.cluster_offsets[c + 1] - off - 1;  // seedcid is free from here on (its last read is above)
    a.cid[sd] = -2 - c;
    a.cluster_idxs[2 * off] = c;
    a.cluster_idxs[2 * off + 1] = sd;
    filled[c] = 1;
    fstart[c] = c;
  }
  grid_sync(a.bar, G);

  int32_t* F = a.F0;
  int32_t* Fn = a.F1;
  int fsize = nC;
  int lvl = 0;
  // Three device-wide barriers per level: A claim | B count (child side) | C scan + emit.
  while (fsize > 0) {
    const int per = (fsize + G - 1) / G;  // frontier nodes per block (contiguous ranges)
    const int32_t* fs_cur = fstart + (lvl & 1) * fstride;
    const int32_t* fl_cur = filled + (lvl & 1) * fstride;
    // ---- A: every frontier node claims its unvisited neighbours (lowest queue position wins)
    for (int bb = gtid; bb < G; bb += nthreads) a.blocksum[bb] = 0;
    for (int f = gwarp; f < fsize; f += nwarps) {
      const int u = F[f];
      const int cu = -2 - a.cid[u];
      const int s = __ldg(a.start_len + 2 * u), l = __ldg(a.start_len + 2 * u + 1);
      if (lane == 0) a.cnt[f] = 0;
      // every node of this cluster has been placed already: nothing left to claim.  On the shifted coordinates (lists
      // of ~400 neighbours, 2-3 BFS levels) the last frontier holds most of a cluster and all of its scans are skipped.
      // (remaining[] is updated in phase C, one device-wide barrier before this read; the per-level bookkeeping in
      // fstart / filled is not yet complete when phase A starts)
      if (__ldcg(remaining + cu) == 0) continue;
      // four independent (neighbour -> state) load chains per lane: the loop is latency-bound otherwise
      for (int e0 = lane; e0 < l; e0 += 128) {
        int w[4], st[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) w[k] = (e0 + 32 * k < l) ? __ldg(a.nbr_idx + s + e0 + 32 * k) : -1;
#pragma unroll
        for (int k = 0; k < 4; ++k) st[k] = (w[k] >= 0) ? __ldcg(a.cid + w[k]) : -1;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (w[k] >= 0 && st[k] == cu) atomicMin(a.parent + w[k], f);
      }
    }
    grid_sync(a.bar, G);
    // ---- B: children per frontier node, counted from the child side (one pass over the nodes, not the edges)
    for (int v = gtid; v < n; v += nthreads) {
      if (__ldcg(a.cid + v) < 0) continue;
      const int f = __ldcg(a.parent + v);
      if (f == 0x7fffffff) continue;
      atomicAdd(a.cnt + f, 1);
      atomicAdd(a.blocksum + f / per, 1);
    }
    grid_sync(a.bar, G);
    // ---- C: exclusive scan of cnt over the block's frontier range (block prefix from blocksum), then emit the
    // next frontier + final positions.  The only base value a block needs from outside its range is the one at
    // the frontier start of the cluster that spans its first node: recomputed locally from blocksum and cnt.
    const int fb = min(fsize, (int)blockIdx.x * per), fe = min(fsize, fb + per);
    int pre = 0, tot = 0;
    for (int b = tid; b < G; b += CL_THREADS) {
      int v = __ldcg(a.blocksum + b);
      tot += v;
      if (b < (int)blockIdx.x) pre += v;
    }
    pre = block_sum(pre, s_red);
    const int total = block_sum(tot, s_red);
    int c0 = -1, c0_start = 0, c0_base = 0;
    if (fb < fe) {
      c0 = -2 - a.cid[F[fb]];
      c0_start = fs_cur[c0];
      if (c0_start < fb) {  // the cluster starts in an earlier block
        const int b0 = c0_start / per;
        int part = 0;
        for (int b = tid; b < b0; b += CL_THREADS) part += __ldcg(a.blocksum + b);
        for (int f = b0 * per + tid; f < c0_start; f += CL_THREADS) part += __ldcg(a.cnt + f);
        c0_base = block_sum(part, s_red);
      }
    }
    int running = pre;
    for (int f0 = fb; f0 < fe; f0 += CL_THREADS) {
      int f = f0 + tid;
      int v = (f < fe) ? __ldcg(a.cnt + f) : 0;
      int inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      __syncthreads();
      if (lane == 31) s_scan[wib] = inc;
      __syncthreads();
      int woff = 0, ctot = 0;
      for (int i = 0; i < CL_THREADS / 32; ++i) {
        int t = s_scan[i];
        if (i < wib) woff += t;
        ctot += t;
      }
      if (f < fe) a.base[f] = running + woff + inc - v;
      running += ctot;
    }
    if (blockIdx.x == 0 && tid == 0) a.base[fsize] = total;
    __syncthreads();  // base[fb, fe) written by this block is read below
    for (int f = fb + wib; f < fe; f += CL_THREADS / 32) {
      const int nchild = __ldcg(a.cnt + f);
      if (nchild == 0) continue;  // no child claimed through this node: its list need not be read again
      const int u = F[f];
      const int c = -2 - a.cid[u];
      if (lane == 0) atomicSub(remaining + c, nchild);
      const int s = __ldg(a.start_len + 2 * u), l = __ldg(a.start_len + 2 * u + 1);
      const int fsc = fs_cur[c];
      const int cstart = (fsc >= fb) ? a.base[fsc] : c0_base;  // fsc < fb only for c == c0
      const int pbase = a.cluster_offsets[c] + fl_cur[c] - cstart;
      int j = a.base[f];
      for (int e0 = 0; e0 < l; e0 += 128) {
        int w[4], pr[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int e = e0 + 32 * k + lane;
          w[k] = (e < l) ? __ldg(a.nbr_idx + s + e) : -1;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) pr[k] = (w[k] >= 0) ? __ldcg(a.parent + w[k]) : -1;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const bool child = (pr[k] == f) && (__ldcg(a.cid + w[k]) == c);  // pr == f implies w >= 0
          const unsigned m = __ballot_sync(0xffffffffu, child);
          if (child) {
            const int jj = j + __popc(m & ((1u << lane) - 1));
            Fn[jj] = w[k];
            const int p = pbase + jj;
            a.cluster_idxs[2 * p] = c;
            a.cluster_idxs[2 * p + 1] = w[k];
            a.cid[w[k]] = -2 - c;  // visited; only this warp sees parent[w] == f, so the race with readers is benign
          }
          j += __popc(m);
        }
      }
    }
    grid_sync(a.bar, G);
    // per-cluster bookkeeping for the next level (reads this level's buffers and base, writes the other halves;
    // the next level's phase C reads them two barriers later)
    {
      int32_t* fs_nxt = fstart + ((lvl + 1) & 1) * fstride;
      int32_t* fl_nxt = filled + ((lvl + 1) & 1) * fstride;
      for (int c = gtid; c <= nC; c += nthreads) {
        const int fsn = __ldcg(a.base + fs_cur[c]);
        fs_nxt[c] = fsn;
        if (c < nC) fl_nxt[c] = fl_cur[c] + __ldcg(a.base + fs_cur[c + 1]) - fsn;
      }
    }
    int32_t* t = F;
    F = Fn;
    Fn = t;
    fsize = total;
    ++lvl;
  }
}

// ------------------------------------------------------------------------------------------
// HAIS centres: sequential fp32 sums in BFS order (hierarchical_aggregation.cpp:13-37,85-89)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
    cluster_centers_kernel(const int32_t* __restrict__ cluster_idxs,
                           const int32_t* __restrict__ cluster_offsets, int n_cluster,
                           const float* __restrict__ coords, const int16_t* __restrict__ labels,
                           const uint8_t* __restrict__ batch_idxs, float* __restrict__ centers) {
  int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (c >= n_cluster) return;
  int b = cluster_offsets[c], e = cluster_offsets[c + 1];
  float sx = 0.f, sy = 0.f, sz = 0.f;
  for (int i0 = b; i0 < e; i0 += 32) {
    int i = i0 + lane;
    float x = 0.f, y = 0.f, z = 0.f;
    if (i < e) {
      int p = cluster_idxs[2 * i + 1];
      x = coords[3 * p];
      y = coords[3 * p + 1];
      z = coords[3 * p + 2];
    }
    int m = min(32, e - i0);
    for (int j = 0; j < m; ++j) {  // strictly sequential adds, same order as the CPU BFS
      sx = __fadd_rn(sx, __shfl_sync(0xffffffffu, x, j));
      sy = __fadd_rn(sy, __shfl_sync(0xffffffffu, y, j));
      sz = __fadd_rn(sz, __shfl_sync(0xffffffffu, z, j));
    }
  }
  if (lane == 0) {
    float cntf = (float)(e - b);
    int seed = cluster_idxs[2 * b + 1];
    centers[5 * c + 0] = __fdiv_rn(sx, cntf);
    centers[5 * c + 1] = __fdiv_rn(sy, cntf);
    centers[5 * c + 2] = __fdiv_rn(sz, cntf);
    centers[5 * c + 3] = (float)labels[seed];
    centers[5 * c + 4] = (float)batch_idxs[seed];
  }
}

// ------------------------------------------------------------------------------------------
// HAIS set aggregation (hierarchical_aggregation.cu:20-64): nearest same-class same-scene
// primary in double precision, strict '<' so the lowest primary index wins ties.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    ha_assign_kernel(const float* __restrict__ fc, int n_frag, const float* __restrict__ pc,
                     const int32_t* __restrict__ prim_offsets, int n_prim,
                     const float* __restrict__ radius_avg, int32_t* __restrict__ assign) {
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n_frag) return;
  float fx = fc[5 * f], fy = fc[5 * f + 1], fz = fc[5 * f + 2], fcls = fc[5 * f + 3], fb = fc[5 * f + 4];
  float nearest = 10000.f;
  int best = -1;
  for (int i = 0; i < n_prim; ++i) {
    if (fabsf(pc[5 * i + 3] - fcls) > 0.1f) continue;
    if (fabsf(pc[5 * i + 4] - fb) > 0.1f) continue;
    double dx = (double)(pc[5 * i] - fx), dy = (double)(pc[5 * i + 1] - fy), dz = (double)(pc[5 * i + 2] - fz);
    float d2 = (float)(dx * dx + dy * dy + dz * dz);
    if (d2 < nearest) {
      nearest = d2;
      best = i;
    }
  }
  int res = -1;
  if (best >= 0) {
    int np = prim_offsets[best + 1] - prim_offsets[best];
    float r_size = (float)(0.01 * (double)sqrtf((float)np));
    float r_cls = radius_avg[(int)fcls];
    float r_set = fmaxf(r_size, r_cls);
    if (nearest < __fmul_rn(r_set, r_set)) res = best;
  }
  assign[f] = res;
}

constexpr int HA_MAX_FRAG = 1024;   // MAX_PER_PRIMARY_ABSORB_FRAGMENT_NUM
constexpr int HA_MAX_POINT = 8192;  // MAX_PER_PRIMARY_ABSORB_POINT_NUM

// warp per primary; pass 0 counts, pass 1 writes.  Absorbed fragments are taken in ascending
// fragment index (the reference's order is atomicAdd arrival order, .cu:60).
__global__ void __launch_bounds__(128)
    ha_concat_kernel(const int32_t* __restrict__ frag_idxs, const int32_t* __restrict__ frag_offsets,
                     int n_frag, const int32_t* __restrict__ prim_idxs,
                     const int32_t* __restrict__ prim_offsets, int n_prim,
                     const int32_t* __restrict__ assign, int32_t* __restrict__ totals,
                     const int32_t* __restrict__ out_base, int32_t* __restrict__ out_idxs,
                     int32_t* __restrict__ out_offsets, int pass) {
  int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (p >= n_prim) return;
  int pb = prim_offsets[p], pe = prim_offsets[p + 1];
  int o = 0;
  if (pass == 1) {
    o = out_base[p];
    for (int i = pb + lane; i < pe; i += 32) {
      out_idxs[2 * (o + i - pb)] = p;
      out_idxs[2 * (o + i - pb) + 1] = prim_idxs[2 * i + 1];
    }
    o += pe - pb;
  }
  int nf = 0, npts = 0;
  for (int f0 = 0; f0 < n_frag && nf < HA_MAX_FRAG; f0 += 32) {
    int f = f0 + lane;
    bool mine = (f < n_frag) && (assign[f] == p);
    unsigned m = __ballot_sync(0xffffffffu, mine);
    while (m && nf < HA_MAX_FRAG) {
      int j = __ffs(m) - 1;
      m &= m - 1;
      int ff = f0 + j;
      int fb = frag_offsets[ff], fe = frag_offsets[ff + 1];
      int take = min(fe - fb, HA_MAX_POINT - npts);
      if (pass == 1) {
        for (int i = lane; i < take; i += 32) {
          out_idxs[2 * (o + npts + i)] = p;
          out_idxs[2 * (o + npts + i) + 1] = frag_idxs[2 * (fb + i) + 1];
        }
      }
      npts += take;
      ++nf;
    }
  }
  if (lane == 0) {
    if (pass == 0) totals[p] = (pe - pb) + npts;
    else out_offsets[p + 1] = o + npts;
  }
  if (pass == 1 && p == 0 && lane == 0) out_offsets[0] = 0;
}

static int coop_grid(const void* kernel, int threads, int64_t work_items) {
  // device / occupancy queries cost ~40 us each: cache them per (device, kernel)
  static const void* cached_kernel[B2S_MAX_DEVICES][4] = {};
  static int cached_blocks[B2S_MAX_DEVICES][4] = {};
  const int dev = current_device();
  int max_blocks = 0;
  for (int i = 0; i < 4; ++i)
    if (cached_kernel[dev][i] == kernel) max_blocks = cached_blocks[dev][i];
  if (max_blocks == 0) {
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0);
    if (per_sm < 1) per_sm = 1;
    max_blocks = sm_count() * std::min(per_sm, 2);
    for (int i = 0; i < 4; ++i)
      if (cached_kernel[dev][i] == nullptr) {
        cached_kernel[dev][i] = kernel;
        cached_blocks[dev][i] = max_blocks;
        break;
      }
  }
  int g = max_blocks;
  int64_t want = cdiv(work_items, threads);
  if (want < 1) want = 1;
  if (want < g) g = (int)want;
  return g;
}

}  // namespace b2s

using namespace b2s;

extern "C" {

size_t b2s_cluster_ws_bytes(int64_t n) {
  if (n < 1) n = 1;
  // order kernel: 11 int arrays of n(+2) + blocksum + barrier; select: 5 arrays + scan
  return 13 * align_up((size_t)(n + 2) * 4) + scan_ws_bytes(n) + 64 * 1024;
}

int b2s_cluster_label(const int32_t* nbr_idx, const int32_t* start_len, const int16_t* labels,
                      int64_t n, int32_t* comp, void* ws, size_t ws_bytes, b2s_stream_t stream) {
  if (n < 0 || n > 0x3fffffff) {
    set_error("cluster_label: invalid n");
    return B2S_E_INVALID;
  }
  if (n == 0) return B2S_OK;
  Workspace w(ws, ws_bytes);
  unsigned* bar = w.take<unsigned>(64);
  int32_t* root = w.take<int32_t>(n);
  int32_t* m = w.take<int32_t>(n);
  int32_t* lastv = w.take<int32_t>(n);
  uint8_t* asym = w.take<uint8_t>(n);
  if (!bar || !asym) {
    set_error("cluster_label: workspace too small");
    return B2S_E_WORKSPACE;
  }
  int* flags = (int*)(bar + 2);
  int* n_asym = (int*)(bar + 8);
  cudaMemsetAsync(bar, 0, 64 * 4, stream);
  int* n_capped = (int*)(bar + 9);
  cc_init_kernel<<<(unsigned)cdiv(n, 256), 256, 0, stream>>>(nbr_idx, start_len, (int)n, root, m, lastv, asym, n_capped);
  cc_prehook_kernel<<<(unsigned)cdiv(n, 256), 256, 0, stream>>>(nbr_idx, start_len, labels, lastv, (int)n, root);
  // The pre-hook leaves chains (node -> lowest neighbour -> its lowest neighbour ...) as long as a cluster is wide in
  // hops (~100 on raw coordinates); flattened here, every find of the hooking pass is one or two reads.  ncu of the
  // raw-coordinate hook before this: 520 us for 1.17 M edges, 300 warps stalled on the scoreboard per issue
  // (profiles/r02_cluster_kernels_ncu.txt).
  cc_compress_kernel<<<(unsigned)cdiv(n, 256), 256, 0, stream>>>((int)n, root);
  cc_hook_kernel<<<(unsigned)cdiv(n * 32, 256), 256, 0, stream>>>(nbr_idx, start_len, labels, lastv, (int)n, root, asym,
                                                                  n_asym, n_capped);
  cc_compress_kernel<<<(unsigned)cdiv(n, 256), 256, 0, stream>>>((int)n, root);
  int grid = coop_grid((const void*)cc_label_kernel, CL_THREADS, n * 32);
  int ni = (int)n;
  void* args[] = {(void*)&nbr_idx, (void*)&start_len, (void*)&labels, (void*)&lastv, (void*)&root, (void*)&asym,
                  (void*)&n_asym, (void*)&ni,     (void*)&m,      (void*)&comp,  (void*)&bar,  (void*)&flags};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)cc_label_kernel, dim3(grid), dim3(CL_THREADS), args, 0, stream);
  if (e != cudaSuccess) {
    set_error(cudaGetErrorString(e));
    return B2S_E_LAUNCH;
  }
  return check_launch("cluster_label");
}

int b2s_cluster_select(const int32_t* comp, const int16_t* labels, int64_t n, int32_t mode,
                       int32_t thr_i, float thr_f, const float* point_num_avg, int32_t group,
                       int32_t* cluster_offsets, int32_t* seeds, int32_t* d_count, void* ws,
                       size_t ws_bytes, b2s_stream_t stream) {
  if (n < 0 || mode < 0 || mode > 3 || (mode >= 2 && (!labels || !point_num_avg))) {
    set_error("cluster_select: invalid argument");
    return B2S_E_INVALID;
  }
  cudaMemsetAsync(d_count, 0, 8, stream);
  cudaMemsetAsync(cluster_offsets, 0, 4, stream);
  if (n == 0) return check_launch("cluster_select(empty)");
  Workspace w(ws, ws_bytes);
  int32_t* size = w.take<int32_t>(n);
  int32_t* flag = w.take<int32_t>(n);
  int32_t* ksize = w.take<int32_t>(n);
  int32_t* cidx = w.take<int32_t>(n);
  int32_t* coff = w.take<int32_t>(n);
  size_t sbytes = scan_ws_bytes(n);
  char* sws = w.take<char>(sbytes);
  if (!sws) {
    set_error("cluster_select: workspace too small");
    return B2S_E_WORKSPACE;
  }
  int grid = (int)cdiv(n, 256);
  cudaMemsetAsync(size, 0, (size_t)n * 4, stream);
  cl_hist_kernel<<<grid, 256, 0, stream>>>(comp, (int)n, size);
  cl_keep_kernel<<<grid, 256, 0, stream>>>(comp, labels, (int)n, size, mode, thr_i, thr_f, point_num_avg, group, flag, ksize);
  int rc = exclusive_scan_i32(flag, cidx, n, sws, sbytes, stream);
  if (rc) return rc;
  rc = exclusive_scan_i32(ksize, coff, n, sws, sbytes, stream);
  if (rc) return rc;
  cl_emit_kernel<<<grid, 256, 0, stream>>>((int)n, flag, ksize, cidx, coff, cluster_offsets, seeds, d_count);
  return check_launch("cluster_select");
}

int b2s_cluster_order(const int32_t* nbr_idx, const int32_t* start_len, const int16_t* labels,
                      const int32_t* comp, int64_t n, int64_t n_active, const int32_t* cluster_offsets,
                      const int32_t* seeds, int32_t n_cluster, int32_t* cluster_idxs, void* ws,
                      size_t ws_bytes, b2s_stream_t stream) {
  if (n < 0 || n_cluster < 0) {
    set_error("cluster_order: invalid argument");
    return B2S_E_INVALID;
  }
  if (n == 0 || n_cluster == 0) return B2S_OK;
  Workspace w(ws, ws_bytes);
  // measured on the benchmark batch (136k foreground points): device-wide variant ~1.0 ms (raw coordinates, deep
  // sparse graph) / ~0.9 ms (shifted, 55 M edges); the per-cluster variant was 1.7 ms / 9 ms when last compared ->
  // it is only used for small inputs, where a cooperative launch of 296 CTAs is all overhead
  (void)n_active;
  const bool dense = n >= 8192;
  if (!dense) {
    OrderArgs a;
    a.nbr_idx = nbr_idx;
    a.start_len = start_len;
    a.labels = labels;
    a.comp = comp;
    a.cluster_offsets = cluster_offsets;
    a.seeds = seeds;
    a.cluster_idxs = cluster_idxs;
    a.n = (int)n;
    a.n_cluster = n_cluster;
    a.cid = w.take<int32_t>(n);
    a.seedcid = w.take<int32_t>(n);
    a.pos = w.take<int32_t>(n);
    a.parent = w.take<int32_t>(n);
    a.cnt = w.take<int32_t>(n + 1);
    if (!a.cnt) {
      set_error("cluster_order: workspace too small");
      return B2S_E_WORKSPACE;
    }
    const unsigned gn = (unsigned)cdiv(n, 256);
    bfs_prep_kernel<<<gn, 256, 0, stream>>>(a, 0);
    bfs_prep_kernel<<<(unsigned)cdiv(n_cluster, 256), 256, 0, stream>>>(a, 1);
    bfs_prep_kernel<<<gn, 256, 0, stream>>>(a, 2);
    const int grid = std::min<int>(n_cluster, 8 * B2S_SM_COUNT);
    bfs_order_kernel<<<grid, CL_THREADS, 0, stream>>>(a);
    return check_launch("cluster_order");
  }
  GridOrderArgs a;
  a.nbr_idx = nbr_idx;
  a.start_len = start_len;
  a.labels = labels;
  a.comp = comp;
  a.cluster_offsets = cluster_offsets;
  a.seeds = seeds;
  a.cluster_idxs = cluster_idxs;
  a.n = (int)n;
  a.n_cluster = n_cluster;
  a.bar = w.take<unsigned>(64);
  a.cid = w.take<int32_t>(n);
  a.seedcid = w.take<int32_t>(n);
  a.parent = w.take<int32_t>(n);
  a.F0 = w.take<int32_t>(n);
  a.F1 = w.take<int32_t>(n);
  a.cnt = w.take<int32_t>(n + 1);
  a.base = w.take<int32_t>(n + 1);
  a.fstart = w.take<int32_t>(2 * (n + 2));  // double-buffered per level
  a.filled = w.take<int32_t>(2 * (n + 2));
  a.blocksum = w.take<int32_t>(4096);
  if (!a.blocksum) {
    set_error("cluster_order: workspace too small");
    return B2S_E_WORKSPACE;
  }
  cudaMemsetAsync(a.bar, 0, 64 * 4, stream);
  int grid = coop_grid((const void*)bfs_order_grid_kernel, CL_THREADS, n * 8);
  if (grid > 4096) grid = 4096;
  // (halving the grid for sparse, deep graphs was measured slower: 1.15 -> 1.64 ms; the levels are bound by the
  // per-level node sweep and frontier work, not by the barrier itself)
  void* args[] = {(void*)&a};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)bfs_order_grid_kernel, dim3(grid), dim3(CL_THREADS), args, 0, stream);
  if (e != cudaSuccess) {
    set_error(cudaGetErrorString(e));
    return B2S_E_LAUNCH;
  }
  return check_launch("cluster_order");
}

int b2s_cluster_centers(const int32_t* cluster_idxs, const int32_t* cluster_offsets,
                        int32_t n_cluster, const float* coords, const int16_t* labels,
                        const uint8_t* batch_idxs, float* centers, b2s_stream_t stream) {
  if (n_cluster <= 0) return B2S_OK;
  cluster_centers_kernel<<<(unsigned)cdiv((int64_t)n_cluster * 32, 128), 128, 0, stream>>>(
      cluster_idxs, cluster_offsets, n_cluster, coords, labels, batch_idxs, centers);
  return check_launch("cluster_centers");
}

int b2s_ha_assign(const float* frag_centers, int32_t n_frag, const float* prim_centers,
                  const int32_t* prim_offsets, int32_t n_prim, const float* radius_avg,
                  int32_t* assign, b2s_stream_t stream) {
  if (n_frag <= 0) return B2S_OK;
  ha_assign_kernel<<<(unsigned)cdiv(n_frag, 256), 256, 0, stream>>>(frag_centers, n_frag, prim_centers,
                                                                    prim_offsets, n_prim, radius_avg, assign);
  return check_launch("ha_assign");
}

size_t b2s_ha_concat_ws_bytes(int32_t n_frag, int32_t n_prim) {
  (void)n_frag;
  int64_t n = n_prim > 0 ? n_prim : 1;
  return 2 * align_up((size_t)n * 4) + scan_ws_bytes(n) + 1024;
}

int b2s_ha_concat(const int32_t* frag_idxs, const int32_t* frag_offsets, int32_t n_frag,
                  const int32_t* prim_idxs, const int32_t* prim_offsets, int32_t n_prim,
                  const int32_t* assign, int32_t* out_idxs, int32_t* out_offsets, void* ws,
                  size_t ws_bytes, b2s_stream_t stream) {
  if (n_prim <= 0) {
    cudaMemsetAsync(out_offsets, 0, 4, stream);
    return check_launch("ha_concat(empty)");
  }
  Workspace w(ws, ws_bytes);
  int32_t* totals = w.take<int32_t>(n_prim);
  int32_t* base = w.take<int32_t>(n_prim);
  size_t sbytes = scan_ws_bytes(n_prim);
  char* sws = w.take<char>(sbytes);
  if (!sws) {
    set_error("ha_concat: workspace too small");
    return B2S_E_WORKSPACE;
  }
  unsigned grid = (unsigned)cdiv((int64_t)n_prim * 32, 128);
  ha_concat_kernel<<<grid, 128, 0, stream>>>(frag_idxs, frag_offsets, n_frag, prim_idxs, prim_offsets, n_prim,
                                             assign, totals, base, out_idxs, out_offsets, 0);
  int rc = exclusive_scan_i32(totals, base, n_prim, sws, sbytes, stream);
  if (rc) return rc;
  ha_concat_kernel<<<grid, 128, 0, stream>>>(frag_idxs, frag_offsets, n_frag, prim_idxs, prim_offsets, n_prim,
                                             assign, totals, base, out_idxs, out_offsets, 1);
  return check_launch("ha_concat");
}

}  // extern "C"
