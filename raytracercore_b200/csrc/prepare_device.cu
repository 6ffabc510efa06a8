// prepare_device.cu — Scene.Prepare on the device (replaces BVH.Construct, Acceleration/BVH.cs:50-236,
// and the host-side flattening that follows it here).
//
// Two parts, both level-synchronous (a handful of small kernels per tree level, one host read-back of the next level's
// size per level):
//
//  1. build_bvh_sah_device: the top-down 16-bin SAH builder of host/bvh_builder.cpp made data-parallel. All segments (a
//     subtree under construction = a range of the primitive index array) of one level are processed together: centroid
//     bounds and bin boxes by atomic min / max (floats and doubles through an order-preserving integer map; a block whose
//     2048 positions lie in one segment pre-reduces in shared memory, so the top levels do not serialise on a few
//     addresses), the split sweep by one thread per segment, the partition as a stable scatter behind one prefix sum over
//     the whole index array. Every decision is the host builder's, in the host builder's arithmetic (this unit is built
//     with -fmad=false, the host with -ffp-contract=off): bins are order-independent (counts, min / max), the SAH sweep is
//     evaluated in the same order with the same strict comparison, the median fall-back (two primitives, or coincident
//     centroids) orders by (centroid, primitive ID) like the host's nth_element comparator — which fixes the *set* on each
//     side — and a subtree over k primitives owns the same 2k-1 consecutive nodes. The result is the host tree, node for
//     node and bit for bit (tests/test_gpu_prepare.py), so trace rates are those of the host SAH tree by construction.
//
//  2. flatten_device: the collapse of the reference-shaped binary tree into the quantised 8-wide device tree and the
//     packing of the primitive / material records, for the f32 mode. Levels of the binary tree by a breadth-first pass
//     (which also validates it: every node reached once, every primitive in exactly one leaf), then bottom-up per level the
//     bounded-leaf counts, finite boxes and the optimal-collapse table (prepare_common.h: dp_node), then the wide nodes
//     breadth-first, one level per step: children, grid, octant slots and quantised bounds per node (make_cnode), the two
//     base indices by a prefix sum over the level — which is exactly the order in which the serial host pass hands them
//     out. The arithmetic is prepare_common.h's, shared with the host flatten in rtc_api.cu, so the device image equals the
//     host image byte for byte (same test).
//
// Library use: cub::DeviceScan / cub::DeviceRadixSort (CCCL as shipped with the toolkit); everything else is hand-written.
#include <cub/cub.cuh>
#include <math_constants.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "prepare_common.h"
#include "rtc_internal.h"

namespace rtc {
namespace {

constexpr int kBins = 16;  // host/bvh_builder.cpp: kBins
constexpr int kPosThreads = 256;
constexpr int kPosPerBlock = 2048;

// order-preserving maps: float <-> uint32 (unsigned compare), double <-> int64 (signed compare)
__device__ __forceinline__ uint32_t ord_f32(float f) {
  const uint32_t u = __float_as_uint(f);
  return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float unord_f32(uint32_t k) { return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xFFFFFFFFu)); }
__device__ __forceinline__ long long ord_f64(double d) {
  const long long b = __double_as_longlong(d);
  return b ^ ((b >> 63) & 0x7FFFFFFFFFFFFFFFll);
}
__device__ __forceinline__ double unord_f64(long long k) { return __longlong_as_double(k ^ ((k >> 63) & 0x7FFFFFFFFFFFFFFFll)); }

struct Seg {  // a subtree under construction: root node index, range [b, e) of the index array
  int32_t base, b, e, pad;
};
struct SegWork {
  uint32_t omin[3], omax[3];  // centroid bounds, order-mapped (atomic targets)
  float cmin[3], scale[3];
  int32_t axis, bin;  // SAH split (axis < 0: median split along faxis)
  int32_t faxis, kl;
};
struct Bin {
  unsigned int n, pad;
  long long lo[3], hi[3];  // order-mapped doubles
};

__device__ __forceinline__ int bin_of(float c, float cmin, float scale) {
  int j = (int)((c - cmin) * scale);
  return min(max(j, 0), kBins - 1);
}

__global__ void k_sb_init(int32_t m, const double* __restrict__ boxes, float* cen, int32_t* idx, int32_t* seg_of) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  for (int a = 0; a < 3; a++) cen[(size_t)i * 3 + a] = (float)((boxes[(size_t)i * 6 + a] + boxes[(size_t)i * 6 + 3 + a]) * 0.5);
  idx[i] = i;
  seg_of[i] = m >= 2 ? 0 : -1;
}

__global__ void k_seg_reset(int32_t nseg, SegWork* work) {
  const int32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nseg) return;
  for (int a = 0; a < 3; a++) {
    work[s].omin[a] = 0xFFFFFFFFu;
    work[s].omax[a] = 0u;
  }
}

// centroid bounds of every segment
__global__ void __launch_bounds__(kPosThreads) k_seg_bounds(int32_t m, const int32_t* __restrict__ idx, const int32_t* __restrict__ seg_of,
                                                            const float* __restrict__ cen, SegWork* work) {
  const int32_t b0 = blockIdx.x * kPosPerBlock, b1 = min(m, b0 + kPosPerBlock);
  const int32_t s_first = seg_of[b0], s_last = seg_of[b1 - 1];
  if (s_first >= 0 && s_first == s_last) {  // the whole block lies in one segment (segments are contiguous and ordered)
    __shared__ uint32_t sh[6];
    if (threadIdx.x < 3) sh[threadIdx.x] = 0xFFFFFFFFu;
    else if (threadIdx.x < 6) sh[threadIdx.x] = 0u;
    __syncthreads();
    uint32_t mn[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, mx[3] = {0u, 0u, 0u};
    for (int32_t i = b0 + threadIdx.x; i < b1; i += kPosThreads) {
      const int32_t p = idx[i];
      for (int a = 0; a < 3; a++) {
        const uint32_t k = ord_f32(cen[(size_t)p * 3 + a]);
        mn[a] = min(mn[a], k);
        mx[a] = max(mx[a], k);
      }
    }
    for (int a = 0; a < 3; a++) {
      const uint32_t wmn = __reduce_min_sync(0xFFFFFFFFu, mn[a]), wmx = __reduce_max_sync(0xFFFFFFFFu, mx[a]);
      if ((threadIdx.x & 31) == 0) {
        atomicMin(&sh[a], wmn);
        atomicMax(&sh[3 + a], wmx);
      }
    }
    __syncthreads();
    if (threadIdx.x < 3) atomicMin(&work[s_first].omin[threadIdx.x], sh[threadIdx.x]);
    else if (threadIdx.x < 6) atomicMax(&work[s_first].omax[threadIdx.x - 3], sh[threadIdx.x]);
    return;
  }
  for (int32_t i = b0 + threadIdx.x; i < b1; i += kPosThreads) {
    const int32_t s = seg_of[i];
    if (s < 0) continue;
    const int32_t p = idx[i];
    for (int a = 0; a < 3; a++) {
      const uint32_t k = ord_f32(cen[(size_t)p * 3 + a]);
      atomicMin(&work[s].omin[a], k);
      atomicMax(&work[s].omax[a], k);
    }
  }
}

__global__ void k_seg_prepare(int32_t nseg, SegWork* work) {
  const int32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nseg) return;
  SegWork& w = work[s];
  float cmin[3], cmax[3];
  for (int a = 0; a < 3; a++) {
    cmin[a] = unord_f32(w.omin[a]);
    cmax[a] = unord_f32(w.omax[a]);
    const float ext = cmax[a] - cmin[a];
    w.cmin[a] = cmin[a];
    w.scale[a] = ext > 0 ? (kBins * (1.0f - 1e-6f)) / ext : 0.0f;
  }
  int a = 0;
  if (cmax[1] - cmin[1] > cmax[a] - cmin[a]) a = 1;
  if (cmax[2] - cmin[2] > cmax[a] - cmin[a]) a = 2;
  w.faxis = a;
}

__global__ void k_bins_init(int64_t nb, Bin* bins) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb) return;
  Bin b;
  b.n = 0;
  b.pad = 0;
  const long long pinf = ord_f64(CUDART_INF), ninf = ord_f64(-CUDART_INF);
  for (int c = 0; c < 3; c++) {
    b.lo[c] = pinf;
    b.hi[c] = ninf;
  }
  bins[i] = b;
}

// bins of the segments [s0, s1) with more than two primitives; positions [p0, p1) cover them
__global__ void __launch_bounds__(kPosThreads) k_bin(int32_t p0, int32_t p1, int32_t s0, int32_t s1, const int32_t* __restrict__ idx,
                                                     const int32_t* __restrict__ seg_of, const float* __restrict__ cen,
                                                     const double* __restrict__ boxes,
                                                     const Seg* __restrict__ segs, const SegWork* __restrict__ work, Bin* bins) {
  const int32_t b0 = p0 + blockIdx.x * kPosPerBlock, b1 = min(p1, b0 + kPosPerBlock);
  if (b0 >= b1) return;
  const int32_t s_first = seg_of[b0], s_last = seg_of[b1 - 1];
  __shared__ Bin sh[3 * kBins];
  const bool uniform = s_first >= 0 && s_first == s_last;
  if (uniform) {
    if (s_first < s0 || s_first >= s1 || segs[s_first].e - segs[s_first].b <= 2) return;
    const long long pinf = ord_f64(CUDART_INF), ninf = ord_f64(-CUDART_INF);
    for (int t = threadIdx.x; t < 3 * kBins; t += kPosThreads) {
      sh[t].n = 0;
      for (int c = 0; c < 3; c++) {
        sh[t].lo[c] = pinf;
        sh[t].hi[c] = ninf;
      }
    }
    __syncthreads();
  }
  for (int32_t i = b0 + threadIdx.x; i < b1; i += kPosThreads) {
    const int32_t s = seg_of[i];
    if (s < s0 || s >= s1) continue;
    if (!uniform && segs[s].e - segs[s].b <= 2) continue;
    const SegWork& w = work[s];
    const int32_t p = idx[i];
    long long l[3], h[3];
    for (int c = 0; c < 3; c++) {
      l[c] = ord_f64(boxes[(size_t)p * 6 + c]);
      h[c] = ord_f64(boxes[(size_t)p * 6 + 3 + c]);
    }
    for (int a = 0; a < 3; a++) {
      const float sc = w.scale[a];
      if (sc == 0) continue;
      const int j = bin_of(cen[(size_t)p * 3 + a], w.cmin[a], sc);
      Bin* bn = uniform ? &sh[a * kBins + j] : &bins[(size_t)(s - s0) * (3 * kBins) + a * kBins + j];
      atomicAdd(&bn->n, 1u);
      for (int c = 0; c < 3; c++) {
        atomicMin(&bn->lo[c], l[c]);
        atomicMax(&bn->hi[c], h[c]);
      }
    }
  }
  if (uniform) {
    __syncthreads();
    Bin* g = &bins[(size_t)(s_first - s0) * (3 * kBins)];
    for (int t = threadIdx.x; t < 3 * kBins; t += kPosThreads) {
      if (sh[t].n == 0) continue;
      atomicAdd(&g[t].n, sh[t].n);
      for (int c = 0; c < 3; c++) {
        atomicMin(&g[t].lo[c], sh[t].lo[c]);
        atomicMax(&g[t].hi[c], sh[t].hi[c]);
      }
    }
  }
}

__device__ __forceinline__ double sah_area(const double* l, const double* h) {
  const double dx = h[0] - l[0], dy = h[1] - l[1], dz = h[2] - l[2];
  return (dx * dy + dy * dz + dz * dx) * 2;  // AABB.GetSurfaceArea, AABB.cs:204-207
}

// the split of every segment in [s0, s1): the host builder's sweep (bvh_builder.cpp: Builder::build), one thread per segment
__global__ void k_sah(int32_t s0, int32_t s1, const Seg* __restrict__ segs, SegWork* work, const Bin* __restrict__ bins, rtc_bvh_node* nodes,
                      int32_t* nchild, int32_t* level_bases, int32_t* flags_out) {
  const int32_t s = s0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= s1) return;
  const Seg sg = segs[s];
  SegWork& w = work[s];
  const int32_t k = sg.e - sg.b;
  int best_axis = -1, best_bin = -1;
  int32_t best_ln = 0;
  double best_cost = CUDART_INF;
  if (k > 2) {
    const Bin* sb = &bins[(size_t)(s - s0) * (3 * kBins)];
    for (int a = 0; a < 3; a++) {
      if (w.scale[a] == 0) continue;
      const Bin* ab = sb + a * kBins;
      double rl[kBins][3], rh[kBins][3];
      int32_t rn[kBins];
      double l3[3] = {CUDART_INF, CUDART_INF, CUDART_INF}, h3[3] = {-CUDART_INF, -CUDART_INF, -CUDART_INF};
      int32_t cnt = 0;
      for (int j = kBins - 1; j >= 1; j--) {
        cnt += (int32_t)ab[j].n;
        for (int c = 0; c < 3; c++) {
          l3[c] = fmin(l3[c], unord_f64(ab[j].lo[c]));
          h3[c] = fmax(h3[c], unord_f64(ab[j].hi[c]));
          rl[j][c] = l3[c];
          rh[j][c] = h3[c];
        }
        rn[j] = cnt;
      }
      double ll[3] = {CUDART_INF, CUDART_INF, CUDART_INF}, lh[3] = {-CUDART_INF, -CUDART_INF, -CUDART_INF};
      int32_t ln = 0;
      for (int j = 0; j < kBins - 1; j++) {
        ln += (int32_t)ab[j].n;
        for (int c = 0; c < 3; c++) {
          ll[c] = fmin(ll[c], unord_f64(ab[j].lo[c]));
          lh[c] = fmax(lh[c], unord_f64(ab[j].hi[c]));
        }
        if (ln == 0 || rn[j + 1] == 0) continue;
        const double cost = sah_area(ll, lh) * ln + sah_area(rl[j + 1], rh[j + 1]) * rn[j + 1];
        if (cost < best_cost) {
          best_cost = cost;
          best_axis = a;
          best_bin = j;
          best_ln = ln;
        }
      }
    }
  }
  int32_t kl;
  if (best_axis >= 0) {
    kl = best_ln;
  } else {
    kl = k / 2;  // median split along the widest centroid axis, ties by primitive ID
    if (k > 2) atomicOr(flags_out, 1);  // needs the ordering pass
  }
  w.axis = best_axis;
  w.bin = best_bin;
  w.kl = kl;
  rtc_bvh_node& nd = nodes[sg.base];
  nd.left = sg.base + 1;
  nd.right = sg.base + 2 * kl;
  nd.prim = -1;
  nd.pad = 0;
  nchild[s] = (kl >= 2 ? 1 : 0) + (k - kl >= 2 ? 1 : 0);
  level_bases[s] = sg.base;
}

__global__ void k_level_total(int32_t nseg, const int32_t* nchild, const int32_t* child0, int32_t* flags_out) {
  flags_out[1] = child0[nseg - 1] + nchild[nseg - 1];
}

// ordering pass for median splits of more than two primitives: (centroid on faxis, primitive index) ascending inside the
// segment, everything else stays where it is. Two stable radix sorts: by the secondary key first.
__global__ void k_sortkey(int32_t m, int pass, const int32_t* __restrict__ idx, const int32_t* __restrict__ seg_of, const float* __restrict__ cen,
                          const Seg* __restrict__ segs, const SegWork* __restrict__ work, uint64_t* keys) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int32_t s = seg_of[i];
  uint64_t key = (uint64_t)(uint32_t)i << 32;
  if (s >= 0) {
    const Seg sg = segs[s];
    if (work[s].axis < 0 && sg.e - sg.b > 2) {
      const int32_t p = idx[i];
      const uint32_t low = pass == 0 ? (uint32_t)p : ord_f32(cen[(size_t)p * 3 + work[s].faxis] + 0.0f);  // (-0 orders as +0)
      key = ((uint64_t)(uint32_t)sg.b << 32) | low;
    }
  }
  keys[i] = key;
}

__device__ __forceinline__ bool median_less(const float* cen, int a, int32_t p, int32_t q) {
  const float cp = cen[(size_t)p * 3 + a], cq = cen[(size_t)q * 3 + a];
  return cp < cq || (cp == cq && p < q);
}

__global__ void k_side(int32_t m, const int32_t* __restrict__ idx, const int32_t* __restrict__ seg_of, const float* __restrict__ cen,
                       const Seg* __restrict__ segs, const SegWork* __restrict__ work, uint32_t* flag) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int32_t s = seg_of[i];
  uint32_t f = 0;
  if (s >= 0) {
    const Seg sg = segs[s];
    const SegWork& w = work[s];
    const int32_t p = idx[i];
    if (w.axis >= 0) {
      f = bin_of(cen[(size_t)p * 3 + w.axis], w.cmin[w.axis], w.scale[w.axis]) <= w.bin;
    } else if (sg.e - sg.b == 2) {
      const bool first_less = median_less(cen, w.faxis, idx[sg.b], idx[sg.b + 1]);
      f = (i == sg.b) ? first_less : !first_less;
    } else {
      f = (i - sg.b) < w.kl;  // ordered by the sort pass
    }
  }
  flag[i] = f;
}

__global__ void k_scatter(int32_t m, const int32_t* __restrict__ idx, const int32_t* __restrict__ seg_of, const uint32_t* __restrict__ flag,
                          const uint32_t* __restrict__ pre, const Seg* __restrict__ segs, const SegWork* __restrict__ work,
                          const int32_t* __restrict__ child0, const double* __restrict__ boxes,
                          const int32_t* __restrict__ prim_ids, int32_t* idx_out, int32_t* seg_out, rtc_bvh_node* nodes) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int32_t s = seg_of[i];
  const int32_t p = idx[i];
  if (s < 0) {
    idx_out[i] = p;
    seg_out[i] = -1;
    return;
  }
  const Seg sg = segs[s];
  const int32_t kl = work[s].kl, k = sg.e - sg.b;
  const int32_t r = (int32_t)(pre[i] - pre[sg.b]);
  const bool left = flag[i] != 0;
  const int32_t np = left ? sg.b + r : sg.b + kl + (i - sg.b - r);
  idx_out[np] = p;
  const int32_t ck = left ? kl : k - kl;
  if (ck == 1) {
    rtc_bvh_node nd;
    for (int c = 0; c < 3; c++) {
      nd.bmin[c] = boxes[(size_t)p * 6 + c];
      nd.bmax[c] = boxes[(size_t)p * 6 + 3 + c];
    }
    nd.left = nd.right = -1;
    nd.prim = prim_ids ? prim_ids[p] : p;
    nd.pad = 0;
    nodes[left ? sg.base + 1 : sg.base + 2 * kl] = nd;
    seg_out[np] = -1;
  } else {
    seg_out[np] = child0[s] + ((!left && kl >= 2) ? 1 : 0);
  }
}

__global__ void k_children(int32_t nseg, const Seg* __restrict__ segs, const SegWork* __restrict__ work, const int32_t* __restrict__ child0,
                           Seg* next) {
  const int32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nseg) return;
  const Seg sg = segs[s];
  const int32_t kl = work[s].kl, k = sg.e - sg.b;
  int32_t c = child0[s];
  if (kl >= 2) next[c++] = Seg{sg.base + 1, sg.b, sg.b + kl, 0};
  if (k - kl >= 2) next[c] = Seg{sg.base + 2 * kl, sg.b + kl, sg.e, 0};
}

__global__ void k_union_level(int32_t n, const int32_t* __restrict__ bases, rtc_bvh_node* nodes) {
  const int32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  rtc_bvh_node& nd = nodes[bases[t]];
  const rtc_bvh_node& l = nodes[nd.left];
  const rtc_bvh_node& r = nodes[nd.right];
  for (int c = 0; c < 3; c++) {  // AABB.Combine, AABB.cs:38-43
    nd.bmin[c] = fmin(l.bmin[c], r.bmin[c]);
    nd.bmax[c] = fmax(l.bmax[c], r.bmax[c]);
  }
}

__global__ void k_single_leaf(const double* boxes, const int32_t* prim_ids, rtc_bvh_node* nodes) {
  rtc_bvh_node nd;
  for (int c = 0; c < 3; c++) {
    nd.bmin[c] = boxes[c];
    nd.bmax[c] = boxes[3 + c];
  }
  nd.left = nd.right = -1;
  nd.prim = prim_ids ? prim_ids[0] : 0;
  nd.pad = 0;
  nodes[0] = nd;
}

// Work arrays of one call, carved out of the caller's arena (one cudaMalloc per rtc_prepare_device instead of forty); what
// does not fit is allocated on its own and freed when the pool goes out of scope.
struct DevPool {
  PrepareArena* arena;
  std::vector<void*> extra;
  cudaError_t err = cudaSuccess;
  explicit DevPool(PrepareArena* a) : arena(a) {}
  template <typename T>
  T* get(size_t count) {
    const size_t bytes = (std::max<size_t>(count * sizeof(T), 16) + 255) & ~(size_t)255;
    if (arena && arena->used + bytes <= arena->cap) {
      T* p = (T*)(arena->base + arena->used);
      arena->used += bytes;
      return p;
    }
    void* p = nullptr;
    if (err == cudaSuccess) err = cudaMalloc(&p, bytes);
    if (err != cudaSuccess) return nullptr;
    extra.push_back(p);
    return (T*)p;
  }
  ~DevPool() {
    for (void* p : extra) cudaFree(p);
  }
};

inline int grid_for(int64_t n, int t) { return (int)std::max<int64_t>(1, (n + t - 1) / t); }

}  // namespace

namespace {
__global__ void k_gather_boxes(int32_t m, const int32_t* __restrict__ ids, const double* __restrict__ all, double* out) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  for (int c = 0; c < 6; c++) out[(size_t)i * 6 + c] = all[(size_t)ids[i] * 6 + c];
}
}  // namespace

cudaError_t launch_gather_boxes(cudaStream_t stream, int32_t m, const int32_t* d_ids, const double* d_all, double* d_out) {
  k_gather_boxes<<<grid_for(m, 256), 256, 0, stream>>>(m, d_ids, d_all, d_out);
  return cudaGetLastError();
}

#define PCU(call)                       \
  do {                                  \
    cudaError_t e__ = (call);           \
    if (e__ != cudaSuccess) return e__; \
  } while (0)

size_t build_bvh_sah_scratch_bytes(int32_t m) {
  const size_t mz = (size_t)std::max(m, 1), max_seg = mz / 2 + 1, bin_cap = std::min<size_t>(max_seg, 65536);
  return mz * (12 + 8 + 8 + 4 + 4 + 4) + max_seg * (2 * sizeof(Seg) + sizeof(SegWork) + 8) + bin_cap * 3 * kBins * sizeof(Bin) + mz * 8 +
         (size_t)64 * 256 + ((size_t)1 << 20);  // (+ alignment slack and the scans' temporary storage)
}

cudaError_t build_bvh_sah_device(cudaStream_t stream, PrepareArena* arena, int32_t* h_pin, int32_t m, const double* d_boxes,
                                 const int32_t* d_prim_ids, rtc_bvh_node* d_nodes, int32_t* levels_out) {
  *levels_out = 0;
  if (m <= 0) return cudaErrorInvalidValue;
  if (m == 1) {
    k_single_leaf<<<1, 1, 0, stream>>>(d_boxes, d_prim_ids, d_nodes);
    return cudaGetLastError();
  }
  DevPool pool(arena);
  const int32_t max_seg = m / 2 + 1;
  const int32_t bin_cap = std::min<int32_t>(max_seg, 65536);  // segments binned per batch (3 x 16 bins of 56 B each: <= 176 MB)
  float* cen = pool.get<float>((size_t)m * 3);
  int32_t* idx[2] = {pool.get<int32_t>(m), pool.get<int32_t>(m)};
  int32_t* seg_of[2] = {pool.get<int32_t>(m), pool.get<int32_t>(m)};
  Seg* segs[2] = {pool.get<Seg>(max_seg), pool.get<Seg>(max_seg)};
  SegWork* work = pool.get<SegWork>(max_seg);
  int32_t* nchild = pool.get<int32_t>(max_seg);
  int32_t* child0 = pool.get<int32_t>(max_seg);
  Bin* bins = pool.get<Bin>((size_t)bin_cap * 3 * kBins);
  uint32_t* flag = pool.get<uint32_t>(m);
  uint32_t* pre = pool.get<uint32_t>(m);
  int32_t* level_bases = pool.get<int32_t>(m);  // inner nodes, level after level (m - 1 in all)
  int32_t* d_flags = pool.get<int32_t>(2);      // [0] = some segment needs the ordering pass, [1] = segments of the next level
  uint64_t* keys[2] = {nullptr, nullptr};       // ordering pass only (allocated on first use)
  size_t tmp_scan_a = 0, tmp_scan_b = 0, tmp_sort = 0;
  PCU(cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan_a, flag, pre, m, stream));
  PCU(cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan_b, nchild, child0, max_seg, stream));
  size_t tmp_bytes = std::max(tmp_scan_a, tmp_scan_b);
  void* d_tmp = pool.get<char>(tmp_bytes);
  void* d_tmp_sort = nullptr;
  int32_t* h_flags = h_pin;
  if (pool.err != cudaSuccess) return pool.err;

  const int gm = grid_for(m, 256);
  const int gpos = grid_for(m, kPosPerBlock);
  k_sb_init<<<gm, 256, 0, stream>>>(m, d_boxes, cen, idx[0], seg_of[0]);
  const Seg root_seg{0, 0, m, 0};
  PCU(cudaMemcpyAsync(segs[0], &root_seg, sizeof(Seg), cudaMemcpyHostToDevice, stream));
  std::vector<int32_t> level_off{0};
  int32_t nseg = 1, cur = 0, levels = 0;
  while (nseg > 0) {
    if (++levels > 100000) return cudaErrorUnknown;  // (every split leaves both sides non-empty: at most m levels)
    const int gs = grid_for(nseg, 128);
    k_seg_reset<<<gs, 128, 0, stream>>>(nseg, work);
    PCU(cudaMemsetAsync(d_flags, 0, 2 * sizeof(int32_t), stream));
    k_seg_bounds<<<gpos, kPosThreads, 0, stream>>>(m, idx[cur], seg_of[cur], cen, work);
    k_seg_prepare<<<gs, 128, 0, stream>>>(nseg, work);
    int32_t* bases = level_bases + level_off.back();
    for (int32_t s0 = 0; s0 < nseg; s0 += bin_cap) {
      const int32_t s1 = std::min(nseg, s0 + bin_cap);
      const int64_t nb = (int64_t)(s1 - s0) * 3 * kBins;
      k_bins_init<<<grid_for(nb, 256), 256, 0, stream>>>(nb, bins);
      // (a batch's positions are not known on the host: the kernel walks all of them and skips the other batches' segments)
      k_bin<<<gpos, kPosThreads, 0, stream>>>(0, m, s0, s1, idx[cur], seg_of[cur], cen, d_boxes, segs[cur], work, bins);
      k_sah<<<grid_for(s1 - s0, 64), 64, 0, stream>>>(s0, s1, segs[cur], work, bins, d_nodes, nchild, bases, d_flags);
    }
    size_t tb = tmp_bytes;
    PCU(cub::DeviceScan::ExclusiveSum(d_tmp, tb, nchild, child0, nseg, stream));
    k_level_total<<<1, 1, 0, stream>>>(nseg, nchild, child0, d_flags);
    PCU(cudaMemcpyAsync(h_flags, d_flags, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    PCU(cudaStreamSynchronize(stream));
    if (h_flags[0]) {
      if (!keys[0]) {
        keys[0] = pool.get<uint64_t>(m);
        keys[1] = pool.get<uint64_t>(m);
        PCU(cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, keys[0], keys[1], idx[0], idx[1], m, 0, 64, stream));
        d_tmp_sort = pool.get<char>(tmp_sort);
        if (pool.err != cudaSuccess) return pool.err;
      }
      int32_t* a = idx[cur];
      int32_t* b = idx[cur ^ 1];
      for (int pass = 0; pass < 2; pass++) {
        k_sortkey<<<gm, 256, 0, stream>>>(m, pass, a, seg_of[cur], cen, segs[cur], work, keys[0]);
        size_t ts = tmp_sort;
        PCU(cub::DeviceRadixSort::SortPairs(d_tmp_sort, ts, keys[0], keys[1], a, b, m, 0, 64, stream));
        std::swap(a, b);
      }  // two passes: the ordered array is back in idx[cur]
    }
    k_side<<<gm, 256, 0, stream>>>(m, idx[cur], seg_of[cur], cen, segs[cur], work, flag);
    tb = tmp_bytes;
    PCU(cub::DeviceScan::ExclusiveSum(d_tmp, tb, flag, pre, m, stream));
    k_scatter<<<gm, 256, 0, stream>>>(m, idx[cur], seg_of[cur], flag, pre, segs[cur], work, child0, d_boxes, d_prim_ids, idx[cur ^ 1],
                                      seg_of[cur ^ 1], d_nodes);
    k_children<<<gs, 128, 0, stream>>>(nseg, segs[cur], work, child0, segs[cur ^ 1]);
    level_off.push_back(level_off.back() + nseg);
    nseg = h_flags[1];
    cur ^= 1;
  }
  for (size_t l = level_off.size() - 1; l-- > 0;) {  // inner boxes, deepest level first
    const int32_t cnt = level_off[l + 1] - level_off[l];
    k_union_level<<<grid_for(cnt, 256), 256, 0, stream>>>(cnt, level_bases + level_off[l], d_nodes);
  }
  PCU(cudaGetLastError());
  PCU(cudaStreamSynchronize(stream));  // (the pool is released on return)
  *levels_out = levels;
  return cudaSuccess;
}

// =====================================================================================================================
// flatten (f32 mode)
// =====================================================================================================================
namespace {

enum { FERR_TWICE = 1, FERR_PRIM_RANGE = 2, FERR_PRIM_TWICE = 4, FERR_CHILD_RANGE = 8 };

// one breadth-first step: nodes order[lo, hi) -> their children appended at order[*count ...]
__global__ void k_bfs(int32_t lo, int32_t hi, int32_t nn, int32_t n, const rtc_bvh_node* __restrict__ nodes, int32_t* order, int32_t* node_seen,
                      int32_t* prim_seen, int32_t* counters /* [0] append cursor, [1] leaves, [2] error bits */) {
  const int32_t t = lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= hi) return;
  const int32_t i = order[t];
  const int32_t prim = nodes[i].prim, l = nodes[i].left, r = nodes[i].right;
  if (prim >= 0) {
    if (prim >= n) {
      atomicOr(&counters[2], FERR_PRIM_RANGE);
      return;
    }
    if (atomicExch(&prim_seen[prim], 1)) atomicOr(&counters[2], FERR_PRIM_TWICE);
    atomicAdd(&counters[1], 1);
    return;
  }
  if (l < 0 || l >= nn || r < 0 || r >= nn) {
    atomicOr(&counters[2], FERR_CHILD_RANGE);
    return;
  }
  const int32_t kids[2] = {l, r};
  for (int c = 0; c < 2; c++) {
    if (atomicExch(&node_seen[kids[c]], 1)) {
      atomicOr(&counters[2], FERR_TWICE);
      continue;
    }
    order[atomicAdd(&counters[0], 1)] = kids[c];
  }
}

struct UpArrays {
  int32_t* nf;
  int32_t* nl;
  double* fmin;
  double* fmax;
  float* T;
  uint8_t* cut;
  int32_t* unb_nodes;
  int32_t* unb_count;
};

// bottom-up step over order[lo, hi): bounded-leaf count, leaf count, finite box, collapse table
__global__ void k_up(int32_t lo, int32_t hi, const int32_t* __restrict__ order, const rtc_bvh_node* __restrict__ nodes, UpArrays u) {
  const int32_t t = lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= hi) return;
  const int32_t i = order[t];
  const rtc_bvh_node& nd = nodes[i];
  double* mn = u.fmin + (size_t)i * 3;
  double* mx = u.fmax + (size_t)i * 3;
  if (nd.prim >= 0) {
    u.nl[i] = 1;
    if (prep::finite_box(nd)) {
      u.nf[i] = 1;
      for (int a = 0; a < 3; a++) {
        mn[a] = nd.bmin[a];
        mx[a] = nd.bmax[a];
      }
    } else {
      u.nf[i] = 0;
      for (int a = 0; a < 3; a++) {
        mn[a] = CUDART_INF;
        mx[a] = -CUDART_INF;
      }
      u.unb_nodes[atomicAdd(u.unb_count, 1)] = i;
    }
    return;
  }
  const int32_t l = nd.left, r = nd.right;
  const int32_t nfl = u.nf[l], nfr = u.nf[r];
  u.nf[i] = nfl + nfr;
  u.nl[i] = u.nl[l] + u.nl[r];
  for (int a = 0; a < 3; a++) {
    const double la = u.fmin[(size_t)l * 3 + a], ra = u.fmin[(size_t)r * 3 + a];
    const double lb = u.fmax[(size_t)l * 3 + a], rb = u.fmax[(size_t)r * 3 + a];
    mn[a] = ra < la ? ra : la;  // std::min / std::max of the host pass
    mx[a] = lb < rb ? rb : lb;
  }
  if (nfl == 0 || nfr == 0) return;  // transparent (or unbounded only): resolve() skips it
  const int32_t rl = prep::resolve(nodes, u.nf, l), rr = prep::resolve(nodes, u.nf, r);
  prep::dp_node(u.T + (size_t)rl * 8, u.T + (size_t)rr * 8, prep::box_area(mn, mx), u.T + (size_t)i * 8, u.cut + (size_t)i * 9);
}

// top-down step: left-first leaf rank of the first leaf below every node
__global__ void k_rank(int32_t lo, int32_t hi, const int32_t* __restrict__ order, const rtc_bvh_node* __restrict__ nodes, const int32_t* __restrict__ nl,
                       int32_t* first) {
  const int32_t t = lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= hi) return;
  const int32_t i = order[t];
  const rtc_bvh_node& nd = nodes[i];
  if (nd.prim >= 0) return;
  first[nd.left] = first[i];
  first[nd.right] = first[i] + nl[nd.left];
}

__global__ void k_unb_fetch(int32_t n_unb, const int32_t* __restrict__ unb_nodes, const rtc_bvh_node* __restrict__ nodes, const int32_t* first,
                            int32_t* out /* n_unb x (rank, prim) */) {
  const int32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_unb) return;
  const int32_t i = unb_nodes[t];
  out[2 * t] = first ? first[i] : 0;
  out[2 * t + 1] = nodes[i].prim;
}

__global__ void k_resolve_root(const rtc_bvh_node* nodes, const int32_t* nf, int32_t root, int32_t* queue) {
  queue[0] = prep::resolve(nodes, nf, root);
}

// wide level, phase A: children, grid, octant slots, quantised bounds of the wide nodes queue[lo, hi)
__global__ void __launch_bounds__(64) k_wide_a(int32_t lo, int32_t hi, const int32_t* __restrict__ queue, prep::TreeView tv, CNode* qn, int32_t* kids,
                                               int8_t* cis, unsigned long long* cnt) {
  const int32_t q = lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= hi) return;
  const int32_t bnode = queue[q];
  int32_t k8[8];
  int8_t c8[8];
  int nk;
  if (tv.nodes[bnode].prim >= 0) {
    k8[0] = bnode;  // tree of a single bounded primitive
    nk = 1;
  } else {
    nk = prep::gather_children(tv, bnode, k8);
  }
  for (int c = nk; c < 8; c++) k8[c] = -1;
  CNode cn;
  prep::make_cnode(tv, k8, nk, cn, c8);
  qn[q] = cn;
  for (int c = 0; c < 8; c++) {
    kids[(size_t)q * 8 + c] = k8[c];
    cis[(size_t)q * 8 + c] = c8[c];
  }
  const unsigned ni = __popc(cn.e_imask >> 24), nlf = __popc(cn.lmask);
  cnt[q - lo] = ((unsigned long long)ni << 32) | nlf;
}

// phase B: base indices in emission order, leaf slots, next level's queue entries in slot order
__global__ void k_wide_b(int32_t lo, int32_t hi, int32_t next_slot0, const unsigned long long* __restrict__ scan, const int32_t* __restrict__ kids,
                         const int8_t* __restrict__ cis, const rtc_bvh_node* __restrict__ nodes, CNode* qn, int32_t* queue, int32_t* slot_prim,
                         unsigned long long* total /* last element's inclusive sum */, const unsigned long long* __restrict__ cnt) {
  const int32_t q = lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= hi) return;
  const unsigned long long ex = scan[q - lo];
  uint32_t child = (uint32_t)hi + (uint32_t)(ex >> 32);
  uint32_t slot = (uint32_t)next_slot0 + (uint32_t)(ex & 0xFFFFFFFFull);
  qn[q].child_base = child;
  qn[q].prim_base = slot;
  for (int s = 0; s < 8; s++) {
    const int c = cis[(size_t)q * 8 + s];
    if (c < 0) continue;
    const int32_t k = kids[(size_t)q * 8 + c];
    const int32_t prim = nodes[k].prim;
    if (prim >= 0)
      slot_prim[slot++] = prim;
    else
      queue[child++] = k;
  }
  if (q == hi - 1) *total = ex + cnt[q - lo];
}

struct RecordArrays {
  const StagedPrim* staged;  // input order
  const int32_t* slot_prim;
  DPrim<float>* prims;
  DMat<float>* mats;
  int32_t* aux;
  int32_t* prim_id;
  int32_t* id_to_slot;
  V4<float>* sgeom;
};

// the records of slot s (rtc_api.cu: build_device_scene, "primitive + material records", f32 mode), from the staged input
// record of its primitive: the float conversions and the half packing were done by the host code the host flatten uses
__global__ void k_records(int32_t n, RecordArrays r) {
  const int32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const int32_t p = r.slot_prim[s];
  r.prim_id[s] = p;
  r.id_to_slot[p] = s;
  const StagedPrim sp = r.staged[p];
  const float* g = sp.geom;
  const uint8_t k = (uint8_t)(sp.kind_flags & 0xFF), f = (uint8_t)(sp.kind_flags >> 8);
  const int32_t xf = sp.xform;
  DPrim<float> d;
  memset(&d, 0, sizeof(d));
  int32_t aux = -1;
  V4<float> sg;
  if (k == RTC_KIND_TRIANGLE) {
    d.a.x = g[0]; d.a.y = g[1]; d.a.z = g[2]; d.a.w = g[9];
    d.b.x = g[3]; d.b.y = g[4]; d.b.z = g[5]; d.b.w = g[10];
    d.c.x = g[6]; d.c.y = g[7]; d.c.z = g[8];
    if ((f & RTC_FLAG_VNORMALS) && xf >= 0) aux = xf | (int32_t)REF_VNORMALS_AUX;
    sg = V4<float>{g[9], g[10], g[11], __uint_as_float(aux >= 0 ? 1u : 0u)};
  } else {
    d.a.x = g[0]; d.a.y = g[1]; d.a.z = g[2]; d.a.w = g[3];
    if (k == RTC_KIND_SPHERE && (f & RTC_FLAG_TRANSFORMED) && xf >= 0) aux = xf;
    sg = V4<float>{g[0], g[1], g[2], k == RTC_KIND_SPHERE ? g[3] : 0.0f};
  }
  d.c.w = __uint_as_float(prep::leaf_ref_of(k, f, xf, (uint32_t)s));
  r.prims[s] = d;
  r.aux[s] = aux;
  r.sgeom[s] = sg;
  DMat<float> dm;
  for (int i = 0; i < 8; i++) dm.w[i] = sp.mat[i];
  r.mats[s] = dm;
}

}  // namespace

size_t flatten_scratch_bytes(int32_t n_nodes, int32_t n_prims) {
  const size_t nn = (size_t)std::max(n_nodes, 1), n = (size_t)std::max(n_prims, 1);
  return nn * (4 + 4 + 4 + 4 + 24 + 24 + 32 + 9 + 4) + n * (4 + 4 + 8) + n * (4 + 32 + 8 + 8 + 8 + sizeof(CNode)) + (size_t)64 * 256 + ((size_t)1 << 20);
}

int flatten_device(cudaStream_t stream, PrepareArena* arena, int32_t* h_pin, const FlattenInput& in, FlattenOutput& out, std::string& err) {
  auto cuda_fail = [&](cudaError_t e, const char* what) {
    err = std::string(what) + ": " + cudaGetErrorString(e);
    return RTC_ERR_CUDA;
  };
#define FCU(call)                                       \
  do {                                                  \
    cudaError_t e__ = (call);                           \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
  } while (0)
  const int32_t nn = in.n_nodes, n = in.n_prims;
  const rtc_bvh_node* nodes = in.nodes;
  DevPool pool(arena);
  int32_t* order = pool.get<int32_t>(nn);
  int32_t* node_seen = pool.get<int32_t>(nn);
  int32_t* prim_seen = pool.get<int32_t>(n);
  int32_t* counters = pool.get<int32_t>(4);
  UpArrays up;
  up.nf = pool.get<int32_t>(nn);
  up.nl = pool.get<int32_t>(nn);
  up.fmin = pool.get<double>((size_t)nn * 3);
  up.fmax = pool.get<double>((size_t)nn * 3);
  up.T = pool.get<float>((size_t)nn * 8);
  up.cut = pool.get<uint8_t>((size_t)nn * 9);
  up.unb_nodes = pool.get<int32_t>(n);
  up.unb_count = counters + 3;
  int32_t* h_counters = h_pin;  // (pinned, 8 x int32)
  if (pool.err != cudaSuccess) return cuda_fail(pool.err, "cudaMalloc (flatten work arrays)");

  // ---- levels of the binary tree + validation -------------------------------------------------------------------
  FCU(cudaMemsetAsync(node_seen, 0, (size_t)nn * sizeof(int32_t), stream));
  FCU(cudaMemsetAsync(prim_seen, 0, (size_t)n * sizeof(int32_t), stream));
  FCU(cudaMemsetAsync(up.T, 0, (size_t)nn * 8 * sizeof(float), stream));
  FCU(cudaMemsetAsync(up.cut, 0, (size_t)nn * 9, stream));
  {
    const int32_t init[4] = {1, 0, 0, 0};
    FCU(cudaMemcpyAsync(counters, init, sizeof(init), cudaMemcpyHostToDevice, stream));
    const int32_t one = 1;
    FCU(cudaMemcpyAsync(order, &in.root, sizeof(int32_t), cudaMemcpyHostToDevice, stream));
    FCU(cudaMemcpyAsync(node_seen + in.root, &one, sizeof(int32_t), cudaMemcpyHostToDevice, stream));
  }
  std::vector<int32_t> lvl{0, 1};
  while (lvl[lvl.size() - 1] > lvl[lvl.size() - 2]) {
    const int32_t lo = lvl[lvl.size() - 2], hi = lvl.back();
    k_bfs<<<grid_for(hi - lo, 256), 256, 0, stream>>>(lo, hi, nn, n, nodes, order, node_seen, prim_seen, counters);
    FCU(cudaMemcpyAsync(h_counters, counters, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    FCU(cudaStreamSynchronize(stream));
    if (h_counters[2]) break;
    lvl.push_back(h_counters[0]);
  }
  if (h_counters[2] & FERR_TWICE) { err = "BVH is not a tree (node reached twice)"; return RTC_ERR_INVALID; }
  if (h_counters[2] & FERR_PRIM_RANGE) { err = "BVH leaf references a primitive out of range"; return RTC_ERR_INVALID; }
  if (h_counters[2] & FERR_PRIM_TWICE) { err = "primitive referenced by two BVH leaves"; return RTC_ERR_INVALID; }
  if (h_counters[2] & FERR_CHILD_RANGE) { err = "BVH child index out of range"; return RTC_ERR_INVALID; }
  if (h_counters[1] != n) { err = "BVH does not reference every primitive exactly once"; return RTC_ERR_INVALID; }
  lvl.pop_back();  // (the last entry repeats its predecessor: the empty level that ended the loop)
  const int n_levels = (int)lvl.size() - 1;

  // ---- bottom-up: counts, finite boxes, collapse table ------------------------------------------------------------
  for (int l = n_levels - 1; l >= 0; l--)
    k_up<<<grid_for(lvl[l + 1] - lvl[l], 128), 128, 0, stream>>>(lvl[l], lvl[l + 1], order, nodes, up);
  FCU(cudaMemcpyAsync(h_counters, up.nf + in.root, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  FCU(cudaMemcpyAsync(h_counters + 1, up.unb_count, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  FCU(cudaStreamSynchronize(stream));
  const int32_t n_bounded = h_counters[0], n_unb = h_counters[1];
  if (n_bounded + n_unb != n) { err = "internal: bounded + unbounded leaves != primitives"; return RTC_ERR_INVALID; }

  // ---- unbounded primitives in left-first order -----------------------------------------------------------------------
  std::vector<std::pair<int32_t, int32_t>> unb;  // (rank, prim)
  if (n_unb > 0) {
    int32_t* first = nullptr;
    if (n_unb > 1) {
      first = pool.get<int32_t>(nn);
      if (pool.err != cudaSuccess) return cuda_fail(pool.err, "cudaMalloc (leaf ranks)");
      FCU(cudaMemsetAsync(first + in.root, 0, sizeof(int32_t), stream));
      for (int l = 0; l < n_levels; l++)
        k_rank<<<grid_for(lvl[l + 1] - lvl[l], 256), 256, 0, stream>>>(lvl[l], lvl[l + 1], order, nodes, up.nl, first);
    }
    int32_t* d_pairs = pool.get<int32_t>((size_t)n_unb * 2);
    if (pool.err != cudaSuccess) return cuda_fail(pool.err, "cudaMalloc (unbounded list)");
    k_unb_fetch<<<grid_for(n_unb, 256), 256, 0, stream>>>(n_unb, up.unb_nodes, nodes, first, d_pairs);
    std::vector<int32_t> h_pairs((size_t)n_unb * 2);
    FCU(cudaMemcpyAsync(h_pairs.data(), d_pairs, h_pairs.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    FCU(cudaStreamSynchronize(stream));
    for (int32_t t = 0; t < n_unb; t++) unb.push_back({h_pairs[2 * t], h_pairs[2 * t + 1]});
    std::sort(unb.begin(), unb.end());
  }

  // ---- the wide tree, breadth-first ---------------------------------------------------------------------------------------
  int32_t n_wide = 0, max_depth = 0;
  CNode* qn = nullptr;
  if (n_bounded > 0) {
    const int32_t nw_max = std::max(n_bounded - 1, 1);
    qn = pool.get<CNode>(nw_max);
    int32_t* queue = pool.get<int32_t>(nw_max);
    int32_t* kids = pool.get<int32_t>((size_t)nw_max * 8);
    int8_t* cis = pool.get<int8_t>((size_t)nw_max * 8);
    unsigned long long* cnt = pool.get<unsigned long long>(nw_max);
    unsigned long long* scan = pool.get<unsigned long long>(nw_max);
    unsigned long long* total = pool.get<unsigned long long>(1);
    size_t tmp_bytes = 0;
    FCU(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cnt, scan, nw_max, stream));
    void* d_tmp = pool.get<char>(tmp_bytes);
    if (pool.err != cudaSuccess) return cuda_fail(pool.err, "cudaMalloc (wide-tree work arrays)");
    unsigned long long* h_total = (unsigned long long*)(h_counters + 4);
    const prep::TreeView tv{nodes, up.nf, up.fmin, up.fmax, up.T, up.cut};
    k_resolve_root<<<1, 1, 0, stream>>>(nodes, up.nf, in.root, queue);
    int32_t lo = 0, hi = 1, next_slot = 0;
    while (hi > lo) {
      max_depth++;
      const int32_t cntl = hi - lo;
      k_wide_a<<<grid_for(cntl, 64), 64, 0, stream>>>(lo, hi, queue, tv, qn, kids, cis, cnt);
      size_t tb = tmp_bytes;
      FCU(cub::DeviceScan::ExclusiveSum(d_tmp, tb, cnt, scan, cntl, stream));
      k_wide_b<<<grid_for(cntl, 128), 128, 0, stream>>>(lo, hi, next_slot, scan, kids, cis, nodes, qn, queue, out.slot_prim, total, cnt);
      FCU(cudaMemcpyAsync(h_total, total, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
      FCU(cudaStreamSynchronize(stream));
      const int32_t n_inner = (int32_t)(*h_total >> 32), n_leaf = (int32_t)(*h_total & 0xFFFFFFFFull);
      if ((int64_t)hi + n_inner > nw_max) { err = "internal: wide-node count exceeds its bound"; return RTC_ERR_INVALID; }
      next_slot += n_leaf;
      lo = hi;
      hi += n_inner;
    }
    n_wide = hi;
    if (next_slot != n_bounded) { err = "internal: wide tree does not hold every bounded primitive"; return RTC_ERR_INVALID; }
  }
  if (n_wide > 0) {  // the wide nodes into the caller's buffer of their own size
    int rcq = in.alloc_qnodes(in.alloc_ctx, (size_t)n_wide * sizeof(CNode), (void**)&out.qnodes);
    if (rcq) { err = "device allocation for the wide nodes failed"; return rcq; }
    FCU(cudaMemcpyAsync(out.qnodes, qn, (size_t)n_wide * sizeof(CNode), cudaMemcpyDeviceToDevice, stream));
  }
  out.n_qnodes = n_wide;
  out.depth = max_depth;
  out.n_bounded = n_bounded;
  out.unbounded_prims.clear();
  for (auto& u : unb) out.unbounded_prims.push_back(u.second);
  if (n_unb > 0)
    FCU(cudaMemcpyAsync(out.slot_prim + n_bounded, out.unbounded_prims.data(), (size_t)n_unb * sizeof(int32_t), cudaMemcpyHostToDevice, stream));

  // ---- records ---------------------------------------------------------------------------------------------------------------
  if (in.records_ready) FCU(cudaStreamWaitEvent(stream, in.records_ready, 0));
  RecordArrays r{in.staged, out.slot_prim, (DPrim<float>*)out.prims, (DMat<float>*)out.mats, out.aux, out.prim_id, out.id_to_slot,
                 (V4<float>*)out.sgeom};
  k_records<<<grid_for(n, 128), 128, 0, stream>>>(n, r);
  FCU(cudaGetLastError());
  FCU(cudaStreamSynchronize(stream));  // (work arrays are released on return)
  return RTC_OK;
#undef FCU
}

// With lazy module loading the first launch of every kernel pays for loading it (a millisecond or so each, ~30 kernels here and in
// CUB): that is a cost of starting the process, not of Scene.Prepare, so rtc_create asks for the kernels' attributes once per
// device, which loads them.
void prepare_device_preload() {
  cudaFuncAttributes a;
  const void* fns[] = {(const void*)k_sb_init,      (const void*)k_seg_reset,  (const void*)k_seg_bounds,  (const void*)k_seg_prepare,
                       (const void*)k_bins_init,    (const void*)k_bin,        (const void*)k_sah,         (const void*)k_level_total,
                       (const void*)k_sortkey,      (const void*)k_side,       (const void*)k_scatter,     (const void*)k_children,
                       (const void*)k_union_level,  (const void*)k_single_leaf, (const void*)k_gather_boxes, (const void*)k_bfs,
                       (const void*)k_up,           (const void*)k_rank,       (const void*)k_unb_fetch,   (const void*)k_resolve_root,
                       (const void*)k_wide_a,       (const void*)k_wide_b,     (const void*)k_records};
  for (const void* f : fns) cudaFuncGetAttributes(&a, f);
  // the three scans the passes use (u32, i32, u64), on one element each
  void* buf = nullptr;
  if (cudaMalloc(&buf, 4096) == cudaSuccess) {
    char* b = (char*)buf;
    size_t tb = 2048;
    cub::DeviceScan::ExclusiveSum(b + 2048, tb, (uint32_t*)b, (uint32_t*)(b + 64), 1);
    tb = 2048;
    cub::DeviceScan::ExclusiveSum(b + 2048, tb, (int32_t*)b, (int32_t*)(b + 64), 1);
    tb = 2048;
    cub::DeviceScan::ExclusiveSum(b + 2048, tb, (unsigned long long*)b, (unsigned long long*)(b + 64), 1);
    cudaDeviceSynchronize();
    cudaFree(buf);
  }
  cudaGetLastError();
}

}  // namespace rtc
