// bvh_build.cu — BVH.Construct on the device.
//
// The reference builds its tree by agglomerative clustering with the surface area of the merged box as the distance
// (Acceleration/BVH.cs:12-21): a heap of nearest pairs for N <= 200 000 (:89-191) and, above that, the locally-ordered
// variant of Walter et al. 2008 driven by k-d tree nearest-neighbour queries (:50-87). Both are strictly sequential —
// one merge at a time, each followed by k-d tree removals and insertions — and are the wall-clock blocker of
// Scene.Prepare for scenes of 10^6 primitives and more.
//
// This is the same clustering made data-parallel (PLOC, Meister & Bittner 2018): primitives are ordered along a Morton
// curve; every round each cluster looks `radius` positions to either side for the neighbour whose union with it has the
// smallest surface area, clusters that chose each other merge into a new inner node, and the survivors are compacted
// for the next round. One primitive per leaf, like the reference (BVH.cs:256-264). Inner boxes are exact f64 unions of
// the leaf boxes, so the result is a valid input for rtc_upload_bvh's flattening in either arithmetic mode.
//
// Library use: cub::DeviceRadixSort for the 64-bit Morton keys and cub::DeviceScan for the compaction offsets (CCCL as
// shipped with the toolkit); the clustering kernels are hand-written.
#include <cub/cub.cuh>
#include <math_constants.h>

#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <vector>

#include "rtc_internal.h"

namespace rtc {
namespace {

struct Box6 {
  double lo[3], hi[3];
};

__device__ __forceinline__ uint64_t spread21(uint64_t v) {  // 21 bits -> every third bit
  v &= 0x1FFFFFull;
  v = (v | (v << 32)) & 0x1F00000000FFFFull;
  v = (v | (v << 16)) & 0x1F0000FF0000FFull;
  v = (v | (v << 8)) & 0x100F00F00F00F00Full;
  v = (v | (v << 4)) & 0x10C30C30C30C30C3ull;
  v = (v | (v << 2)) & 0x1249249249249249ull;
  return v;
}

__global__ void k_morton(int32_t m, const Box6* boxes, double cx, double cy, double cz, double sx, double sy, double sz,
                         uint64_t* keys, uint32_t* vals) {
  int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const Box6& b = boxes[i];
  // centroid on a 2^21 grid over the centroid bounds
  double x = ((b.lo[0] + b.hi[0]) * 0.5 - cx) * sx, y = ((b.lo[1] + b.hi[1]) * 0.5 - cy) * sy, z = ((b.lo[2] + b.hi[2]) * 0.5 - cz) * sz;
  uint64_t qx = (uint64_t)fmin(fmax(x, 0.0), 2097151.0), qy = (uint64_t)fmin(fmax(y, 0.0), 2097151.0),
           qz = (uint64_t)fmin(fmax(z, 0.0), 2097151.0);
  keys[i] = (spread21(qx) << 2) | (spread21(qy) << 1) | spread21(qz);
  vals[i] = (uint32_t)i;
}

// leaves in Morton order: node i = the primitive at sorted position i
__global__ void k_leaves(int32_t m, const uint32_t* order, const Box6* boxes, const int32_t* prim_ids, rtc_bvh_node* nodes,
                         int32_t* cluster) {
  int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const uint32_t k = order[i];
  rtc_bvh_node nd;
  for (int a = 0; a < 3; a++) {
    nd.bmin[a] = boxes[k].lo[a];
    nd.bmax[a] = boxes[k].hi[a];
  }
  nd.left = nd.right = -1;
  nd.prim = prim_ids[k];
  nd.pad = 0;
  nodes[i] = nd;
  cluster[i] = i;
}

// the reference's merge cost: surface area of the union (BVH.cs:12-21, AABB.GetSurfaceArea AABB.cs:204-207)
__device__ __forceinline__ float union_area(const float* a, const float* b) {
  const float dx = fmaxf(a[3], b[3]) - fminf(a[0], b[0]), dy = fmaxf(a[4], b[4]) - fminf(a[1], b[1]),
              dz = fmaxf(a[5], b[5]) - fminf(a[2], b[2]);
  return dx * dy + dy * dz + dz * dx;
}

constexpr int kNNThreads = 128;
constexpr int kMaxRadius = 32;

// Nearest neighbour within `radius` positions. Candidate pairs are ordered by (area, not-buddy, min index, max index) with
// buddy = positions differing only in their lowest bit: a strict, symmetric total order on pairs, so the globally smallest
// pair is always mutual (progress every round), and runs of identical boxes pair up (i, i^1) instead of chaining.
__global__ void __launch_bounds__(kNNThreads) k_nearest(int32_t n, int radius, const int32_t* cluster, const rtc_bvh_node* nodes,
                                                        int32_t* nearest) {
  __shared__ float s_box[kNNThreads + 2 * kMaxRadius][6];
  const int32_t first = blockIdx.x * kNNThreads - radius;
  for (int t = threadIdx.x; t < kNNThreads + 2 * radius; t += kNNThreads) {
    const int32_t j = first + t;
    if (j >= 0 && j < n) {
      const rtc_bvh_node& nd = nodes[cluster[j]];
      // outward-rounded floats: the clustering only compares areas
      for (int a = 0; a < 3; a++) {
        s_box[t][a] = __double2float_rd(nd.bmin[a]);
        s_box[t][3 + a] = __double2float_ru(nd.bmax[a]);
      }
    }
  }
  __syncthreads();
  const int32_t i = blockIdx.x * kNNThreads + threadIdx.x;
  if (i >= n) return;
  const float* me = s_box[threadIdx.x + radius];
  float best = CUDART_INF_F;
  int32_t best_j = -1;
  bool best_buddy = false;
  for (int dlt = -radius; dlt <= radius; dlt++) {
    const int32_t j = i + dlt;
    if (dlt == 0 || j < 0 || j >= n) continue;
    const float a = union_area(me, s_box[threadIdx.x + radius + dlt]);
    const bool buddy = (i ^ j) == 1;
    // (area, !buddy, min(i,j), max(i,j)) ascending; j ascends through the window, so among equal (area, buddy) the pair with
    // the smaller min index and then the smaller max index is the first one met for j < i ... and for j > i min = i is
    // fixed and max = j ascends: the first met wins in both halves, and a j < i always beats a j > i (min(i,j) = j < i).
    const bool better = (a < best) || (a == best && buddy && !best_buddy);
    if (best_j < 0 || better) {
      best = a;
      best_j = j;
      best_buddy = buddy;
    }
  }
  nearest[i] = best_j;
}

__global__ void k_merge(int32_t n, const int32_t* cluster, const int32_t* nearest, rtc_bvh_node* nodes, int32_t* next_node,
                        int32_t* merged, uint32_t* keep) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t j = nearest[i];
  if (j >= 0 && nearest[j] == i) {
    if (i < j) {
      const int32_t id = atomicAdd(next_node, 1);
      const int32_t l = cluster[i], r = cluster[j];
      rtc_bvh_node nd;
      for (int a = 0; a < 3; a++) {
        nd.bmin[a] = fmin(nodes[l].bmin[a], nodes[r].bmin[a]);
        nd.bmax[a] = fmax(nodes[l].bmax[a], nodes[r].bmax[a]);
      }
      nd.left = l;
      nd.right = r;
      nd.prim = -1;
      nd.pad = 0;
      nodes[id] = nd;
      merged[i] = id;
      keep[i] = 1;
    } else {
      keep[i] = 0;
    }
  } else {
    merged[i] = cluster[i];
    keep[i] = 1;
  }
}

__global__ void k_compact_clusters(int32_t n, const int32_t* merged, const uint32_t* keep, const uint32_t* pos, int32_t* cluster_out,
                                   int32_t* n_out) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (keep[i]) cluster_out[pos[i]] = merged[i];
  if (i == n - 1) *n_out = (int32_t)(pos[i] + keep[i]);
}

}  // namespace

#define BCU(call)                        \
  do {                                   \
    cudaError_t e__ = (call);            \
    if (e__ != cudaSuccess) {            \
      release();                         \
      return e__;                        \
    }                                    \
  } while (0)

// Builds the tree over m bounded primitives. boxes: m x (lo[3], hi[3]) f64, prim_ids: m primitive IDs. nodes_out receives
// 2m-1 nodes (leaves first, in Morton order), *root_out the root index, *rounds_out the number of clustering rounds.
cudaError_t build_bvh_ploc(cudaStream_t stream, int32_t m, const double* boxes, const int32_t* prim_ids, int radius,
                           rtc_bvh_node* nodes_out, int32_t* root_out, int32_t* rounds_out) {
  radius = radius < 1 ? 1 : (radius > kMaxRadius ? kMaxRadius : radius);
  const bool verbose = std::getenv("RTC_B200_VERBOSE") != nullptr;
  const auto t_start = std::chrono::steady_clock::now();
  auto since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count(); };
  Box6* d_boxes = nullptr;
  int32_t *d_prim = nullptr, *d_cl[2] = {nullptr, nullptr}, *d_nn = nullptr, *d_merged = nullptr, *d_counters = nullptr;
  uint64_t *d_keys = nullptr, *d_keys2 = nullptr;
  uint32_t *d_vals = nullptr, *d_vals2 = nullptr, *d_keep = nullptr, *d_pos = nullptr;
  rtc_bvh_node* d_nodes = nullptr;
  void* d_tmp = nullptr;
  auto release = [&]() {
    cudaFree(d_boxes); cudaFree(d_prim); cudaFree(d_cl[0]); cudaFree(d_cl[1]); cudaFree(d_nn); cudaFree(d_merged);
    cudaFree(d_counters); cudaFree(d_keys); cudaFree(d_keys2); cudaFree(d_vals); cudaFree(d_vals2); cudaFree(d_keep);
    cudaFree(d_pos); cudaFree(d_nodes); cudaFree(d_tmp);
  };
  *rounds_out = 0;
  if (m == 1) {
    rtc_bvh_node nd;
    for (int a = 0; a < 3; a++) {
      nd.bmin[a] = boxes[a];
      nd.bmax[a] = boxes[3 + a];
    }
    nd.left = nd.right = -1;
    nd.prim = prim_ids[0];
    nd.pad = 0;
    nodes_out[0] = nd;
    *root_out = 0;
    return cudaSuccess;
  }
  // centroid bounds (host: the boxes were just produced there)
  double cmin[3] = {1e300, 1e300, 1e300}, cmax[3] = {-1e300, -1e300, -1e300};
  for (int32_t i = 0; i < m; i++)
    for (int a = 0; a < 3; a++) {
      const double c = (boxes[(size_t)i * 6 + a] + boxes[(size_t)i * 6 + 3 + a]) * 0.5;
      cmin[a] = c < cmin[a] ? c : cmin[a];
      cmax[a] = c > cmax[a] ? c : cmax[a];
    }
  double scale[3];
  for (int a = 0; a < 3; a++) scale[a] = cmax[a] > cmin[a] ? 2097151.0 / (cmax[a] - cmin[a]) : 0.0;

  const size_t n_nodes = (size_t)2 * m - 1;
  BCU(cudaMalloc((void**)&d_boxes, (size_t)m * sizeof(Box6)));
  BCU(cudaMalloc((void**)&d_prim, (size_t)m * sizeof(int32_t)));
  BCU(cudaMalloc((void**)&d_cl[0], (size_t)m * sizeof(int32_t)));
  BCU(cudaMalloc((void**)&d_cl[1], (size_t)m * sizeof(int32_t)));
  BCU(cudaMalloc((void**)&d_nn, (size_t)m * sizeof(int32_t)));
  BCU(cudaMalloc((void**)&d_merged, (size_t)m * sizeof(int32_t)));
  BCU(cudaMalloc((void**)&d_counters, 2 * sizeof(int32_t)));
  BCU(cudaMalloc((void**)&d_keys, (size_t)m * sizeof(uint64_t)));
  BCU(cudaMalloc((void**)&d_keys2, (size_t)m * sizeof(uint64_t)));
  BCU(cudaMalloc((void**)&d_vals, (size_t)m * sizeof(uint32_t)));
  BCU(cudaMalloc((void**)&d_vals2, (size_t)m * sizeof(uint32_t)));
  BCU(cudaMalloc((void**)&d_keep, (size_t)m * sizeof(uint32_t)));
  BCU(cudaMalloc((void**)&d_pos, (size_t)m * sizeof(uint32_t)));
  BCU(cudaMalloc((void**)&d_nodes, n_nodes * sizeof(rtc_bvh_node)));
  size_t tmp_sort = 0, tmp_scan = 0;
  BCU(cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, d_keys, d_keys2, d_vals, d_vals2, m, 0, 63, stream));
  BCU(cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan, d_keep, d_pos, m, stream));
  const size_t tmp_bytes = tmp_sort > tmp_scan ? tmp_sort : tmp_scan;
  BCU(cudaMalloc(&d_tmp, tmp_bytes));

  BCU(cudaMemcpyAsync(d_boxes, boxes, (size_t)m * sizeof(Box6), cudaMemcpyHostToDevice, stream));
  BCU(cudaMemcpyAsync(d_prim, prim_ids, (size_t)m * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
  const int T = 256;
  const int G = (m + T - 1) / T;
  k_morton<<<G, T, 0, stream>>>(m, d_boxes, cmin[0], cmin[1], cmin[2], scale[0], scale[1], scale[2], d_keys, d_vals);
  size_t tb = tmp_bytes;
  BCU(cub::DeviceRadixSort::SortPairs(d_tmp, tb, d_keys, d_keys2, d_vals, d_vals2, m, 0, 63, stream));
  k_leaves<<<G, T, 0, stream>>>(m, d_vals2, d_boxes, d_prim, d_nodes, d_cl[0]);
  int32_t init[2] = {m, m};  // next free node, active clusters
  BCU(cudaMemcpyAsync(d_counters, init, sizeof(init), cudaMemcpyHostToDevice, stream));

  double t_sorted = 0;
  if (verbose) {
    cudaStreamSynchronize(stream);
    t_sorted = since();
  }
  int32_t n = m, cur = 0, rounds = 0;
  while (n > 1) {
    if (++rounds > 100000) {
      release();
      return cudaErrorUnknown;  // cannot happen: every round merges at least the globally smallest pair
    }
    k_nearest<<<(n + kNNThreads - 1) / kNNThreads, kNNThreads, 0, stream>>>(n, radius, d_cl[cur], d_nodes, d_nn);
    const int g = (n + T - 1) / T;
    k_merge<<<g, T, 0, stream>>>(n, d_cl[cur], d_nn, d_nodes, d_counters, d_merged, d_keep);
    tb = tmp_bytes;
    BCU(cub::DeviceScan::ExclusiveSum(d_tmp, tb, d_keep, d_pos, n, stream));
    k_compact_clusters<<<g, T, 0, stream>>>(n, d_merged, d_keep, d_pos, d_cl[cur ^ 1], d_counters + 1);
    BCU(cudaMemcpyAsync(&n, d_counters + 1, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    BCU(cudaStreamSynchronize(stream));
    cur ^= 1;
  }
  const double t_clustered = since();
  int32_t root = -1;
  BCU(cudaMemcpyAsync(&root, d_cl[cur], sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  BCU(cudaMemcpyAsync(nodes_out, d_nodes, n_nodes * sizeof(rtc_bvh_node), cudaMemcpyDeviceToHost, stream));
  BCU(cudaStreamSynchronize(stream));
  BCU(cudaGetLastError());
  *root_out = root;
  *rounds_out = rounds;
  release();
  if (verbose)
    std::fprintf(stderr, "[rtcore_b200] device BVH build, %d primitives, radius %d: upload + Morton sort %.1f ms, %d clustering rounds %.1f ms, "
                         "tree read-back %.1f ms\n", m, radius, t_sorted, rounds, t_clustered - t_sorted, since() - t_clustered);
  return cudaSuccess;
}

}  // namespace rtc
