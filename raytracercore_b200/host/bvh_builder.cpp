// bvh_builder.cpp — replacement for the reference's BVH.Construct (Raytracing/Acceleration/BVH.cs:50-236).
//
// The reference builds agglomeratively (brute force for N<=20, heap + k-d tree up to 200k, locally-ordered
// clustering above) at a cost that makes the 1M/10M-primitive scenes impractical, and its tie-breaking depends
// on runtime object hash codes, so its topology is not reproducible even run to run. This builder keeps the
// reference's *shape* — a binary tree, exactly one primitive per leaf (BVH.cs:256-264), inner Volume = the
// Combine of the children's boxes (AABB.cs:38-43) — and chooses the topology with a top-down 16-bin SAH split.
// Closest-hit results do not depend on topology except on exact-distance ties.
//
// Layout: a subtree over k primitives owns exactly 2k-1 consecutive nodes, root first, left subtree next. That
// makes node allocation position-deterministic, so subtrees can be built by independent threads.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <thread>

#include "scene.h"

namespace rtcore {
namespace {

constexpr int kBins = 16;

struct Builder {
  const double* lo;
  const double* hi;
  std::vector<int32_t> idx;
  std::vector<float> cen;  // centroids, n*3 (f32 is enough to bin)
  rtc_bvh_node* nodes;
  int max_spawn_depth;

  static void leaf_box(rtc_bvh_node& nd, const double* l, const double* h) {
    for (int k = 0; k < 3; k++) {
      nd.bmin[k] = l[k];
      nd.bmax[k] = h[k];
    }
  }

  void build(int32_t base, int32_t b, int32_t e, int depth) {
    rtc_bvh_node& nd = nodes[base];
    nd.pad = 0;
    int32_t k = e - b;
    if (k == 1) {
      int32_t p = idx[b];
      leaf_box(nd, lo + (size_t)p * 3, hi + (size_t)p * 3);
      nd.left = nd.right = -1;
      nd.prim = p;
      return;
    }
    // centroid bounds
    float cmin[3] = {INFINITY, INFINITY, INFINITY}, cmax[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int32_t i = b; i < e; i++) {
      const float* c = &cen[(size_t)idx[i] * 3];
      for (int a = 0; a < 3; a++) {
        cmin[a] = std::min(cmin[a], c[a]);
        cmax[a] = std::max(cmax[a], c[a]);
      }
    }
    int32_t mid = -1;
    int best_axis = -1, best_bin = -1;
    double best_cost = std::numeric_limits<double>::infinity();
    float scale[3];
    for (int a = 0; a < 3; a++) {
      float ext = cmax[a] - cmin[a];
      scale[a] = ext > 0 ? (kBins * (1.0f - 1e-6f)) / ext : 0.0f;
    }
    if (k > 2) {
      struct Bin {
        double lo[3], hi[3];
        int32_t n;
      };
      Bin bins[3][kBins];
      for (int a = 0; a < 3; a++)
        for (int j = 0; j < kBins; j++) {
          bins[a][j].n = 0;
          for (int c = 0; c < 3; c++) {
            bins[a][j].lo[c] = std::numeric_limits<double>::infinity();
            bins[a][j].hi[c] = -std::numeric_limits<double>::infinity();
          }
        }
      for (int32_t i = b; i < e; i++) {
        int32_t p = idx[i];
        const float* c = &cen[(size_t)p * 3];
        const double* l = lo + (size_t)p * 3;
        const double* h = hi + (size_t)p * 3;
        for (int a = 0; a < 3; a++) {
          if (scale[a] == 0) continue;
          int j = (int)((c[a] - cmin[a]) * scale[a]);
          j = std::min(std::max(j, 0), kBins - 1);
          Bin& bn = bins[a][j];
          bn.n++;
          for (int c2 = 0; c2 < 3; c2++) {
            bn.lo[c2] = std::min(bn.lo[c2], l[c2]);
            bn.hi[c2] = std::max(bn.hi[c2], h[c2]);
          }
        }
      }
      auto area = [](const double* l, const double* h) {
        double dx = h[0] - l[0], dy = h[1] - l[1], dz = h[2] - l[2];
        return (dx * dy + dy * dz + dz * dx) * 2;  // AABB.GetSurfaceArea, AABB.cs:204-207
      };
      for (int a = 0; a < 3; a++) {
        if (scale[a] == 0) continue;
        double rl[kBins][3], rh[kBins][3];
        int32_t rn[kBins];
        double l3[3] = {INFINITY, INFINITY, INFINITY}, h3[3] = {-INFINITY, -INFINITY, -INFINITY};
        int32_t cnt = 0;
        for (int j = kBins - 1; j >= 1; j--) {
          const Bin& bn = bins[a][j];
          cnt += bn.n;
          for (int c = 0; c < 3; c++) {
            l3[c] = std::min(l3[c], bn.lo[c]);
            h3[c] = std::max(h3[c], bn.hi[c]);
            rl[j][c] = l3[c];
            rh[j][c] = h3[c];
          }
          rn[j] = cnt;
        }
        double ll[3] = {INFINITY, INFINITY, INFINITY}, lh[3] = {-INFINITY, -INFINITY, -INFINITY};
        int32_t ln = 0;
        for (int j = 0; j < kBins - 1; j++) {
          const Bin& bn = bins[a][j];
          ln += bn.n;
          for (int c = 0; c < 3; c++) {
            ll[c] = std::min(ll[c], bn.lo[c]);
            lh[c] = std::max(lh[c], bn.hi[c]);
          }
          if (ln == 0 || rn[j + 1] == 0) continue;
          double cost = area(ll, lh) * ln + area(rl[j + 1], rh[j + 1]) * rn[j + 1];
          if (cost < best_cost) {
            best_cost = cost;
            best_axis = a;
            best_bin = j;
          }
        }
      }
    }
    if (best_axis >= 0) {
      int a = best_axis;
      float cm = cmin[a], sc = scale[a];
      auto it = std::partition(idx.begin() + b, idx.begin() + e, [&](int32_t p) {
        int j = (int)((cen[(size_t)p * 3 + a] - cm) * sc);
        j = std::min(std::max(j, 0), kBins - 1);
        return j <= best_bin;
      });
      mid = (int32_t)(it - idx.begin());
    }
    if (mid <= b || mid >= e) {
      // no usable SAH split (k == 2, or coincident centroids): median split along the widest centroid axis
      int a = 0;
      if (cmax[1] - cmin[1] > cmax[a] - cmin[a]) a = 1;
      if (cmax[2] - cmin[2] > cmax[a] - cmin[a]) a = 2;
      mid = b + k / 2;
      std::nth_element(idx.begin() + b, idx.begin() + mid, idx.begin() + e, [&](int32_t p, int32_t q) {
        float cp = cen[(size_t)p * 3 + a], cq = cen[(size_t)q * 3 + a];
        return cp < cq || (cp == cq && p < q);
      });
    }
    int32_t kl = mid - b;
    int32_t left = base + 1, right = base + 2 * kl;
    if (depth < max_spawn_depth && k > 32768) {
      std::thread t([&, left, b, mid, depth]() { build(left, b, mid, depth + 1); });
      build(right, mid, e, depth + 1);
      t.join();
    } else {
      build(left, b, mid, depth + 1);
      build(right, mid, e, depth + 1);
    }
    nd.left = left;
    nd.right = right;
    nd.prim = -1;
    for (int c = 0; c < 3; c++) {  // AABB.Combine, AABB.cs:38-43
      nd.bmin[c] = std::fmin(nodes[left].bmin[c], nodes[right].bmin[c]);
      nd.bmax[c] = std::fmax(nodes[left].bmax[c], nodes[right].bmax[c]);
    }
  }
};

}  // namespace

int BuildBVH(int32_t n, const double* bmin, const double* bmax, std::vector<rtc_bvh_node>& nodes, int threads) {
  nodes.clear();
  if (n <= 0) return -1;
  Builder bld;
  bld.lo = bmin;
  bld.hi = bmax;
  std::vector<int32_t> unbounded;
  bld.idx.reserve(n);
  bld.cen.assign((size_t)n * 3, 0.0f);
  for (int32_t i = 0; i < n; i++) {
    bool finite = true;
    for (int k = 0; k < 3; k++) finite = finite && std::isfinite(bmin[(size_t)i * 3 + k]) && std::isfinite(bmax[(size_t)i * 3 + k]);
    if (!finite) {
      unbounded.push_back(i);
      continue;
    }
    bld.idx.push_back(i);
    for (int k = 0; k < 3; k++)
      bld.cen[(size_t)i * 3 + k] = (float)((bmin[(size_t)i * 3 + k] + bmax[(size_t)i * 3 + k]) * 0.5);
  }
  int32_t m = (int32_t)bld.idx.size();
  size_t total = (m > 0 ? (size_t)2 * m - 1 : 0) + (m > 0 ? 2 * unbounded.size() : (unbounded.empty() ? 0 : 2 * unbounded.size() - 1));
  nodes.resize(total);
  bld.nodes = nodes.data();
  int d = 0;
  while ((1 << d) < std::max(1, threads)) d++;
  bld.max_spawn_depth = d + 1;
  int32_t root = -1;
  int32_t next = 0;
  if (m > 0) {
    bld.build(0, 0, m, 0);
    root = 0;
    next = 2 * m - 1;
  }
  // Unbounded primitives (planes, Plane.cs:68-74) become leaves chained above the root, first plane outermost-left
  // so the left-first leaf order (BVH.cs:314-315) lists planes in ID order before everything else.
  for (size_t j = unbounded.size(); j-- > 0;) {
    int32_t p = unbounded[j];
    rtc_bvh_node& leaf = nodes[next];
    std::memset(&leaf, 0, sizeof(leaf));
    for (int k = 0; k < 3; k++) {
      leaf.bmin[k] = bmin[(size_t)p * 3 + k];
      leaf.bmax[k] = bmax[(size_t)p * 3 + k];
    }
    leaf.left = leaf.right = -1;
    leaf.prim = p;
    if (root < 0) {
      root = next++;
      continue;
    }
    rtc_bvh_node& par = nodes[next + 1];
    std::memset(&par, 0, sizeof(par));
    par.left = next;
    par.right = root;
    par.prim = -1;
    for (int k = 0; k < 3; k++) {
      par.bmin[k] = std::fmin(leaf.bmin[k], nodes[root].bmin[k]);
      par.bmax[k] = std::fmax(leaf.bmax[k], nodes[root].bmax[k]);
    }
    root = next + 1;
    next += 2;
  }
  nodes.resize(next);
  return root;
}

}  // namespace rtcore
