// prepare_common.h — the arithmetic of Scene.Prepare's second half for the f32 mode: collapsing the reference-shaped binary
// tree (one primitive per leaf, BVH.cs:239-264) into the quantised 8-wide device tree. Every function here is
// __host__ __device__ and is the ONE statement of its step: the host flatten in rtc_api.cu (f64 mode's neighbour, also the
// checker of the device path in tests) and the per-level kernels in prepare_device.cu both call it, so the two produce the same
// image bit for bit. Both translation units are compiled without FMA contraction of these expressions (plain IEEE double
// and float operations; the kernels' unit is built with -fmad=false, the host side targets baseline x86-64).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "rtc_internal.h"

#ifdef __CUDACC__
#define RTC_HD __host__ __device__ inline
#else
#define RTC_HD inline
#endif

namespace rtc {
namespace prep {

// What the collapse reads per binary node, as raw arrays (host: std::vector storage, device: global memory).
struct TreeView {
  const rtc_bvh_node* nodes;
  const int32_t* nf;     // bounded (finite-box) leaves below the node
  const double* fmin;    // [node][3] union of the finite leaf boxes below the node
  const double* fmax;
  const float* T;        // [node][8] collapse cost table (see dp_node)
  const uint8_t* cut;    // [node][9] left side's share of j slots
};

RTC_HD double pos_inf() { return HUGE_VAL; }

RTC_HD bool finite_box(const rtc_bvh_node& nd) {
  bool fin = true;
  for (int a = 0; a < 3; a++) fin = fin && isfinite(nd.bmin[a]) && isfinite(nd.bmax[a]);
  return fin;
}

RTC_HD double box_area(const double* lo, const double* hi) {
  const double dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
  const double a = (dx * dy + dy * dz + dz * dx) * 2;
  return isfinite(a) ? a : 1.7976931348623157e308;
}

// f64 -> f32 toward -inf, then one more ulp: the f32 slab test evaluates fma(bound, 1/d, -o/d), whose rounding is equivalent
// to moving the bound by up to an ulp of the coordinate magnitude
RTC_HD float round_down_f32(double x) {
  float f = (float)x;
  if ((double)f > x) f = nextafterf(f, -HUGE_VALF);
  return nextafterf(f, -HUGE_VALF);
}

// a binary node with bounded leaves on one side only is transparent
RTC_HD int32_t resolve(const rtc_bvh_node* nodes, const int32_t* nf, int32_t i) {
  while (nodes[i].prim < 0) {
    const int32_t l = nodes[i].left, r = nodes[i].right;
    if (nf[l] == 0) i = r;
    else if (nf[r] == 0) i = l;
    else break;
  }
  return i;
}

// Optimal collapse (after Ylitie, Karras & Laine 2017, sec. 3): T(m, i) = least total surface area of the wide nodes needed
// below binary node m when m's subtree may occupy at most i child slots of its parent wide node. A leaf costs nothing; an
// inner node either becomes a wide node itself (its area + its two sides spread over 8 slots) or hands its slots on to its
// two sides. Evaluated children-first; cut[j] is the left side's share when j slots are split. Tl / Tr: the tables of the
// (resolved) left and right child, index 0 unused.
#ifndef RTC_COLLAPSE_WIDTH
#define RTC_COLLAPSE_WIDTH 8  // children per wide node the collapse aims for (tuning experiments: narrower trees in the 8-slot node)
#endif
RTC_HD void dp_node(const float* Tl, const float* Tr, double area, float* Ti, uint8_t* cut_i) {
  float D[9];
  for (int j = 2; j <= 8; j++) {
    float bestv = HUGE_VALF;
    int bestk = 1;
    for (int k = 1; k < j; k++) {
      const int kl = k < 7 ? k : 7, kr = (j - k) < 7 ? (j - k) : 7;
      const float v = Tl[kl] + Tr[kr];
      if (v < bestv) {
        bestv = v;
        bestk = k;
      }
    }
    D[j] = bestv;
    cut_i[j] = (uint8_t)bestk;
  }
  const float as_node = (float)area + D[RTC_COLLAPSE_WIDTH];
  Ti[1] = as_node;
  for (int j = 2; j <= 7; j++) Ti[j] = as_node < D[j] ? as_node : D[j];
}

// fills kids[] with the children (binary nodes) of the wide node rooted at binary node m; returns their number (<= 8)
RTC_HD int gather_children(const TreeView& t, int32_t m, int32_t* kids) {
  struct It {
    int32_t node;
    int8_t slots, force_split;
  };
  It st[32];
  int sp = 0, nk = 0;
  st[sp].node = m;
  st[sp].slots = RTC_COLLAPSE_WIDTH;
  st[sp].force_split = 1;
  sp++;
  while (sp > 0) {
    const It it = st[--sp];
    const rtc_bvh_node& nd = t.nodes[it.node];
    if (nd.prim >= 0) {
      kids[nk++] = it.node;
      continue;
    }
    const float* Ti = t.T + (size_t)it.node * 8;
    const int sl = it.slots < 7 ? it.slots : 7;
    if (!it.force_split && (it.slots == 1 || Ti[sl] >= Ti[1])) {
      kids[nk++] = it.node;  // stays a wide node of its own
      continue;
    }
    const int k = t.cut[(size_t)it.node * 9 + it.slots];
    st[sp].node = resolve(t.nodes, t.nf, nd.right);
    st[sp].slots = (int8_t)(it.slots - k);
    st[sp].force_split = 0;
    sp++;
    st[sp].node = resolve(t.nodes, t.nf, nd.left);
    st[sp].slots = (int8_t)k;
    st[sp].force_split = 0;
    sp++;
  }
  return nk;
}

// smallest e with 2^e * `steps` >= ext (ext > 0), exactly (no libm logarithm: host and device must agree)
RTC_HD int grid_exponent(double ext, double steps) {
  int k;
  const double m = frexp(ext / steps, &k);  // ext / steps = m 2^k, m in [0.5, 1)
  int e = (m == 0.5) ? k - 1 : k;
  while (ldexp(1.0, e) * steps < ext) e++;
  return e;
}

// The wide node over kids[0..nk): box, per-axis power-of-two grid (origin one step below the box minimum, so the one-step
// padding of the children never clamps at 0; step = smallest power of two that spans the box in 250 steps), octant slot
// assignment (child i goes to the free slot whose octant signs best match its offset from the centre, greedily by the
// largest match), quantised child bounds padded by one step, masks. child_base / prim_base are left 0.
RTC_HD void make_cnode(const TreeView& t, const int32_t* kids, int nk, CNode& cn, int8_t* child_in_slot) {
  double lo[3], hi[3];
  for (int a = 0; a < 3; a++) {
    lo[a] = pos_inf();
    hi[a] = -pos_inf();
    for (int c = 0; c < nk; c++) {
      const double l = t.fmin[(size_t)kids[c] * 3 + a], h = t.fmax[(size_t)kids[c] * 3 + a];
      lo[a] = l < lo[a] ? l : lo[a];
      hi[a] = h > hi[a] ? h : hi[a];
    }
  }
  memset(&cn, 0, sizeof(cn));
  float p[3];
  int ex[3];
  double step[3];
  for (int a = 0; a < 3; a++) {
    const double ext = hi[a] - lo[a];
    int e = ext > 0 ? grid_exponent(ext, 250.0) : -100;
    e = e < -120 ? -120 : (e > 120 ? 120 : e);
    while (ldexp(1.0, e) * 250.0 < ext) e++;
    p[a] = round_down_f32(lo[a] - ldexp(1.0, e));
    while (ldexp(1.0, e) * 253.0 < hi[a] - (double)p[a]) {
      e++;
      p[a] = round_down_f32(lo[a] - ldexp(1.0, e));
    }
    ex[a] = e;
    step[a] = ldexp(1.0, e);
  }
  cn.px = p[0];
  cn.py = p[1];
  cn.pz = p[2];
  int slot_of[8];
  {
    double ctr[3];
    for (int a = 0; a < 3; a++) ctr[a] = 0.5 * (lo[a] + hi[a]);
    double cost[8][8];
    for (int c = 0; c < nk; c++)
      for (int s = 0; s < 8; s++) {
        double v = 0;
        for (int a = 0; a < 3; a++) {
          const double off = 0.5 * (t.fmin[(size_t)kids[c] * 3 + a] + t.fmax[(size_t)kids[c] * 3 + a]) - ctr[a];
          v += ((s >> a) & 1) ? off : -off;
        }
        cost[c][s] = v;
      }
    uint32_t cu = 0, su = 0;
    for (int it = 0; it < nk; it++) {
      int bc = -1, bs = -1;
      double bv = -pos_inf();
      for (int c = 0; c < nk; c++)
        if (!((cu >> c) & 1u))
          for (int s = 0; s < 8; s++)
            if (!((su >> s) & 1u) && cost[c][s] > bv) {
              bv = cost[c][s];
              bc = c;
              bs = s;
            }
      if (bc < 0) {  // every remaining cost is -inf or NaN (degenerate boxes): first free child into the first free slot
        for (int c = 0; c < nk && bc < 0; c++)
          if (!((cu >> c) & 1u)) bc = c;
        for (int s = 0; s < 8 && bs < 0; s++)
          if (!((su >> s) & 1u)) bs = s;
      }
      cu |= 1u << bc;
      su |= 1u << bs;
      slot_of[bc] = bs;
    }
  }
  for (int s = 0; s < 8; s++) child_in_slot[s] = -1;
  for (int c = 0; c < nk; c++) child_in_slot[slot_of[c]] = (int8_t)c;
  uint32_t imask = 0, lmask = 0;
  uint8_t qb[6][8];
  memset(qb, 0, sizeof(qb));
  for (int s = 0; s < 8; s++) {
    const int c = child_in_slot[s];
    if (c < 0) {
      for (int a = 0; a < 3; a++) {  // empty slot: inverted box, never hit
        qb[a][s] = 255;
        qb[3 + a][s] = 0;
      }
      continue;
    }
    const int32_t k = kids[c];
    for (int a = 0; a < 3; a++) {
      const double ql = floor((t.fmin[(size_t)k * 3 + a] - (double)p[a]) / step[a]) - 1.0;
      const double qh = ceil((t.fmax[(size_t)k * 3 + a] - (double)p[a]) / step[a]) + 1.0;
      qb[a][s] = (uint8_t)(ql < 0.0 ? 0.0 : (ql > 255.0 ? 255.0 : ql));
      qb[3 + a][s] = (uint8_t)(qh < 0.0 ? 0.0 : (qh > 255.0 ? 255.0 : qh));
    }
    if (t.nodes[k].prim >= 0)
      lmask |= 1u << s;
    else
      imask |= 1u << s;
  }
  cn.e_imask = (uint32_t)(ex[0] + 127) | ((uint32_t)(ex[1] + 127) << 8) | ((uint32_t)(ex[2] + 127) << 16) | (imask << 24);
  cn.lmask = lmask;
  for (int r = 0; r < 6; r++)
    for (int s = 0; s < 8; s++) cn.q[r * 2 + (s >> 2)] |= (uint32_t)qb[r][s] << (8 * (s & 3));
}

// The leaf reference of primitive p in device slot `slot` (kind, flags, slot)
RTC_HD uint32_t leaf_ref_of(uint8_t kind, uint8_t flags, int32_t xform, uint32_t slot) {
  const uint32_t dk = kind == RTC_KIND_TRIANGLE ? DK_TRI
                      : kind == RTC_KIND_PLANE  ? DK_PLANE
                      : ((flags & RTC_FLAG_TRANSFORMED) && xform >= 0) ? DK_XSPHERE
                                                                       : DK_SPHERE;
  uint32_t r = REF_LEAF | (dk << REF_KIND_SHIFT) | slot;
  if (flags & RTC_FLAG_MIRROR) r |= REF_MIRROR;
  if (flags & RTC_FLAG_TWOSIDED) r |= REF_TWOSIDED;
  if (flags & RTC_FLAG_INVERT) r |= REF_INVERT;
  return r;
}

}  // namespace prep
}  // namespace rtc
