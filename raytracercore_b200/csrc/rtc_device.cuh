// rtc_device.cuh — the wavefront path tracer's device code, templated on the arithmetic type R.
//
// R = double (kernels_f64.cu, compiled -fmad=false): a restatement of the reference's f64 arithmetic, operation by
//     operation, with an explicit fma() exactly where the reference's AVX path fuses (SIMDHelpers.Cross, the
//     Triangle/Sphere position FMAs). Reference files are cited per function (paths relative to RaytracerCore/).
// R = float  (kernels_f32.cu, compiled with the default -fmad=true): same algorithm, throughput-oriented forms of
//     the box and sphere tests, tolerance 1e-4 against the f64 oracle.
#pragma once
#include <math_constants.h>

#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>
#include <type_traits>
#include <utility>

#include "rtc_internal.h"

namespace rtc {

// ---------------------------------------------------------------------------------------------------------
// scalar helpers
// ---------------------------------------------------------------------------------------------------------
template <typename R>
struct Num;
template <>
struct Num<float> {
  static __device__ __forceinline__ float inf() { return CUDART_INF_F; }
  static __device__ __forceinline__ float nan() { return CUDART_NAN_F; }
  static constexpr bool is_f64 = false;
};
template <>
struct Num<double> {
  static __device__ __forceinline__ double inf() { return CUDART_INF; }
  static __device__ __forceinline__ double nan() { return CUDART_NAN; }
  static constexpr bool is_f64 = true;
};

__device__ __forceinline__ float rfma(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double rfma(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ float rsqrt_(float a) { return sqrtf(a); }
__device__ __forceinline__ double rsqrt_(double a) { return sqrt(a); }
__device__ __forceinline__ float rrcp(float a) { return __frcp_rn(a); }
__device__ __forceinline__ double rrcp(double a) { return 1.0 / a; }
// Intersection arithmetic of the f32 mode: the hardware approximations (one MUFU.RCP / MUFU.SQRT, <= 2 ulp, subnormals
// flushed), which have no out-of-line slow path -- an IEEE division or square root compiles to a subroutine call whose calling convention
// constrains register allocation around the whole traversal loop. f64 mode keeps the exact operations.
__device__ __forceinline__ float xrcp(float a) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
}
__device__ __forceinline__ double xrcp(double a) { return 1.0 / a; }
__device__ __forceinline__ float xsqrt(float a) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
}
__device__ __forceinline__ double xsqrt(double a) { return sqrt(a); }
__device__ __forceinline__ float xdiv(float a, float b) { return a * xrcp(b); }
__device__ __forceinline__ double xdiv(double a, double b) { return a / b; }
__device__ __forceinline__ float clamped_rcp(float a) {
  float r = xrcp(a);
  return fabsf(r) <= 0x1.0p64f ? r : copysignf(0x1.0p64f, a);  // also maps the NaN of 1/NaN away
}
__device__ __forceinline__ double clamped_rcp(double a) { return 1.0 / a; }
__device__ __forceinline__ float rpow(float a, float b) { return powf(a, b); }
__device__ __forceinline__ double rpow(double a, double b) { return pow(a, b); }
__device__ __forceinline__ float racos(float a) { return acosf(a); }
__device__ __forceinline__ double racos(double a) { return acos(a); }
// f32: the angles are u * 2 pi with u in [0, 1); sincospif has no large-argument slow path (sincosf carries a Payne-Hanek
// reduction with a local-memory table that is never needed here)
__device__ __forceinline__ void rsincos(float a, float* s, float* c) { sincospif(a * 0.31830988618379067f, s, c); }
__device__ __forceinline__ void rsincos(double a, double* s, double* c) { *s = sin(a); *c = cos(a); }
__device__ __forceinline__ bool rsignbit(float a) { return (__float_as_uint(a) >> 31) != 0; }
__device__ __forceinline__ bool rsignbit(double a) { return (__double_as_longlong(a) < 0); }
__device__ __forceinline__ bool risnan(float a) { return a != a; }
__device__ __forceinline__ bool risnan(double a) { return a != a; }
// MAXPD / MINPD as used by Sse2.Max/Min: (a > b) ? a : b — the second operand wins on NaN.
template <typename R>
__device__ __forceinline__ R sse_max(R a, R b) { return a > b ? a : b; }
template <typename R>
__device__ __forceinline__ R sse_min(R a, R b) { return a < b ? a : b; }

__device__ __forceinline__ uint32_t code_of(float w) { return __float_as_uint(w); }
__device__ __forceinline__ uint32_t code_of(double w) { return (uint32_t)__double_as_longlong(w); }
__device__ __forceinline__ void set_code(float& w, uint32_t c) { w = __uint_as_float(c); }
__device__ __forceinline__ void set_code(double& w, uint32_t c) { w = __longlong_as_double((long long)c); }

template <typename R>
struct V3 {
  R x, y, z;
};
template <typename R>
__device__ __forceinline__ V3<R> mk3(R x, R y, R z) { V3<R> v; v.x = x; v.y = y; v.z = z; return v; }
template <typename R>
__device__ __forceinline__ V3<R> xyz(const V4<R>& a) { return mk3(a.x, a.y, a.z); }
template <typename R>
__device__ __forceinline__ V3<R> operator+(const V3<R>& a, const V3<R>& b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename R>
__device__ __forceinline__ V3<R> operator-(const V3<R>& a, const V3<R>& b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
template <typename R>
__device__ __forceinline__ V3<R> operator*(const V3<R>& a, R s) { return mk3(a.x * s, a.y * s, a.z * s); }
template <typename R>
__device__ __forceinline__ V3<R> operator/(const V3<R>& a, R s) { return mk3(a.x / s, a.y / s, a.z / s); }
template <typename R>
__device__ __forceinline__ V3<R> neg3(const V3<R>& a) { return mk3(-a.x, -a.y, -a.z); }
// Vec4D.Dot (Vec4D.cs:341-347) / SIMDHelpers.Dot (SIMDHelpers.cs:70-100) with a zero W product: (x+y)+z
template <typename R>
__device__ __forceinline__ R dot3(const V3<R>& a, const V3<R>& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
// Vec4D.Cross (Vec4D.cs:355-364): plain products
template <typename R>
__device__ __forceinline__ V3<R> cross3(const V3<R>& a, const V3<R>& b) {
  return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// SIMDHelpers.Cross (SIMDHelpers.cs:44-61): Fma.MultiplySubtract(leftA, rightA, leftB * rightB)
template <typename R>
__device__ __forceinline__ V3<R> scross3(const V3<R>& a, const V3<R>& b) {
  return mk3(rfma(a.y, b.z, -(a.z * b.y)), rfma(a.z, b.x, -(a.x * b.z)), rfma(a.x, b.y, -(a.y * b.x)));
}
// SIMDHelpers.Normalize (SIMDHelpers.cs:332-335): v / sqrt((x²+y²)+(z²+w²))
template <typename R>
__device__ __forceinline__ V3<R> normalize3(const V3<R>& a) {
  R l = rsqrt_((a.x * a.x + a.y * a.y) + a.z * a.z);
  return mk3(a.x / l, a.y / l, a.z / l);
}
template <typename R>
__device__ __forceinline__ V3<R> xnormalize3(const V3<R>& a) {  // the same with the intersection arithmetic
  R l = xsqrt((a.x * a.x + a.y * a.y) + a.z * a.z);
  return mk3(xdiv(a.x, l), xdiv(a.y, l), xdiv(a.z, l));
}
// Mat4x4D * Vec4D through SIMDHelpers.MultiplyMatrixVector (Mat4x4D.cs:171-180, SIMDHelpers.cs:111-127,222-237):
// (m0 x + m1 y) + (m2 z + m3 w), for a point (w = 1) and for a direction (w = 0).
template <typename R>
__device__ __forceinline__ V3<R> xf_point(const V4<R>* rows, const V3<R>& p) {
  return mk3((rows[0].x * p.x + rows[0].y * p.y) + (rows[0].z * p.z + rows[0].w),
             (rows[1].x * p.x + rows[1].y * p.y) + (rows[1].z * p.z + rows[1].w),
             (rows[2].x * p.x + rows[2].y * p.y) + (rows[2].z * p.z + rows[2].w));
}
template <typename R>
__device__ __forceinline__ V3<R> xf_dir(const V4<R>* rows, const V3<R>& d) {
  return mk3((rows[0].x * d.x + rows[0].y * d.y) + rows[0].z * d.z, (rows[1].x * d.x + rows[1].y * d.y) + rows[1].z * d.z,
             (rows[2].x * d.x + rows[2].y * d.y) + rows[2].z * d.z);
}

template <typename R>
__device__ __forceinline__ V4<R> ldg4(const V4<R>* p) {
  if constexpr (std::is_same<R, float>::value) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    V4<R> r; r.x = t.x; r.y = t.y; r.z = t.z; r.w = t.w;
    return r;
  } else {
    const double2* q = reinterpret_cast<const double2*>(p);
    double2 a = __ldg(q), b = __ldg(q + 1);
    V4<R> r; r.x = a.x; r.y = a.y; r.z = b.x; r.w = b.y;
    return r;
  }
}
template <typename R>
__device__ __forceinline__ V4<R> ld4(const V4<R>* p) { return *p; }
template <typename R>
__device__ __forceinline__ void st4(V4<R>* p, R x, R y, R z, R w) {
  V4<R> v; v.x = x; v.y = y; v.z = z; v.w = w;
  *p = v;
}

// One row of a wide BVH node (W values of R) with the widest loads the row size allows.
__device__ __forceinline__ void unpack4(const uint4& t, float* o) {
  o[0] = __uint_as_float(t.x); o[1] = __uint_as_float(t.y); o[2] = __uint_as_float(t.z); o[3] = __uint_as_float(t.w);
}
__device__ __forceinline__ void unpack4(const uint4& t, double* o) {
  o[0] = __hiloint2double((int)t.y, (int)t.x);
  o[1] = __hiloint2double((int)t.w, (int)t.z);
}
template <typename R, int W>
__device__ __forceinline__ void load_row(const R* p, R (&out)[W]) {
  constexpr int bytes = W * (int)sizeof(R);
  if constexpr (bytes % 16 == 0) {
#pragma unroll
    for (int i = 0; i < bytes / 16; i++) {
      uint4 t = __ldg(reinterpret_cast<const uint4*>(p) + i);
      unpack4(t, &out[i * (16 / (int)sizeof(R))]);
    }
  } else {
    static_assert(bytes == 8, "unsupported node row size");
    uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
    out[0] = __uint_as_float(t.x);
    out[1] = __uint_as_float(t.y);
  }
}
template <int W>
__device__ __forceinline__ void load_children(const uint32_t* p, uint32_t (&out)[W]) {
  if constexpr (W % 4 == 0) {
#pragma unroll
    for (int i = 0; i < W / 4; i++) {
      uint4 t = __ldg(reinterpret_cast<const uint4*>(p) + i);
      out[4 * i] = t.x; out[4 * i + 1] = t.y; out[4 * i + 2] = t.z; out[4 * i + 3] = t.w;
    }
  } else {
    uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
    out[0] = t.x;
    out[1] = t.y;
  }
}

// compare-exchange networks that sort W (key, value) pairs ascending by key
template <typename R>
__device__ __forceinline__ void cswap(R& ka, uint32_t& va, R& kb, uint32_t& vb) {
  const bool s = kb < ka;
  const R tk = s ? kb : ka;
  const uint32_t tv = s ? vb : va;
  kb = s ? ka : kb;
  vb = s ? va : vb;
  ka = tk;
  va = tv;
}
template <typename R, int W>
__device__ __forceinline__ void sort_children(R (&k)[W], uint32_t (&v)[W]) {
#define RTC_CS(a, b) cswap(k[a], v[a], k[b], v[b])
  if constexpr (W == 2) {
    RTC_CS(0, 1);
  } else if constexpr (W == 4) {
    RTC_CS(0, 1); RTC_CS(2, 3); RTC_CS(0, 2); RTC_CS(1, 3); RTC_CS(1, 2);
  } else {
    static_assert(W == 8, "unsupported BVH width");
    RTC_CS(0, 1); RTC_CS(2, 3); RTC_CS(4, 5); RTC_CS(6, 7);
    RTC_CS(0, 2); RTC_CS(1, 3); RTC_CS(4, 6); RTC_CS(5, 7);
    RTC_CS(1, 2); RTC_CS(5, 6); RTC_CS(0, 4); RTC_CS(3, 7);
    RTC_CS(1, 5); RTC_CS(2, 6);
    RTC_CS(1, 4); RTC_CS(3, 6);
    RTC_CS(2, 4); RTC_CS(3, 5);
    RTC_CS(3, 4);
  }
#undef RTC_CS
}

// ---------------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011): counter = (pixel, sample, stage, block), key = seed. Replaces the unseeded
// System.Random of Raytracer.cs:48 so that GPU ranks and the CPU oracle draw identical streams.
//   stage 0      camera ray: block 0 -> subX, subY ; block 1 -> lens radius, lens angle   (Raytracer.cs:265-273)
//   stage 1 + i  bounce i:   block 0 -> shine z, shine theta (:53-54) ; block 1 -> lobe pick (:178), diffuse z (:215)
//                            block 2 -> diffuse theta (:216)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    if (r) {
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
  }
  out[0] = c0;
  out[1] = c1;
  out[2] = c2;
  out[3] = c3;
}
template <typename R>
__device__ __forceinline__ void uniforms2(uint32_t seed_lo, uint32_t seed_hi, uint32_t pixel, uint32_t sample, uint32_t stage,
                                          uint32_t block, R& u0, R& u1) {
  uint32_t r[4];
  philox4x32_10(pixel, sample, stage, block, seed_lo, seed_hi, r);
  if constexpr (Num<R>::is_f64) {
    u0 = (double)((((unsigned long long)r[1] << 32) | r[0]) >> 11) * 0x1.0p-53;
    u1 = (double)((((unsigned long long)r[3] << 32) | r[2]) >> 11) * 0x1.0p-53;
  } else {  // the top 24 bits of the same 53-bit draws
    u0 = (float)(r[1] >> 8) * 0x1.0p-24f;
    u1 = (float)(r[3] >> 8) * 0x1.0p-24f;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Vec4D.CreateHorizontal / CreateHorizon (Vec4D.cs:33-58) with MatrixTransforms.Rotate (MatrixTransforms.cs:25-38)
// ---------------------------------------------------------------------------------------------------------
template <typename R>
__device__ __forceinline__ V3<R> create_horizon(const V3<R>& pole, R z, R theta) {
  V3<R> c = cross3(pole, mk3(R(0), R(0), R(1)));
  if (c.x == 0 && c.y == 0 && c.z == 0)
    c = mk3(R(1), R(0), R(0));
  else
    c = normalize3(c);
  R s, co;
  rsincos(theta, &s, &co);
  R cosOpp = 1 - co;
  V3<R> v = (pole * z) + (c * rsqrt_(1 - z * z));
  const V3<R>& a = pole;
  R m00 = co + a.x * a.x * cosOpp, m01 = a.x * a.y * cosOpp - a.z * s, m02 = a.x * a.z * cosOpp + a.y * s;
  R m10 = a.y * a.x * cosOpp + a.z * s, m11 = co + a.y * a.y * cosOpp, m12 = a.y * a.z * cosOpp - a.x * s;
  R m20 = a.z * a.x * cosOpp - a.y * s, m21 = a.z * a.y * cosOpp + a.x * s, m22 = co + a.z * a.z * cosOpp;
  return mk3((m00 * v.x + m01 * v.y) + m02 * v.z, (m10 * v.x + m11 * v.y) + m12 * v.z, (m20 * v.x + m21 * v.y) + m22 * v.z);
}

// ---------------------------------------------------------------------------------------------------------
// Primitive tests. A candidate is one entry of the Hit[] a reference DoRayTrace returns (closest first).
// ---------------------------------------------------------------------------------------------------------
template <typename R>
struct Cand {
  R t;
  V3<R> pos;     // only filled when with_pos<R, WITH_NORMAL>()
  V3<R> normal;  // only filled when WITH_NORMAL
  bool inside;
};
// The f32 traversal does not carry hit positions: its self-hit rule tolerates origin + t * direction, which
// consider_cand forms on the rare occasion a candidate lies on the skip primitive. f64 keeps the reference's exact forms.
template <typename R, bool WITH_NORMAL>
__device__ __forceinline__ constexpr bool with_pos() { return WITH_NORMAL || Num<R>::is_f64; }

// Triangle.RayTraceAVXFaster + GetNormal (Primitives/Triangle.cs:77-146, 209-224)
// FORCE (finalize_hit): the traversal has already accepted this candidate; recompute its record without re-deciding the
// hit (the f32 units are compiled with FMA contraction, so the same test inlined in two kernels may round differently
// on an edge) and with the inside flag the traversal reported.
template <typename R, bool WITH_NORMAL, bool FORCE = false>
__device__ __forceinline__ int tri_hits(const SceneView<R>& sc, uint32_t ref, const V4<R>& A, const V4<R>& B, const V4<R>& C,
                                        const V3<R>& o, const V3<R>& d, Cand<R>* out, bool forced_inside = false) {
  const uint32_t slot = ref & REF_SLOT_MASK;
  V3<R> v0 = xyz(A), e1 = xyz(B), e2 = xyz(C);
  V3<R> off = o - v0;                // :84
  V3<R> s1 = scross3(off, e1);       // :85
  V3<R> s2 = scross3(d, e2);         // :86
  R u = dot3(off, s2);               // :89-97
  R v = dot3(d, s1);
  R dist = dot3(e2, s1);
  R det = dot3(e1, s2);
  R inv = xrcp(det);                 // :107
  if (risnan(inv)) inv = 0;          // :108-110
  u = u * inv;                       // :112
  v = v * inv;
  dist = dist * inv;                 // :113
  bool reject = (u < 0) | (v < 0);   // :116
  if (ref & REF_MIRROR)
    reject |= (u > 1) | (v > 1);     // :117-118
  else
    reject |= (u + v) > 1;
  reject |= dist < 0;                // :120
  if (!FORCE && reject) return 0;
  bool inside = FORCE ? forced_inside : (inv < 0);  // :126
  out[0].t = dist;
  out[0].inside = inside;
  if (with_pos<R, WITH_NORMAL>())
    out[0].pos = mk3(rfma(e1.x, u, rfma(e2.x, v, v0.x)), rfma(e1.y, u, rfma(e2.y, v, v0.y)), rfma(e1.z, u, rfma(e2.z, v, v0.z)));  // :130
  if (WITH_NORMAL) {
    V3<R> N = xyz(ldg4(&sc.sgeom[slot]));
    int32_t ax = sc.aux[slot];
    if (ax >= 0 && (ax & REF_VNORMALS_AUX)) {  // :211-219 (weights and the zero face normal are the reference's)
      const DXform<R>* xf = sc.xforms + (ax & 0x3FFFFFFF);
      V3<R> n0 = xyz(ldg4(&xf->r[0])), n1 = xyz(ldg4(&xf->r[1])), n2 = xyz(ldg4(&xf->r[2]));
      V3<R> nn = xnormalize3(((n0 * u) + (n1 * v)) + (n2 * (u + v)));  // (f64: the exact sqrt and divisions)
      if (inside)
        nn = nn - (N * xdiv(2 * (dot3(nn, N)), dot3(N, N)));
      out[0].normal = nn;
    } else {
      out[0].normal = inside ? (N * R(-1)) : N;  // :221-223
    }
  }
  return 1;
}

// Sphere.RayTraceAVX (Primitives/Sphere.cs:50-155). XF = the sphere is transformed and `x` holds its matrix rows. XF is a
// template parameter so that the rows are defined and consumed inside one straight-line region: a run-time flag tested
// twice makes them conditionally defined, and the register allocator then keeps all of them alive around the whole
// traversal loop (24 registers).
// dlen (f32 mode only): |d|^2 - 1 as the host measured it on the f64 ray it handed in (rtc_trace_closest), 0 inside the render
// loop, where the f32 mode re-normalises every bounce direction.
template <typename R, bool WITH_NORMAL, bool FORCE, bool XF>
__device__ __forceinline__ int sphere_hits(const DXform<R>* x, const V4<R>& A, const V3<R>& o, const V3<R>& d, Cand<R>* out, R dlen) {
  V3<R> C = xyz(A);
  R radius = A.w;
  constexpr bool xf = XF;
  V3<R> oo = o, od = d;
  V4<R> rows[9];
  if (xf) {  // :58-76
#pragma unroll
    for (int i = 0; i < 9; i++) rows[i] = ldg4(&x->r[i]);
    oo = xf_point(rows, o);
    od = xnormalize3(xf_dir(rows, d));
  }
  V3<R> off = oo - C;  // :79
  R t_far, t_close;
  if constexpr (Num<R>::is_f64) {
    R b = -2 * dot3(off, od);                     // :80,84
    R c = dot3(off, off) - radius * radius;       // :81,85 (RadiusSqr = value*value, Sphere.cs:44-45)
    R radix = xsqrt((b * b) - (4 * c));          // :86
    t_far = (b + radix) / 2;                      // :89
    t_close = (b - radix) / 2;                    // :90
  } else {
    // f32: same roots, better-conditioned algebra (perpendicular-distance discriminant, c/q for the small root). The
    // reference's quadratic assumes |od| = 1 (Sphere.cs:80-90) although its bounce directions are only re-normalised every
    // third bounce (Raytracer.cs:74-75) and drift up to 1e-5 off unit length in between: its discriminant is bp^2 - c, while
    // r^2 - |l|^2 = bp^2 (2 - |od|^2) - c. The dlen term restores the reference's value for such a ray (it moves t by 1e-4
    // on small distant spheres). dlen comes from the f64 ray: the f32 components cannot tell a defect from their own rounding.
    R bp = -dot3(off, od);
    V3<R> l = mk3(rfma(bp, od.x, off.x), rfma(bp, od.y, off.y), rfma(bp, od.z, off.z));
    R disc = rfma(bp * bp, XF ? R(0) : dlen, radius * radius - dot3(l, l));
    if (FORCE) disc = fmaxf(disc, R(0));  // an accepted grazing hit keeps a real root
    R c = dot3(off, off) - radius * radius;
    R sq = xsqrt(disc);  // NaN when the ray misses
    const bool fwd = bp >= 0;  // (-0 counts as forward: q must not cancel)
    R q = fwd ? bp + sq : bp - sq;
    R other = xdiv(c, q);
    if (fwd) {
      t_far = q;
      t_close = other;
    } else {
      t_close = q;
      t_far = other;
    }
    if (risnan(sq)) t_far = Num<R>::nan();
  }
  if (!FORCE && !(t_far >= 0) && !xf) return 0;  // (for transformed spheres the test applies to the re-measured distances)
  V3<R> pf = mk3(rfma(t_far, od.x, oo.x), rfma(t_far, od.y, oo.y), rfma(t_far, od.z, oo.z));      // :94
  V3<R> pc = mk3(rfma(t_close, od.x, oo.x), rfma(t_close, od.y, oo.y), rfma(t_close, od.z, oo.z));  // :97
  V3<R> nf, nc;
  if (WITH_NORMAL || xf) {
    const V3<R> df = pf - C, dc = pc - C;
    nf = mk3(xdiv(df.x, radius), xdiv(df.y, radius), xdiv(df.z, radius));  // :95
    nc = mk3(xdiv(dc.x, radius), xdiv(dc.y, radius), xdiv(dc.z, radius));  // :98
  }
  if (xf) {  // :100-139
    pf = xf_point(rows + 3, pf);
    pc = xf_point(rows + 3, pc);
    if (WITH_NORMAL) {
      nf = xnormalize3(xf_dir(rows + 6, nf));
      nc = xnormalize3(xf_dir(rows + 6, nc));
    }
    t_far = dot3(d, pf - o);
    t_close = dot3(d, pc - o);
    if (!FORCE && !(t_far >= 0)) return 0;  // :145-146
  }
  if (WITH_NORMAL) nf = neg3(nf);  // :142
  if (!FORCE && !(t_close >= 0)) {  // :148-149 (FORCE: always [near, far]; finalize_hit picks by the reported inside flag)
    out[0].t = t_far;
    if (with_pos<R, WITH_NORMAL>()) out[0].pos = pf;
    out[0].inside = true;
    if (WITH_NORMAL) out[0].normal = nf;
    return 1;
  }
  out[0].t = t_close;  // :151-154
  out[0].inside = false;
  out[1].t = t_far;
  out[1].inside = true;
  if (with_pos<R, WITH_NORMAL>()) {
    out[0].pos = pc;
    out[1].pos = pf;
  }
  if (WITH_NORMAL) {
    out[0].normal = nc;
    out[1].normal = nf;
  }
  return 2;
}

// Util.NearlyEqual (Util.cs:41-56)
template <typename R>
__device__ __forceinline__ bool nearly_equal(R a, R b, R delta) {
  if (delta == 0) return true;
  delta = delta < 0 ? -delta : delta;
  R mx = (risnan(a) || risnan(b)) ? Num<R>::nan() : (a > b ? a : b);  // Math.Max
  if constexpr (Num<R>::is_f64)
    return delta <= 4.94065645841247e-317 || delta / mx < 1e-24;  // double.Epsilon * 1e7 ; Util.NearEnough
  else
    return xdiv(delta, mx) < 1e-9f;  // f32 stand-in for NearEnough (self-hit rule in float mode)
}

// Plane.DoRayTrace (Primitives/Plane.cs:36-66)
template <typename R, bool WITH_NORMAL, bool FORCE = false>
__device__ __forceinline__ int plane_hits(const V4<R>& A, const V3<R>& o, const V3<R>& d, Cand<R>* out) {
  V3<R> N = xyz(A);
  R od = A.w;
  R ray_dist = dot3(o, N);   // :38
  R denom = dot3(d, N);      // :39
  if (nearly_equal(denom, R(0), denom - R(0)) && nearly_equal(od, ray_dist, od - ray_dist)) {  // :41-42
    out[0].t = 0;
    if (with_pos<R, WITH_NORMAL>()) out[0].pos = o;
    out[0].inside = true;
    if (WITH_NORMAL) out[0].normal = N;
    return 1;
  }
  if (!FORCE && denom == 0) return 0;  // :44-45
  R dist = xdiv(od - ray_dist, denom);  // :47
  if (FORCE || dist >= (Num<R>::is_f64 ? R(-1e-24) : R(0))) {  // :49
    V3<R> hp = o + (d * dist);  // :51
    bool inside = dot3(N, d) > 0;  // :56
    V3<R> dl = hp - o;
    out[0].t = xsqrt((dl.x * dl.x + dl.y * dl.y) + dl.z * dl.z);  // :62 (hitPos - ray.Origin).Length
    if (with_pos<R, WITH_NORMAL>()) out[0].pos = hp;
    out[0].inside = inside;
    if (WITH_NORMAL) out[0].normal = inside ? neg3(N) : N;
    return 1;
  }
  return 0;
}

// The primitive record is fetched whole (3 x 16-byte loads issued together); its third vector carries the leaf reference
// (kind, flags, slot) in the w lane, so a leaf test costs one memory round trip.
template <typename R>
struct PrimRec {
  V4<R> a, b, c;
};
template <typename R>
__device__ __forceinline__ PrimRec<R> load_prim(const SceneView<R>& sc, uint32_t slot) {
  const DPrim<R>* pr = sc.prims + slot;
  PrimRec<R> p;
  p.a = ldg4(&pr->a);
  p.b = ldg4(&pr->b);
  p.c = ldg4(&pr->c);
  return p;
}
template <typename R>
__device__ __forceinline__ uint32_t ref_of(const PrimRec<R>& p) { return code_of(p.c.w); }

template <typename R, bool WITH_NORMAL, bool FORCE = false>
__device__ __forceinline__ int prim_hits(const SceneView<R>& sc, uint32_t ref, const PrimRec<R>& p, const V3<R>& o, const V3<R>& d,
                                         Cand<R>* out, bool forced_inside = false, R dlen = R(0)) {
  const int kind = (ref >> REF_KIND_SHIFT) & 3;
  if (kind == DK_TRI) return tri_hits<R, WITH_NORMAL, FORCE>(sc, ref, p.a, p.b, p.c, o, d, out, forced_inside);
  if (kind == DK_PLANE) return plane_hits<R, WITH_NORMAL, FORCE>(p.a, o, d, out);
  if (kind == DK_XSPHERE)
    return sphere_hits<R, WITH_NORMAL, FORCE, true>(sc.xforms + sc.aux[ref & REF_SLOT_MASK], p.a, o, d, out, R(0));
  return sphere_hits<R, WITH_NORMAL, FORCE, false>(nullptr, p.a, o, d, out, dlen);
}

// The skip hit (previous bounce's Hit) as the trace kernel sees it: primitive slot and inside flag live in
// registers, the rest (position, distance, normal) stays in the path's previous-hit record and is only read on
// the rare occasion a candidate lies on the same primitive. Everything is passed BY VALUE to the out-of-line
// skip_matches: a by-reference argument would pin the caller's copy (and the candidate arrays) in local memory.
template <typename R>
struct Skip {
  uint32_t code;  // the previous hit's code (slot | HIT_INSIDE | HIT_SECOND), HIT_MISS when there is none
  __device__ __forceinline__ uint32_t slot() const { return code == HIT_MISS ? REF_SLOT_MASK + 1 : (code & REF_SLOT_MASK); }
  __device__ __forceinline__ bool inside() const { return (code & HIT_INSIDE) != 0; }
};
template <typename R>
struct SkipSrc {          // warp-uniform: where the skip records of the wavefront live
  const V4<R>* hpos;      // [path] xyz = skip position (= ray origin in the render loop), w = skip Hit.Distance
  const V4<R>* hnrm;      // [path] xyz = skip normal
  const V4<R>* spos;      // [path] explicit skip position (rtc_trace_closest) or nullptr
};

// Util.RayHitMatches (Util.cs:179-192) for a candidate on the same primitive as the skip hit.
template <typename R>
__device__ __forceinline__ bool skip_matches_impl(const SceneView<R>* scp, uint32_t ref, V3<R> o, V3<R> d, int which, bool cand_inside,
                                                  R cand_t, V3<R> cand_pos, bool skip_inside, const V4<R>* hpos, const V4<R>* hnrm,
                                                  const V4<R>* sposp) {
  V4<R> hp = ld4(hpos), hn = ld4(hnrm);
  V3<R> spos = sposp ? xyz(ld4(sposp)) : xyz(hp);
  V3<R> snrm = xyz(hn);
  if constexpr (Num<R>::is_f64) {
    // `a == b` (Hit.cs:44-59): identical primitive, position, distance, normal and inside flag
    if (cand_pos.x == spos.x && cand_pos.y == spos.y && cand_pos.z == spos.z && cand_t == hp.w && cand_inside == skip_inside) {
      Cand<R> full[2];
      const PrimRec<R> pr = load_prim(*scp, ref & REF_SLOT_MASK);
      prim_hits<R, true>(*scp, ref, pr, o, d, full);
      const V3<R> n = which ? full[1].normal : full[0].normal;
      if (n.x == snrm.x && n.y == snrm.y && n.z == snrm.z) return true;
    }
  }
  // Vec4D.NearlyEquals (Vec4D.cs:439-442): squared lengths include W = 1
  R la = ((cand_pos.x * cand_pos.x + cand_pos.y * cand_pos.y) + cand_pos.z * cand_pos.z) + 1;
  R lb = ((spos.x * spos.x + spos.y * spos.y) + spos.z * spos.z) + 1;
  V3<R> dl = cand_pos - spos;
  R ld = (dl.x * dl.x + dl.y * dl.y) + dl.z * dl.z;
  if (!nearly_equal(la, lb, ld)) return false;
  if (dot3(d, snrm) > 0) return cand_inside != skip_inside;
  return cand_inside == skip_inside;
}
// f64: out of line (it re-evaluates the primitive with normals); f32: a dozen operations, inlined so that the traversal
// kernel contains no call at all.
static __device__ __noinline__ bool skip_matches(const SceneView<double>* scp, uint32_t ref, V3<double> o, V3<double> d, int which,
                                          bool cand_inside, double cand_t, V3<double> cand_pos, bool skip_inside,
                                          const V4<double>* hpos, const V4<double>* hnrm, const V4<double>* sposp) {
  return skip_matches_impl<double>(scp, ref, o, d, which, cand_inside, cand_t, cand_pos, skip_inside, hpos, hnrm, sposp);
}
static __device__ __forceinline__ bool skip_matches(const SceneView<float>* scp, uint32_t ref, V3<float> o, V3<float> d, int which,
                                             bool cand_inside, float cand_t, V3<float> cand_pos, bool skip_inside,
                                             const V4<float>* hpos, const V4<float>* hnrm, const V4<float>* sposp) {
  return skip_matches_impl<float>(scp, ref, o, d, which, cand_inside, cand_t, cand_pos, skip_inside, hpos, hnrm, sposp);
}

// ---------------------------------------------------------------------------------------------------------
// Ray / two-box test. f64: AABB.IntersectAVX (Acceleration/AABB.cs:107-142) per child, lane for lane.
// ---------------------------------------------------------------------------------------------------------
template <typename R>
__device__ __forceinline__ void slab_ref(R lo, R hi, R o, R d, R inv, R& n, R& f) {
  if (d == 0 && o >= lo && o <= hi) {  // :117-123
    lo = -Num<R>::inf();
    hi = Num<R>::inf();
  }
  bool sgn = rsignbit(d);  // :126-127
  R a = sgn ? hi : lo, b = sgn ? lo : hi;
  n = (a - o) * inv;  // :129-131
  f = (b - o) * inv;
}

template <typename R>
__device__ __forceinline__ bool box_test(R lox, R hix, R loy, R hiy, R loz, R hiz, const V3<R>& o, const V3<R>& d,
                                         const V3<R>& inv, R& near_out) {
  if constexpr (Num<R>::is_f64) {
    R nx, fx, ny, fy, nz, fz;
    slab_ref(lox, hix, o.x, d.x, inv.x, nx, fx);
    slab_ref(loy, hiy, o.y, d.y, inv.y, ny, fy);
    slab_ref(loz, hiz, o.z, d.z, inv.z, nz, fz);
    // :133-136 with the W lane (-inf, +inf): Max(lower, upper) then MaxScalar(x, swap(x))
    R nr = sse_max(sse_max(nx, nz), sse_max(ny, -Num<R>::inf()));
    R fr = sse_min(sse_min(fx, fz), sse_min(fy, Num<R>::inf()));
    near_out = nr;
    // :138 miss iff near > far or far < 0 ; BVH.cs:302 then requires far >= 0
    return !(nr > fr) && (fr >= 0);
  } else {
    R t0x = (lox - o.x) * inv.x, t1x = (hix - o.x) * inv.x;
    R t0y = (loy - o.y) * inv.y, t1y = (hiy - o.y) * inv.y;
    R t0z = (loz - o.z) * inv.z, t1z = (hiz - o.z) * inv.z;
    // fminf/fmaxf drop NaN (0 * inf on a slab whose plane contains the origin) == the reference's (-inf,+inf) lane
    R nr = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fminf(t0z, t1z));
    R fr = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fmaxf(t0z, t1z));
    fr *= 1.0000004f;  // conservative far plane against f32 rounding (Ize 2013)
    near_out = nr;
    return (nr <= fr) && (fr >= 0);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Scene.RayTrace (Raytracing/Scene.cs:65-120): closest hit over the BVH.
//
// The reference collects every pierced leaf, stable-sorts by box Near and scans until Near > previous.Far, accepting
// with a strict `<`; the result is the global closest hit, ties resolved in favour of the smaller (Near, left-first
// leaf order). This traversal visits nodes front to back with distance pruning and applies exactly that order as a
// tie key (t, leaf Near, slot), slot being the left-first leaf index, so the answer is independent of visit order.
// ---------------------------------------------------------------------------------------------------------
template <typename R>
struct Best {
  R t, near_;
  uint32_t code;  // the hit code the path pool stores: slot | HIT_INSIDE | HIT_SECOND, HIT_MISS = none
};

// One candidate of Primitive.RayTrace (Primitives/Primitive.cs:46-75). Returns true when the candidate is acceptable
// (the reference's `return curHit`), whether or not it then beats the best hit so far.
template <typename R>
__device__ __forceinline__ bool consider_cand(const SceneView<R>& sc, uint32_t ref, int which, bool cand_inside, R t,
                                              const V3<R>& pos, R leaf_near, const V3<R>& o, const V3<R>& d, const Skip<R>& sk,
                                              const SkipSrc<R>& src, uint32_t path, Best<R>& best) {
  const uint32_t slot = ref & REF_SLOT_MASK;
  const bool inside = cand_inside ^ ((ref & REF_INVERT) != 0);  // :60-61, Hit.cs:39-42
  if (inside && !(ref & REF_TWOSIDED)) return false;            // :63-64
  if (slot == sk.slot()) {                                      // :66
    bool same;
    const int kind = (ref >> REF_KIND_SHIFT) & 3;
    if (!Num<R>::is_f64 && (kind == DK_TRI || kind == DK_PLANE)) {
      // f32 mode: a flat primitive can only re-hit itself at the ray origin, so the same primitive is always the
      // self-hit; the positional rule is kept for spheres (their far hit is a legitimate second hit).
      same = true;
    } else {
      const V3<R> cpos = Num<R>::is_f64 ? pos : o + (d * t);
      same = skip_matches(Num<R>::is_f64 ? &sc : nullptr, ref, o, d, which, inside, t, cpos, sk.inside(), src.hpos + path, src.hnrm + path,
                             src.spos ? src.spos + path : nullptr);
    }
    if (same) return false;
  }
  // Scene.cs:85-86 strict `<`; equal distances fall back to the reference's scan order (Near, then leaf order).
  // A candidate whose distance is NaN or +inf (the reference's det == 0 artefacts) is never taken.
  bool take = t < best.t;
  if (!take && t == best.t && best.code != HIT_MISS) {
    uint32_t bslot = best.code & REF_SLOT_MASK;
    if constexpr (Num<R>::is_f64)
      take = (leaf_near < best.near_) || (leaf_near == best.near_ && slot < bslot);
    else
      take = slot < bslot;  // f32 mode: exact-distance ties go to the lower slot (the box near is not tracked)
  }
  if (take) {
    best.t = t;
    best.near_ = leaf_near;
    // slot | kind (ref bits 29-30 -> 26-27) | Primitive.Invert (ref bit 26 -> 28) | which | inside
    best.code = slot | ((ref >> (REF_KIND_SHIFT - HIT_KIND_SHIFT)) & (3u << HIT_KIND_SHIFT)) | ((ref & REF_INVERT) << 2) |
                (inside ? HIT_INSIDE : 0u) | (which ? HIT_SECOND : 0u);
  }
  return true;
}

template <typename R>
__device__ __forceinline__ void test_leaf(const SceneView<R>& sc, uint32_t ref, const PrimRec<R>& pr, R leaf_near, const V3<R>& o,
                                          const V3<R>& d, R dlen, const Skip<R>& sk, const SkipSrc<R>& src, uint32_t path, Best<R>& best) {
  Cand<R> c[2];  // indexed with constants only: stays in registers
  const int n = prim_hits<R, false>(sc, ref, pr, o, d, c, false, dlen);
  bool accepted = false;
  if (n > 0) accepted = consider_cand<R>(sc, ref, 0, c[0].inside, c[0].t, c[0].pos, leaf_near, o, d, sk, src, path, best);
  if (n > 1 && !accepted) consider_cand<R>(sc, ref, 1, c[1].inside, c[1].t, c[1].pos, leaf_near, o, d, sk, src, path, best);  // :68-70
}

// Completes a trace result (distance, slot, inside, which) into the full Hit record (Hit.cs:14-20) by re-evaluating the
// winning primitive once. Runs in the streaming kernels (k_shade, k_export_hits), not in the traversal loop.
template <typename R>
__device__ __forceinline__ void finalize_hit(const SceneView<R>& sc, uint32_t code, const V3<R>& o, const V3<R>& d, R dlen, V3<R>& pos,
                                             V3<R>& normal, R& t) {
  const PrimRec<R> pr = load_prim(sc, code & REF_SLOT_MASK);
  const uint32_t ref = ref_of(pr);
  // the primitive's own inside flag (before Primitive.Invert, Primitive.cs:60-61), as the traversal saw it
  const bool raw_inside = ((code & HIT_INSIDE) != 0) != ((ref & REF_INVERT) != 0);
  const int kind = (ref >> REF_KIND_SHIFT) & 3;
  Cand<R> c[2];
  c[0].t = c[1].t = Num<R>::nan();
  c[0].pos = c[1].pos = c[0].normal = c[1].normal = mk3(Num<R>::nan(), Num<R>::nan(), Num<R>::nan());
  prim_hits<R, true, true>(sc, ref, pr, o, d, c, raw_inside, dlen);
  // forced sphere records are [near (outside), far (inside)]; the other kinds have one record
  const bool second = (kind == DK_SPHERE || kind == DK_XSPHERE) ? raw_inside : false;
  pos = second ? c[1].pos : c[0].pos;
  normal = second ? c[1].normal : c[0].normal;
  t = second ? c[1].t : c[0].t;
}

// ---------------------------------------------------------------------------------------------------------
// Cameras and Raytracer.GetCameraRay (Raytracing/Raytracer.cs:262-282, Cameras/FrustumCamera.cs:33-41,
// Cameras/OrthoCamera.cs:33-38, Vectors/Ray.cs:21-24,53-74)
// ---------------------------------------------------------------------------------------------------------
template <typename R>
__device__ __forceinline__ void camera_get_ray(const CameraView<R>& c, R x, R y, V3<R>& o, V3<R>& d) {
  V3<R> look = mk3(c.look[0], c.look[1], c.look[2]), side = mk3(c.side[0], c.side[1], c.side[2]),
        up = mk3(c.up[0], c.up[1], c.up[2]), pos = mk3(c.position[0], c.position[1], c.position[2]);
  if (c.kind == RTC_CAMERA_FRUSTUM) {
    R off_x = c.tan_fov_x2 * ((x - c.w2) / c.w2);
    R off_y = c.tan_fov_y2 * ((y - c.h2) / c.h2);
    V3<R> dir = (look + (side * off_x)) + (up * off_y);
    o = pos;
    d = normalize3(dir);
  } else {
    o = (pos + (side * ((x - c.w2) * c.h_mult))) + (up * ((y - c.h2) * c.v_mult));
    d = normalize3(look);
  }
}

template <typename R>
__device__ __forceinline__ void get_camera_ray(const CameraView<R>& c, const ParamsView<R>& par, int x, int y, uint32_t sample,
                                               V3<R>& o, V3<R>& d) {
  const uint32_t pixel = (uint32_t)(y * par.width + x);
  R u0, u1;
  uniforms2<R>(par.seed_lo, par.seed_hi, pixel, sample, 0, 0, u0, u1);
  R sub_x = R(x) + u0;
  R sub_y = R(y) + u1;
  camera_get_ray(c, sub_x, sub_y, o, d);
  o = o + (d * c.image_plane);  // Ray.Offset
  if (c.dof_amount != 0) {
    V3<R> focus = o + (d * (c.focal_length - c.image_plane));  // Ray.GetPoint
    R l0, l1;
    uniforms2<R>(par.seed_lo, par.seed_hi, pixel, sample, 0, 1, l0, l1);
    R dist = rsqrt_(l0) * c.dof_amount;
    R angle = l1 * R(3.14159265358979323846) * 2;
    R sn, cs;
    rsincos(angle, &sn, &cs);
    R off_x = cs * dist;
    R off_y = sn * dist;
    V3<R> o2, d2;
    camera_get_ray(c, sub_x + off_x, sub_y + off_y, o2, d2);
    o2 = o2 + (d2 * c.image_plane);
    o = o2;
    d = normalize3(focus - o2);  // PointingTowards -> FromTo
  }
}

// Band-local pixel index -> pixel. Pixels are numbered tile by tile, 8 x 4 pixels per tile, so that the 32 camera rays of a
// warp cover a compact patch of the image instead of a 32 x 1 strip (closer rays fetch the same nodes: fewer L1 wavefronts);
// a band whose width or height is no multiple of the tile falls back to row-major order. The Philox counter uses the
// pixel's own coordinates, so the image does not depend on this numbering.
__device__ __forceinline__ void band_pix_xy(const Band& b, uint32_t pix, int& x, int& y) {
  const uint32_t rw = (uint32_t)(b.x1 - b.x0), rh = (uint32_t)(b.y1 - b.y0);
  if (((rw & 7u) | (rh & 3u)) == 0) {
    const uint32_t tile = pix >> 5, in = pix & 31u, tiles_per_row = rw >> 3;
    const uint32_t ty = tile / tiles_per_row, tx = tile - ty * tiles_per_row;
    x = b.x0 + (int)(tx * 8u + (in & 7u));
    y = b.y0 + (int)(ty * 4u + (in >> 3));
  } else {
    const uint32_t py = pix / rw;
    x = b.x0 + (int)(pix - py * rw);
    y = b.y0 + (int)py;
  }
}

__device__ __forceinline__ void band_pixel(const Band& b, uint32_t path, int& x, int& y, uint32_t& sample) {
  uint32_t s_local = path / b.n_pix;
  uint32_t pix = path - s_local * b.n_pix;
  band_pix_xy(b, pix, x, y);
  sample = b.first_sample + s_local;
}

// ---------------------------------------------------------------------------------------------------------
// Kernels
// ---------------------------------------------------------------------------------------------------------
constexpr int kStreamThreads = 256;
#ifndef RTC_TRACE_THREADS
#define RTC_TRACE_THREADS 128
#endif
#ifndef RTC_TRACE_MIN_BLOCKS
#define RTC_TRACE_MIN_BLOCKS 7
#endif
#ifndef RTC_SHADE_THREADS
#define RTC_SHADE_THREADS 256
#endif
constexpr int kShadeThreads = RTC_SHADE_THREADS;
#ifndef RTC_SHADE_MIN_BLOCKS
#define RTC_SHADE_MIN_BLOCKS 3
#endif
constexpr int kTraceThreads = RTC_TRACE_THREADS;
constexpr int kTraceMinBlocks = RTC_TRACE_MIN_BLOCKS;

// raygen: one thread per path of the band. Writes the bounce-0 ray (already re-normalised as GetColor does for
// i % 3 == 0, Raytracer.cs:74-75), tint = 1 and an empty skip hit.
template <typename R>
__global__ void __launch_bounds__(kStreamThreads) k_raygen(CameraView<R> cam, ParamsView<R> par, Band band, PathView<R> pv) {
  uint32_t path = blockIdx.x * blockDim.x + threadIdx.x;
  if (path == 0) {
    pv.ctl->count[0] = band.n_paths;
    pv.ctl->count[1] = 0;
    pv.ctl->work_trace = 0;
  }
  if (path >= band.n_paths) return;
  int x, y;
  uint32_t sample;
  band_pixel(band, path, x, y, sample);
  V3<R> o, d;
  get_camera_ray(cam, par, x, y, sample, o, d);
  d = normalize3(d);  // Ray.Directional at i == 0
  R w;
  set_code(w, HIT_MISS);  // no skip hit
  st4(&pv.dir[path], d.x, d.y, d.z, R(0));
  st4(&pv.tint[path], R(1), R(1), R(1), R(0));
  st4(&pv.hpos[path], o.x, o.y, o.z, Num<R>::is_f64 ? R(0) : w);
  if (Num<R>::is_f64) st4(&pv.hnrm[path], R(0), R(0), R(0), w);
}

template <typename R>
__global__ void __launch_bounds__(kStreamThreads) k_camera_rays(CameraView<R> cam, ParamsView<R> par, int64_t n,
                                                                 const int32_t* xy, const uint32_t* sample, rtc_ray* out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  V3<R> o, d;
  get_camera_ray(cam, par, xy[2 * i], xy[2 * i + 1], sample[i], o, d);
  out[i].origin[0] = o.x; out[i].origin[1] = o.y; out[i].origin[2] = o.z;
  out[i].dir[0] = d.x; out[i].dir[1] = d.y; out[i].dir[2] = d.z;
}

// What one trace launch reads and writes, resolved on the host so that every pointer is a kernel parameter (constant
// bank operand) instead of a value selected on the device that would occupy registers for the whole traversal.
template <typename R>
struct TraceIO {
  const V4<R>* dir;       // [path] ray direction (w, f32 mode: its norm defect)
  const V4<R>* in_hpos;   // [path] xyz = ray origin = previous hit position, w = f64: previous Hit.Distance, f32: the skip hit's code
  const V4<R>* in_hnrm;   // [path] xyz = previous hit normal, w = previous hit code (the skip hit)
  const V4<R>* skip_pos;  // [path] explicit skip position (rtc_trace_closest) or nullptr
  THit<R>* out;           // [path] the answer: Hit.Distance, hit code
  const uint32_t* queue;  // live path ids, or nullptr for the identity queue (bounce 0)
  const uint32_t* count;  // number of queue entries
  Control* ctl;
};

// The same for one shade launch (dynamic indexing of pointer arrays inside a kernel parameter forces a local-memory copy of
// the whole parameter block and a local load per access).
template <typename R>
struct ShadeIO {
  V4<R>* dir;             // [path] in: ray direction, out: next direction
  V4<R>* tint;            // [path] throughput
  const THit<R>* thit;    // [path] this bounce's trace result
  V4<R>* hpos;            // [path] in: xyz = this bounce's ray origin; out: the hit position (| Hit.Distance or code)
  V4<R>* hnrm;            // [path] out: the hit normal | code
  V4<R>* radiance;        // [path] out for finished paths
  const uint32_t* queue;  // live path ids of this bounce, or nullptr for the identity queue (bounce 0)
  const uint32_t* count;  // their number
  uint32_t* queue_out;    // survivors are appended here ...
  uint32_t* count_out;    // ... and counted here
  int32_t* dbg_type;      // optional per-path BounceType (rtc_debug_trace)
  R* dbg_fresnel;         // optional per-path FresnelRatio
};

// trace: persistent warps, one ray per lane, scheduled warp-synchronously. On sm_70+ a per-lane `while` nest is not
// re-converged at loop exits and degenerates into lanes issuing one at a time, so the kernel is written as ONE
// warp-uniform loop whose every iteration runs exactly one of three bodies, chosen by warp vote:
//   refill  lanes whose ray is finished write its Hit record and take the next queue entries (one warp-aggregated
//           atomicAdd on a global cursor); taken when more than 32 - kRefill lanes are idle (Aila & Laine 2009);
//   (f32 kernel: a lane may carry one postponed leaf group, see k_trace_q8 -- a leaf step then runs once RTC_LEAF_T lanes
//   cannot go on without one, and every lane with pending leaf hits takes part: 14 lanes per leaf step instead of 10,
//   26.5 per node step instead of 24, for 4 % more node visits; +2.7 % rays/s on the 1 M-triangle soup)
//   node    lanes standing on an inner node fetch it, test its child boxes (f64 mode: four, with the reference's own slab
//           arithmetic), descend to the nearest hit child and push the others, farthest first;
//   leaf    lanes standing on a leaf test its primitive and pop their next stack entry.
// node vs leaf is greedy: whichever has more lanes ready runs, the other lanes wait. Leaves are ~1 in 16 steps of a
// ray, so a plain while-while loop (all lanes reach a leaf before any is tested) leaves ~3/4 of the lanes idle.
// The per-lane stack holds (node, box near) pairs so stale entries are discarded without fetching the node.
// k_trace below is the f64 parity kernel (uncompressed 4-wide nodes); the f32 production kernel k_trace_q8 further down
// uses the same scheduler over quantised 8-wide nodes.
#ifndef RTC_REFILL
#define RTC_REFILL 26
#endif
#ifndef RTC_LEAF_T
#define RTC_LEAF_T 4
#endif
#ifndef RTC_POSTPONE
#define RTC_POSTPONE 1
#endif
constexpr int kRefill = RTC_REFILL;
constexpr uint32_t kNone = 0xFFFFFFFFu;

// The f64 kernel's traversal stack: (node, box near) pairs in dynamic shared memory, one column per thread
// ([entry][thread] nodes, then [entry][thread] nears), sized per scene to the tree's depth -- a per-thread array would live
// in local memory (1.5 KB per thread, every push and pop a trip through L1).
template <typename R>
struct TraceStack {
  uint32_t node_at, near_at;  // shared-window byte addresses of this thread's entry 0
  __device__ __forceinline__ void init(uint32_t base, int entries) {
    node_at = base + threadIdx.x * 4u;
    near_at = base + (uint32_t)entries * kTraceThreads * 4u + threadIdx.x * (uint32_t)sizeof(R);
  }
  __device__ __forceinline__ void push(int sp, uint32_t node, R nr) const {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(node_at + (uint32_t)sp * kTraceThreads * 4u), "r"(node));
    if constexpr (sizeof(R) == 8)
      asm volatile("st.shared.f64 [%0], %1;" ::"r"(near_at + (uint32_t)sp * kTraceThreads * 8u), "d"(nr));
    else
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(near_at + (uint32_t)sp * kTraceThreads * 4u), "f"(nr));
  }
  __device__ __forceinline__ uint32_t node(int sp) const {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(node_at + (uint32_t)sp * kTraceThreads * 4u));
    return v;
  }
  __device__ __forceinline__ R nearv(int sp) const {
    R v;
    if constexpr (sizeof(R) == 8)
      asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(near_at + (uint32_t)sp * kTraceThreads * 8u));
    else
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(near_at + (uint32_t)sp * kTraceThreads * 4u));
    return v;
  }
};

template <typename R>
__device__ __forceinline__ void stack_pop(const TraceStack<R>& stk, int& sp, R best_t, uint32_t& cur, R& cur_near) {
  cur = kNone;
  while (sp > 0) {
    sp--;
    const R nr = stk.nearv(sp);
    if (!(nr > best_t)) {
      cur = stk.node(sp);
      cur_near = nr;
      break;
    }
  }
}

template <typename R, bool COUNT>
__global__ void __launch_bounds__(kTraceThreads, 2) k_trace(SceneView<R> sc, TraceIO<R> io) {
  const uint32_t count = *io.count;
  const int lane = threadIdx.x & 31;
  const uint32_t lt_mask = (1u << lane) - 1u;
  uint32_t n_nodes = 0, n_prims = 0;
  extern __shared__ uint2 s_stack[];
  TraceStack<R> stk;
  stk.init((uint32_t)__cvta_generic_to_shared(s_stack), sc.q_stack);
  bool active = false;      // the lane holds an unfinished ray
  bool finished = false;    // the lane holds a finished ray whose Hit record is not written yet
  bool exhausted = false;   // warp-uniform: the queue has no more entries
  uint32_t path = 0;
  V3<R> o = mk3(R(0), R(0), R(0)), d = o, inv = o;
  Skip<R> sk;
  sk.code = HIT_MISS;
  SkipSrc<R> src;
  src.hpos = io.in_hpos;
  src.hnrm = io.in_hnrm;
  src.spos = io.skip_pos;
  Best<R> best;
  best.t = Num<R>::inf();
  best.near_ = 0;
  best.code = HIT_MISS;
  int sp = 0;
  uint32_t cur = kNone;
  R cur_near = 0;

  for (;;) {
    const unsigned m_idle = __ballot_sync(0xFFFFFFFFu, !active);
    if (m_idle == 0xFFFFFFFFu || (!exhausted && __popc(m_idle) > 32 - kRefill)) {
      // ---- refill (f64) ------------------------------------------------------------------------------------
      if (finished) {  // (distance, hit code): position and normal are completed by k_shade / k_export_hits
        THit<R> th;
        th.t = best.t;
        th.code = best.code;
        th.pad = 0;
        io.out[path] = th;
        finished = false;
      }
      if (exhausted) {
        if (m_idle == 0xFFFFFFFFu) break;
        continue;
      }
      uint32_t base = 0;
      const int leader = __ffs(m_idle) - 1;
      if (lane == leader) base = atomicAdd(&io.ctl->work_trace, (uint32_t)__popc(m_idle));
      base = __shfl_sync(0xFFFFFFFFu, base, leader);
      bool got = true;
      if (!active) {
        const uint32_t idx = base + __popc(m_idle & lt_mask);
        got = idx < count;
        if (got) {
          path = io.queue ? io.queue[idx] : idx;
          V4<R> dv = ld4(&io.dir[path]);
          V4<R> op = ld4(&io.in_hpos[path]);
          const uint32_t code = code_of(io.in_hnrm[path].w);
          o = xyz(op);
          d = xyz(dv);
          inv = mk3(R(1) / d.x, R(1) / d.y, R(1) / d.z);  // AABB.cs:129
          sk.code = code;
          best.t = Num<R>::inf();
          best.near_ = 0;
          best.code = HIT_MISS;
          sp = 0;
          cur = sc.root;
          cur_near = 0;
          active = true;
        }
      }
      exhausted = !__all_sync(0xFFFFFFFFu, got);
      continue;
    }

    const unsigned m_node = __ballot_sync(0xFFFFFFFFu, active && !(cur & REF_LEAF));
    const unsigned m_leaf = ~(m_node | m_idle);
    if (__popc(m_node) >= __popc(m_leaf)) {
      // ---- node step ---------------------------------------------------------------------------------------
      if (active && !(cur & REF_LEAF)) {
        constexpr int W = Width<R>::value;
        const DNode<R>* np = sc.nodes + cur;
        uint32_t ch[W];
        R key[W];
        int nh = 0;
        {
          R lo[3][W], hi[3][W];
#pragma unroll
          for (int a = 0; a < 3; a++) {
            load_row<R, W>(np->lo[a], lo[a]);
            load_row<R, W>(np->hi[a], hi[a]);
          }
          load_children<W>(np->child, ch);
#pragma unroll
          for (int c = 0; c < W; c++) {
            R nr;
            bool h = box_test<R>(lo[0][c], hi[0][c], lo[1][c], hi[1][c], lo[2][c], hi[2][c], o, d, inv, nr);
            h = h && (ch[c] != REF_EMPTY) && !(nr > best.t);
            if (h && risnan(nr)) nr = -Num<R>::inf();
            key[c] = h ? nr : Num<R>::inf();
            nh += h ? 1 : 0;
          }
        }
        if (COUNT) n_nodes++;
        sort_children<R, W>(key, ch);  // hits first, nearest first (misses carry +inf keys)
#pragma unroll
        for (int c = W - 1; c >= 1; c--) {
          if (c < nh) {  // farthest pushed first, so the nearest pending child is popped first
            stk.push(sp, ch[c], key[c]);
            sp++;
          }
        }
        if (nh > 0) {
          cur = ch[0];
          cur_near = key[0];
        } else {
          stack_pop<R>(stk, sp, best.t, cur, cur_near);
        }
      }
    } else {
      // ---- leaf step ---------------------------------------------------------------------------------------
      if (active && (cur & REF_LEAF)) {
        if (COUNT) n_prims++;
        test_leaf<R>(sc, cur, load_prim(sc, cur & REF_SLOT_MASK), cur_near, o, d, R(0), sk, src, path, best);
        stack_pop<R>(stk, sp, best.t, cur, cur_near);
      }
    }
    if (active && cur == kNone) {
      active = false;
      finished = true;
    }
  }
  if (COUNT) {
    atomicAdd(&io.ctl->nodes_visited, (unsigned long long)n_nodes);
    atomicAdd(&io.ctl->prims_tested, (unsigned long long)n_prims);
  }
}

// ---------------------------------------------------------------------------------------------------------
// trace, f32 production mode: the same warp-synchronous scheduler over the quantised 8-wide tree (CNode).
//   node step: pop the highest-priority pending child of the current inner group, fetch its node with
//              2 x LDG.256 + 1 x LDG.128, decode the 8 child boxes (one IDP.4A per bound: byte q becomes the float
//              2^23 + q, the 2^23 is folded into the FMA addend; one FFMA2 per bound PAIR) and intersect them; the hits
//              form one inner group and one leaf group, addressed implicitly (base + popcount), so at most ONE stack
//              entry is pushed;
//   leaf step: pop the highest-priority pending leaf of the leaf group and test its primitive.
// Register budget (72 -> 7 CTAs = 28 warps per SM; 64 registers / 8 CTAs measures 2.7 % less with the postponed leaf group, which
// then spills 8 bytes): everything a lane needs only now and then lives in
// shared memory, one 4-byte column per thread and field -- the stack [entry][thread] (conflict-free for any per-lane
// depth), then direction, path id, skip code, origin and reciprocal direction -- and is addressed from ONE per-thread
// shared-window byte address (`sm`); the kernel contains no call (see xrcp/xsqrt) and no conditionally defined values
// that would stay live around the loop (see sphere_hits).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ldg256(const void* p, uint32_t (&w)[8]) {
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
               : "l"(p));
}
// float(2^23 + byte k of w): dp4a(w, 1 << 8k, 2^23 as float bits) = one IDP.4A on the FMA-heavy pipe instead of a PRMT on
// the (half-rate, busier) ALU pipe
__device__ __forceinline__ float qbyte(uint32_t w, int k) { return __uint_as_float(__dp4a(w, 1u << (8 * k), 0x4B000000u)); }

// Packed f32 pairs (sm_100 FFMA2 / FADD2): one issue slot for two children's slab distances.
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b, float c) {
  unsigned long long A, B, C, D;
  asm("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %1};" : "=l"(B) : "f"(b));
  asm("mov.b64 %0, {%1, %1};" : "=l"(C) : "f"(c));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(D) : "l"(A), "l"(B), "l"(C));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(D));
}
__device__ __forceinline__ void fsub2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  unsigned long long A, B, D;
  asm("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(B) : "f"(b0), "f"(b1));
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(D) : "l"(A), "l"(B));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(D));
}
// shared-window accesses through a 32-bit byte address (no generic pointer, no base recomputation per access)
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v)); }
__device__ __forceinline__ uint2 lds64(uint32_t a) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts64(uint32_t a, uint32_t x, uint32_t y) {
  asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y));
}

#ifdef RTC_TRACE_MAXNREG  // tuning: an explicit register budget instead of one derived from the resident-CTA target
#define RTC_Q8_BOUNDS __maxnreg__(RTC_TRACE_MAXNREG)
#else
#define RTC_Q8_BOUNDS __launch_bounds__(kTraceThreads, kTraceMinBlocks)
#endif
constexpr int kQ8StateSlots = 6;  // per-thread cold state in front of the stack, 8 bytes per slot: o.xyz, inv.xyz, d.xyz, norm defect, path, skip code

template <bool COUNT>
__global__ void RTC_Q8_BOUNDS k_trace_q8(SceneView<float> sc, TraceIO<float> io) {
  using R = float;
  // [kQ8StateSlots + sc.q_stack][kTraceThreads] x 8 bytes: first the cold ray state of each thread, two words per slot (one
  // LDS.64 / STS.64 each), then its stack entries. Every access is this thread's column address `sm` plus an immediate (state)
  // or plus sp * stride + an immediate (stack): one address register for all of it.
  extern __shared__ uint2 s_stack[];
  constexpr uint32_t kStackStride = kTraceThreads * 8u;
  const uint32_t sm = (uint32_t)__cvta_generic_to_shared(s_stack) + threadIdx.x * 8u;
  enum { S_OXY = 0, S_OZIX, S_IYZ, S_DXY, S_DZL, S_PATHSKIP };  // (o.x, o.y) (o.z, inv.x) (inv.y, inv.z) (d.x, d.y) (d.z, norm defect) (path, skip code)
  constexpr uint32_t kStackBase = kQ8StateSlots * kStackStride;
  const uint32_t count = *io.count;
  uint32_t n_nodes = 0, n_prims = 0, n_node_steps = 0, n_leaf_steps = 0;
  // lane state in `sp`: >= 0 traversing (= stack depth), kFinished = result not written yet, kEmpty = no ray
  constexpr int kFinished = -1, kEmpty = -2;
  int sp = kEmpty;
  uint32_t exhausted = 0;      // warp-uniform: the queue has no more entries
  uint32_t lanes_changed = 1;  // warp-uniform: some lane finished its ray since the idle lanes were last counted
  uint32_t octinv = 0;
  SkipSrc<R> src;
  src.hpos = io.in_hpos;
  src.hnrm = io.in_hnrm;
  src.spos = io.skip_pos;
  Best<R> best;
  best.t = Num<R>::inf();
  best.near_ = 0;
  best.code = HIT_MISS;
  // current groups: inner (igx = child_base, igy = hits << 8 | imask) and leaf (lgx = prim_base, lgy = hits << 8 | lmask);
  // hit bits are stored at position slot ^ octinv so that the highest set bit is the child to visit first
  uint32_t igx = 0, igy = 0, lgx = 0, lgy = 0;
  // postponed leaf group (Aila & Laine's speculative traversal, one group deep): a lane whose node step produced leaf hits
  // does not wait for the warp's next leaf step while it still has inner children to visit; it parks the group here and
  // goes on, and must stop only when a second leaf group arrives or its inner work runs out. Leaf steps then find more
  // lanes ready, node steps fewer lanes waiting.
  uint32_t pgx = 0, pgy = 0;

  for (;;) {
    unsigned m_idle = 0;
    if (lanes_changed) m_idle = __ballot_sync(0xFFFFFFFFu, sp < 0);
    lanes_changed = 0;
    if (m_idle == 0xFFFFFFFFu || (!exhausted && __popc(m_idle) > 32 - kRefill)) {
      // ---- refill ------------------------------------------------------------------------------------------
      lanes_changed = 1;  // recount after the refill (and keep coming back here once the queue is exhausted)
      if (sp == kFinished) {  // (distance, hit code) in one 8-byte store: position and normal are completed by k_shade
        const uint32_t fpath = lds32(sm + S_PATHSKIP * kStackStride);
        THit<R> th;
        th.t = best.t;
        th.code = best.code;
        io.out[fpath] = th;
        sp = kEmpty;
      }
      if (exhausted) {
        if (m_idle == 0xFFFFFFFFu) break;
        continue;
      }
      const int lane = threadIdx.x & 31;
      uint32_t base = 0;
      const int leader = __ffs(m_idle) - 1;
      if (lane == leader) base = atomicAdd(&io.ctl->work_trace, (uint32_t)__popc(m_idle));
      base = __shfl_sync(0xFFFFFFFFu, base, leader);
      bool got = true;
      if (sp < 0) {
        const uint32_t idx = base + __popc(m_idle & ((1u << lane) - 1u));
        got = idx < count;
        if (got) {
          const uint32_t npath = io.queue ? io.queue[idx] : idx;
          V4<R> dv = ld4(&io.dir[npath]);      // direction | norm defect
          V4<R> op = ld4(&io.in_hpos[npath]);  // origin | skip code: the whole ray in two 16-byte loads
          const uint32_t code = code_of(op.w);
          const V3<R> o = xyz(op), d = xyz(dv);
          const V3<R> inv = mk3(clamped_rcp(d.x), clamped_rcp(d.y), clamped_rcp(d.z));
          sts64(sm + S_OXY * kStackStride, __float_as_uint(o.x), __float_as_uint(o.y));
          sts64(sm + S_OZIX * kStackStride, __float_as_uint(o.z), __float_as_uint(inv.x));
          sts64(sm + S_IYZ * kStackStride, __float_as_uint(inv.y), __float_as_uint(inv.z));
          sts64(sm + S_DXY * kStackStride, __float_as_uint(d.x), __float_as_uint(d.y));
          sts64(sm + S_DZL * kStackStride, __float_as_uint(d.z), __float_as_uint(dv.w));
          sts64(sm + S_PATHSKIP * kStackStride, npath, code);
          Skip<R> sk;
          sk.code = code;
          const uint32_t oct = (rsignbit(d.x) ? 1u : 0u) | (rsignbit(d.y) ? 2u : 0u) | (rsignbit(d.z) ? 4u : 0u);
          octinv = 7u ^ oct;
          best.t = Num<R>::inf();
          best.near_ = 0;
          best.code = HIT_MISS;
          sp = 0;
          // a ray with a non-finite component (a degenerate bounce) hits nothing; without this guard its NaN slab
          // distances would pass every box test and walk the whole tree
          const bool finite = (fabsf(o.x) + fabsf(o.y) + fabsf(o.z)) + (fabsf(d.x) + fabsf(d.y) + fabsf(d.z)) < Num<R>::inf();
          // primitives without a finite box (planes) are tested up front, in leaf order
          for (int i = 0; finite && i < sc.n_unbounded; i++) {
            if (COUNT) n_prims++;
            const uint32_t uref = sc.unbounded[i];
            test_leaf<R>(sc, uref, load_prim(sc, uref & REF_SLOT_MASK), -Num<R>::inf(), o, d, dv.w, sk, src, npath, best);
          }
          // virtual root group: one inner child in slot 0 = node 0
          igx = 0;
          igy = (sc.qnodes && finite) ? (((1u << (0u ^ octinv)) << 8) | 1u) : 0u;
          lgx = 0;
          lgy = 0;
          pgy = 0;
        }
      }
      exhausted = __all_sync(0xFFFFFFFFu, got) ? 0u : 1u;
      continue;
    }

#if RTC_POSTPONE
    const bool has_l = sp >= 0 && (lgy >> 8) != 0, has_p = sp >= 0 && (pgy >> 8) != 0;
    const bool want_node = sp >= 0 && (igy >> 8) != 0 && !(has_l && has_p);  // a free leaf-group slot for the node's leaves
    const bool want_leaf = has_l || has_p;                                   // takes part in a leaf step
    const unsigned m_leaf = __ballot_sync(0xFFFFFFFFu, want_leaf && !want_node);  // lanes that cannot go on without one
    const unsigned m_node = __ballot_sync(0xFFFFFFFFu, want_node);
    if (m_node && __popc(m_leaf) < RTC_LEAF_T) {
#else
    const bool has_p = false;
    const bool want_leaf = sp >= 0 && (lgy >> 8) != 0;
    const bool want_node = sp >= 0 && !want_leaf && (igy >> 8) != 0;
    const unsigned m_leaf = __ballot_sync(0xFFFFFFFFu, want_leaf);
    const unsigned m_node = __ballot_sync(0xFFFFFFFFu, want_node);
    if (m_node && __popc(m_leaf) < RTC_LEAF_T) {
#endif
      // ---- node step ---------------------------------------------------------------------------------------
      if (COUNT && (threadIdx.x & 31) == 0) n_node_steps++;
      if (want_node) {
        const uint32_t hits = igy >> 8;
        const uint32_t b = 31u - (uint32_t)__clz((int)hits);
        const uint32_t s = b ^ octinv;
        const uint32_t imask_g = igy & 0xFFu;
        const uint32_t node = igx + (uint32_t)__popc(imask_g & ((1u << s) - 1u));
        igy &= ~(0x100u << b);
        if ((igy >> 8) != 0) {  // siblings still pending: the group goes to the stack
          sts64(sm + kStackBase + (uint32_t)sp * kStackStride, igx, igy);
          sp++;
        }
        const CNode* np = sc.qnodes + node;
        uint32_t w0[8], w1[8];
        ldg256(np, w0);
        ldg256(reinterpret_cast<const char*>(np) + 32, w1);
        const uint4 w2 = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(np) + 64));
        if (COUNT) n_nodes++;
        const uint32_t em = w0[3];
        const float sx = __uint_as_float((em & 0xFFu) << 23), sy = __uint_as_float(((em >> 8) & 0xFFu) << 23),
                    sz = __uint_as_float(((em >> 16) & 0xFFu) << 23);
        const uint32_t imask = em >> 24, lmask = w0[6] & 0xFFu;
        // t = (p + q * step - o) * inv = q * (step * inv) + (p - o) * inv ; q enters as 2^23 + q, so the addend carries
        // -2^23 * step * inv (its rounding is half a grid step: the builder pads every box by one step)
        const uint2 s0 = lds64(sm + S_OXY * kStackStride), s1 = lds64(sm + S_OZIX * kStackStride), s2 = lds64(sm + S_IYZ * kStackStride);
        const V3<R> ro = mk3(__uint_as_float(s0.x), __uint_as_float(s0.y), __uint_as_float(s1.x));
        const V3<R> ri = mk3(__uint_as_float(s1.y), __uint_as_float(s2.x), __uint_as_float(s2.y));
        const float ax = sx * ri.x, ay = sy * ri.y, az = sz * ri.z;
        const float bx = fmaf(-8388608.0f, ax, (__uint_as_float(w0[0]) - ro.x) * ri.x);
        const float by = fmaf(-8388608.0f, ay, (__uint_as_float(w0[1]) - ro.y) * ri.y);
        const float bz = fmaf(-8388608.0f, az, (__uint_as_float(w0[2]) - ro.z) * ri.z);
        // near / far byte rows by direction sign: w1 = {lox0,lox1,loy0,loy1,loz0,loz1,hix0,hix1}, w2 = {hiy0,hiy1,hiz0,hiz1}
        const bool nx_ = !(octinv & 1u), ny_ = !(octinv & 2u), nz_ = !(octinv & 4u);  // sign of the direction per axis
        const uint32_t nxw[2] = {nx_ ? w1[6] : w1[0], nx_ ? w1[7] : w1[1]}, fxw[2] = {nx_ ? w1[0] : w1[6], nx_ ? w1[1] : w1[7]};
        const uint32_t nyw[2] = {ny_ ? w2.x : w1[2], ny_ ? w2.y : w1[3]}, fyw[2] = {ny_ ? w1[2] : w2.x, ny_ ? w1[3] : w2.y};
        const uint32_t nzw[2] = {nz_ ? w2.z : w1[4], nz_ ? w2.w : w1[5]}, fzw[2] = {nz_ ? w1[4] : w2.z, nz_ ? w1[5] : w2.w};
        // two children per FFMA2; the hit mask is gathered from the sign bits of (far - near) with one funnel shift per
        // child, children taken from slot 7 down so that child c ends at bit c
        uint32_t acc = 0;
#pragma unroll
        for (int c = 6; c >= 0; c -= 2) {
          const int wi = c >> 2, k = c & 3;
          float tnx0, tnx1, tny0, tny1, tnz0, tnz1, tfx0, tfx1, tfy0, tfy1, tfz0, tfz1;
          ffma2(tnx0, tnx1, qbyte(nxw[wi], k), qbyte(nxw[wi], k + 1), ax, bx);
          ffma2(tny0, tny1, qbyte(nyw[wi], k), qbyte(nyw[wi], k + 1), ay, by);
          ffma2(tnz0, tnz1, qbyte(nzw[wi], k), qbyte(nzw[wi], k + 1), az, bz);
          ffma2(tfx0, tfx1, qbyte(fxw[wi], k), qbyte(fxw[wi], k + 1), ax, bx);
          ffma2(tfy0, tfy1, qbyte(fyw[wi], k), qbyte(fyw[wi], k + 1), ay, by);
          ffma2(tfz0, tfz1, qbyte(fzw[wi], k), qbyte(fzw[wi], k + 1), az, bz);
          const float nr0 = fmaxf(fmaxf(fmaxf(tnx0, tny0), tnz0), 0.0f), nr1 = fmaxf(fmaxf(fmaxf(tnx1, tny1), tnz1), 0.0f);
          const float fr0 = fminf(fminf(fminf(tfx0, tfy0), tfz0), best.t), fr1 = fminf(fminf(fminf(tfx1, tfy1), tfz1), best.t);
          float df0, df1;
          fsub2(df0, df1, fr0, fr1, nr0, nr1);  // negative <=> near > far (x - x = +0; no NaN: every term is finite)
          acc = __funnelshift_l(__float_as_uint(df1), acc, 1);
          acc = __funnelshift_l(__float_as_uint(df0), acc, 1);
        }
        const uint32_t hitbits = ~acc;  // bit c set <=> child c is hit (bits 8.. are garbage, masked below)
        uint32_t hb = (hitbits & imask) | ((hitbits & lmask) << 8);
        // slot order -> visit order: bit s moves to position s ^ octinv (three conditional swaps)
        if (octinv & 4u) hb = ((hb & 0x0F0Fu) << 4) | ((hb >> 4) & 0x0F0Fu);
        if (octinv & 2u) hb = ((hb & 0x3333u) << 2) | ((hb >> 2) & 0x3333u);
        if (octinv & 1u) hb = ((hb & 0x5555u) << 1) | ((hb >> 1) & 0x5555u);
        igx = w0[4];
        igy = ((hb & 0xFFu) << 8) | imask;
#if RTC_POSTPONE
        if ((lgy >> 8) != 0) {  // the pending leaf group is parked (the slot is free: see want_node)
          pgx = lgx;
          pgy = lgy;
        }
#endif
        lgx = w0[5];
        lgy = (hb & 0xFF00u) | lmask;
      }
    } else {
      // ---- leaf step ---------------------------------------------------------------------------------------
      if (COUNT && (threadIdx.x & 31) == 0) n_leaf_steps++;
      if (want_leaf) {
        const uint32_t gx = has_p ? pgx : lgx, gy = has_p ? pgy : lgy;  // the parked (older, nearer) group first
        const uint32_t hits = gy >> 8;
        const uint32_t b = 31u - (uint32_t)__clz((int)hits);
        const uint32_t s = b ^ octinv;
        const uint32_t slot = gx + (uint32_t)__popc((gy & 0xFFu) & ((1u << s) - 1u));
        if (has_p)
          pgy = gy & ~(0x100u << b);
        else
          lgy = gy & ~(0x100u << b);
        if (COUNT) n_prims++;
        const PrimRec<R> pr = load_prim(sc, slot);
        Skip<R> sk;
        const uint2 s0 = lds64(sm + S_OXY * kStackStride), s3 = lds64(sm + S_DXY * kStackStride), s4 = lds64(sm + S_DZL * kStackStride),
                    s5 = lds64(sm + S_PATHSKIP * kStackStride);
        sk.code = s5.y;
        const V3<R> o = mk3(__uint_as_float(s0.x), __uint_as_float(s0.y), __uint_as_float(lds32(sm + S_OZIX * kStackStride)));
        const V3<R> d = mk3(__uint_as_float(s3.x), __uint_as_float(s3.y), __uint_as_float(s4.x));
        test_leaf<R>(sc, ref_of(pr), pr, R(0), o, d, __uint_as_float(s4.y), sk, src, s5.x, best);
      }
    }
    bool done_now = false;
#if RTC_POSTPONE
    if (sp >= 0 && (igy >> 8) == 0) {  // the next inner group is fetched even while leaf hits are pending
      if (sp > 0) {
        sp--;
        const uint2 g = lds64(sm + kStackBase + (uint32_t)sp * kStackStride);
        igx = g.x;
        igy = g.y;
      } else if (((lgy | pgy) >> 8) == 0) {
        sp = kFinished;
        done_now = true;
      }
    }
#else
    if (sp >= 0 && ((lgy | igy) >> 8) == 0) {
      if (sp > 0) {
        sp--;
        const uint2 g = lds64(sm + kStackBase + (uint32_t)sp * kStackStride);
        igx = g.x;
        igy = g.y;
      } else {
        sp = kFinished;
        done_now = true;
      }
    }
#endif
    lanes_changed = __any_sync(0xFFFFFFFFu, done_now) ? 1u : 0u;
  }
  if (COUNT) {
    atomicAdd(&io.ctl->nodes_visited, (unsigned long long)n_nodes);
    atomicAdd(&io.ctl->prims_tested, (unsigned long long)n_prims);
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(&io.ctl->node_steps, (unsigned long long)n_node_steps);
      atomicAdd(&io.ctl->leaf_steps, (unsigned long long)n_leaf_steps);
    }
  }
}

// DoubleColor.Luminance (DoubleColor.cs:76-81)
template <typename R>
__device__ __forceinline__ R luminance(R r, R g, R b) { return (R(0.299) * r + R(0.587) * g) + R(0.114) * b; }

// ---------------------------------------------------------------------------------------------------------
// Shading arithmetic of the f32 mode. Like the traversal (xrcp / xsqrt above), k_shade<float> contains no call: IEEE
// division, sqrtf, powf, acosf and sincosf all carry out-of-line slow paths whose calling convention costs registers
// around the whole kernel (80 -> 56 registers without them). Everything here is accurate to a few f32 ulp, three orders
// below the 1 % image tolerance; the f64 mode keeps the exact libm forms of the reference.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float xlog2(float a) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
}
__device__ __forceinline__ float xexp2(float a) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
}
__device__ __forceinline__ float xrsqrt(float a) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
}
// RandomShine's polar cosine z = u^(1/shininess) (Raytracer.cs:53) together with sqrt(1 - z^2), which CreateHorizon needs
// (Vec4D.cs:56). For mirror-like shininess (1e5, 1e6) z is within 1e-5 of 1 and 1 - z^2 computed from a rounded z has no
// significant bits left in f32, so the complement w = 1 - z is carried instead: for a small exponent e = ln(u) / shininess,
// w = -expm1(e) by its series, and 1 - z^2 = w (2 - w).
__device__ __forceinline__ void shine_polar(float u, float shininess, float& z, float& sn) {
  const float x = xlog2(u) * xrcp(shininess);  // log2 of the result; u = 0 gives -inf -> z = 0 like pow(0, p)
  const float e = x * 0.6931471805599453f;
  float w;
  if (fabsf(e) < 0.25f) {
    w = -e * fmaf(e, fmaf(e, fmaf(e, fmaf(e, fmaf(e, 1.0f / 720.0f, 1.0f / 120.0f), 1.0f / 24.0f), 1.0f / 6.0f), 0.5f), 1.0f);
    z = 1.0f - w;
  } else {
    z = xexp2(x);
    w = 1.0f - z;
  }
  sn = xsqrt(w * (2.0f - w));  // NaN for z > 1 (negative shininess), like the reference's sqrt(1 - z^2)
}
// acos on [0, 1) (the diffuse lobe's 2 acos(u) / pi, Raytracer.cs:215): sqrt(1 - x) * p(x), Abramowitz & Stegun 4.4.46,
// |error| <= 2e-8.
__device__ __forceinline__ float xacos01(float x) {
  float p = fmaf(x, -0.0012624911f, 0.0066700901f);
  p = fmaf(x, p, -0.0170881256f);
  p = fmaf(x, p, 0.0308918810f);
  p = fmaf(x, p, -0.0501743046f);
  p = fmaf(x, p, 0.0889789874f);
  p = fmaf(x, p, -0.2145988016f);
  p = fmaf(x, p, 1.5707963050f);
  return xsqrt(1.0f - x) * p;
}

// CreateHorizon with the polar sine handed in (f32 mode): same construction as create_horizon above, approximate rsqrt /
// sin / cos (angle = 2 pi u, u in [0, 1): one MUFU each after the range scaling the instruction does itself).
__device__ __forceinline__ V3<float> create_horizon_f32(const V3<float>& pole, float z, float sn, float u_theta) {
  V3<float> c = mk3(pole.y, -pole.x, 0.0f);  // pole x (0, 0, 1)
  const float l2 = c.x * c.x + c.y * c.y;
  if (l2 == 0.0f) {
    c = mk3(1.0f, 0.0f, 0.0f);
  } else {
    const float il = xrsqrt(l2);
    c = mk3(c.x * il, c.y * il, 0.0f);
  }
  const float theta = u_theta * 6.283185307179586f;
  const float s = __sinf(theta), co = __cosf(theta);
  const float cosOpp = 1.0f - co;
  const V3<float> v = (pole * z) + (c * sn);
  const V3<float>& a = pole;
  const float m00 = co + a.x * a.x * cosOpp, m01 = a.x * a.y * cosOpp - a.z * s, m02 = a.x * a.z * cosOpp + a.y * s;
  const float m10 = a.y * a.x * cosOpp + a.z * s, m11 = co + a.y * a.y * cosOpp, m12 = a.y * a.z * cosOpp - a.x * s;
  const float m20 = a.z * a.x * cosOpp - a.y * s, m21 = a.z * a.y * cosOpp + a.x * s, m22 = co + a.z * a.z * cosOpp;
  return mk3((m00 * v.x + m01 * v.y) + m02 * v.z, (m10 * v.x + m11 * v.y) + m12 * v.z, (m20 * v.x + m21 * v.y) + m22 * v.z);
}

// The material of one slot: colours, ior, shininess (Primitive.cs:16-129; Specular / Refraction already black when the
// primitive is not reflective). f64: four 32-byte vectors; f32: one 256-bit load of 12 halfs + 2 floats.
template <typename R>
struct Mat {
  V3<R> emis, diff, spec, refr;
  R ior, shininess;
};
__device__ __forceinline__ void unpack_half2(uint32_t w, float& lo, float& hi) {
  asm("{.reg .b16 l, h; mov.b32 {l, h}, %2; cvt.f32.f16 %0, l; cvt.f32.f16 %1, h;}" : "=f"(lo), "=f"(hi) : "r"(w));
}
__device__ __forceinline__ Mat<float> load_mat(const DMat<float>* mp) {
  uint32_t w[8];
  ldg256(mp, w);
  Mat<float> m;
  unpack_half2(w[0], m.emis.x, m.emis.y);
  unpack_half2(w[1], m.emis.z, m.diff.x);
  unpack_half2(w[2], m.diff.y, m.diff.z);
  unpack_half2(w[3], m.spec.x, m.spec.y);
  unpack_half2(w[4], m.spec.z, m.refr.x);
  unpack_half2(w[5], m.refr.y, m.refr.z);
  m.ior = __uint_as_float(w[6]);
  m.shininess = __uint_as_float(w[7]);
  return m;
}
__device__ __forceinline__ Mat<double> load_mat(const DMat<double>* mp) {
  const V4<double> e = ldg4(&mp->emis_ior), d = ldg4(&mp->diff_shin), sp = ldg4(&mp->spec), rf = ldg4(&mp->refr);
  Mat<double> m;
  m.emis = xyz(e);
  m.diff = xyz(d);
  m.spec = xyz(sp);
  m.refr = xyz(rf);
  m.ior = e.w;
  m.shininess = d.w;
  return m;
}

// Completes the trace kernel's answer (distance, code) into position and normal of the Hit (Hit.cs:14-20). f32 mode: flat
// triangles and plain spheres -- all of BASELINE's synthetic scenes -- need only the 16-byte sgeom record: the position is
// origin + t * direction (Sphere.cs:94,97 compute exactly that; for a triangle it equals Triangle.cs:130's v0 + u e1 + v e2
// to f32 rounding), the normal is +-N or +-(P - C) / r by the primitive's own inside flag. Planes, transformed spheres and
// vertex-normal triangles, and everything in f64 mode, re-evaluate the primitive (finalize_hit).
template <typename R>
__device__ __forceinline__ void complete_hit(const SceneView<R>& sc, uint32_t code, const V3<R>& o, const V3<R>& d, R dlen, R t_in,
                                             V3<R>& pos, V3<R>& normal, R& t) {
  if constexpr (!Num<R>::is_f64) {
    const uint32_t kind = (code >> HIT_KIND_SHIFT) & 3u;
    const bool raw_inside = (((code >> 30) ^ (code >> 28)) & 1u) != 0;
    if (kind == DK_TRI || kind == DK_SPHERE) {
      const V4<R> sg = ldg4(&sc.sgeom[code & REF_SLOT_MASK]);
      if (kind == DK_SPHERE || code_of(sg.w) == 0u) {
        t = t_in;
        pos = mk3(rfma(t_in, d.x, o.x), rfma(t_in, d.y, o.y), rfma(t_in, d.z, o.z));
        V3<R> n = xyz(sg);
        if (kind == DK_SPHERE) {
          const R ir = xrcp(sg.w);
          n = mk3((pos.x - sg.x) * ir, (pos.y - sg.y) * ir, (pos.z - sg.z) * ir);
        }
        normal = raw_inside ? neg3(n) : n;
        return;
      }
    }
  }
  finalize_hit<R>(sc, code, o, d, dlen, pos, normal, t);
}

// shade: one bounce of Raytracer.GetColor (Raytracing/Raytracer.cs:71-245) for every live path: terminal tests,
// RandomShine, Fresnel / total internal reflection, lobe roulette, next ray, tint. Terminated paths write their
// radiance; the survivors overwrite their ray (hpos = origin, dir) and tint in place and are appended to the next bounce's
// queue by a warp-aggregated stream compaction (ballot + popc + one atomicAdd per warp) -- the loop's `break`/`return` of
// the reference.
template <typename R>
__global__ void __launch_bounds__(kShadeThreads, RTC_SHADE_MIN_BLOCKS) k_shade(SceneView<R> sc, ParamsView<R> par, Band band, ShadeIO<R> io,
                                                           int bounce) {
  constexpr bool F64 = Num<R>::is_f64;
  const uint32_t count = *io.count;
  const uint32_t* queue = io.queue;
  uint32_t* qout = io.queue_out;
  uint32_t* out_count = io.count_out;
  const int identity_queue = io.queue == nullptr;
  const uint32_t stride = gridDim.x * blockDim.x;
  const uint32_t rounds = (count + stride - 1) / stride;  // warp-uniform trip count: every lane takes part in the ballots
  uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  for (uint32_t rnd = 0; rnd < rounds; rnd++, idx += stride) {
    bool alive = false;
    uint32_t path = 0;
    if (idx < count) {
    path = identity_queue ? idx : queue[idx];
    const THit<R> th = io.thit[path];
    const uint32_t code = th.code;
    bool done = false;
    R out_r = 0, out_g = 0, out_b = 0;
    int dbg = 0;  // BounceType.Skipped
    R dbg_f = Num<R>::nan();
    if (code == HIT_MISS) {  // :81-91
      dbg = 8;
      if (bounce == 0) {
        out_r = out_g = out_b = R(-1);  // DoubleColor.Placeholder
      } else {
        out_r = par.ambient[0];
        out_g = par.ambient[1];
        out_b = par.ambient[2];
      }
      done = true;
    } else {
      // every load that depends on (path, slot) only is issued up front, behind one another: the kernel is bound by gather
      // latency, not by arithmetic
      const V4<R> op = ld4(&io.hpos[path]);
      const V4<R> dv = ld4(&io.dir[path]);
      const V4<R> tv = ld4(&io.tint[path]);
      const Mat<R> m = load_mat(sc.mats + (code & REF_SLOT_MASK));
      const V3<R> d = xyz(dv);
      V3<R> pos, normal;
      R hit_t;
      complete_hit<R>(sc, code, xyz(op), d, F64 ? R(0) : dv.w, th.t, pos, normal, hit_t);  // also the next ray's origin and skip hit
      const bool hit_inside = (code & HIT_INSIDE) != 0;
      bool have_out = false;
      V3<R> out_dir = mk3(R(0), R(0), R(0));
      R nt_r = 0, nt_g = 0, nt_b = 0, total_l = 0;

      if (par.debug_geom) {  // :93-98
        dbg = 9;
        out_r = (m.spec.x + m.diff.x) + m.emis.x;
        out_g = (m.spec.y + m.diff.y) + m.emis.y;
        out_b = (m.spec.z + m.diff.z) + m.emis.z;
        done = true;
      } else if (bounce >= par.recursion) {  // :100-104
        dbg = 7;
      } else {
        int x, y;
        uint32_t sample;
        band_pixel(band, path, x, y, sample);
        const uint32_t pixel = (uint32_t)(y * par.width + x);
        const uint32_t stage = 1u + (uint32_t)bounce;
        R u1, u2;
        uniforms2<R>(par.seed_lo, par.seed_hi, pixel, sample, stage, 0, u1, u2);
        // RandomShine, :51-56
        V3<R> rough;
        if constexpr (F64) {
          R zs = (m.shininess == Num<R>::inf()) ? R(1) : rpow(u1, 1 / m.shininess);
          rough = create_horizon(normal, zs, u2 * R(3.14159265358979323846) * 2);  // :108
        } else {
          float zs = 1.0f, sn = 0.0f;
          if (m.shininess != Num<R>::inf()) shine_polar(u1, m.shininess, zs, sn);
          rough = create_horizon_f32(normal, zs, sn, u2);
        }
        R diff_l = luminance(m.diff.x, m.diff.y, m.diff.z), spec_l = luminance(m.spec.x, m.spec.y, m.spec.z),
          refr_l = luminance(m.refr.x, m.refr.y, m.refr.z), emis_l = luminance(m.emis.x, m.emis.y, m.emis.z);  // :110-113
        R cosv = -dot3(rough, d);  // :115
        R cos_out = 0, ior_ratio = 0;
        if (((refr_l > 0) | (spec_l > 0)) && m.ior != 0 && cosv >= 0) {  // :120
          R ior_in, ior_out;
          if (hit_inside) {
            ior_in = m.ior;
            ior_out = par.air_ior;
          } else {
            ior_in = par.air_ior;
            ior_out = m.ior;
          }
          ior_ratio = F64 ? ior_in / ior_out : xdiv(ior_in, ior_out);   // :136
          R sin_out = ior_ratio * xsqrt(1 - (cosv * cosv));             // :137
          if (sin_out >= 1) {                                           // :140-145
            refr_l = 0;
            dbg_f = 1;
          } else {
            cos_out = xsqrt(1 - (sin_out * sin_out));                   // :148
            const R rs_n = (ior_out * cosv) - (ior_in * cos_out), rs_d = (ior_out * cosv) + (ior_in * cos_out);  // :149
            const R rp_n = (ior_in * cosv) - (ior_out * cos_out), rp_d = (ior_in * cosv) + (ior_out * cos_out);  // :150
            const R rs = F64 ? rs_n / rs_d : xdiv(rs_n, rs_d), rp = F64 ? rp_n / rp_d : xdiv(rp_n, rp_d);
            R ratio = ((rs * rs) + (rp * rp)) * R(0.5);                 // :151 (x / 2 == x * 0.5 exactly)
            spec_l *= ratio;
            refr_l *= 1 - ratio;
            dbg_f = ratio;
          }
        } else {
          refr_l = 0;  // :158-161
        }
        total_l = ((diff_l + spec_l) + refr_l) + emis_l;  // :163
        if (total_l <= 0) {  // :165-169
          dbg = 6;
        } else {
          R u3, u4;
          uniforms2<R>(par.seed_lo, par.seed_hi, pixel, sample, stage, 1, u3, u4);
          R ray_rand = u3 * total_l;  // :178
          if (refr_l != 0 && (ray_rand -= refr_l) <= 0) {  // :181-193
            dbg = 4;
            out_dir = (rough * -cos_out) + ((d + (rough * cosv)) * ior_ratio);
            have_out = true;
            if (hit_inside) {
              nt_r = nt_g = nt_b = 1;
            } else {
              nt_r = m.refr.x; nt_g = m.refr.y; nt_b = m.refr.z;
            }
          } else if (spec_l != 0 && (ray_rand -= spec_l) <= 0) {  // :194-209
            dbg = 3;
            V3<R> od = d + (rough * (cosv * 2));  // Reflection, :58-61
            if (dot3(od, normal) > 0) {
              dbg = 2;
              out_dir = od;
              have_out = true;
              nt_r = m.spec.x; nt_g = m.spec.y; nt_b = m.spec.z;
            }
          } else if (diff_l != 0 && (ray_rand -= diff_l) <= 0) {  // :210-219
            dbg = 1;
            R u5, u6;
            uniforms2<R>(par.seed_lo, par.seed_hi, pixel, sample, stage, 2, u5, u6);
            if constexpr (F64) {
              R z = (2 * racos(u4)) / R(3.14159265358979323846);
              out_dir = create_horizon(normal, z, u5 * R(3.14159265358979323846) * 2);
            } else {
              const float z = xacos01(u4) * 0.6366197723675814f;  // 2 acos(u) / pi
              out_dir = create_horizon_f32(normal, z, xsqrt(1.0f - z * z), u5);
            }
            have_out = true;
            nt_r = m.diff.x; nt_g = m.diff.y; nt_b = m.diff.z;
          } else {
            dbg = 5;  // :220-229 emission picked
          }
          // :231-232 `outRay == Ray.Zero`
          if (have_out && pos.x == 0 && pos.y == 0 && pos.z == 0 && out_dir.x == 0 && out_dir.y == 0 && out_dir.z == 0)
            have_out = false;
        }
      }
      if (have_out) {
        R mx = risnan(total_l) ? total_l : (total_l > 1 ? total_l : R(1));  // Math.Max(totalLum, 1), :238
        R tr = tv.x * (nt_r * mx), tg = tv.y * (nt_g * mx), tb = tv.z * (nt_b * mx);  // :238-240
        if constexpr (F64) {
          if ((bounce + 1) % 3 == 0) out_dir = normalize3(out_dir);  // :74-75 of the next iteration
        } else {
          // f32 mode keeps every ray direction at unit length: the reference's sphere test silently assumes it (Sphere.cs:80-90),
          // and the drift it tolerates between its every-third-bounce normalisations is far below f32 resolution anyway
          const float il = xrsqrt((out_dir.x * out_dir.x + out_dir.y * out_dir.y) + out_dir.z * out_dir.z);
          out_dir = out_dir * il;
        }
        st4(&io.dir[path], out_dir.x, out_dir.y, out_dir.z, R(0));
        st4(&io.tint[path], tr, tg, tb, R(0));
      } else if (!done) {
        out_r = tv.x * m.emis.x;  // :245
        out_g = tv.y * m.emis.y;
        out_b = tv.z * m.emis.z;
        done = true;
      }
      if (have_out || io.dbg_type) {  // the completed Hit: the next ray's origin and skip hit (and rtc_debug_trace's record)
        R w;
        set_code(w, code);
        st4(&io.hpos[path], pos.x, pos.y, pos.z, F64 ? hit_t : w);
        // the skip hit's normal: the f32 traversal reads it only for the positional self-hit rule of spheres (a flat primitive equal
        // to the skip primitive is always the self-hit, and the skip code travels in hpos.w): 16 bytes per survivor not written
        const uint32_t hk = (code >> HIT_KIND_SHIFT) & 3u;
        if (F64 || io.dbg_type || hk == DK_SPHERE || hk == DK_XSPHERE) st4(&io.hnrm[path], normal.x, normal.y, normal.z, w);
      }
    }
    if (io.dbg_type) {
      io.dbg_type[path] = dbg;
      io.dbg_fresnel[path] = dbg_f;
    }
    if (done)
      st4(&io.radiance[path], out_r, out_g, out_b, R(0));
    alive = !done;
    }
    const uint32_t mask = __ballot_sync(0xFFFFFFFFu, alive);
    if (mask) {
      const int lane = threadIdx.x & 31;
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(out_count, (uint32_t)__popc(mask));
      base = __shfl_sync(0xFFFFFFFFu, base, 0);
      if (alive) qout[base + __popc(mask & ((1u << lane) - 1u))] = path;
    }
  }
}

// Bookkeeping between bounces (one thread): ray counter, reset of the trace cursor, zeroing of the queue length just
// consumed (the compaction itself is fused into k_shade).
static __global__ void k_end_bounce(Control* ctl, int q) {
  ctl->rays += ctl->count[q];
  ctl->count[q] = 0;
  ctl->work_trace = 0;
}

// accumulate: FullRaytracer's drain loop (FullRaytracer.cs:326-339) + SampleSet.AddSample/AddMiss (SampleSet.cs:32-44),
// one thread per pixel of the band, samples added in ascending order (deterministic, no atomics).
template <typename R>
__global__ void __launch_bounds__(kStreamThreads) k_accumulate(ParamsView<R> par, Band band, PathView<R> pv, double* rgb_sum,
                                                                uint32_t* samples, uint32_t* misses) {
  uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= band.n_pix) return;
  int px, py;
  band_pix_xy(band, pix, px, py);
  size_t g = (size_t)py * par.width + px;
  double r = rgb_sum[g * 3 + 0], gg = rgb_sum[g * 3 + 1], b = rgb_sum[g * 3 + 2];
  uint32_t ns = samples[g], nm = misses[g];
  for (uint32_t s = 0; s < band.n_samples; s++) {
    V4<R> c = ld4(&pv.radiance[(size_t)s * band.n_pix + pix]);
    if (c.x == R(-1) && c.y == R(-1) && c.z == R(-1)) {  // == DoubleColor.Placeholder -> AddMiss
      nm++;
    } else {
      r += (double)c.x;
      gg += (double)c.y;
      b += (double)c.z;
      ns++;
    }
  }
  rgb_sum[g * 3 + 0] = r;
  rgb_sum[g * 3 + 1] = gg;
  rgb_sum[g * 3 + 2] = b;
  samples[g] = ns;
  misses[g] = nm;
}

// ---- rtc_trace_closest plumbing --------------------------------------------------------------------------
template <typename R>
__global__ void __launch_bounds__(kStreamThreads) k_import_rays(SceneView<R> sc, int64_t n, const rtc_ray* rays,
                                                                 const rtc_hit* skip, const int32_t* id_to_slot,
                                                                 PathView<R> pv) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) {
    pv.ctl->count[0] = (uint32_t)n;
    pv.ctl->count[1] = 0;
    pv.ctl->work_trace = 0;
  }
  if (i >= n) return;
  const rtc_ray& r = rays[i];
  // f32 mode: how far the f64 direction is off unit length (the reference's sphere test assumes it is not, Sphere.cs:80-90;
  // sphere_hits needs the defect to land on the reference's roots, and f32 components cannot carry one below 1e-7)
  const double dl = ((r.dir[0] * r.dir[0] + r.dir[1] * r.dir[1]) + r.dir[2] * r.dir[2]) - 1.0;
  st4(&pv.dir[i], (R)r.dir[0], (R)r.dir[1], (R)r.dir[2], Num<R>::is_f64 ? R(0) : (R)dl);
  R t = 0, w;
  R nx = 0, ny = 0, nz = 0, px = 0, py = 0, pz = 0;
  uint32_t code = HIT_MISS;
  if (skip && skip[i].prim >= 0 && skip[i].prim < sc.n_prims) {
    const rtc_hit& s = skip[i];
    code = (uint32_t)id_to_slot[s.prim] | (s.inside ? HIT_INSIDE : 0u);
    t = (R)s.t;
    nx = (R)s.normal[0]; ny = (R)s.normal[1]; nz = (R)s.normal[2];
    px = (R)s.position[0]; py = (R)s.position[1]; pz = (R)s.position[2];
  }
  set_code(w, code);
  st4(&pv.hpos[i], (R)r.origin[0], (R)r.origin[1], (R)r.origin[2], Num<R>::is_f64 ? t : w);
  st4(&pv.hnrm[i], nx, ny, nz, w);
  st4(&pv.skip_pos[i], px, py, pz, R(0));
}

// finalize != 0: straight after a trace launch (rtc_trace_closest) -- the Hit record is completed here from the ray still in
// hpos / dir; finalize == 0: after k_shade (rtc_debug_trace), which has already stored position and normal.
template <typename R>
__global__ void __launch_bounds__(kStreamThreads) k_export_hits(SceneView<R> sc, int64_t n, PathView<R> pv, rtc_hit* out,
                                                                 int finalize) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const THit<R> th = pv.thit[i];
  const uint32_t code = th.code;
  V4<R> hp = ld4(&pv.hpos[i]);
  V3<R> fp = xyz(hp), fn = mk3(R(0), R(0), R(0));
  R ft = th.t;
  if (code != HIT_MISS) {
    if (finalize) {
      const V4<R> dv = ld4(&pv.dir[i]);
      finalize_hit<R>(sc, code, xyz(hp), xyz(dv), Num<R>::is_f64 ? R(0) : dv.w, fp, fn, ft);
    } else {
      fn = xyz(ld4(&pv.hnrm[i]));
    }
  }
  rtc_hit h;
  if (code == HIT_MISS) {
    h.prim = -1;
    h.inside = 0;
    h.t = 0;
    h.position[0] = h.position[1] = h.position[2] = 0;
    h.normal[0] = h.normal[1] = h.normal[2] = 0;
  } else {
    h.prim = sc.prim_id[code & REF_SLOT_MASK];
    h.inside = (code & HIT_INSIDE) ? 1 : 0;
    h.t = (double)ft;
    h.position[0] = (double)fp.x; h.position[1] = (double)fp.y; h.position[2] = (double)fp.z;
    h.normal[0] = (double)fn.x; h.normal[1] = (double)fn.y; h.normal[2] = (double)fn.z;
  }
  out[i] = h;
}

template <typename R>
__global__ void __launch_bounds__(kStreamThreads) k_export_radiance(Band band, ParamsView<R> par, PathView<R> pv, double* out_rgb) {
  uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= band.n_pix) return;
  int px, py;
  band_pix_xy(band, pix, px, py);
  size_t g = (size_t)py * par.width + px;
  V4<R> c = ld4(&pv.radiance[pix]);
  out_rgb[g * 3 + 0] = (double)c.x;
  out_rgb[g * 3 + 1] = (double)c.y;
  out_rgb[g * 3 + 2] = (double)c.z;
}

// ---- DebugRaycaster overlay (Raytracing/DebugRaycaster.cs:170-265) -----------------------------------------
template <typename R>
__global__ void __launch_bounds__(kStreamThreads) k_overlay_rays(CameraView<R> cam, ParamsView<R> par, Band band, PathView<R> pv) {
  uint32_t path = blockIdx.x * blockDim.x + threadIdx.x;
  if (path == 0) {
    pv.ctl->count[0] = band.n_paths;
    pv.ctl->count[1] = 0;
    pv.ctl->work_trace = 0;
  }
  if (path >= band.n_paths) return;
  int x, y;
  uint32_t sample;
  band_pixel(band, path, x, y, sample);
  V3<R> o, d;
  camera_get_ray(cam, R(x), R(y), o, d);  // camera.GetRay(x, y)
  o = o + (d * cam.image_plane);          // .Offset(camera.imagePlane), DebugRaycaster.cs:236
  R w;
  set_code(w, HIT_MISS);
  st4(&pv.dir[path], d.x, d.y, d.z, R(0));
  st4(&pv.hpos[path], o.x, o.y, o.z, Num<R>::is_f64 ? R(0) : w);
  if (Num<R>::is_f64) st4(&pv.hnrm[path], R(0), R(0), R(0), w);
}

template <typename R>
__global__ void __launch_bounds__(kStreamThreads) k_overlay_prims(SceneView<R> sc, ParamsView<R> par, Band band, PathView<R> pv,
                                                                   int32_t* out) {
  uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= band.n_pix) return;
  int px, py;
  band_pix_xy(band, pix, px, py);
  size_t g = (size_t)py * par.width + px;
  const uint32_t code = pv.thit[pix].code;
  out[g] = code == HIT_MISS ? -1 : sc.prim_id[code & REF_SLOT_MASK];
}

// ---------------------------------------------------------------------------------------------------------
// Launchers
// ---------------------------------------------------------------------------------------------------------
inline int div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

template <typename R>
int Kernels<R>::trace_blocks_per_sm(size_t smem) {
  int nb = 0;
  if constexpr (Num<R>::is_f64)
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_trace<R, false>, kTraceThreads, smem);
  else
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_trace_q8<false>, kTraceThreads, smem);
  return nb > 0 ? nb : 1;
}

template <typename R>
cudaError_t Kernels<R>::raygen(const LaunchCfg& cfg, const CameraView<R>& cam, const ParamsView<R>& par, const Band& band,
                               const PathView<R>& pv) {
  k_raygen<R><<<div_up(band.n_paths, kStreamThreads), kStreamThreads, 0, cfg.stream>>>(cam, par, band, pv);
  return cudaGetLastError();
}

template <typename R>
cudaError_t Kernels<R>::camera_rays(const LaunchCfg& cfg, const CameraView<R>& cam, const ParamsView<R>& par, int64_t n,
                                    const int32_t* xy, const uint32_t* sample, rtc_ray* out) {
  if (n == 0) return cudaSuccess;
  k_camera_rays<R><<<div_up(n, kStreamThreads), kStreamThreads, 0, cfg.stream>>>(cam, par, n, xy, sample, out);
  return cudaGetLastError();
}

// Vec4D.CreateHorizon (Vec4D.cs:33-58) for n (pole, z, theta) tuples, as k_shade evaluates it in this arithmetic mode
__device__ __forceinline__ V3<double> horizon_of(const V3<double>& pole, double z, double theta) { return create_horizon<double>(pole, z, theta); }
__device__ __forceinline__ V3<float> horizon_of(const V3<float>& pole, float z, float theta) {
  return create_horizon_f32(pole, z, xsqrt(1.0f - z * z), theta * 0.15915494309189535f);
}
template <typename R>
__global__ void k_horizon(int64_t n, const double* in, double* out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const V3<R> v = horizon_of(mk3((R)in[5 * i], (R)in[5 * i + 1], (R)in[5 * i + 2]), (R)in[5 * i + 3], (R)in[5 * i + 4]);
  out[3 * i] = v.x;
  out[3 * i + 1] = v.y;
  out[3 * i + 2] = v.z;
}
template <typename R>
cudaError_t Kernels<R>::horizon(const LaunchCfg& cfg, int64_t n, const double* in, double* out) {
  if (n == 0) return cudaSuccess;
  k_horizon<R><<<div_up(n, kStreamThreads), kStreamThreads, 0, cfg.stream>>>(n, in, out);
  return cudaGetLastError();
}

template <typename R>
cudaError_t Kernels<R>::trace(const LaunchCfg& cfg, const SceneView<R>& sc, const PathView<R>& pv, int q, bool identity_queue) {
  const size_t smem = Num<R>::is_f64 ? (size_t)sc.q_stack * kTraceThreads * (4 + sizeof(R))
                                     : (size_t)(sc.q_stack + kQ8StateSlots) * kTraceThreads * sizeof(uint2);
  // Function attributes and occupancy are per (device, kernel) for the whole process, whatever thread launches: the largest
  // dynamic shared-memory size asked for so far is kept per device (raised, never lowered: a smaller launch is always legal
  // under a larger limit) and the resident-CTA count per (device, size), under one mutex.
  static std::mutex mu;
  static std::map<int, size_t> limit_set;
  static std::map<std::pair<int, size_t>, int> resident;
  int dev = 0;
  cudaGetDevice(&dev);
  int per_sm = 1;
  {
    std::lock_guard<std::mutex> g(mu);
    size_t& lim = limit_set[dev];
    if (smem > lim && (Num<R>::is_f64 || smem > 48 * 1024)) {  // beyond the default 48 KB of dynamic shared memory: opt in
      cudaError_t e;
      if constexpr (Num<R>::is_f64) {
        e = cudaFuncSetAttribute(k_trace<R, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_trace<R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      } else {
        e = cudaFuncSetAttribute(k_trace_q8<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_trace_q8<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      }
      if (e != cudaSuccess) return e;  // (a tree too deep for the SM's shared memory surfaces here, through fail())
      lim = smem;
    }
    auto it = resident.find({dev, smem});
    if (it == resident.end()) it = resident.emplace(std::make_pair(dev, smem), trace_blocks_per_sm(smem)).first;
    per_sm = it->second;
  }
  if (const char* cap = std::getenv("RTC_TRACE_BLOCKS_PER_SM"))  // tuning aid: fewer resident CTAs than fit
    per_sm = std::max(1, std::min(per_sm, std::atoi(cap)));
  int grid = cfg.sm_count * per_sm;
  TraceIO<R> io;
  io.dir = pv.dir;
  io.in_hpos = pv.hpos;
  io.in_hnrm = pv.hnrm;
  io.skip_pos = pv.skip_pos;
  io.out = pv.thit;
  io.queue = identity_queue ? nullptr : pv.queue[q];
  io.count = &pv.ctl->count[q];
  io.ctl = pv.ctl;
  if constexpr (Num<R>::is_f64) {
    if (cfg.counters)
      k_trace<R, true><<<grid, kTraceThreads, smem, cfg.stream>>>(sc, io);
    else
      k_trace<R, false><<<grid, kTraceThreads, smem, cfg.stream>>>(sc, io);
  } else {
    if (cfg.counters)
      k_trace_q8<true><<<grid, kTraceThreads, smem, cfg.stream>>>(sc, io);
    else
      k_trace_q8<false><<<grid, kTraceThreads, smem, cfg.stream>>>(sc, io);
  }
  return cudaGetLastError();
}

template <typename R>
cudaError_t Kernels<R>::shade(const LaunchCfg& cfg, const SceneView<R>& sc, const ParamsView<R>& par, const Band& band,
                              const PathView<R>& pv, int q, int bounce, bool identity_queue) {
  int grid = cfg.sm_count * 8;
  ShadeIO<R> io;
  io.dir = pv.dir;
  io.tint = pv.tint;
  io.thit = pv.thit;
  io.hpos = pv.hpos;
  io.hnrm = pv.hnrm;
  io.radiance = pv.radiance;
  io.queue = identity_queue ? nullptr : pv.queue[q];
  io.count = &pv.ctl->count[q];
  io.queue_out = pv.queue[q ^ 1];
  io.count_out = &pv.ctl->count[q ^ 1];
  io.dbg_type = pv.dbg_type;
  io.dbg_fresnel = pv.dbg_fresnel;
  k_shade<R><<<grid * (kStreamThreads / kShadeThreads), kShadeThreads, 0, cfg.stream>>>(sc, par, band, io, bounce);
  return cudaGetLastError();
}

template <typename R>
cudaError_t Kernels<R>::compact(const LaunchCfg& cfg, const PathView<R>& pv, int q, bool) {
  k_end_bounce<<<1, 1, 0, cfg.stream>>>(pv.ctl, q);
  return cudaGetLastError();
}

template <typename R>
cudaError_t Kernels<R>::accumulate(const LaunchCfg& cfg, const ParamsView<R>& par, const Band& band, const PathView<R>& pv,
                                   double* rgb_sum, uint32_t* samples, uint32_t* misses) {
  k_accumulate<R><<<div_up(band.n_pix, kStreamThreads), kStreamThreads, 0, cfg.stream>>>(par, band, pv, rgb_sum, samples, misses);
  return cudaGetLastError();
}

template <typename R>
cudaError_t Kernels<R>::import_rays(const LaunchCfg& cfg, const SceneView<R>& sc, int64_t n, const rtc_ray* rays,
                                    const rtc_hit* skip, const int32_t* id_to_slot, const PathView<R>& pv) {
  k_import_rays<R><<<div_up(n, kStreamThreads), kStreamThreads, 0, cfg.stream>>>(sc, n, rays, skip, id_to_slot, pv);
  return cudaGetLastError();
}

template <typename R>
cudaError_t Kernels<R>::export_hits(const LaunchCfg& cfg, const SceneView<R>& sc, int64_t n, const PathView<R>& pv,
                                    rtc_hit* out, bool finalize) {
  k_export_hits<R><<<div_up(n, kStreamThreads), kStreamThreads, 0, cfg.stream>>>(sc, n, pv, out, finalize ? 1 : 0);
  return cudaGetLastError();
}

template <typename R>
cudaError_t Kernels<R>::overlay_rays(const LaunchCfg& cfg, const CameraView<R>& cam, const ParamsView<R>& par, const Band& band,
                                     const PathView<R>& pv) {
  k_overlay_rays<R><<<div_up(band.n_paths, kStreamThreads), kStreamThreads, 0, cfg.stream>>>(cam, par, band, pv);
  return cudaGetLastError();
}

template <typename R>
cudaError_t Kernels<R>::overlay_prims(const LaunchCfg& cfg, const SceneView<R>& sc, const ParamsView<R>& par, const Band& band,
                                      const PathView<R>& pv, int32_t* out) {
  k_overlay_prims<R><<<div_up(band.n_pix, kStreamThreads), kStreamThreads, 0, cfg.stream>>>(sc, par, band, pv, out);
  return cudaGetLastError();
}

template <typename R>
cudaError_t Kernels<R>::export_radiance(const LaunchCfg& cfg, const Band& band, const ParamsView<R>& par, const PathView<R>& pv,
                                        double* out_rgb) {
  k_export_radiance<R><<<div_up(band.n_pix, kStreamThreads), kStreamThreads, 0, cfg.stream>>>(band, par, pv, out_rgb);
  return cudaGetLastError();
}

}  // namespace rtc
