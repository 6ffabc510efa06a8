// rtc_oracle.cpp — CPU parity oracle (TEST INFRASTRUCTURE; see rtc_oracle.h for the rules of use).
//
// A from-scratch restatement, in IEEE f64, of the reference CPU render path of Zaggy1024/RaytracerCore.
// Every function cites the reference lines it follows (paths relative to RaytracerCore/). The reference's
// taken path on any AVX2+FMA host is the AVX one (Vectors/SIMDHelpers.cs:15), so the summation orders and the
// places where an FMA is used below are the AVX path's. Build with -ffp-contract=off: every fused operation
// in this file is an explicit std::fma.
//
// Pinned against the reference's two published renders (tests/test_screenshots.py) and analytic known answers; per-ray
// quantities remain UNPINNED by the reference (it has no tests / golden vectors and cannot run here); see rtc_oracle.h.

#include "rtc_oracle.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace {

constexpr double kInf = std::numeric_limits<double>::infinity();
constexpr double kNaN = std::numeric_limits<double>::quiet_NaN();
constexpr double kNearEnough = 1e-24;  // Util.cs:18
constexpr double kPi = 3.14159265358979323846;  // Math.PI

// ------------------------------------------------------------------------------------------------------
// Vec4D (Vectors/Vec4D.cs). The class-level SIMD switches are off except Normalize (Vec4D.cs:20-26), so the
// arithmetic below is the scalar code of the operators, evaluated left to right without contraction.
// ------------------------------------------------------------------------------------------------------
struct V4 {
  double x, y, z, w;
};

inline V4 v4(double x, double y, double z, double w) { return V4{x, y, z, w}; }
inline V4 operator+(const V4& a, const V4& b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }  // Vec4D.cs:93-112
inline V4 operator-(const V4& a, const V4& b) { return {a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }  // :114-133
inline V4 operator*(const V4& a, double s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }             // :178-198
inline V4 operator/(const V4& a, double s) { return {a.x / s, a.y / s, a.z / s, a.w / s}; }             // :227-247
inline V4 neg(const V4& a) { return {-a.x, -a.y, -a.z, -a.w}; }                                        // :156-176
inline double dot(const V4& a, const V4& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }    // :341-347
inline V4 cross(const V4& a, const V4& b) {                                                            // :355-364
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x, 0};
}
inline double sqlen(const V4& a) { return a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w; }              // :280-298
inline double length(const V4& a) { return std::sqrt(sqlen(a)); }                                     // :303-321
inline bool eq3(const V4& a, const V4& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }            // :463-481

// SIMDHelpers (Vectors/SIMDHelpers.cs)
inline double sdot(const V4& a, const V4& b) {  // PreDot + Add2, :70-100: (x+y)+(z+w)
  return (a.x * b.x + a.y * b.y) + (a.z * b.z + a.w * b.w);
}
inline V4 scross(const V4& a, const V4& b) {  // Cross, :44-61: Fma.MultiplySubtract(leftA, rightA, leftB*rightB)
  return {std::fma(a.y, b.z, -(a.z * b.y)), std::fma(a.z, b.x, -(a.x * b.z)), std::fma(a.x, b.y, -(a.y * b.x)),
          std::fma(a.w, b.w, -(a.w * b.w))};
}
inline V4 snormalize(const V4& a) {  // Normalize, :332-335 -> Length4 -> LengthSquared4 -> Sum4(v), :134-143
  double s = (a.x * a.x + a.y * a.y) + (a.z * a.z + a.w * a.w);
  double l = std::sqrt(s);
  return {a.x / l, a.y / l, a.z / l, a.w / l};
}
inline V4 normalize(const V4& a) { return snormalize(a); }  // Vec4D.Normalize, Vec4D.cs:326-340 (SIMDNormalize = true)

// Mat4x4D (Vectors/Mat4x4D.cs), row-major D00..D33. mat*vec goes through SIMDHelpers.MultiplyMatrixVector
// (SIMDHelpers.cs:222-237) whose Sum4 (:111-127) yields (p0+p1)+(p2+p3) per component.
struct M4 {
  double m[16];
};
inline V4 matvec(const M4& M, const V4& v) {  // Mat4x4D.cs:171-180
  const double* m = M.m;
  return {(m[0] * v.x + m[1] * v.y) + (m[2] * v.z + m[3] * v.w), (m[4] * v.x + m[5] * v.y) + (m[6] * v.z + m[7] * v.w),
          (m[8] * v.x + m[9] * v.y) + (m[10] * v.z + m[11] * v.w),
          (m[12] * v.x + m[13] * v.y) + (m[14] * v.z + m[15] * v.w)};
}
inline M4 rotate(double angle, const V4& axis) {  // MatrixTransforms.Rotate, MatrixTransforms.cs:25-38
  double c = std::cos(angle);
  double s = std::sin(angle);
  double co = 1 - c;
  M4 r;
  double* m = r.m;
  m[0] = c + axis.x * axis.x * co;
  m[1] = axis.x * axis.y * co - axis.z * s;
  m[2] = axis.x * axis.z * co + axis.y * s;
  m[3] = 0;
  m[4] = axis.y * axis.x * co + axis.z * s;
  m[5] = c + axis.y * axis.y * co;
  m[6] = axis.y * axis.z * co - axis.x * s;
  m[7] = 0;
  m[8] = axis.z * axis.x * co - axis.y * s;
  m[9] = axis.z * axis.y * co + axis.x * s;
  m[10] = c + axis.z * axis.z * co;
  m[11] = 0;
  m[12] = 0;
  m[13] = 0;
  m[14] = 0;
  m[15] = 1;
  return r;
}

V4 create_horizontal(const V4& v) {  // Vec4D.CreateHorizontal, Vec4D.cs:33-43
  V4 c = cross(v, v4(0, 0, 1, 0));
  if (eq3(c, v4(0, 0, 0, 0))) return v4(1, 0, 0, 0);
  return normalize(c);
}
V4 create_horizon(const V4& pole, double z, double theta) {  // Vec4D.CreateHorizon, Vec4D.cs:52-58
  V4 c = create_horizontal(pole);
  return matvec(rotate(theta, pole), (pole * z) + (c * std::sqrt(1 - z * z)));
}

// Util.NearlyEqual (Util.cs:41-56)
inline bool nearly_equal(double a, double b, double delta) {
  const double min_normal = std::numeric_limits<double>::denorm_min() * 1e7;  // double.Epsilon * 1e7
  if (delta == 0) return true;
  delta = std::fabs(delta);
  double mx = (std::isnan(a) || std::isnan(b)) ? kNaN : std::max(a, b);  // Math.Max
  return delta <= min_normal || delta / mx < kNearEnough;
}
inline bool nearly_equal(double a, double b) { return nearly_equal(a, b, a - b); }
inline bool nearly_equals(const V4& a, const V4& b) {  // Vec4D.NearlyEquals, Vec4D.cs:439-442
  return nearly_equal(sqlen(a), sqlen(b), sqlen(a - b));
}

// x86 MAXPD/MINPD semantics as used by Sse2.Max/Min (second operand wins on NaN / equality)
inline double sse_max(double a, double b) { return a > b ? a : b; }
inline double sse_min(double a, double b) { return a < b ? a : b; }
inline double clamp_sse(double v, double lo, double hi) { return sse_min(sse_max(v, lo), hi); }  // Util.cs:126-134

// ------------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11) — our replacement for the unseeded System.Random of Raytracer.cs:48.
// ------------------------------------------------------------------------------------------------------
inline void philox_round(uint32_t c[4], const uint32_t k[2]) {
  uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
  uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
  uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
  uint32_t n1 = (uint32_t)p1;
  uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
  uint32_t n3 = (uint32_t)p0;
  c[0] = n0;
  c[1] = n1;
  c[2] = n2;
  c[3] = n3;
}
inline void philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
  uint32_t k[2] = {key[0], key[1]};
  for (int r = 0; r < 10; r++) {
    if (r) {
      k[0] += 0x9E3779B9u;
      k[1] += 0xBB67AE85u;
    }
    philox_round(c, k);
  }
  out[0] = c[0];
  out[1] = c[1];
  out[2] = c[2];
  out[3] = c[3];
}
// Two uniforms in [0,1) per block: 53 bits each. counter = (pixel, sample, stage, block), key = seed.
// stage 0 = camera ray (block 0: subX, subY; block 1: lens radius, lens angle);
// stage 1+i = bounce i (block 0: shine z, shine theta; block 1: lobe pick, diffuse z; block 2: diffuse theta).
inline void uniforms(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t stage, uint32_t block, double out[2]) {
  uint32_t ctr[4] = {pixel, sample, stage, block};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t r[4];
  philox(ctr, key, r);
  out[0] = (double)((((uint64_t)r[1] << 32) | r[0]) >> 11) * 0x1.0p-53;
  out[1] = (double)((((uint64_t)r[3] << 32) | r[2]) >> 11) * 0x1.0p-53;
}

// ------------------------------------------------------------------------------------------------------
// Scene data
// ------------------------------------------------------------------------------------------------------
struct Ray {  // Vectors/Ray.cs:27-29
  V4 o, d;
};

struct Hit {  // Raytracing/Hit.cs:14-20; prim < 0 == null
  int prim = -1;
  V4 pos{0, 0, 0, 0};
  double dist = 0;
  V4 normal{0, 0, 0, 0};
  bool inside = false;
};
inline bool hit_equal(const Hit& a, const Hit& b) {  // Hit.operator==, Hit.cs:44-59
  if (a.prim < 0) return b.prim < 0;
  if (b.prim < 0) return false;
  return a.prim == b.prim && eq3(a.pos, b.pos) && a.dist == b.dist && eq3(a.normal, b.normal) && a.inside == b.inside;
}

struct Prim {
  uint8_t kind, flags;
  // triangle
  V4 v0, e1, e2, n;
  V4 vn0, vn1, vn2;
  // sphere
  V4 center;
  double radius, radius_sqr;
  M4 to_world, to_object, to_normal;  // reference names (Sphere.cs:17-19): to_world is world->object
  // plane
  V4 pnormal;
  double pdist;
  // material (Primitive.cs:16-129)
  double emission[3], diffuse[3], specular[3], refraction[3], ior, shininess;
  bool mirror() const { return flags & RTC_FLAG_MIRROR; }
  bool two_sided() const { return flags & RTC_FLAG_TWOSIDED; }
  bool invert() const { return flags & RTC_FLAG_INVERT; }
  bool transformed() const { return flags & RTC_FLAG_TRANSFORMED; }
  bool vnormals() const { return flags & RTC_FLAG_VNORMALS; }
};

struct Node {  // Acceleration/BVH.cs:239-254
  V4 bmin, bmax;
  int left, right, prim;
  bool skip_volume;
};

struct BI {  // Acceleration/BoundingIntersection.cs:3-17
  int node;
  double near_, far_;
};

}  // namespace

struct orc_scene {
  std::vector<Prim> prims;
  std::vector<Node> nodes;
  int root = -1;
  rtc_camera cam{};
  rtc_params par{};
};

namespace {

// ------------------------------------------------------------------------------------------------------
// AABB.IntersectAVX (Acceleration/AABB.cs:107-142). Lanes x,y,z,w; the w lane (box W = 1, origin W = 1,
// direction W = 0) always ends up (-inf, +inf) and is kept so NaN handling follows MAXPD/MINPD exactly.
// ------------------------------------------------------------------------------------------------------
inline bool aabb_intersect(const V4& bmin, const V4& bmax, const Ray& ray, double& near_out, double& far_out) {
  const double o[4] = {ray.o.x, ray.o.y, ray.o.z, ray.o.w};
  const double d[4] = {ray.d.x, ray.d.y, ray.d.z, ray.d.w};
  const double mn[4] = {bmin.x, bmin.y, bmin.z, bmin.w};
  const double mx[4] = {bmax.x, bmax.y, bmax.z, bmax.w};
  double n4[4], f4[4];
  for (int i = 0; i < 4; i++) {
    double lo = mn[i], hi = mx[i];
    if (d[i] == 0 && o[i] >= lo && o[i] <= hi) {  // :117-123
      lo = -kInf;
      hi = kInf;
    }
    bool sgn = std::signbit(d[i]);  // BlendVariable on the sign bit, :126-127
    double a = sgn ? hi : lo;
    double b = sgn ? lo : hi;
    double inv = 1.0 / d[i];  // :129
    n4[i] = (a - o[i]) * inv;
    f4[i] = (b - o[i]) * inv;
  }
  // :133-136: Max(lower, upper) then MaxScalar(x, swap(x))
  double nl0 = sse_max(n4[0], n4[2]), nl1 = sse_max(n4[1], n4[3]);
  double nr = sse_max(nl0, nl1);
  double fl0 = sse_min(f4[0], f4[2]), fl1 = sse_min(f4[1], f4[3]);
  double fr = sse_min(fl0, fl1);
  if ((nr > fr) | (fr < 0)) return false;  // :138 (ordered compares: NaN -> false)
  near_out = nr;
  far_out = fr;
  return true;
}

// ------------------------------------------------------------------------------------------------------
// Triangle.RayTraceAVXFaster + GetNormal (Primitives/Triangle.cs:77-146, 209-224)
// ------------------------------------------------------------------------------------------------------
inline int triangle_hits(const Prim& p, int id, const Ray& ray, Hit out[2]) {
  V4 offset = ray.o - p.v0;                                         // :84
  V4 side1 = scross(offset, p.e1);                                  // :85
  V4 side2 = scross(ray.d, p.e2);                                   // :86
  double u = sdot(offset, side2);                                   // :89,93,97
  double v = sdot(ray.d, side1);                                    // :90,93,97
  double dist = sdot(p.e2, side1);                                  // :91,94,98
  double det = sdot(p.e1, side2);                                   // :92,95,99
  double inv = 1.0 / det;                                           // :107
  if (std::isnan(inv)) inv = 0;                                     // :108-110 (And with CompareOrdered)
  u = u * inv;                                                      // :112
  v = v * inv;
  dist = dist * inv;                                                // :113
  bool reject = (u < 0) | (v < 0);                                  // :116
  if (p.mirror())
    reject |= (u > 1) | (v > 1);                                    // :117-118
  else
    reject |= (u + v) > 1;
  reject |= dist < 0;                                               // :120
  if (reject) return 0;                                             // :123-124
  bool inside = inv < 0;                                            // :126
  V4 pos = {std::fma(p.e1.x, u, std::fma(p.e2.x, v, p.v0.x)), std::fma(p.e1.y, u, std::fma(p.e2.y, v, p.v0.y)),
            std::fma(p.e1.z, u, std::fma(p.e2.z, v, p.v0.z)), std::fma(p.e1.w, u, std::fma(p.e2.w, v, p.v0.w))};  // :130
  V4 normal;
  if (p.vnormals()) {                                               // :211-219 (bug-compatible weights)
    V4 nn = normalize(((p.vn0 * u) + (p.vn1 * v)) + (p.vn2 * (u + v)));
    if (inside)
      normal = nn - (p.n * (2 * (dot(nn, p.n)) / dot(p.n, p.n)));
    else
      normal = nn;
  } else {
    normal = inside ? (p.n * -1.0) : p.n;                           // :221-223
  }
  out[0].prim = id;
  out[0].pos = pos;
  out[0].dist = dist;
  out[0].normal = normal;
  out[0].inside = inside;
  return 1;
}

// ------------------------------------------------------------------------------------------------------
// Sphere.RayTraceAVX (Primitives/Sphere.cs:50-155)
// ------------------------------------------------------------------------------------------------------
inline int sphere_hits(const Prim& p, int id, const Ray& ray, Hit out[2]) {
  V4 obj_o = ray.o, obj_d = ray.d;
  const V4 world_o = ray.o, world_d = ray.d;
  if (p.transformed()) {                                            // :58-76
    obj_o = matvec(p.to_world, obj_o);
    obj_d = snormalize(matvec(p.to_world, obj_d));
  }
  V4 offset = obj_o - p.center;                                     // :79
  double b = -2 * sdot(offset, obj_d);                              // :80,84
  double c = sdot(offset, offset) - p.radius_sqr;                   // :81,85
  double radix = std::sqrt((b * b) - (4 * c));                      // :86
  double dist_far = (b + radix) / 2;                                // :89
  double dist_close = (b - radix) / 2;                              // :90
  auto fma4 = [](double t, const V4& d, const V4& o) {
    return V4{std::fma(t, d.x, o.x), std::fma(t, d.y, o.y), std::fma(t, d.z, o.z), std::fma(t, d.w, o.w)};
  };
  V4 pos_far = fma4(dist_far, obj_d, obj_o);                        // :94
  V4 n_far = (pos_far - p.center) / p.radius;                       // :95
  V4 pos_close = fma4(dist_close, obj_d, obj_o);                    // :97
  V4 n_close = (pos_close - p.center) / p.radius;                   // :98
  if (p.transformed()) {                                            // :100-139
    pos_far = matvec(p.to_object, pos_far);
    pos_close = matvec(p.to_object, pos_close);
    n_far = snormalize(matvec(p.to_normal, n_far));
    dist_far = sdot(world_d, pos_far - world_o);
    n_close = snormalize(matvec(p.to_normal, n_close));
    dist_close = sdot(world_d, pos_close - world_o);
  }
  n_far = neg(n_far);                                               // :142
  if (!(dist_far >= 0)) return 0;                                   // :145-146
  if (!(dist_close >= 0)) {                                         // :148-149
    out[0] = Hit{id, pos_far, dist_far, n_far, true};
    return 1;
  }
  out[0] = Hit{id, pos_close, dist_close, n_close, false};          // :151-154
  out[1] = Hit{id, pos_far, dist_far, n_far, true};
  return 2;
}

// ------------------------------------------------------------------------------------------------------
// Plane.DoRayTrace (Primitives/Plane.cs:36-66)
// ------------------------------------------------------------------------------------------------------
inline int plane_hits(const Prim& p, int id, const Ray& ray, Hit out[2]) {
  double ray_dist = dot(ray.o, p.pnormal);                          // :38
  double denom = dot(ray.d, p.pnormal);                             // :39
  if (nearly_equal(denom, 0) && nearly_equal(p.pdist, ray_dist)) {  // :41-42
    out[0] = Hit{id, ray.o, 0, p.pnormal, true};
    return 1;
  }
  if (denom == 0) return 0;                                         // :44-45
  double dist = (p.pdist - ray_dist) / denom;                       // :47
  if (dist >= -kNearEnough) {                                       // :49
    V4 hp = ray.o + (ray.d * dist);                                 // :51, Ray.cs:53-56
    V4 hn = p.pnormal;
    bool inside = false;
    if (dot(p.pnormal, ray.d) > 0) {                                // :56-60
      hn = neg(hn);
      inside = true;
    }
    out[0] = Hit{id, hp, length(hp - ray.o), hn, inside};           // :62
    return 1;
  }
  return 0;
}

// Util.RayHitMatches (Util.cs:179-192)
// orc_set_selfhit_mode: 0 = the reference's rule (default). 1 = the documented self-hit rule of the library's f32 mode
// (deviation 1), restated here in f64 so that a test can tell that deviation apart from everything else:
// a flat primitive equal to the skip hit's primitive is always the self-hit; spheres keep the positional rule with a relative
// threshold of 1e-9 on squared distances instead of Util.NearEnough = 1e-24.
int g_selfhit_mode = 0;

inline bool ray_hit_matches(const Ray& ray, const Hit& a, const Hit& b, int kind) {
  if (g_selfhit_mode == 1) {
    if (a.prim < 0 || b.prim < 0 || a.prim != b.prim) return false;
    if (kind != RTC_KIND_SPHERE) return true;
    const double la = sqlen(a.pos), lb = sqlen(b.pos), ld = sqlen(a.pos - b.pos);
    if (!(ld == 0 || ld / std::max(la, lb) < 1e-9)) return false;
    if (dot(ray.d, b.normal) > 0) return a.inside != b.inside;
    return a.inside == b.inside;
  }
  if (hit_equal(a, b)) return true;
  if (a.prim < 0 || b.prim < 0) return false;
  if (a.prim != b.prim) return false;
  if (!nearly_equals(a.pos, b.pos)) return false;
  if (dot(ray.d, b.normal) > 0) return a.inside != b.inside;
  return a.inside == b.inside;
}

// Primitive.RayTrace (Primitives/Primitive.cs:46-75)
inline Hit primitive_ray_trace(const Prim& p, int id, const Ray& ray, const Hit& skip) {
  Hit hits[2];
  int n;
  switch (p.kind) {
    case RTC_KIND_TRIANGLE: n = triangle_hits(p, id, ray, hits); break;
    case RTC_KIND_SPHERE: n = sphere_hits(p, id, ray, hits); break;
    default: n = plane_hits(p, id, ray, hits); break;
  }
  for (int i = 0; i < n; i++) {
    Hit cur = hits[i];
    if (p.invert()) cur.inside = !cur.inside;                       // :60-61, Hit.cs:39-42
    if (cur.inside && !p.two_sided()) continue;                     // :63-64
    if (!ray_hit_matches(ray, cur, skip, p.kind)) return cur;       // :66-70
  }
  return Hit{};
}

// BVH<T>.IntersectLeaves, recursive overload (Acceleration/BVH.cs:295-316)
void intersect_leaves(const orc_scene& s, int ni, const Ray& ray, std::vector<BI>& list, double near_, double far_) {
  const Node& nd = s.nodes[ni];
  if (!nd.skip_volume) {
    double n, f;
    if (!aabb_intersect(nd.bmin, nd.bmax, ray, n, f)) return;  // (NaN,NaN) -> !(far >= 0)
    if (!(f >= 0)) return;                                      // :302
    near_ = n;
    far_ = f;
  }
  if (nd.prim >= 0) {
    list.push_back(BI{ni, near_, far_});
    return;
  }
  intersect_leaves(s, nd.left, ray, list, near_, far_);
  intersect_leaves(s, nd.right, ray, list, near_, far_);
}

inline int compare_to(double a, double b) {  // double.CompareTo
  if (a < b) return -1;
  if (a > b) return 1;
  if (a == b) return 0;
  if (std::isnan(a)) return std::isnan(b) ? 0 : -1;
  return 1;
}

// Scene.RayTracePrimitives, accelerator branch (Raytracing/Scene.cs:71-92)
Hit scene_ray_trace_bvh(const orc_scene& s, const Ray& ray, const Hit& skip, std::vector<BI>& list) {
  Hit hit;
  list.clear();
  if (s.root < 0) return hit;
  intersect_leaves(s, s.root, ray, list, 0, 0);  // BVH.cs:323-327 (distances = default)
  // Util.InsertSort (Util.cs:262-280), comparer a.Near.CompareTo(b.Near) (BVH.cs:241)
  for (size_t i = 1; i < list.size(); i++) {
    BI a = list[i];
    long j = (long)i - 1;
    while (j >= 0 && compare_to(a.near_, list[j].near_) < 0) {
      list[j + 1] = list[j];
      j--;
    }
    list[j + 1] = a;
  }
  int previous = -1;
  for (size_t i = 0; i < list.size(); i++) {
    const BI& cur = list[i];
    if (previous >= 0 && cur.near_ > list[previous].far_) break;  // :80-81
    int pid = s.nodes[cur.node].prim;
    Hit h = primitive_ray_trace(s.prims[pid], pid, ray, skip);    // :83
    if (h.prim >= 0 && (hit.prim < 0 || h.dist < hit.dist)) {     // :85-86
      hit = h;
      previous = (int)i;
    }
  }
  return hit;
}

// Scene.RayTracePrimitives, accelerator-less branch (Scene.cs:95-107)
Hit scene_ray_trace_all(const orc_scene& s, const Ray& ray, const Hit& skip) {
  Hit hit;
  for (size_t i = 0; i < s.prims.size(); i++) {
    Hit h = primitive_ray_trace(s.prims[i], (int)i, ray, skip);
    if (h.prim >= 0 && (hit.prim < 0 || h.dist < hit.dist)) hit = h;
  }
  return hit;
}

// ------------------------------------------------------------------------------------------------------
// Cameras (Raytracing/Cameras/{FrustumCamera,OrthoCamera}.cs) and Raytracer.GetCameraRay
// ------------------------------------------------------------------------------------------------------
inline V4 cv3(const double a[3], double w) { return V4{a[0], a[1], a[2], w}; }

Ray camera_get_ray(const rtc_camera& c, double x, double y) {
  V4 look = cv3(c.look, 0), side = cv3(c.side, 0), up = cv3(c.up, 0), pos = cv3(c.position, 1);
  if (c.kind == RTC_CAMERA_FRUSTUM) {  // FrustumCamera.cs:33-41
    double off_x = c.tan_fov_x2 * ((x - c.w2) / c.w2);
    double off_y = c.tan_fov_y2 * ((y - c.h2) / c.h2);
    V4 dir = look + (side * off_x) + (up * off_y);
    return Ray{pos, normalize(dir)};  // Ray.Directional, Ray.cs:21-24
  }
  // OrthoCamera.cs:33-38
  V4 start = pos + (side * ((x - c.w2) * c.h_mult)) + (up * ((y - c.h2) * c.v_mult));
  return Ray{start, normalize(look)};
}
inline Ray ray_offset(const Ray& r, double dist) { return Ray{r.o + (r.d * dist), r.d}; }  // Ray.cs:53-62

Ray get_camera_ray(const orc_scene& s, int x, int y, uint32_t sample) {  // Raytracer.cs:262-282
  const rtc_camera& c = s.cam;
  uint32_t pixel = (uint32_t)(y * s.par.width + x);
  double u[2];
  uniforms(s.par.seed, pixel, sample, 0, 0, u);
  double sub_x = x + u[0];
  double sub_y = y + u[1];
  Ray ray = ray_offset(camera_get_ray(c, sub_x, sub_y), c.image_plane);
  if (c.dof_amount != 0) {
    V4 focus = ray.o + (ray.d * (c.focal_length - c.image_plane));  // Ray.GetPoint
    double l[2];
    uniforms(s.par.seed, pixel, sample, 0, 1, l);
    double dist = std::sqrt(l[0]) * c.dof_amount;
    double angle = l[1] * kPi * 2;
    double off_x = std::cos(angle) * dist;
    double off_y = std::sin(angle) * dist;
    Ray r2 = ray_offset(camera_get_ray(c, sub_x + off_x, sub_y + off_y), c.image_plane);
    ray = Ray{r2.o, normalize(focus - r2.o)};  // PointingTowards -> FromTo, Ray.cs:15-18,71-74
  }
  return ray;
}

// ------------------------------------------------------------------------------------------------------
// Raytracer.GetColor (Raytracing/Raytracer.cs:65-246)
// ------------------------------------------------------------------------------------------------------
inline double luminance(const double c[3]) { return 0.299 * c[0] + 0.587 * c[1] + 0.114 * c[2]; }  // DoubleColor.cs:76-81

struct RaySeg {  // one Scene.RayTrace call of a path, as GetColor made it (Raytracer.cs:77): ray, skip hit, answer, bounce
  Ray ray;
  Hit skip, hit;
  int bounce;
};

struct PathCtx {
  std::vector<BI> list;
  uint64_t rays = 0;
  std::vector<RaySeg>* dump = nullptr;  // orc_dump_path_rays: every segment is recorded here
};

enum BounceType { Skipped, Diffuse, Specular, SpecularFail, Transmitted, Emission, PureBlack, RecursionComplete, Missed, Debug };

void get_color(const orc_scene& s, Ray ray, uint32_t pixel, uint32_t sample, PathCtx& ctx, double out[3],
               rtc_debug_ray* debug, int debug_cap, int* debug_n) {
  static const double black[3] = {0, 0, 0};
  Hit prev_hit, hit;
  double tint[3] = {1, 1, 1};
  const int recursion = s.par.recursion;
  if (debug_n) *debug_n = 0;
  for (int i = 0; i <= recursion; i++) {
    if (i % 3 == 0) ray = Ray{ray.o, normalize(ray.d)};  // :74-75
    hit = scene_ray_trace_bvh(s, ray, prev_hit, ctx.list);  // :77
    ctx.rays++;
    if (ctx.dump) ctx.dump->push_back(RaySeg{ray, prev_hit, hit, i});
    rtc_debug_ray* dr = nullptr;
    if (debug && i < debug_cap) {
      dr = &debug[i];
      *debug_n = i + 1;
      dr->hit.prim = hit.prim;
      dr->hit.inside = hit.inside;
      dr->hit.t = hit.dist;
      dr->hit.position[0] = hit.pos.x; dr->hit.position[1] = hit.pos.y; dr->hit.position[2] = hit.pos.z;
      dr->hit.normal[0] = hit.normal.x; dr->hit.normal[1] = hit.normal.y; dr->hit.normal[2] = hit.normal.z;
      dr->type = Skipped;
      dr->pad = 0;
      dr->fresnel_ratio = kNaN;
    }
    if (hit.prim < 0) {  // :81-91
      if (dr) dr->type = Missed;
      if (i == 0) {
        out[0] = out[1] = out[2] = -1;  // DoubleColor.Placeholder
        return;
      }
      out[0] = s.par.ambient[0]; out[1] = s.par.ambient[1]; out[2] = s.par.ambient[2];
      return;
    }
    const Prim& p = s.prims[hit.prim];
    const bool reflective = p.shininess > 0;  // Primitive.IsReflective, Primitive.cs:106
    const double* specular = reflective ? p.specular : black;      // Primitive.cs:111-115
    const double* refraction = reflective ? p.refraction : black;  // Primitive.cs:120-124
    if (s.par.debug_geom) {  // :93-98
      if (dr) dr->type = Debug;
      for (int k = 0; k < 3; k++) out[k] = specular[k] + p.diffuse[k] + p.emission[k];
      return;
    }
    if (i >= recursion) {  // :100-104
      if (dr) dr->type = RecursionComplete;
      break;
    }
    const uint32_t stage = 1 + (uint32_t)i;
    double u01[2];
    uniforms(s.par.seed, pixel, sample, stage, 0, u01);
    // RandomShine, :51-56
    double zs = (p.shininess == kInf) ? 1 : std::pow(u01[0], 1 / p.shininess);
    double theta_s = u01[1] * kPi * 2;
    V4 rough = create_horizon(hit.normal, zs, theta_s);  // :108
    double diff_l = luminance(p.diffuse), spec_l = luminance(specular), refr_l = luminance(refraction),
           emis_l = luminance(p.emission);  // :110-113
    double cosv = -dot(rough, ray.d);  // :115
    double cos_out = 0, ior_ratio = 0;
    if (((refr_l > 0) | (spec_l > 0)) && p.ior != 0 && cosv >= 0) {  // :120
      double ior_in, ior_out;
      if (hit.inside) {
        ior_in = p.ior;
        ior_out = s.par.air_ior;
      } else {
        ior_in = s.par.air_ior;
        ior_out = p.ior;
      }
      ior_ratio = ior_in / ior_out;  // :136
      double sin_out = ior_ratio * std::sqrt(1 - (cosv * cosv));  // :137
      if (sin_out >= 1) {  // :140-145
        refr_l = 0;
        if (dr) dr->fresnel_ratio = 1;
      } else {
        cos_out = std::sqrt(1 - (sin_out * sin_out));  // :148
        double rs = ((ior_out * cosv) - (ior_in * cos_out)) / ((ior_out * cosv) + (ior_in * cos_out));  // :149
        double rp = ((ior_in * cosv) - (ior_out * cos_out)) / ((ior_in * cosv) + (ior_out * cos_out));  // :150
        double ratio = ((rs * rs) + (rp * rp)) / 2;  // :151
        spec_l *= ratio;
        refr_l *= 1 - ratio;
        if (dr) dr->fresnel_ratio = ratio;
      }
    } else {
      refr_l = 0;  // :158-161
    }
    double total_l = diff_l + spec_l + refr_l + emis_l;  // :163
    if (total_l <= 0) {  // :165-169
      if (dr) dr->type = PureBlack;
      break;
    }
    double u23[2];
    uniforms(s.par.seed, pixel, sample, stage, 1, u23);
    double new_tint[3] = {0, 0, 0};
    double ray_rand = u23[0] * total_l;  // :178
    bool have_out = false;
    V4 out_dir{0, 0, 0, 0};
    if (refr_l != 0 && (ray_rand -= refr_l) <= 0) {  // :181-193
      if (dr) dr->type = Transmitted;
      out_dir = (rough * -cos_out) + ((ray.d + (rough * cosv)) * ior_ratio);
      have_out = true;
      for (int k = 0; k < 3; k++) new_tint[k] = hit.inside ? 1.0 : refraction[k];
    } else if (spec_l != 0 && (ray_rand -= spec_l) <= 0) {  // :194-209
      if (dr) dr->type = SpecularFail;
      V4 od = ray.d + (rough * (cosv * 2));  // Reflection, :58-61
      if (dot(od, hit.normal) > 0) {
        if (dr) dr->type = Specular;
        out_dir = od;
        have_out = true;
        for (int k = 0; k < 3; k++) new_tint[k] = specular[k];
      }
    } else if (diff_l != 0 && (ray_rand -= diff_l) <= 0) {  // :210-219
      if (dr) dr->type = Diffuse;
      double u4[2];
      uniforms(s.par.seed, pixel, sample, stage, 2, u4);
      double z = (2 * std::acos(u23[1])) / kPi;
      double theta = u4[0] * kPi * 2;
      out_dir = create_horizon(hit.normal, z, theta);
      have_out = true;
      for (int k = 0; k < 3; k++) new_tint[k] = p.diffuse[k];
    } else {  // :220-229
      if (dr) dr->type = Emission;
      break;
    }
    // :231-232 — Ray.Zero comparison: a ray whose origin and direction are both (0,0,0) also terminates.
    if (!have_out || (eq3(hit.pos, v4(0, 0, 0, 0)) && eq3(out_dir, v4(0, 0, 0, 0)))) break;
    prev_hit = hit;
    ray = Ray{hit.pos, out_dir};
    double m = (std::isnan(total_l)) ? kNaN : std::max(total_l, 1.0);  // Math.Max, :238
    for (int k = 0; k < 3; k++) {
      double nt = new_tint[k] * m;
      tint[k] = tint[k] * nt;  // :240
    }
  }
  const Prim& lp = s.prims[hit.prim];
  for (int k = 0; k < 3; k++) out[k] = tint[k] * lp.emission[k];  // :245
}

void to_rtc_hit(const Hit& h, rtc_hit& o) {
  o.prim = h.prim;
  o.inside = h.inside ? 1 : 0;
  o.t = h.dist;
  o.position[0] = h.pos.x; o.position[1] = h.pos.y; o.position[2] = h.pos.z;
  o.normal[0] = h.normal.x; o.normal[1] = h.normal.y; o.normal[2] = h.normal.z;
}
Hit from_rtc_hit(const rtc_hit& h) {
  Hit o;
  o.prim = h.prim;
  o.inside = h.inside != 0;
  o.dist = h.t;
  o.pos = V4{h.position[0], h.position[1], h.position[2], 1};
  o.normal = V4{h.normal[0], h.normal[1], h.normal[2], 0};
  return o;
}

template <class F>
void parallel_for(int64_t n, int threads, int64_t grain, F f) {
  if (threads <= 1 || n <= grain) {
    f(0, n, 0);
    return;
  }
  std::atomic<int64_t> next{0};
  std::vector<std::thread> ts;
  for (int t = 0; t < threads; t++) {
    ts.emplace_back([&, t]() {
      for (;;) {
        int64_t b = next.fetch_add(grain);
        if (b >= n) break;
        f(b, std::min(n, b + grain), t);
      }
    });
  }
  for (auto& t : ts) t.join();
}

}  // namespace

// ========================================================================================================
extern "C" {

orc_scene* orc_scene_create(const rtc_scene_desc* d, int32_t n_nodes, const rtc_bvh_node* nodes, int32_t root,
                            const rtc_camera* camera, const rtc_params* params) {
  orc_scene* s = new orc_scene();
  s->prims.resize(d->n_prims);
  for (int i = 0; i < d->n_prims; i++) {
    Prim& p = s->prims[i];
    std::memset(&p, 0, sizeof(p));
    p.kind = d->kind[i];
    p.flags = d->flags[i];
    const double* g = d->geom + (size_t)i * RTC_GEOM_STRIDE;
    const double* x = (d->xform && d->xform[i] >= 0) ? d->xforms + (size_t)d->xform[i] * RTC_XFORM_STRIDE : nullptr;
    if (p.kind == RTC_KIND_TRIANGLE) {
      p.v0 = V4{g[0], g[1], g[2], 1};
      p.e1 = V4{g[3], g[4], g[5], 0};
      p.e2 = V4{g[6], g[7], g[8], 0};
      p.n = V4{g[9], g[10], g[11], 0};
      if (p.vnormals() && x) {
        p.vn0 = V4{x[0], x[1], x[2], 0};
        p.vn1 = V4{x[3], x[4], x[5], 0};
        p.vn2 = V4{x[6], x[7], x[8], 0};
      }
    } else if (p.kind == RTC_KIND_SPHERE) {
      p.center = V4{g[0], g[1], g[2], 1};
      p.radius = g[3];
      p.radius_sqr = g[4];
      if (p.transformed() && x) {
        std::memcpy(p.to_world.m, x, 16 * sizeof(double));
        std::memcpy(p.to_object.m, x + 16, 16 * sizeof(double));
        std::memcpy(p.to_normal.m, x + 32, 16 * sizeof(double));
      }
    } else {
      p.pnormal = V4{g[0], g[1], g[2], 0};
      p.pdist = g[3];
    }
    const double* m = d->material + (size_t)i * RTC_MATERIAL_STRIDE;
    for (int k = 0; k < 3; k++) {
      p.emission[k] = m[k];
      p.diffuse[k] = m[3 + k];
      p.specular[k] = m[6 + k];
      p.refraction[k] = m[9 + k];
    }
    p.ior = m[12];
    p.shininess = m[13];
  }
  s->nodes.resize(n_nodes);
  for (int i = 0; i < n_nodes; i++) {
    Node& n = s->nodes[i];
    n.bmin = V4{nodes[i].bmin[0], nodes[i].bmin[1], nodes[i].bmin[2], 1};
    n.bmax = V4{nodes[i].bmax[0], nodes[i].bmax[1], nodes[i].bmax[2], 1};
    n.left = nodes[i].left;
    n.right = nodes[i].right;
    n.prim = nodes[i].prim;
    n.skip_volume = false;
  }
  // BVH.MakeParent (BVH.cs:44-48): child.SkipVolume = child.Volume.Equals(parent.Volume) (AABB.cs:231-240)
  for (int i = 0; i < n_nodes; i++) {
    Node& n = s->nodes[i];
    if (n.prim >= 0) continue;
    for (int c : {n.left, n.right}) {
      Node& ch = s->nodes[c];
      ch.skip_volume = eq3(ch.bmin, n.bmin) && eq3(ch.bmax, n.bmax);
    }
  }
  s->root = root;
  if (camera) s->cam = *camera;
  if (params) s->par = *params;
  return s;
}

void orc_scene_destroy(orc_scene* s) { delete s; }
void orc_set_camera(orc_scene* s, const rtc_camera* camera) { s->cam = *camera; }
void orc_set_params(orc_scene* s, const rtc_params* params) { s->par = *params; }
void orc_set_selfhit_mode(int mode) { g_selfhit_mode = mode; }

int64_t orc_trace_closest(orc_scene* s, int64_t n, const rtc_ray* rays, const rtc_hit* skip, rtc_hit* out, int mode,
                          int check_both, int threads) {
  std::atomic<int64_t> diff{0};
  parallel_for(n, threads, 4096, [&](int64_t b, int64_t e, int) {
    std::vector<BI> list;
    list.reserve(64);
    int64_t local = 0;
    for (int64_t i = b; i < e; i++) {
      Ray r{V4{rays[i].origin[0], rays[i].origin[1], rays[i].origin[2], 1}, V4{rays[i].dir[0], rays[i].dir[1], rays[i].dir[2], 0}};
      Hit sk;
      if (skip && skip[i].prim >= 0) sk = from_rtc_hit(skip[i]);
      Hit h = mode == 0 ? scene_ray_trace_bvh(*s, r, sk, list) : scene_ray_trace_all(*s, r, sk);
      if (check_both) {
        Hit h2 = mode == 0 ? scene_ray_trace_all(*s, r, sk) : scene_ray_trace_bvh(*s, r, sk, list);
        if (!hit_equal(h, h2)) local++;
      }
      to_rtc_hit(h, out[i]);
    }
    diff += local;
  });
  return diff.load();
}

void orc_camera_rays(orc_scene* s, int64_t n, const int32_t* xy, const uint32_t* sample, rtc_ray* out) {
  for (int64_t i = 0; i < n; i++) {
    Ray r = get_camera_ray(*s, xy[2 * i], xy[2 * i + 1], sample[i]);
    out[i].origin[0] = r.o.x; out[i].origin[1] = r.o.y; out[i].origin[2] = r.o.z;
    out[i].dir[0] = r.d.x; out[i].dir[1] = r.d.y; out[i].dir[2] = r.d.z;
  }
}

uint64_t orc_render(orc_scene* s, int32_t x0, int32_t y0, int32_t x1, int32_t y1, uint32_t first_sample,
                    uint32_t n_samples, int threads, double* rgb_sum, uint32_t* samples, uint32_t* misses) {
  // Work items are image rows of the rectangle; each pixel is owned by one item, and its samples are taken in
  // ascending order, so the per-pixel sum order matches FullRaytracer's pass-by-pass accumulation (:326-339).
  const int w = s->par.width;
  std::atomic<uint64_t> rays{0};
  parallel_for(y1 - y0, threads, 1, [&](int64_t b, int64_t e, int) {
    PathCtx ctx;
    ctx.list.reserve(64);
    for (int64_t yy = b; yy < e; yy++) {
      int y = y0 + (int)yy;
      for (int x = x0; x < x1; x++) {
        size_t px = (size_t)y * w + x;
        for (uint32_t k = 0; k < n_samples; k++) {
          uint32_t smp = first_sample + k;
          Ray r = get_camera_ray(*s, x, y, smp);
          double c[3];
          get_color(*s, r, (uint32_t)px, smp, ctx, c, nullptr, 0, nullptr);
          if (c[0] == -1 && c[1] == -1 && c[2] == -1) {  // == Placeholder -> AddMiss (FullRaytracer.cs:334-335)
            misses[px]++;
          } else {  // SampleSet.AddSample, SampleSet.cs:32-36
            rgb_sum[px * 3 + 0] += c[0];
            rgb_sum[px * 3 + 1] += c[1];
            rgb_sum[px * 3 + 2] += c[2];
            samples[px]++;
          }
        }
      }
    }
    rays += ctx.rays;
  });
  return rays.load();
}

// The same loop over a lattice of the whole frame: pixels (off_x + i * stride_x, off_y + j * stride_y). A bounded sample of a
// frame that keeps its mix of rays (silhouettes, background, every depth of the scene) -- bench.py's CPU legs.
uint64_t orc_render_lattice(orc_scene* s, int32_t stride_x, int32_t stride_y, int32_t off_x, int32_t off_y, uint32_t first_sample,
                            uint32_t n_samples, int threads, double* rgb_sum, uint32_t* samples, uint32_t* misses) {
  const int w = s->par.width, h = s->par.height;
  if (stride_x < 1 || stride_y < 1 || off_x < 0 || off_y < 0) return 0;
  const int rows = off_y < h ? (h - off_y + stride_y - 1) / stride_y : 0;
  std::atomic<uint64_t> rays{0};
  parallel_for(rows, threads, 1, [&](int64_t b, int64_t e, int) {
    PathCtx ctx;
    ctx.list.reserve(64);
    for (int64_t j = b; j < e; j++) {
      const int y = off_y + (int)j * stride_y;
      for (int x = off_x; x < w; x += stride_x) {
        const size_t px = (size_t)y * w + x;
        for (uint32_t k = 0; k < n_samples; k++) {
          const uint32_t smp = first_sample + k;
          Ray r = get_camera_ray(*s, x, y, smp);
          double c[3];
          get_color(*s, r, (uint32_t)px, smp, ctx, c, nullptr, 0, nullptr);
          if (c[0] == -1 && c[1] == -1 && c[2] == -1) {
            misses[px]++;
          } else {
            rgb_sum[px * 3 + 0] += c[0];
            rgb_sum[px * 3 + 1] += c[1];
            rgb_sum[px * 3 + 2] += c[2];
            samples[px]++;
          }
        }
      }
    }
    rays += ctx.rays;
  });
  return rays.load();
}

void orc_render_samples(orc_scene* s, uint32_t sample, int threads, double* out_rgb) {
  const int w = s->par.width, h = s->par.height;
  parallel_for(h, threads, 1, [&](int64_t b, int64_t e, int) {
    PathCtx ctx;
    for (int64_t y = b; y < e; y++)
      for (int x = 0; x < w; x++) {
        size_t px = (size_t)y * w + x;
        Ray r = get_camera_ray(*s, x, (int)y, sample);
        get_color(*s, r, (uint32_t)px, sample, ctx, out_rgb + px * 3, nullptr, 0, nullptr);
      }
  });
}

void orc_debug_trace(orc_scene* s, int32_t x, int32_t y, uint32_t sample, int32_t capacity, rtc_debug_ray* out,
                     int32_t* n) {
  PathCtx ctx;
  Ray r = get_camera_ray(*s, x, y, sample);
  double c[3];
  int cnt = 0;
  get_color(*s, r, (uint32_t)(y * s->par.width + x), sample, ctx, c, out, capacity, &cnt);
  *n = cnt;
}

int64_t orc_dump_path_rays(orc_scene* s, int64_t n, const int32_t* xy, const uint32_t* sample, int threads, int64_t capacity,
                           rtc_ray* rays, rtc_hit* skip, rtc_hit* hits, int32_t* bounce) {
  // paths are traced in parallel, their segments stored in path order (deterministic for any thread count)
  std::vector<std::vector<RaySeg>> per((size_t)n);
  parallel_for(n, threads, 64, [&](int64_t b, int64_t e, int) {
    PathCtx ctx;
    ctx.list.reserve(64);
    for (int64_t i = b; i < e; i++) {
      ctx.dump = &per[(size_t)i];
      const int x = xy[2 * i], y = xy[2 * i + 1];
      Ray r = get_camera_ray(*s, x, y, sample[i]);
      double c[3];
      get_color(*s, r, (uint32_t)(y * s->par.width + x), sample[i], ctx, c, nullptr, 0, nullptr);
    }
  });
  int64_t k = 0;
  for (int64_t i = 0; i < n; i++)
    for (const RaySeg& g : per[(size_t)i]) {
      if (k < capacity) {
        rays[k].origin[0] = g.ray.o.x; rays[k].origin[1] = g.ray.o.y; rays[k].origin[2] = g.ray.o.z;
        rays[k].dir[0] = g.ray.d.x; rays[k].dir[1] = g.ray.d.y; rays[k].dir[2] = g.ray.d.z;
        to_rtc_hit(g.skip, skip[k]);
        to_rtc_hit(g.hit, hits[k]);
        bounce[k] = g.bounce;
      }
      k++;
    }
  return k;
}

static int intersection_count(const orc_scene& s, int ni, const Ray& ray) {  // BVH.GetIntersectionCount, BVH.cs:352-363
  const Node& nd = s.nodes[ni];
  double n, f;
  bool hit = aabb_intersect(nd.bmin, nd.bmax, ray, n, f) && f >= 0;
  if (!hit) return 0;
  if (nd.prim >= 0) return 1;
  return 1 + intersection_count(s, nd.left, ray) + intersection_count(s, nd.right, ray);
}

void orc_debug_raycast(orc_scene* s, int32_t mode, int32_t* out) {
  const int w = s->par.width, h = s->par.height;
  std::vector<BI> list;
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      Ray ray = ray_offset(camera_get_ray(s->cam, x, y), s->cam.image_plane);  // DebugRaycaster.cs:236
      int v;
      if (mode == 0) {
        Hit hit = scene_ray_trace_bvh(*s, ray, Hit{}, list);  // :194
        v = hit.prim;
      } else {
        v = s->root >= 0 ? intersection_count(*s, s->root, ray) : 0;  // :204
      }
      out[(size_t)y * w + x] = v;
    }
}

void orc_tonemap(int32_t w, int32_t h, const double* rgb_sum, const uint32_t* samples, const uint32_t* misses,
                 double exposure, const double back[3], double back_a, uint32_t* argb) {
  auto code = [](double r, double g, double b, double a) -> uint32_t {  // SampleSet.GetColorCode, :47-53
    return ((uint32_t)(int)(clamp_sse(a, 0, 1) * 255) << 24) | ((uint32_t)(int)(clamp_sse(r, 0, 1) * 255) << 16) |
           ((uint32_t)(int)(clamp_sse(g, 0, 1) * 255) << 8) | ((uint32_t)(int)(clamp_sse(b, 0, 1) * 255) << 0);
  };
  for (size_t i = 0; i < (size_t)w * h; i++) {
    uint32_t S = samples[i], M = misses[i];
    if (S == 0) {  // :57-58
      argb[i] = code(back[0] * exposure, back[1] * exposure, back[2] * exposure, back_a);
      continue;
    }
    double total = (double)(uint32_t)(S + M);  // :85
    double mult = exposure / S;                // :86
    double r = rgb_sum[i * 3] * mult, g = rgb_sum[i * 3 + 1] * mult, b = rgb_sum[i * 3 + 2] * mult, a = 1;
    double back_alpha_amt = M / total;  // :93
    double back_amt = back_alpha_amt * back_a;
    r += (back[0] - r) * back_amt;
    g += (back[1] - g) * back_amt;
    b += (back[2] - b) * back_amt;
    a += (back_a - a) * back_alpha_amt;
    const double gamma = 1 / 2.2;  // :101
    r = std::pow(r, gamma);
    g = std::pow(g, gamma);
    b = std::pow(b, gamma);
    argb[i] = code(r, g, b, a);
  }
}

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { philox(ctr, key, out); }
void orc_uniforms(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t stage, uint32_t block, double out[2]) {
  uniforms(seed, pixel, sample, stage, block, out);
}
void orc_create_horizon(const double pole[3], double z, double theta, double out[3]) {
  V4 r = create_horizon(V4{pole[0], pole[1], pole[2], 0}, z, theta);
  out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
int orc_aabb_intersect(const double bmin[3], const double bmax[3], const rtc_ray* ray, double* near_out, double* far_out) {
  Ray r{V4{ray->origin[0], ray->origin[1], ray->origin[2], 1}, V4{ray->dir[0], ray->dir[1], ray->dir[2], 0}};
  double n = kNaN, f = kNaN;
  bool ok = aabb_intersect(V4{bmin[0], bmin[1], bmin[2], 1}, V4{bmax[0], bmax[1], bmax[2], 1}, r, n, f);
  *near_out = n;
  *far_out = f;
  return ok ? 1 : 0;
}

}  // extern "C"
