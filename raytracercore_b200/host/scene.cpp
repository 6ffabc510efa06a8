// scene.cpp — host-side Scene / Primitive / Camera mirror and the flattening into the C-ABI structs.
// Reference lines are cited per function (paths relative to RaytracerCore/). Compile with -ffp-contract=off.
#include "scene.h"

#include <cmath>
#include <cstring>
#include <limits>
#include <thread>

namespace rtcore {

// ---------------------------------------------------------------------------------------------------------
// Primitive construction
// ---------------------------------------------------------------------------------------------------------
static Vec4D VertexNormal(const Vec4D& n) { return Normalize(n); }  // Vertex ctor, Primitives/Vertex.cs:10-14

void Primitive::Recalculate() {  // Triangle.cs:54-66
  Edge0to1 = Vert1 - Vert0;
  Edge0to2 = Vert2 - Vert0;
  if (!HasNormals) {
    Normal = Normalize(Cross(Edge0to1, Edge0to2));
    Norm0 = Norm1 = Norm2 = VertexNormal(Normal);
  }
}

Primitive Primitive::MakeTriangle(const Vec4D& p0, const Vec4D& p1, const Vec4D& p2, bool mirror) {
  Primitive t;
  t.Kind = RTC_KIND_TRIANGLE;
  t.Vert0 = p0;
  t.Vert1 = p1;
  t.Vert2 = p2;
  t.Norm0 = t.Norm1 = t.Norm2 = VertexNormal(Vec4D(0, 0, 1, 0));  // Vertex(pos), Vertex.cs:16-18
  t.Recalculate();
  t.Mirror = mirror;
  return t;
}

Primitive Primitive::MakeTriangle(const Vec4D& p0, const Vec4D& n0, const Vec4D& p1, const Vec4D& n1, const Vec4D& p2,
                                  const Vec4D& n2) {
  Primitive t;
  t.Kind = RTC_KIND_TRIANGLE;
  t.Vert0 = p0;
  t.Vert1 = p1;
  t.Vert2 = p2;
  t.Norm0 = VertexNormal(n0);
  t.Norm1 = VertexNormal(n1);
  t.Norm2 = VertexNormal(n2);
  t.HasNormals = true;
  t.Recalculate();  // edges only; the face Normal stays (0,0,0,0) exactly as in the reference (Triangle.cs:59)
  return t;
}

Primitive Primitive::MakeRectangle(const Vec4D& upOrigin, const Vec4D& upDir, const Vec4D& normal, double width,
                                   double height) {  // Triangle.CreateRectangle, Triangle.cs:13-20
  Vec4D side = Normalize(Cross(upDir, normal));
  Vec4D v0 = upOrigin + (upDir * (-height / 2)) + (side * (-width / 2));
  Vec4D v1 = v0 + (side * width);
  Vec4D v2 = v0 + (upDir * height);
  return MakeTriangle(v0, v1, v2, true);
}

Primitive Primitive::MakeSphere(const Vec4D& center, double radius) {  // Sphere.cs:23-27,39-47
  Primitive s;
  s.Kind = RTC_KIND_SPHERE;
  s.Center = center;
  s.RadiusValue = radius;
  s.RadiusSqr = radius * radius;
  return s;
}

Primitive Primitive::MakePlane(double origin, const Vec4D& normal) {  // Plane.cs:18-22
  Primitive p;
  p.Kind = RTC_KIND_PLANE;
  p.OriginDistance = origin;
  p.PlaneNormal = Normalize(normal);
  return p;
}

void Primitive::Transform(const Mat4x4D& forward, const Mat4x4D& inverse) {
  switch (Kind) {
    case RTC_KIND_TRIANGLE: {  // Triangle.cs:68-74, Vertex.Transformed, Vertex.cs:25-28
      Vert0 = forward * Vert0;
      Vert1 = forward * Vert1;
      Vert2 = forward * Vert2;
      Norm0 = VertexNormal(Normalize(forward * Norm0));
      Norm1 = VertexNormal(Normalize(forward * Norm1));
      Norm2 = VertexNormal(Normalize(forward * Norm2));
      Recalculate();
      break;
    }
    case RTC_KIND_SPHERE: {  // Sphere.cs:29-37
      if (forward != Mat4x4D::Identity()) Transformed = true;
      MatrixToObject = MatrixToObject * forward;
      MatrixToWorld = inverse * MatrixToWorld;
      MatrixToNormal = MatrixToWorld.Transpose3x3();
      break;
    }
    default: {  // Plane.cs:29-34
      Vec4D center = forward * GetCenter();
      PlaneNormal = Normalize(inverse.Transpose3x3() * PlaneNormal);
      OriginDistance = Dot(center, PlaneNormal);
      break;
    }
  }
}

Vec4D Primitive::GetCenter() const {
  switch (Kind) {
    case RTC_KIND_TRIANGLE: return (Vert0 + Vert1 + Vert2) / 3;  // Triangle.cs:226-229
    case RTC_KIND_SPHERE: return Transformed ? MatrixToObject * Center : Center;  // Sphere.cs:212-218
    default: return Vec4D(0, 0, 0, 1) + PlaneNormal * OriginDistance;  // Plane.cs:24-27
  }
}

double Primitive::GetMaxCenterDistance(const Vec4D& direction) const {
  switch (Kind) {
    case RTC_KIND_TRIANGLE: {  // Triangle.cs:231-263
      Vec4D center = GetCenter();
      double dist = 0;
      Vec4D v0 = Vert0 - center, v1 = Vert1 - center, v2 = Vert2 - center, v3;
      if (Mirror) v3 = Vert0 + Edge0to1 + Edge0to2 - center;
      dist = std::fmax(Dot(v0, direction), dist);
      dist = std::fmax(Dot(v1, direction), dist);
      dist = std::fmax(Dot(v2, direction), dist);
      if (v3 != Vec4D()) dist = std::fmax(Dot(v3, direction), dist);
      return dist;
    }
    case RTC_KIND_SPHERE: {  // Sphere.cs:220-232
      if (Transformed) {
        double sin = std::sqrt(1 - direction.X * direction.X);
        Vec4D vec(direction.X, direction.Y * sin, direction.Z * sin, 0);
        return Length(MatrixToObject.Transpose3x3() * vec) * RadiusValue;
      }
      return RadiusValue;
    }
    default:  // Plane.cs:68-74
      if (std::fabs(Dot(PlaneNormal, direction)) == 1) return 0;
      return std::numeric_limits<double>::infinity();
  }
}

void PrimitiveBounds(const Primitive& p, double bmin[3], double bmax[3]) {  // AABB.CreateFromBounded, AABB.cs:20-36
  Vec4D c = p.GetCenter();
  Vec4D lo = c - Vec4D(p.GetMaxCenterDistance(Vec4D(-1, 0, 0, 0)), p.GetMaxCenterDistance(Vec4D(0, -1, 0, 0)),
                       p.GetMaxCenterDistance(Vec4D(0, 0, -1, 0)), 0);
  Vec4D hi = c + Vec4D(p.GetMaxCenterDistance(Vec4D(1, 0, 0, 0)), p.GetMaxCenterDistance(Vec4D(0, 1, 0, 0)),
                       p.GetMaxCenterDistance(Vec4D(0, 0, 1, 0)), 0);
  bmin[0] = lo.X; bmin[1] = lo.Y; bmin[2] = lo.Z;
  bmax[0] = hi.X; bmax[1] = hi.Y; bmax[2] = hi.Z;
}

// (fills the fields the bounds read, into a caller-provided object: a Primitive with its three matrices is ~900 bytes, and the
// bounds of a million flattened triangles are computed one after the other)
static void PrimitiveFromDesc(const rtc_scene_desc& d, int i, Primitive& p) {
  const double* g = d.geom + (size_t)i * RTC_GEOM_STRIDE;
  p.Kind = d.kind[i];
  uint8_t f = d.flags[i];
  if (p.Kind == RTC_KIND_TRIANGLE) {
    p.Vert0 = Vec4D(g[0], g[1], g[2], 1);
    p.Edge0to1 = Vec4D(g[3], g[4], g[5], 0);
    p.Edge0to2 = Vec4D(g[6], g[7], g[8], 0);
    p.Vert1 = p.Vert0 + p.Edge0to1;
    p.Vert2 = p.Vert0 + p.Edge0to2;
    p.Mirror = f & RTC_FLAG_MIRROR;
  } else if (p.Kind == RTC_KIND_SPHERE) {
    p.Center = Vec4D(g[0], g[1], g[2], 1);
    p.RadiusValue = g[3];
    p.RadiusSqr = g[4];
    p.Transformed = (f & RTC_FLAG_TRANSFORMED) && d.xform && d.xform[i] >= 0;
    if (p.Transformed) {
      const double* x = d.xforms + (size_t)d.xform[i] * RTC_XFORM_STRIDE;
      std::memcpy(p.MatrixToWorld.D, x, sizeof(double) * 16);
      std::memcpy(p.MatrixToObject.D, x + 16, sizeof(double) * 16);
      std::memcpy(p.MatrixToNormal.D, x + 32, sizeof(double) * 16);
    }
  } else {
    p.PlaneNormal = Vec4D(g[0], g[1], g[2], 0);
    p.OriginDistance = g[3];
  }
}

// The two common cases of DescPrimitiveBounds written out: the same operations on the same values as PrimitiveBounds over
// GetCenter / GetMaxCenterDistance above (a Dot with a signed unit vector is the signed component: the other products are
// zeros), without building a Primitive. Anything else -- transformed spheres, planes, non-finite input -- takes the general
// path; tests/test_bvh_builder.py compares the two on every kind.
static bool FastBounds(const rtc_scene_desc& d, int i, double bmin[3], double bmax[3]) {
  const double* g = d.geom + (size_t)i * RTC_GEOM_STRIDE;
  const uint8_t kind = d.kind[i], f = d.flags[i];
  if (kind == RTC_KIND_TRIANGLE) {
    for (int k = 0; k < 9; k++)
      if (!std::isfinite(g[k])) return false;
    const bool mirror = f & RTC_FLAG_MIRROR;
    double v3[3] = {0, 0, 0}, c[3];
    for (int k = 0; k < 3; k++) {
      const double v0 = g[k], v1 = g[k] + g[3 + k], v2 = g[k] + g[6 + k];
      c[k] = ((v0 + v1) + v2) / 3;  // Triangle.cs:226-229
      if (mirror) v3[k] = ((g[k] + g[3 + k]) + g[6 + k]) - c[k];
    }
    const bool use_v3 = !(v3[0] == 0 && v3[1] == 0 && v3[2] == 0);  // `v3 != Vec4D()`
    for (int k = 0; k < 3; k++) {
      const double a0 = g[k] - c[k], a1 = (g[k] + g[3 + k]) - c[k], a2 = (g[k] + g[6 + k]) - c[k];
      double dn = std::fmax(-a0, 0.0), dp = std::fmax(a0, 0.0);
      dn = std::fmax(-a1, dn);
      dp = std::fmax(a1, dp);
      dn = std::fmax(-a2, dn);
      dp = std::fmax(a2, dp);
      if (use_v3) {
        dn = std::fmax(-v3[k], dn);
        dp = std::fmax(v3[k], dp);
      }
      bmin[k] = std::nextafter(c[k] - dn, -std::numeric_limits<double>::infinity());
      bmax[k] = std::nextafter(c[k] + dp, std::numeric_limits<double>::infinity());
    }
    return true;
  }
  if (kind == RTC_KIND_SPHERE && !((f & RTC_FLAG_TRANSFORMED) && d.xform && d.xform[i] >= 0)) {
    for (int k = 0; k < 4; k++)
      if (!std::isfinite(g[k])) return false;
    for (int k = 0; k < 3; k++) {  // Sphere.cs:212-232
      bmin[k] = g[k] - g[3];
      bmax[k] = g[k] + g[3];
    }
    return true;
  }
  return false;
}

void DescPrimitiveBounds(const rtc_scene_desc& d, int i, double bmin[3], double bmax[3]) {
  if (FastBounds(d, i, bmin, bmax)) return;
  DescPrimitiveBoundsGeneral(d, i, bmin, bmax);
}

void DescPrimitiveBoundsGeneral(const rtc_scene_desc& d, int i, double bmin[3], double bmax[3]) {
  static thread_local Primitive p;
  PrimitiveFromDesc(d, i, p);
  PrimitiveBounds(p, bmin, bmax);
  if (p.Kind == RTC_KIND_TRIANGLE) {
    for (int k = 0; k < 3; k++) {
      bmin[k] = std::nextafter(bmin[k], -std::numeric_limits<double>::infinity());
      bmax[k] = std::nextafter(bmax[k], std::numeric_limits<double>::infinity());
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Cameras
// ---------------------------------------------------------------------------------------------------------
rtc_camera Camera::InitRender(int width, int height) {
  rtc_camera c;
  std::memset(&c, 0, sizeof(c));
  c.kind = Kind;
  // Camera.InitRender, Cameras/Camera.cs:54-63
  c.w2 = width / 2.0;
  c.h2 = height / 2.0;
  Vec4D look = Normalize(lookAt - position);
  Vec4D side = Normalize(Cross(look, -up));
  up = Normalize(Cross(look, side));
  side = -side;
  if (Kind == RTC_CAMERA_FRUSTUM) {  // FrustumCamera.cs:24-31
    double tanY = std::tan(fovY / 2);
    c.tan_fov_x2 = tanY * (width / (double)height);
    c.tan_fov_y2 = -tanY;
  } else {  // OrthoCamera.cs:22-31
    double camW = (1 / c.w2);
    double camH = (1 / c.h2) * (height / (double)width);
    c.h_mult = camW * sizeMult;
    c.v_mult = -camH * sizeMult;
  }
  const Vec4D* src[4] = {&position, &look, &side, &up};
  double* dst[4] = {c.position, c.look, c.side, c.up};
  for (int i = 0; i < 4; i++) {
    dst[i][0] = src[i]->X;
    dst[i][1] = src[i]->Y;
    dst[i][2] = src[i]->Z;
  }
  c.image_plane = imagePlane;
  c.dof_amount = dofAmount;
  c.focal_length = focalLength;
  return c;
}

// ---------------------------------------------------------------------------------------------------------
// Scene
// ---------------------------------------------------------------------------------------------------------
void Scene::AddPrimitive(const Primitive& p) {  // Scene.cs:58-63
  prims_.push_back(p);
  prims_.back().ID = (int)prims_.size() - 1;
  flat_valid_ = false;
  ResetAccelerator();
}

void Scene::AddFlat(uint8_t kind, uint8_t flags, const double geom[12], const double material[14]) {
  flat_only_ = true;
  kind_.push_back(kind);
  flags_.push_back(flags);
  geom_.insert(geom_.end(), geom, geom + RTC_GEOM_STRIDE);
  material_.insert(material_.end(), material, material + RTC_MATERIAL_STRIDE);
  flat_valid_ = false;
  ResetAccelerator();
}

const rtc_scene_desc& Scene::Desc() {
  if (!flat_valid_) {
    if (!flat_only_) {
      size_t n = prims_.size();
      kind_.assign(n, 0);
      flags_.assign(n, 0);
      geom_.assign(n * RTC_GEOM_STRIDE, 0.0);
      material_.assign(n * RTC_MATERIAL_STRIDE, 0.0);
      xform_.assign(n, -1);
      xforms_.clear();
      for (size_t i = 0; i < n; i++) {
        const Primitive& p = prims_[i];
        double* g = &geom_[i * RTC_GEOM_STRIDE];
        uint8_t f = 0;
        if (p.TwoSided) f |= RTC_FLAG_TWOSIDED;
        if (p.Invert) f |= RTC_FLAG_INVERT;
        kind_[i] = p.Kind;
        if (p.Kind == RTC_KIND_TRIANGLE) {
          const Vec4D* v[4] = {&p.Vert0, &p.Edge0to1, &p.Edge0to2, &p.Normal};
          for (int k = 0; k < 4; k++) {
            g[k * 3 + 0] = v[k]->X;
            g[k * 3 + 1] = v[k]->Y;
            g[k * 3 + 2] = v[k]->Z;
          }
          if (p.Mirror) f |= RTC_FLAG_MIRROR;
          if (p.HasNormals) {
            f |= RTC_FLAG_VNORMALS;
            xform_[i] = (int32_t)(xforms_.size() / RTC_XFORM_STRIDE);
            size_t base = xforms_.size();
            xforms_.resize(base + RTC_XFORM_STRIDE, 0.0);
            const Vec4D* nn[3] = {&p.Norm0, &p.Norm1, &p.Norm2};
            for (int k = 0; k < 3; k++) {
              xforms_[base + k * 3 + 0] = nn[k]->X;
              xforms_[base + k * 3 + 1] = nn[k]->Y;
              xforms_[base + k * 3 + 2] = nn[k]->Z;
            }
          }
        } else if (p.Kind == RTC_KIND_SPHERE) {
          g[0] = p.Center.X; g[1] = p.Center.Y; g[2] = p.Center.Z;
          g[3] = p.RadiusValue;
          g[4] = p.RadiusSqr;
          if (p.Transformed) {
            f |= RTC_FLAG_TRANSFORMED;
            xform_[i] = (int32_t)(xforms_.size() / RTC_XFORM_STRIDE);
            size_t base = xforms_.size();
            xforms_.resize(base + RTC_XFORM_STRIDE, 0.0);
            std::memcpy(&xforms_[base], p.MatrixToWorld.D, 16 * sizeof(double));
            std::memcpy(&xforms_[base + 16], p.MatrixToObject.D, 16 * sizeof(double));
            std::memcpy(&xforms_[base + 32], p.MatrixToNormal.D, 16 * sizeof(double));
          }
        } else {
          g[0] = p.PlaneNormal.X; g[1] = p.PlaneNormal.Y; g[2] = p.PlaneNormal.Z;
          g[3] = p.OriginDistance;
        }
        flags_[i] = f;
        double* m = &material_[i * RTC_MATERIAL_STRIDE];
        const DoubleColor* cs[4] = {&p.Emission, &p.Diffuse, &p.Specular, &p.Refraction};
        for (int k = 0; k < 4; k++) {
          m[k * 3 + 0] = cs[k]->R;
          m[k * 3 + 1] = cs[k]->G;
          m[k * 3 + 2] = cs[k]->B;
        }
        m[12] = p.RefractiveIndex;
        m[13] = p.Shininess;
      }
    } else {
      xform_.assign(kind_.size(), -1);
      xforms_.clear();
    }
    desc_.n_prims = (int32_t)kind_.size();
    desc_.n_xforms = (int32_t)(xforms_.size() / RTC_XFORM_STRIDE);
    desc_.kind = kind_.data();
    desc_.flags = flags_.data();
    desc_.geom = geom_.data();
    desc_.xform = xform_.data();
    desc_.xforms = xforms_.empty() ? nullptr : xforms_.data();
    desc_.material = material_.data();
    flat_valid_ = true;
  }
  return desc_;
}

rtc_params Scene::Params(uint64_t seed) const {
  rtc_params p;
  std::memset(&p, 0, sizeof(p));
  p.width = Width;
  p.height = Height;
  p.recursion = Recursion;
  p.debug_geom = DebugGeom ? 1 : 0;
  p.ambient[0] = AmbientRGB.R;
  p.ambient[1] = AmbientRGB.G;
  p.ambient[2] = AmbientRGB.B;
  p.air_ior = AirRefractiveIndex;
  p.seed = seed;
  return p;
}

const std::vector<rtc_bvh_node>& Scene::Accelerator(int* root) {  // Scene.Prepare, Scene.cs:39-49
  if (nodes_.empty() && PrimitiveCount() > 0) {
    const rtc_scene_desc& d = Desc();
    int n = d.n_prims;
    std::vector<double> lo((size_t)n * 3), hi((size_t)n * 3);
    for (int i = 0; i < n; i++) {
      if (flat_only_)
        DescPrimitiveBounds(d, i, &lo[(size_t)i * 3], &hi[(size_t)i * 3]);
      else
        PrimitiveBounds(prims_[i], &lo[(size_t)i * 3], &hi[(size_t)i * 3]);
    }
    int threads = (int)std::thread::hardware_concurrency();
    root_ = BuildBVH(n, lo.data(), hi.data(), nodes_, threads > 0 ? threads : 1);
  }
  if (root) *root = root_;
  return nodes_;
}

// ---------------------------------------------------------------------------------------------------------
// Synthetic scenes: SplitMix64-driven so every consumer builds identical scenes.
// ---------------------------------------------------------------------------------------------------------
namespace {
struct SplitMix64 {
  uint64_t s;
  explicit SplitMix64(uint64_t seed) : s(seed) {}
  uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  double uniform() { return (double)(next() >> 11) * 0x1.0p-53; }
  double range(double a, double b) { return a + (b - a) * uniform(); }
};
}  // namespace

std::unique_ptr<Scene> MakeSynthetic(const std::string& name, int64_t n, uint64_t seed, double jitter) {
  std::unique_ptr<Scene> sc(new Scene());
  SplitMix64 rng(seed);
  sc->AmbientRGB = DoubleColor(1, 1, 1);
  sc->BackgroundRGB = DoubleColor::Black();
  sc->BackgroundAlpha = 0;
  Camera cam;
  cam.Kind = RTC_CAMERA_FRUSTUM;
  cam.initPosition = cam.position = Vec4D(0, 0, -3.5, 1);
  cam.initLookAt = cam.lookAt = Vec4D(0, 0, 0, 1);
  cam.initUp = cam.up = Vec4D(0, 1, 0, 0);
  cam.fovY = toRadians(40);
  cam.focalLength = Length(cam.initLookAt - cam.position);
  sc->Cameras.push_back(cam);
  if (name == "soup") {
    sc->Width = sc->Height = 2048;
    sc->Recursion = 4;
    if (jitter <= 0) jitter = 0.01;
    for (int64_t i = 0; i < n; i++) {
      Vec4D c(rng.range(-1, 1), rng.range(-1, 1), rng.range(-1, 1), 1);
      Vec4D v[3];
      for (int k = 0; k < 3; k++)
        v[k] = Vec4D(c.X + rng.range(-jitter, jitter), c.Y + rng.range(-jitter, jitter), c.Z + rng.range(-jitter, jitter), 1);
      Primitive t = Primitive::MakeTriangle(v[0], v[1], v[2], false);
      double g[12] = {t.Vert0.X, t.Vert0.Y, t.Vert0.Z, t.Edge0to1.X, t.Edge0to1.Y, t.Edge0to1.Z,
                      t.Edge0to2.X, t.Edge0to2.Y, t.Edge0to2.Z, t.Normal.X, t.Normal.Y, t.Normal.Z};
      double m[14] = {0};
      bool emissive = rng.uniform() < 0.01;
      double dr = rng.range(0.2, 0.9), dg = rng.range(0.2, 0.9), db = rng.range(0.2, 0.9);
      if (emissive) {
        m[0] = m[1] = m[2] = 8;
      } else {
        m[3] = dr; m[4] = dg; m[5] = db;
      }
      m[12] = 0;
      m[13] = 100;
      sc->AddFlat(RTC_KIND_TRIANGLE, RTC_FLAG_TWOSIDED, g, m);
    }
  } else if (name == "spheres") {
    sc->Width = 1920;
    sc->Height = 1080;
    sc->Recursion = 8;
    for (int64_t i = 0; i < n; i++) {
      double g[12] = {0};
      g[0] = rng.range(-1, 1); g[1] = rng.range(-1, 1); g[2] = rng.range(-1, 1);
      g[3] = rng.range(0.004, 0.012);
      g[4] = g[3] * g[3];
      double m[14] = {0};
      double dr = rng.range(0.2, 0.9), dg = rng.range(0.2, 0.9), db = rng.range(0.2, 0.9);
      switch (i % 3) {
        case 0:  // mirror
          m[6] = m[7] = m[8] = 0.9;
          m[13] = 1e6;
          break;
        case 1:  // glass
          m[6] = m[7] = m[8] = 0.9;
          m[9] = m[10] = m[11] = 0.9;
          m[12] = 1.52;
          m[13] = 1e5;
          break;
        default:
          m[3] = dr; m[4] = dg; m[5] = db;
          m[13] = 100;
      }
      if (i % 200 == 0) m[0] = m[1] = m[2] = 8;
      sc->AddFlat(RTC_KIND_SPHERE, RTC_FLAG_TWOSIDED, g, m);
    }
  } else {
    return nullptr;
  }
  return sc;
}

}  // namespace rtcore
