// scene.h — host-side mirror of the reference's scene model for the render path: Scene (Raytracing/Scene.cs),
// Primitive + Triangle/Sphere/Plane (Raytracing/Primitives/*.cs), Cube (Raytracing/Objects/Cube.cs), the cameras
// (Raytracing/Cameras/*.cs) and SceneLoader (SceneLoader.cs). Names and member meaning follow the reference; the
// storage is flat so a Scene can be handed to the C ABI (include/rtcore_b200.h) without conversion.
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/rtcore_b200.h"
#include "vecmath.h"

namespace rtcore {

struct DoubleColor {  // DoubleColor.cs
  double R = 0, G = 0, B = 0;
  DoubleColor() = default;
  DoubleColor(double r, double g, double b) : R(r), G(g), B(b) {}
  explicit DoubleColor(double v) : R(v), G(v), B(v) {}
  bool operator==(const DoubleColor& o) const { return R == o.R && G == o.G && B == o.B; }
  bool operator!=(const DoubleColor& o) const { return !(*this == o); }
  static DoubleColor Placeholder() { return DoubleColor(-1); }  // DoubleColor.cs:15
  static DoubleColor Black() { return DoubleColor(0); }
};

// Primitive (Primitives/Primitive.cs) with the geometry of Triangle / Sphere / Plane folded in.
struct Primitive {
  int ID = -1;
  uint8_t Kind = RTC_KIND_TRIANGLE;
  bool TwoSided = false, Invert = false;
  DoubleColor Emission, Diffuse, Specular, Refraction;  // raw backing fields (Primitive.cs:20-21,98-124)
  double Shininess = 100;                               // Primitive.cs:31
  double RefractiveIndex = 0;
  // Triangle (Triangle.cs:22-29)
  Vec4D Vert0, Vert1, Vert2, Norm0, Norm1, Norm2, Edge0to1, Edge0to2, Normal;
  bool Mirror = false, HasNormals = false;
  // Sphere (Sphere.cs:11-21)
  Vec4D Center;
  double RadiusValue = 0, RadiusSqr = 0;
  bool Transformed = false;
  Mat4x4D MatrixToObject, MatrixToWorld, MatrixToNormal;
  // Plane (Plane.cs:13-14)
  Vec4D PlaneNormal;
  double OriginDistance = 0;

  static Primitive MakeTriangle(const Vec4D& p0, const Vec4D& p1, const Vec4D& p2, bool mirror);  // Triangle.cs:31-40
  static Primitive MakeTriangle(const Vec4D& p0, const Vec4D& n0, const Vec4D& p1, const Vec4D& n1, const Vec4D& p2,
                                const Vec4D& n2);                                                 // Triangle.cs:46-52
  static Primitive MakeRectangle(const Vec4D& upOrigin, const Vec4D& upDir, const Vec4D& normal, double width,
                                 double height);                                                  // Triangle.cs:13-20
  static Primitive MakeSphere(const Vec4D& center, double radius);                                // Sphere.cs:23-27
  static Primitive MakePlane(double origin, const Vec4D& normal);                                 // Plane.cs:18-22
  void Transform(const Mat4x4D& forward, const Mat4x4D& inverse);  // Triangle.cs:68-74, Sphere.cs:29-37, Plane.cs:29-34
  void Recalculate();                                              // Triangle.cs:54-66
  // IBoundedObject (Triangle.cs:226-263, Sphere.cs:212-232, Plane.cs:24-27,68-74)
  Vec4D GetCenter() const;
  double GetMaxCenterDistance(const Vec4D& direction) const;
};

struct Camera {  // Cameras/Camera.cs, FrustumCamera.cs, OrthoCamera.cs
  int Kind = RTC_CAMERA_FRUSTUM;
  Vec4D initPosition, initLookAt, initUp;
  Vec4D position, lookAt, up;
  double fovY = 0;      // radians (FrustumCamera.cs:15-22)
  double sizeMult = 0;  // OrthoCamera.cs:7
  double imagePlane = 0, dofAmount = 0, focalLength = 0;
  // Camera.InitRender + subclass InitRender: fills the post-init state the kernels read.
  rtc_camera InitRender(int width, int height);
};

class Scene {  // Raytracing/Scene.cs
 public:
  int Width = 0, Height = 0;
  DoubleColor BackgroundRGB = DoubleColor::Black();
  double BackgroundAlpha = 0;
  DoubleColor AmbientRGB = DoubleColor::Black();
  bool DebugGeom = false;
  int CurrentCamera = 0;
  std::vector<Camera> Cameras;
  int Recursion = 3;
  double AirRefractiveIndex = 1.000293;

  void AddPrimitive(const Primitive& p);  // Scene.cs:58-63
  const std::vector<Primitive>& Primitives() const { return prims_; }
  int PrimitiveCount() const { return flat_only_ ? (int)kind_.size() : (int)prims_.size(); }
  // Bulk path for the synthetic 1M/10M-primitive scenes: append straight to the flat arrays (no Primitive objects).
  void AddFlat(uint8_t kind, uint8_t flags, const double geom[12], const double material[14]);

  // Flattened view for rtc_upload_scene; arrays are owned by the Scene and rebuilt lazily.
  const rtc_scene_desc& Desc();
  rtc_params Params(uint64_t seed) const;

  // Replacement for Scene.Prepare -> BVH.Construct (Scene.cs:39-49): reference-shaped nodes.
  const std::vector<rtc_bvh_node>& Accelerator(int* root);
  void ResetAccelerator() { nodes_.clear(); root_ = -1; }
  bool HasAccelerator() const { return !nodes_.empty(); }  // Scene._Accelerator != null (Scene.cs:41)

 private:
  std::vector<Primitive> prims_;
  bool flat_valid_ = false, flat_only_ = false;
  std::vector<uint8_t> kind_, flags_;
  std::vector<double> geom_, xforms_, material_;
  std::vector<int32_t> xform_;
  rtc_scene_desc desc_{};
  std::vector<rtc_bvh_node> nodes_;
  int root_ = -1;
};

class LoaderException : public std::runtime_error {  // SceneLoader.cs:16-26
 public:
  std::string Command;
  int Line;
  LoaderException(const std::string& command, int line, const std::string& inner)
      : std::runtime_error("Error while parsing command " + command + " on line " + std::to_string(line) + ": " + inner),
        Command(command), Line(line) {}
};

namespace SceneLoader {                                        // SceneLoader.cs:112-441
std::unique_ptr<Scene> FromFile(const std::string& filename);   // nullptr when the file is missing (:430-439)
std::unique_ptr<Scene> FromString(const std::string& text);
}  // namespace SceneLoader

// Leaf box exactly as AABB.CreateFromBounded (AABB.cs:20-36).
void PrimitiveBounds(const Primitive& p, double bmin[3], double bmax[3]);
// Leaf box from the flattened description alone (the ABI carries Vert0/Edge0to1/Edge0to2, not Vert1/Vert2, so
// triangle boxes are rebuilt from v0, v0+e1, v0+e2 and widened by one ulp to stay conservative).
void DescPrimitiveBounds(const rtc_scene_desc& d, int i, double bmin[3], double bmax[3]);
// the same through a full Primitive (what DescPrimitiveBounds falls back to for transformed spheres and planes)
void DescPrimitiveBoundsGeneral(const rtc_scene_desc& d, int i, double bmin[3], double bmax[3]);
// Replacement for BVH.Construct (BVH.cs:50-236): binned-SAH tree with exactly one primitive per leaf
// (BVH.cs:256-264) over the n leaf boxes (bmin/bmax: n*3). Primitives with infinite boxes (planes) are chained
// above the root. Nodes come out in reference shape; returns the root index (-1 for n == 0).
int BuildBVH(int32_t n, const double* bmin, const double* bmax, std::vector<rtc_bvh_node>& nodes, int threads);

// Synthetic scenes of BASELINE.json: "soup" (n triangles), "spheres" (n spheres).
std::unique_ptr<Scene> MakeSynthetic(const std::string& name, int64_t n, uint64_t seed, double jitter);

}  // namespace rtcore

// Handle type of the C façade (include/rtcore_host.h).
struct rtcs_scene {
  std::shared_ptr<rtcore::Scene> scene;  // shared with every rtcs_raytracer created over it: freeing the handle does not pull
                                         // the scene from under a running render
};
