// host_api.cpp — C façade (include/rtcore_host.h) over the C++ host layer: scene loading/flattening/BVH build.
#include <cstdio>
#include <cstring>
#include <thread>

#include "../../include/rtcore_host.h"
#include "scene.h"

using namespace rtcore;

namespace {
void set_err(char* err, int32_t cap, const char* msg) {
  if (err && cap > 0) {
    std::snprintf(err, (size_t)cap, "%s", msg);
  }
}
}  // namespace

extern "C" {

rtcs_scene* rtcs_scene_load(const char* path, char* err, int32_t err_cap) {
  set_err(err, err_cap, "");
  if (!path) {
    set_err(err, err_cap, "path is null");
    return nullptr;
  }
  try {
    std::unique_ptr<Scene> sc = SceneLoader::FromFile(path);
    if (!sc) return nullptr;
    rtcs_scene* s = new rtcs_scene();
    s->scene = std::move(sc);
    return s;
  } catch (const std::exception& e) {
    set_err(err, err_cap, e.what());
    return nullptr;
  }
}

rtcs_scene* rtcs_scene_parse(const char* text, char* err, int32_t err_cap) {
  set_err(err, err_cap, "");
  if (!text) {
    set_err(err, err_cap, "text is null");
    return nullptr;
  }
  try {
    rtcs_scene* s = new rtcs_scene();
    s->scene = SceneLoader::FromString(text);
    return s;
  } catch (const std::exception& e) {
    set_err(err, err_cap, e.what());
    return nullptr;
  }
}

rtcs_scene* rtcs_scene_synthetic(const char* name, int64_t n, uint64_t seed, double jitter) {
  if (!name || n < 0) return nullptr;
  std::unique_ptr<Scene> sc = MakeSynthetic(name, n, seed, jitter);
  if (!sc) return nullptr;
  rtcs_scene* s = new rtcs_scene();
  s->scene = std::move(sc);
  return s;
}

void rtcs_scene_free(rtcs_scene* s) { delete s; }

int rtcs_scene_globals(rtcs_scene* s, rtcs_globals* out) {
  if (!s || !out) return RTC_ERR_INVALID;
  Scene& sc = *s->scene;
  std::memset(out, 0, sizeof(*out));
  out->width = sc.Width;
  out->height = sc.Height;
  out->recursion = sc.Recursion;
  out->debug_geom = sc.DebugGeom ? 1 : 0;
  out->n_cameras = (int32_t)sc.Cameras.size();
  out->current_camera = sc.CurrentCamera;
  out->n_prims = sc.PrimitiveCount();
  out->background[0] = sc.BackgroundRGB.R;
  out->background[1] = sc.BackgroundRGB.G;
  out->background[2] = sc.BackgroundRGB.B;
  out->background_alpha = sc.BackgroundAlpha;
  out->ambient[0] = sc.AmbientRGB.R;
  out->ambient[1] = sc.AmbientRGB.G;
  out->ambient[2] = sc.AmbientRGB.B;
  out->air_ior = sc.AirRefractiveIndex;
  return RTC_OK;
}

int rtcs_scene_override(rtcs_scene* s, int32_t width, int32_t height, int32_t recursion, int32_t current_camera) {
  if (!s) return RTC_ERR_INVALID;
  Scene& sc = *s->scene;
  if (current_camera >= 0) {
    if ((size_t)current_camera >= sc.Cameras.size()) return RTC_ERR_INVALID;
    sc.CurrentCamera = current_camera;
  }
  if (width >= 0) sc.Width = width;
  if (height >= 0) sc.Height = height;
  if (recursion >= 0) sc.Recursion = recursion;
  return RTC_OK;
}

int rtcs_scene_set_ambient(rtcs_scene* s, const double rgb[3]) {
  if (!s || !rgb) return RTC_ERR_INVALID;
  s->scene->AmbientRGB = DoubleColor(rgb[0], rgb[1], rgb[2]);
  return RTC_OK;
}

int rtcs_scene_set_debug_geom(rtcs_scene* s, int32_t on) {
  if (!s) return RTC_ERR_INVALID;
  s->scene->DebugGeom = on != 0;
  return RTC_OK;
}

int rtcs_scene_desc(rtcs_scene* s, rtc_scene_desc* out) {
  if (!s || !out) return RTC_ERR_INVALID;
  *out = s->scene->Desc();
  return RTC_OK;
}

int rtcs_scene_params(rtcs_scene* s, uint64_t seed, rtc_params* out) {
  if (!s || !out) return RTC_ERR_INVALID;
  *out = s->scene->Params(seed);
  return RTC_OK;
}

int rtcs_scene_camera(rtcs_scene* s, int32_t index, int32_t width, int32_t height, rtc_camera* out) {
  if (!s || !out) return RTC_ERR_INVALID;
  Scene& sc = *s->scene;
  if (index < 0 || (size_t)index >= sc.Cameras.size() || width <= 0 || height <= 0) return RTC_ERR_INVALID;
  *out = sc.Cameras[index].InitRender(width, height);
  return RTC_OK;
}

int rtcs_scene_bvh(rtcs_scene* s, const rtc_bvh_node** nodes, int32_t* n_nodes, int32_t* root) {
  if (!s || !nodes || !n_nodes || !root) return RTC_ERR_INVALID;
  int r = -1;
  const std::vector<rtc_bvh_node>& v = s->scene->Accelerator(&r);
  *nodes = v.data();
  *n_nodes = (int32_t)v.size();
  *root = r;
  return RTC_OK;
}

int rtcs_scene_primitive_bounds(rtcs_scene* s, int32_t i, double bmin[3], double bmax[3]) {
  if (!s || !bmin || !bmax) return RTC_ERR_INVALID;
  Scene& sc = *s->scene;
  if (i < 0 || i >= sc.PrimitiveCount()) return RTC_ERR_INVALID;
  if ((size_t)i < sc.Primitives().size())
    PrimitiveBounds(sc.Primitives()[i], bmin, bmax);
  else
    DescPrimitiveBounds(sc.Desc(), i, bmin, bmax);
  return RTC_OK;
}

int rtcs_desc_primitive_bounds(const rtc_scene_desc* d, int32_t i, int32_t general, double bmin[3], double bmax[3]) {
  if (!d || !bmin || !bmax || i < 0 || i >= d->n_prims) return RTC_ERR_INVALID;
  if (general)
    DescPrimitiveBoundsGeneral(*d, i, bmin, bmax);
  else
    DescPrimitiveBounds(*d, i, bmin, bmax);
  return RTC_OK;
}

int rtcs_build_bvh(const rtc_scene_desc* d, int32_t threads, rtc_bvh_node* nodes, int32_t* n_nodes, int32_t* root) {
  if (!d || !nodes || !n_nodes || !root || d->n_prims < 0) return RTC_ERR_INVALID;
  int n = d->n_prims;
  std::vector<double> lo((size_t)n * 3), hi((size_t)n * 3);
  for (int i = 0; i < n; i++) DescPrimitiveBounds(*d, i, &lo[(size_t)i * 3], &hi[(size_t)i * 3]);
  std::vector<rtc_bvh_node> v;
  if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
  *root = BuildBVH(n, lo.data(), hi.data(), v, threads > 0 ? threads : 1);
  *n_nodes = (int32_t)v.size();
  if (!v.empty()) std::memcpy(nodes, v.data(), v.size() * sizeof(rtc_bvh_node));
  return RTC_OK;
}

}  // extern "C"
