/*
 * rtcore_host.h — C façade over the C++ host layer (raytracercore_b200/host/), i.e. the part of the reference that
 * stays on the host either side of the kernels: SceneLoader (SceneLoader.cs), Scene (Raytracing/Scene.cs),
 * Camera.InitRender (Raytracing/Cameras/{Camera,FrustumCamera,OrthoCamera}.cs), the BVH build that replaces BVH.Construct
 * (Raytracing/Acceleration/BVH.cs:193) and a FullRaytracer-shaped progressive renderer
 * (Raytracing/FullRaytracer.cs). Lives in the same shared library as the kernels (librtcore_b200.so). Nothing in
 * the rtcs_scene_* group touches CUDA, so scene loading / flattening / BVH building work on a machine without a GPU; rendering
 * (rtcs_raytracer_*) requires one and fails loudly otherwise.
 */
#ifndef RTCORE_HOST_H
#define RTCORE_HOST_H

#include "rtcore_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rtcs_scene rtcs_scene;

/* Scene-file globals (Scene.cs:16-35) plus counts. */
typedef struct rtcs_globals {
  int32_t width, height, recursion, debug_geom;
  int32_t n_cameras, current_camera, n_prims, pad;
  double background[3];
  double background_alpha;
  double ambient[3];
  double air_ior;
} rtcs_globals;

/* SceneLoader.FromFile (SceneLoader.cs:112). Returns NULL with err == "" when the file does not exist
 * (the reference returns null, :430-439) and NULL with the LoaderException text otherwise. */
rtcs_scene* rtcs_scene_load(const char* path, char* err, int32_t err_cap);
rtcs_scene* rtcs_scene_parse(const char* text, char* err, int32_t err_cap);
/* BASELINE.json synthetic scenes: name "soup" (n triangles, vertex jitter) or "spheres" (n spheres). */
rtcs_scene* rtcs_scene_synthetic(const char* name, int64_t n, uint64_t seed, double jitter);
void rtcs_scene_free(rtcs_scene* s);

int rtcs_scene_globals(rtcs_scene* s, rtcs_globals* out);
/* Harness overrides of the public Scene fields (Scene.cs:16-17,26,33). Negative values leave a field unchanged. */
int rtcs_scene_override(rtcs_scene* s, int32_t width, int32_t height, int32_t recursion, int32_t current_camera);
int rtcs_scene_set_ambient(rtcs_scene* s, const double rgb[3]);
int rtcs_scene_set_debug_geom(rtcs_scene* s, int32_t on);

/* Flattened Scene.Primitives; pointers stay valid until the scene is freed or modified. */
int rtcs_scene_desc(rtcs_scene* s, rtc_scene_desc* out);
int rtcs_scene_params(rtcs_scene* s, uint64_t seed, rtc_params* out);
/* Camera.InitRender(width,height) of camera `index` (FullRaytracer.cs:269). */
int rtcs_scene_camera(rtcs_scene* s, int32_t index, int32_t width, int32_t height, rtc_camera* out);
/* Scene.Prepare (Scene.cs:39-49): builds (once) and returns the accelerator in reference shape. */
int rtcs_scene_bvh(rtcs_scene* s, const rtc_bvh_node** nodes, int32_t* n_nodes, int32_t* root);
/* AABB.CreateFromBounded of primitive i (AABB.cs:20-36). */
int rtcs_scene_primitive_bounds(rtcs_scene* s, int32_t i, double bmin[3], double bmax[3]);

/* AABB.CreateFromBounded of primitive i of an ABI scene description, as the builders compute it (general != 0: through a
 * full Primitive object, the path kept for transformed spheres and planes; the two must agree bit for bit). */
int rtcs_desc_primitive_bounds(const rtc_scene_desc* scene, int32_t i, int32_t general, double bmin[3], double bmax[3]);
/* Stand-alone builder over an ABI scene description; nodes must hold 2*n_prims entries. */
int rtcs_build_bvh(const rtc_scene_desc* scene, int32_t threads, rtc_bvh_node* nodes, int32_t* n_nodes, int32_t* root);

/* ---- FullRaytracer mirror (Raytracing/FullRaytracer.cs) ---------------------------------------------- */
typedef struct rtcs_raytracer rtcs_raytracer;
/* status callback: (user, status text, progress in [0,1)) — FullRaytracer.UpdateStatus (:91-94) without the Bitmap,
 * which the caller pulls with rtcs_raytracer_get_bitmap. Called from the render thread. */
typedef void (*rtcs_status_fn)(void* user, const char* status, double progress);

/* new FullRaytracer(scene, ...) (:66): `device` replaces the thread count; precision is RTC_F32 / RTC_F64. */
rtcs_raytracer* rtcs_raytracer_create(rtcs_scene* scene, int32_t device, int32_t precision, uint64_t seed,
                                      rtcs_status_fn status, void* user, char* err, int32_t err_cap);
void rtcs_raytracer_destroy(rtcs_raytracer* r);
/* Start() (:243) — blocking progressive render; returns when Stop() was requested (or max_samples > 0 reached).
 * samples_per_pass is the number of samples each GPU pass adds to every pixel; 0 = automatic (passes of about 8 Mi paths: 4 at
 * 1920x1080, 18 at 700x700, never more than 64). */
int rtcs_raytracer_start(rtcs_raytracer* r, uint32_t samples_per_pass, uint32_t max_samples);
void rtcs_raytracer_stop(rtcs_raytracer* r);    /* Stop()   (:409) */
void rtcs_raytracer_pause(rtcs_raytracer* r);   /* Pause()  (:377) */
void rtcs_raytracer_resume(rtcs_raytracer* r);  /* Resume() (:403) */
int rtcs_raytracer_is_running(rtcs_raytracer* r);   /* IsRunning  (:375) */
int rtcs_raytracer_is_paused(rtcs_raytracer* r);    /* IsPaused   (:401) */
int rtcs_raytracer_is_stopping(rtcs_raytracer* r);  /* IsStopping (:416) */
void rtcs_raytracer_set_exposure(rtcs_raytracer* r, double exposure); /* Exposure (:36) */
/* GetSampleSet(x,y) (:131-146): Color rgb, Samples, Misses of one pixel (x, y clamped like the reference). */
int rtcs_raytracer_get_sample_set(rtcs_raytracer* r, int32_t x, int32_t y, double rgb[3], uint32_t* samples, uint32_t* misses);
/* GetBitmap() (:179-205): width*height ARGB8, row-major. */
int rtcs_raytracer_get_bitmap(rtcs_raytracer* r, uint32_t* argb);
/* The underlying kernel context (for stats / multi-GPU plumbing). */
rtc_ctx* rtcs_raytracer_ctx(rtcs_raytracer* r);
const char* rtcs_raytracer_last_error(rtcs_raytracer* r);

#ifdef __cplusplus
}
#endif
#endif
