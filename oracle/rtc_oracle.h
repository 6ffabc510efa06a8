/*
 * rtc_oracle.h — C interface of the CPU parity oracle (TEST INFRASTRUCTURE, not product code).
 *
 * The oracle is a from-scratch C++ restatement, in IEEE f64, of the reference's CPU render path
 * (RaytracerCore/Raytracing/{Raytracer,Scene,Hit,SampleSet}.cs, Acceleration/{BVH,AABB}.cs,
 * Primitives/{Primitive,Triangle,Sphere,Plane}.cs, Cameras/{Camera,FrustumCamera,OrthoCamera}.cs, Vectors/{Vec4D,Mat4x4D,Ray,SIMDHelpers,MatrixTransforms}.cs, Util.cs, DoubleColor.cs).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * PINNING: the reference ships no tests, golden vectors or fixtures and cannot be built or run in
 * this image (no .NET SDK; WinForms target). What it does ship is two renders made by the real program
 * (Screenshots/bounce-with-lens.png, Screenshots/die.png); tests/test_screenshots.py checks this oracle against
 * them (tests/golden/screenshots.npz): silhouettes (alpha = hit fraction per pixel, including the depth-of-field
 * blur of die.txt) agree to 1e-3, converged radiance to 1-3 % over the whole image and 10 % per coarse tile, with the
 * UI exposure -- the one setting the bitmaps do not record -- at 1.0 for die.png and 1.5 for bounce-with-lens.png.
 * Beyond that the oracle is pinned by hand-derived analytic known answers (tests/golden/kat.json) and by
 * line-by-line citation of the reference. Per-ray quantities (hit index, inside flag, t, normal) have no
 * reference-made golden data: for those, parity remains UNPINNED by the reference.
 */
#ifndef RTC_ORACLE_H
#define RTC_ORACLE_H

#include "../include/rtcore_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_scene orc_scene;

/* Copies everything. nodes/root is the BVH the queries walk (the reference walks Scene.Accelerator). */
orc_scene* orc_scene_create(const rtc_scene_desc* scene, int32_t n_nodes, const rtc_bvh_node* nodes, int32_t root,
                            const rtc_camera* camera, const rtc_params* params);
void orc_scene_destroy(orc_scene* s);
void orc_set_camera(orc_scene* s, const rtc_camera* camera);
void orc_set_params(orc_scene* s, const rtc_params* params);
/* Process-wide switch of the self-hit rule (Util.RayHitMatches, Util.cs:179-192): 0 = the reference's (default); 1 = the rule
 * of the library's f32 mode (flat primitive == skip primitive is always the self-hit; spheres: positional rule with 1e-9 instead
 * of 1e-24), restated in f64. Only tests/golden/make_shading_fixtures.py uses 1, to produce the fixture that isolates this
 * one documented deviation of the f32 mode on scenes whose reference image depends on f64 rounding noise (BASELINE C4). */
void orc_set_selfhit_mode(int mode);

/* mode 0: Scene.RayTracePrimitives with the accelerator (Scene.cs:71-92): collect every pierced leaf
 *         (BVH.cs:295-331), stable sort by Near, early-out on Near > previous.Far.
 * mode 1: the reference's accelerator-less loop over all primitives in ID order (Scene.cs:95-107).
 * Returns the number of rays whose mode-0 and mode-1 answers differ when check_both != 0 (else 0). */
int64_t orc_trace_closest(orc_scene* s, int64_t n, const rtc_ray* rays, const rtc_hit* skip, rtc_hit* out, int mode,
                          int check_both, int threads);

/* Raytracer.GetCameraRay (Raytracer.cs:262-282) with the Philox stream of (pixel, sample). */
void orc_camera_rays(orc_scene* s, int64_t n, const int32_t* xy, const uint32_t* sample, rtc_ray* out);

/* n_samples passes over [x0,x1)x[y0,y1), accumulated into SampleSet planes (row-major y*width+x) exactly as
 * FullRaytracer.cs:326-339. threads worker threads over tiles (FullRaytracer.cs:271-302). Returns the number of
 * Scene.RayTrace calls made. */
uint64_t orc_render(orc_scene* s, int32_t x0, int32_t y0, int32_t x1, int32_t y1, uint32_t first_sample,
                    uint32_t n_samples, int threads, double* rgb_sum, uint32_t* samples, uint32_t* misses);
/* The same over a lattice of the whole frame, pixels (off_x + i stride_x, off_y + j stride_y): a bounded sample that keeps
 * the frame's mix of rays (bench.py's CPU legs). */
uint64_t orc_render_lattice(orc_scene* s, int32_t stride_x, int32_t stride_y, int32_t off_x, int32_t off_y, uint32_t first_sample,
                            uint32_t n_samples, int threads, double* rgb_sum, uint32_t* samples, uint32_t* misses);

/* One sample of every pixel, raw GetColor output (Placeholder = -1,-1,-1 for misses). out: w*h*3. */
void orc_render_samples(orc_scene* s, uint32_t sample, int threads, double* out_rgb);

/* The ray batches of SURVEY.md appendix C: every Scene.RayTrace call (Raytracer.cs:77) that GetColor makes for the n paths
 * (x, y, sample) -- the ray, the skip hit handed in (prim -1 = none), the oracle's answer and the bounce index -- in path
 * order, bounce order within a path. Returns the number of segments (stores at most `capacity`). */
int64_t orc_dump_path_rays(orc_scene* s, int64_t n, const int32_t* xy, const uint32_t* sample, int threads, int64_t capacity,
                           rtc_ray* rays, rtc_hit* skip, rtc_hit* hits, int32_t* bounce);

/* Raytracer.GetDebugTrace(x,y) (Raytracer.cs:254-260). */
void orc_debug_trace(orc_scene* s, int32_t x, int32_t y, uint32_t sample, int32_t capacity, rtc_debug_ray* out,
                     int32_t* n);

/* DebugRaycaster.GetColor's queries per pixel (DebugRaycaster.cs:170-215,236): mode 0 = hit Primitive.ID or -1,
 * mode 1 = BVH.GetIntersectionCount (BVH.cs:352-363). out: w*h int32. */
void orc_debug_raycast(orc_scene* s, int32_t mode, int32_t* out);

/* SampleSet.GetOutput over a whole image (SampleSet.cs:61-113, FullRaytracer.cs:179-205). */
void orc_tonemap(int32_t w, int32_t h, const double* rgb_sum, const uint32_t* samples, const uint32_t* misses,
                 double exposure, const double back_rgb[3], double back_a, uint32_t* argb);

/* Philox4x32-10 block and the two uniforms in [0,1) the render path derives from it. */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void orc_uniforms(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t stage, uint32_t block, double out[2]);

/* Vec4D.CreateHorizon (Vec4D.cs:52-58) — exposed for known-answer tests. */
void orc_create_horizon(const double pole[3], double z, double theta, double out[3]);
/* AABB.IntersectAVX (AABB.cs:107-142): returns 1 and near/far, or 0 for a miss. */
int orc_aabb_intersect(const double bmin[3], const double bmax[3], const rtc_ray* ray, double* near_out, double* far_out);

#ifdef __cplusplus
}
#endif
#endif
