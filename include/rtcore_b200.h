/*
 * rtcore_b200.h — C ABI of librtcore_b200.so, the B200 (sm_100a) wavefront path-tracing backend that
 * replaces the CPU render loop of Zaggy1024/RaytracerCore.
 *
 * The reference has no FFI of its own; the seam this ABI replaces is the managed call chain
 *   FullRaytracer.Start -> Raytracer.Render -> Raytracer.GetColor -> Scene.RayTrace -> SampleSet
 * (RaytracerCore/Raytracing/FullRaytracer.cs:243, Raytracer.cs:294/:65, Scene.cs:113, SampleSet.cs:32).
 * Each entry point below names the reference member it stands in for. All pointers are plain host
 * pointers unless a name says "device"; arrays are caller-owned and copied during the call. One host
 * thread per handle at a time, one handle per GPU. Every function returns RTC_OK (0) or an RTC_ERR_* code;
 * rtc_last_error() gives the text. There is no CPU fallback: without a CUDA device rtc_create() fails.
 */
#ifndef RTCORE_B200_H
#define RTCORE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTC_ABI_VERSION 1

/* ---- status codes ---------------------------------------------------------------------------------- */
enum {
  RTC_OK = 0,
  RTC_ERR_INVALID = 1,   /* bad argument (null pointer, negative count, index out of range)            */
  RTC_ERR_CUDA = 2,      /* a CUDA runtime call failed; text in rtc_last_error                         */
  RTC_ERR_STATE = 3,     /* call order violated (e.g. render before scene/bvh/camera/params are set)   */
  RTC_ERR_NOMEM = 4,     /* host or device allocation failed                                           */
  RTC_ERR_UNSUPPORTED = 5,
  RTC_ERR_NCCL = 6
};

/* ---- arithmetic mode ------------------------------------------------------------------------------- */
enum {
  RTC_F32 = 0,           /* production mode: float geometry/BVH, tolerance 1e-4 against the f64 oracle  */
  RTC_F64 = 1            /* parity mode: restates the reference's IEEE-f64 arithmetic operation by      */
                         /* operation (FMA exactly where the reference's AVX path fuses)                */
};

/* ---- primitives: Raytracing/Primitives/{Primitive,Triangle,Sphere,Plane}.cs ------------------------- */
enum {
  RTC_KIND_TRIANGLE = 0, /* Triangle.cs — also the parallelogram ("Mirror") used by every Cube face      */
  RTC_KIND_SPHERE = 1,   /* Sphere.cs — optionally affine-transformed (RTC_FLAG_TRANSFORMED)            */
  RTC_KIND_PLANE = 2     /* Plane.cs                                                                    */
};

enum {
  RTC_FLAG_MIRROR = 1,       /* Triangle.Mirror        (Triangle.cs:25)                                 */
  RTC_FLAG_TWOSIDED = 2,     /* Primitive.TwoSided     (Primitive.cs:88)                                */
  RTC_FLAG_INVERT = 4,       /* Primitive.Invert       (Primitive.cs:92)                                */
  RTC_FLAG_TRANSFORMED = 8,  /* Sphere.Transformed     (Sphere.cs:16)                                   */
  RTC_FLAG_VNORMALS = 16     /* Triangle.HasNormals    (Triangle.cs:26); vnormals[] row is used         */
};

#define RTC_GEOM_STRIDE 12
#define RTC_MATERIAL_STRIDE 14
#define RTC_XFORM_STRIDE 48

/*
 * Flattened Scene.Primitives (Scene.cs:160), indexed by Primitive.ID (= insertion order, Scene.cs:58-63).
 *   geom[i*12..]   triangle: Vert0.Position xyz, Edge0to1 xyz, Edge0to2 xyz, Normal xyz (Triangle.cs:22-29,54-66)
 *                  sphere:   Center xyz, RadiusValue, RadiusSqr, 7 unused              (Sphere.cs:11-14)
 *                  plane:    Normal xyz, OriginDistance, 8 unused                      (Plane.cs:13-14)
 *   xform[i]       row in xforms[] for a transformed sphere / vertex-normal triangle, else -1
 *   xforms[j*48..] sphere:   MatrixToWorld (=world->object), MatrixToObject (=object->world), MatrixToNormal,
 *                            each 16 doubles row-major D00..D33 (Sphere.cs:17-19, Mat4x4D.cs:21-39)
 *                  triangle: Vert0.Normal, Vert1.Normal, Vert2.Normal (9 doubles), rest unused
 *   material[i*14..] Emission rgb, Diffuse rgb, Specular rgb, Refraction rgb (raw backing fields),
 *                  RefractiveIndex, Shininess                                          (Primitive.cs:16-129)
 */
typedef struct rtc_scene_desc {
  int32_t n_prims;
  int32_t n_xforms;
  const uint8_t* kind;
  const uint8_t* flags;
  const double* geom;
  const int32_t* xform;    /* may be NULL when n_xforms == 0 */
  const double* xforms;    /* may be NULL when n_xforms == 0 */
  const double* material;
} rtc_scene_desc;

/*
 * One node of the reference-shaped binary BVH (Acceleration/BVH.cs:239-285): Volume (AABB.Minimum/Maximum,
 * AABB.cs:45-46), Left/Right as node indices, and for leaves (IsLeaf) the LeafID = primitive ID. Exactly one
 * primitive per leaf (BVH.cs:256-264). Infinite boxes (planes, Plane.cs:68-74) are allowed.
 */
typedef struct rtc_bvh_node {
  double bmin[3];
  double bmax[3];
  int32_t left;   /* node index, -1 for a leaf */
  int32_t right;  /* node index, -1 for a leaf */
  int32_t prim;   /* primitive ID for a leaf, -1 for an inner node */
  int32_t pad;
} rtc_bvh_node;

/* Camera after Camera.InitRender / FrustumCamera.InitRender / OrthoCamera.InitRender
 * (Cameras/Camera.cs:54-63, FrustumCamera.cs:24-31, OrthoCamera.cs:22-31). */
enum { RTC_CAMERA_FRUSTUM = 0, RTC_CAMERA_ORTHO = 1 };
typedef struct rtc_camera {
  int32_t kind;
  int32_t pad;
  double position[3];
  double look[3];
  double side[3];
  double up[3];
  double w2, h2;
  double tan_fov_x2, tan_fov_y2; /* frustum: tanFOVX2, tanFOVY2 (already negated, FrustumCamera.cs:30) */
  double h_mult, v_mult;         /* ortho: hMult, vMult (OrthoCamera.cs:29-30)                         */
  double image_plane, dof_amount, focal_length; /* Camera.cs:25-27 */
} rtc_camera;

/* Scene globals read by the render loop (Scene.cs:16-35). ambient == (-1,-1,-1) is DoubleColor.Placeholder,
 * i.e. "ambient miss" (SceneLoader.cs:183-188): bounced misses then count as Misses. */
typedef struct rtc_params {
  int32_t width, height;
  int32_t recursion;
  int32_t debug_geom;
  double ambient[3];
  double air_ior;       /* Scene.AirRefractiveIndex, 1.000293 */
  uint64_t seed;        /* Philox4x32-10 key; replaces the unseeded System.Random of Raytracer.cs:48 */
} rtc_params;

/* Ray (Vectors/Ray.cs:27-29) and Hit (Raytracing/Hit.cs:14-20). prim == -1 means "no hit" / "no skip". */
typedef struct rtc_ray {
  double origin[3];
  double dir[3];
} rtc_ray;

typedef struct rtc_hit {
  int32_t prim;
  int32_t inside;
  double t;            /* Hit.Distance */
  double position[3];
  double normal[3];
} rtc_hit;

/* Per-bounce record of Raytracer.DebugRay (Raytracer.cs:28-33); type uses Raytracer.BounceType order (:14-26). */
typedef struct rtc_debug_ray {
  rtc_hit hit;
  int32_t type;
  int32_t pad;
  double fresnel_ratio;
} rtc_debug_ray;

enum {
  /* (RTC_K_COMPACT: the stream compaction itself is fused into the shade kernel; this slot times the one-thread
   *  bookkeeping launch between bounces) */
  RTC_K_RAYGEN = 0, RTC_K_TRACE = 1, RTC_K_SHADE = 2, RTC_K_COMPACT = 3, RTC_K_ACCUMULATE = 4, RTC_K_COUNT = 5
};

typedef struct rtc_stats {
  uint64_t paths;                 /* camera paths started (== Raytracer.GetColor(x,y) calls)            */
  uint64_t rays;                  /* closest-hit queries (== Scene.RayTrace calls), all bounces          */
  uint64_t launches[RTC_K_COUNT]; /* kernel launches per kernel family                                   */
  double ms[RTC_K_COUNT];         /* CUDA-event time per kernel family (only while timing is enabled)    */
  uint64_t nodes_visited;         /* only with rtc_set_option(RTC_OPT_COUNTERS,1): BVH nodes fetched     */
  uint64_t prims_tested;          /*   "                                          primitive tests        */
  uint64_t node_steps;            /*   "   warp-level node iterations of the f32 trace scheduler                */
  uint64_t leaf_steps;            /*   "   warp-level leaf iterations                                           */
} rtc_stats;

enum {
  RTC_OPT_KERNEL_TIMING = 1, /* record CUDA events around every launch (rtc_stats.ms)                  */
  RTC_OPT_COUNTERS = 2,      /* instrumented traversal (nodes_visited / prims_tested)                  */
  RTC_OPT_MAX_PATHS = 3,     /* size of the path pool (paths in flight, all wavefronts together)        */
  RTC_OPT_WAVES = 4,         /* wavefronts in flight per rtc_render: 2 (default; alternate bands of the frame run on two
                                streams, each with half the pool) or 1. RTC_OPT_KERNEL_TIMING implies 1.             */
  RTC_OPT_REORDER = 5        /* RTC_F32: sort the live queue between bounces by the Morton cell of the ray origins (1) or by cell
                                and direction octant (2); 0 = off. Same image bit for bit; pays for scenes beyond the L2.   */
};

/* ---- lifetime -------------------------------------------------------------------------------------- */
typedef struct rtc_ctx rtc_ctx;

int rtc_abi_version(void);
/* Number of CUDA devices visible (0 when there is none or the driver is missing). */
int rtc_device_count(void);
/* FullRaytracer ctor (FullRaytracer.cs:66): one context per GPU. precision is RTC_F32 or RTC_F64. */
int rtc_create(int device, int precision, rtc_ctx** out);
void rtc_destroy(rtc_ctx* ctx);
/* Text of the last error on this context (ctx == NULL: last error of a failed rtc_create on this thread). */
const char* rtc_last_error(rtc_ctx* ctx);
int rtc_set_option(rtc_ctx* ctx, int option, int64_t value);
/* Run all of this context's work on the caller's CUDA stream (a cudaStream_t passed as void*; NULL restores the
 * context's own stream). Lets a host that already owns a stream (and times it with its own events) order the
 * render passes and the per-frame collective with its other work. */
int rtc_set_stream(rtc_ctx* ctx, void* cuda_stream);

/* ---- scene hand-over: what Scene.Prepare + FullRaytracer.Start set up (FullRaytracer.cs:253-269) ----- */
/* Scene.Primitives + materials (Scene.cs:28,160). Invalidates any BVH previously set. */
int rtc_upload_scene(rtc_ctx* ctx, const rtc_scene_desc* scene);
/* Scene.Accelerator as built by the host (BVH.Construct, BVH.cs:193): reference topology, n_nodes nodes. */
int rtc_upload_bvh(rtc_ctx* ctx, int32_t n_nodes, const rtc_bvh_node* nodes, int32_t root);
/* Replacement for BVH.Construct (BVH.cs:50-236): binned-SAH build over the uploaded primitives, leaf boxes
 * exactly as AABB.CreateFromBounded (AABB.cs:20-36); planes are chained above the root. */
int rtc_build_bvh(rtc_ctx* ctx);
/* The same on the device: the reference's agglomerative clustering with the merged box's surface area as the distance
 * (BVH.cs:12-21, 50-191) made data-parallel -- primitives ordered along a Morton curve, every round each cluster picks its
 * cheapest neighbour within `radius` positions (0 = default 16) and mutual choices merge (PLOC). One primitive per leaf,
 * exact f64 boxes; rtc_get_bvh returns the tree in the reference's shape. rounds (may be NULL) receives the number of
 * clustering rounds. */
int rtc_build_bvh_device(rtc_ctx* ctx, int32_t radius, int32_t* rounds);
/* Scene.Prepare (Scene.cs:39-49) wholly on the device: the tree is built there -- RTC_BUILDER_SAH: the binned-SAH tree of
 * rtc_build_bvh, node for node; RTC_BUILDER_PLOC: the clustering of rtc_build_bvh_device with `radius` -- and, in RTC_F32 mode,
 * collapsed, quantised and packed into the device layout there too, byte for byte what rtc_build_bvh / rtc_upload_bvh produce
 * on the host. Nothing comes back over PCIe: rtc_get_bvh and rtc_bake fetch the tree / the image from the device when asked.
 * RTC_F64 mode (the parity mode) builds on the device and flattens on the host. stats may be NULL. */
enum { RTC_BUILDER_SAH = 0, RTC_BUILDER_PLOC = 1 };
typedef struct rtc_prepare_stats {
  double boxes_ms;    /* leaf boxes (AABB.CreateFromBounded) on the host threads                     */
  double build_ms;    /* box upload + tree construction                                               */
  double flatten_ms;  /* collapse + quantisation + records (includes waiting for the scene arrays)  */
  double total_ms;
  int32_t build_levels; /* SAH: tree levels; PLOC: clustering rounds */
  int32_t wide_depth;   /* depth of the 8-wide tree (RTC_F32)         */
  int32_t n_wide_nodes;
  int32_t n_bounded;
} rtc_prepare_stats;
int rtc_prepare_device(rtc_ctx* ctx, int32_t builder, int32_t radius, rtc_prepare_stats* stats);
/* Scene.Prepare caches the accelerator on the host (Scene.cs:39-49); the equivalent here is a host-resident image of
 * the device layout (pinned memory), made once from the current scene + BVH and re-uploaded with plain H2D copies at
 * every FullRaytracer.Start(). A baked image belongs to one arithmetic mode. */
typedef struct rtc_baked rtc_baked;
int rtc_bake(rtc_ctx* ctx, rtc_baked** out);
int rtc_upload_baked(rtc_ctx* ctx, const rtc_baked* baked); /* asynchronous; rtc_baked_free waits for uploads still reading the image */
int64_t rtc_baked_bytes(const rtc_baked* baked);
/* One of the RTC_BAKED_SEGMENTS arrays of the image (read-only view into the baked image's own memory, valid until
 * rtc_baked_free): what a host that caches prepared scenes on disk would write out, and what the tests compare. */
enum { RTC_BAKED_SEGMENTS = 10 };
int rtc_baked_segment(const rtc_baked* baked, int32_t segment, const void** data, int64_t* bytes);
void rtc_baked_free(rtc_baked* baked);
/* Read the current tree back in reference shape (for SceneInspector.cs:226-265 and for the parity oracle). */
int rtc_get_bvh_size(rtc_ctx* ctx, int32_t* n_nodes, int32_t* root);
int rtc_get_bvh(rtc_ctx* ctx, int32_t capacity, rtc_bvh_node* nodes);
/* Scene.Camera after InitRender (FullRaytracer.cs:269). */
int rtc_set_camera(rtc_ctx* ctx, const rtc_camera* camera);
/* Scene.Width/Height/Recursion/AmbientRGB/DebugGeom/AirRefractiveIndex. Changing width/height reallocates
 * and clears the accumulation buffer (new SampleSet[w,h], FullRaytracer.cs:259-266). */
int rtc_set_params(rtc_ctx* ctx, const rtc_params* params);

/* ---- the hot path ---------------------------------------------------------------------------------- */
/* Scene.RayTrace(ray, skipHit) (Scene.cs:113-120) for n rays. skip may be NULL (== default(Hit) for all). */
int rtc_trace_closest(rtc_ctx* ctx, int64_t n, const rtc_ray* rays, const rtc_hit* skip, rtc_hit* out);
/* Raytracer.GetCameraRay (Raytracer.cs:262-282) for n (x, y, sample) triples; xy is n*2 int32. */
int rtc_camera_rays(rtc_ctx* ctx, int64_t n, const int32_t* xy, const uint32_t* sample, rtc_ray* out);
/* Vec4D.CreateHorizon(pole, z, theta) (Vec4D.cs:33-58: the lobe sample of RandomShine / the diffuse bounce, Raytracer.cs:51-61)
 * as the shading kernel of this context's arithmetic mode evaluates it: n tuples (pole.xyz, z, theta) in, n x xyz out. A
 * known-answer hook for the parity tests, like rtc_camera_rays. */
int rtc_debug_create_horizon(rtc_ctx* ctx, int64_t n, const double* pole_z_theta, double* out);
/* n_samples passes of Raytracer.Render over the pixel rectangle [x0,x1) x [y0,y1) (Raytracer.cs:302-327),
 * samples first_sample .. first_sample+n_samples-1 of every pixel, accumulated like FullRaytracer.cs:326-339.
 * Asynchronous on the context's stream; rtc_sync / rtc_read_accum / rtc_tonemap_argb wait for it. */
int rtc_render(rtc_ctx* ctx, int32_t x0, int32_t y0, int32_t x1, int32_t y1, uint32_t first_sample, uint32_t n_samples);
/* One progressive frame end to end: rtc_render over the whole image followed by rtc_read_accum, with the read-back of
 * each band of rows overlapped with the rendering of the next band (the drain of FullRaytracer.cs:326-344 running
 * beside the workers). Host buffers are best pinned; any of them may be NULL. Blocks until they are filled. */
int rtc_render_read(rtc_ctx* ctx, uint32_t first_sample, uint32_t n_samples, double* rgb_sum, uint32_t* samples, uint32_t* misses);
int rtc_sync(rtc_ctx* ctx);

/* ---- SampleSet[,] (SampleSet.cs:7-44): row-major y*width+x ------------------------------------------ */
int rtc_clear_accum(rtc_ctx* ctx);
int rtc_read_accum(rtc_ctx* ctx, double* rgb_sum, uint32_t* samples, uint32_t* misses);
int rtc_write_accum(rtc_ctx* ctx, const double* rgb_sum, const uint32_t* samples, const uint32_t* misses);
/* Device addresses of the accumulation planes (rgb_sum: w*h*3 f64, samples / misses: w*h u32), for the
 * per-frame collective. */
int rtc_accum_device_ptrs(rtc_ctx* ctx, void** rgb_sum, void** samples, void** misses);
/* FullRaytracer.GetBitmap / SampleSet.GetOutput (FullRaytracer.cs:179-205, SampleSet.cs:61-113): ARGB8 of the planes as of the
 * last accumulation issued. Converted on the device into a persistent buffer and copied out through pinned staging on a
 * stream of its own: the render stream is not synchronised, so the UI's 100 ms refresh (FullRaytracer.cs:33,369) does not
 * stall rendering. This call and rtc_read_pixel are the two that may be made from a second host thread while another thread
 * is inside rtc_render / rtc_render_read / rtc_sync on the same handle. */
int rtc_tonemap_argb(rtc_ctx* ctx, double exposure, const double back_rgb[3], double back_a, uint32_t* argb);
/* FullRaytracer.GetSampleSet(x, y) (FullRaytracer.cs:131-146): Color sum, Samples, Misses of one pixel (MainWindow.cs:360-370
 * calls it on every mouse move). Same threading rule as rtc_tonemap_argb. */
int rtc_read_pixel(rtc_ctx* ctx, int32_t x, int32_t y, double rgb_sum[3], uint32_t* samples, uint32_t* misses);

/* ---- inspection ------------------------------------------------------------------------------------ */
/* Raytracer.GetDebugTrace(x, y) (Raytracer.cs:254-260,289-292) for one (pixel, sample); *n <= recursion+1. */
int rtc_debug_trace(rtc_ctx* ctx, int32_t x, int32_t y, uint32_t sample, int32_t capacity, rtc_debug_ray* out, int32_t* n);
/* DebugRaycaster.RenderDebug's per-pixel query (DebugRaycaster.cs:170-265): one unjittered camera ray per pixel
 * (camera.GetRay(x, y).Offset(imagePlane), :236). out is width*height int32, row-major.
 *   RTC_OVERLAY_PRIMITIVES        Primitive.ID of Scene.RayTrace(ray, null), or -1            (DisplayMode.Primitives, :194-199)
 *   RTC_OVERLAY_BOUNDING_VOLUMES  BVH.GetIntersectionCount(ray) over the reference-shaped tree (DisplayMode.BoundingVolumes,
 *                                 :200-212, BVH.cs:352-363); needs a tree from rtc_upload_bvh / rtc_build_bvh
 * The colour mapping (ColorRotation, alpha from the box count) stays on the host. */
enum { RTC_OVERLAY_PRIMITIVES = 0, RTC_OVERLAY_BOUNDING_VOLUMES = 1 };
int rtc_debug_raycast(rtc_ctx* ctx, int32_t mode, int32_t* out);
/* DisplayMode.Selection over primitives (DebugRaycaster.SetDisplayOnly(Primitive / IObject), :140-192: one
 * PrimitiveIntersector per selected primitive, the nearest one wins): per pixel the Primitive.ID of the nearest of the n_sel
 * listed primitives along the same unjittered camera ray, or -1 -- the selection alone, whatever hides it in the full scene.
 * The selection is prepared as a scene of its own on the device, so a whole mesh costs milliseconds where the reference tests
 * every selected primitive for every pixel. Needs the host-side description (rtc_upload_scene on this context). */
int rtc_debug_raycast_selection(rtc_ctx* ctx, int32_t n_sel, const int32_t* prim_ids, int32_t* out);
/* Per-path radiance of one sample pass (== the DoubleColor[w,h] a tile hands to OnTileFinished,
 * Raytracer.cs:305-326), rgb = (-1,-1,-1) for misses. out is w*h*3 doubles over the full image. */
int rtc_render_samples(rtc_ctx* ctx, uint32_t sample, double* out_rgb);
int rtc_get_stats(rtc_ctx* ctx, rtc_stats* stats);
int rtc_reset_stats(rtc_ctx* ctx);

/* ---- multi-GPU: scene replicated, sample ranges per rank, one collective per frame ------------------ */
#define RTC_NCCL_ID_BYTES 128
int rtc_comm_unique_id(void* id128);
int rtc_comm_init(rtc_ctx* ctx, int32_t nranks, int32_t rank, const void* id128);
/* The per-frame collective (ncclReduce / ncclAllReduce over NVLink). A rank's planes hold what it has rendered and not yet
 * handed over; the call moves every rank's contribution into the job's running total:
 *   root >= 0  root's planes += all other ranks' planes, which are then reset to zero on the stream. The progressive loop
 *              `rtc_render; rtc_reduce_accum(0)` therefore accumulates on the root exactly like the single-GPU path (no
 *              clear between frames; rtc_read_accum / rtc_tonemap_argb are meaningful on the root only).
 *   root <  0  all-reduce: every rank ends with the running total. The library keeps a device copy of it and takes it out
 *              again on ranks != 0 before the next collective, so the total is never contributed twice.
 * rtc_clear_accum / rtc_write_accum / a size change start a fresh contribution on the calling rank. Collective: all ranks
 * of the communicator must call it with the same root. */
int rtc_reduce_accum(rtc_ctx* ctx, int32_t root);
/* Multi-GPU Start(): the device scene of `root` (tree, primitive records, materials -- what rtc_upload_bvh, rtc_build_bvh or
 * rtc_upload_baked left there) is copied to every other rank of the communicator with ncclBroadcast over NVLink, so the scene
 * crosses PCIe once per node instead of once per GPU and Scene.Prepare (Scene.cs:39-49) runs on one rank only. The receiving
 * contexts need no rtc_upload_scene / rtc_upload_bvh; like after rtc_upload_baked of a foreign image they hold no host-side
 * description (rtc_bake, rtc_get_bvh and the bounding-volume overlay report RTC_ERR_STATE). Camera and parameters are set
 * per rank as usual. Collective: every rank calls it with the same root, contexts of the same arithmetic mode. */
int rtc_bcast_scene(rtc_ctx* ctx, int32_t root);
int rtc_comm_destroy(rtc_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif
