// rtc_internal.h — device data layout and the launcher interface shared by the C-ABI host code (rtc_api.cu)
// and the two kernel translation units (kernels_f32.cu built with -fmad=true, kernels_f64.cu with -fmad=false).
//
// HBM layout (R = float in RTC_F32 mode, double in RTC_F64 mode; V4<R> is one 16-byte / 32-byte vector load):
//   DNode<R>  f64 mode: 4 children per node (256 B), stored as lo[axis][child] / hi[axis][child] rows + 4 child references,
//             so two 16-byte loads = one bound of all children. (f32 mode uses the quantised 8-wide CNode below.)
//   DPrim<R>  3 x V4 per primitive, stored in left-first DFS leaf order so slot == leaf order (tie-break key):
//               triangle  a=(v0.xyz,N.x) b=(e1.xyz,N.y) c=(e2.xyz,ref)      (Triangle.cs:22-29; the whole N lives in sgeom[])
//               sphere    a=(center.xyz,radius)            c=(0,0,0,ref)     (Sphere.cs:11-14)
//               plane     a=(normal.xyz,originDistance)    c=(0,0,0,ref)     (Plane.cs:13-14)
//             ref = the leaf reference (REF_LEAF | kind | flags | slot) as raw bits in the w lane, so that a leaf test is
//             ONE memory round trip (three 16-byte loads issued together) instead of reference -> record.
//   DXform<R> 9 x V4 for transformed spheres (rows 0-2 of MatrixToWorld, MatrixToObject, MatrixToNormal) or
//             vertex-normal triangles (n0,n1,n2 in rows 0-2).
//   DMat<R>   f64: 4 x V4: (emission, ior) (diffuse, shininess) (specular, 0) (refraction, 0)  (Primitive.cs:16-129)
//             f32: 32 B: the four colours as 12 halfs, then ior and shininess as floats (one 256-bit load)
//   sgeom     V4<R> per slot, what completing a hit record needs without fetching the 48-byte primitive: triangle (N.xyz,
//             flag bits: 1 = vertex normals), sphere (centre.xyz, radius), plane (normal.xyz, 0)
//   path pool: SoA per path: dir (xyz | f32: norm defect), tint, hpos (hit position | f64: Hit.Distance, f32: hit code),
//             hnrm (hit normal | hit code), thit (the trace kernel's output: distance, code), radiance.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/rtcore_b200.h"

namespace rtc {

template <typename R>
struct alignas(sizeof(R) * 4) V4 {
  R x, y, z, w;
};

// child reference encoding
constexpr uint32_t REF_LEAF = 0x80000000u;
constexpr uint32_t REF_EMPTY = 0x7FFFFFFFu;  // inner-node slot that is never hit (padding child of a 1-leaf tree)
constexpr int REF_KIND_SHIFT = 29;           // 2 bits: DK_*
constexpr uint32_t REF_MIRROR = 1u << 28;
constexpr uint32_t REF_TWOSIDED = 1u << 27;
constexpr uint32_t REF_INVERT = 1u << 26;
constexpr uint32_t REF_SLOT_MASK = (1u << 26) - 1;
enum { DK_TRI = 0, DK_SPHERE = 1, DK_XSPHERE = 2, DK_PLANE = 3 };
constexpr uint32_t REF_VNORMALS_AUX = 0x40000000u;  // flag kept in aux[] for vertex-normal triangles

// hit code stored in the w lane of the hit-normal vector
constexpr uint32_t HIT_MISS = 0xFFFFFFFFu;
constexpr uint32_t HIT_INSIDE = 1u << 30;
constexpr uint32_t HIT_SECOND = 1u << 29;  // the hit is the second entry of the primitive's Hit[] (a sphere's far hit)
constexpr uint32_t HIT_INVERT = 1u << 28;  // Primitive.Invert of the hit primitive: its own inside flag is INSIDE ^ INVERT
constexpr int HIT_KIND_SHIFT = 26;         // 2 bits: DK_* of the hit primitive (k_shade completes flat triangles and plain
                                           // spheres from sgeom[] alone)


constexpr int kTraceStack = 128;

// Branching factor of the device BVH. The reference's tree is binary (BVH.cs:239-254); for the device it is collapsed
// into W-wide nodes (children of large-area inner nodes are pulled up) so that a ray makes ~log_W instead of log_2
// dependent memory round trips. Leaves, their boxes and their left-first order are unchanged by the collapse.
template <typename R>
struct Width {
  static constexpr int value = 4;  // measured in f64 mode: 2-wide 321, 4-wide 408, 8-wide 355 Mrays/s on the 1 M-triangle scene
};

constexpr int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

template <typename R, int W>
struct alignas(next_pow2((6 * (int)sizeof(R) + 4) * W)) DNodeW {
  R lo[3][W];  // lo[axis][child]: one vector load fetches one bound of every child
  R hi[3][W];
  uint32_t child[W];  // inner-node index, leaf reference, or REF_EMPTY
};

template <typename R>
using DNode = DNodeW<R, Width<R>::value>;

// f32 production mode: 8-wide node with child boxes quantised to 8 bits per bound on a per-node power-of-two grid
// (after Ylitie, Karras & Laine 2017). 96 B, fetched with two 256-bit loads and one 128-bit load — the trace kernel
// is bound by L1 wavefronts (one per lane and load instruction for incoherent rays), not by bytes. Inner children are
// stored contiguously from child_base, the primitives of leaf children contiguously from prim_base, so a child is
// addressed by base + popcount of the mask bits below its slot. Children sit in the slot whose octant code best
// matches their offset from the node centre: for a ray with octant `oct`, slot s is visited in order of s ^ (7 ^ oct).
struct alignas(32) CNode {
  float px, py, pz;     // grid origin = node box minimum
  uint32_t e_imask;     // ex | ey << 8 | ez << 16 | imask << 24 : grid step 2^(e-127) per axis, inner-child slot mask
  uint32_t child_base;  // node index of the first inner child
  uint32_t prim_base;   // slot of the first leaf child's primitive
  uint32_t lmask;       // leaf-child slot mask
  uint32_t pad0;
  uint32_t q[12];       // qlo x,y,z then qhi x,y,z; 8 bytes (two words) each, byte s = child slot s
  uint32_t pad1[4];
};
static_assert(sizeof(CNode) == 96, "CNode layout");
// The 8-wide traversal stack lives in dynamic shared memory, [entry][thread] x 8 bytes, sized per scene to the tree's
// depth + 1 (one pending sibling group per level): a 1 M-triangle tree needs ~10 entries = 10 KB per 128-thread CTA,
// which leaves most of the SM's 228 KB to the L1 cache the node and primitive fetches run through.
constexpr int kQStackMax = 48;

template <typename R>
struct DPrim {
  V4<R> a, b, c;
};

template <typename R>
struct DXform {
  V4<R> r[9];
};

template <typename R>
struct DMat {
  V4<R> emis_ior, diff_shin, spec, refr;
};
// f32 mode: emission, diffuse, specular, refraction rgb as 12 halfs (w[0..5], low half first), w[6] = ior, w[7] = shininess
// as float bits. Colours are radiance scale factors and selection weights: half precision (2^-11 relative) is three orders
// below the 1 % image tolerance; ior and shininess (up to 1e6 and +inf) keep their float range.
template <>
struct alignas(32) DMat<float> {
  uint32_t w[8];
};

// What a trace launch writes per ray: Hit.Distance and the hit code (slot | kind | invert | second | inside, or HIT_MISS).
template <typename R>
struct THit;
template <>
struct alignas(8) THit<float> {
  float t;
  uint32_t code;
};
template <>
struct alignas(16) THit<double> {
  double t;
  uint32_t code, pad;
};

template <typename R>
struct SceneView {
  const DNode<R>* nodes;
  const DPrim<R>* prims;
  const DXform<R>* xforms;
  const DMat<R>* mats;
  const int32_t* aux;      // per slot: xform row (low 30 bits) | REF_VNORMALS_AUX, or -1
  const int32_t* prim_id;  // per slot: Primitive.ID
  const V4<R>* sgeom;      // per slot: triangle (N.xyz | flags), sphere (centre.xyz | radius), plane (normal.xyz | 0)
  uint32_t root;           // index of the root inner node
  int32_t n_prims;
  const CNode* qnodes;     // f32 mode only: the quantised 8-wide tree over the bounded primitives
  const uint32_t* unbounded;  // f32 mode only: leaf references of primitives with infinite boxes (planes)
  int32_t n_unbounded;
  int32_t q_stack;         // traversal stack entries per lane (f32: 8-wide tree depth + 1; f64: binary tree's need + 2)
};

template <typename R>
struct CameraView {
  int32_t kind;
  R position[3], look[3], side[3], up[3];
  R w2, h2, tan_fov_x2, tan_fov_y2, h_mult, v_mult, image_plane, dof_amount, focal_length;
};

template <typename R>
struct ParamsView {
  int32_t width, height, recursion, debug_geom;
  R ambient[3];
  R air_ior;
  uint32_t seed_lo, seed_hi;
};

// The wavefront band a launch works on: pixels [x0,x1) x [y0,y1), samples first_sample .. +n_samples-1.
// path id = s_local * n_pix + pix_local ; pix_local = (y-y0)*(x1-x0) + (x-x0)
struct Band {
  int32_t x0, y0, x1, y1;
  uint32_t first_sample, n_samples;
  uint32_t n_pix, n_paths;
};

struct Control {            // device-resident launch control block
  uint32_t count[2];        // live queue lengths (ping-pong)
  uint32_t work_trace;      // dynamic work cursor of the persistent trace kernel
  uint32_t pad;
  unsigned long long rays;  // closest-hit queries issued (sum of queue lengths over bounces)
  unsigned long long nodes_visited, prims_tested;  // RTC_OPT_COUNTERS
  unsigned long long node_steps, leaf_steps;       // RTC_OPT_COUNTERS: warp-level scheduler iterations (f32 kernel)
};

template <typename R>
struct PathView {
  V4<R>* dir;       // xyz = direction; w (f32 mode) = |direction|^2 - 1 where the host handed in a ray that is off unit length
  V4<R>* tint;      // rgb
  V4<R>* hpos;      // xyz = last hit position = next ray origin (camera ray origin at bounce 0); w = f64: Hit.Distance, f32: hit code
  V4<R>* hnrm;      // xyz = last hit normal, w = hit code bits
  THit<R>* thit;    // the trace kernel's answer for the current bounce
  V4<R>* radiance;  // rgb of finished paths ((-1,-1,-1) = miss)
  V4<R>* skip_pos;  // optional (rtc_trace_closest only): explicit skip-hit position; nullptr = ray origin
  uint32_t* queue[2];
  Control* ctl;
  int32_t* dbg_type;  // optional per-path BounceType (rtc_debug_trace)
  R* dbg_fresnel;     // optional per-path FresnelRatio
};

struct LaunchCfg {
  cudaStream_t stream;
  int sm_count;
  bool counters;
};

// Launchers, one set per arithmetic mode. All are asynchronous on cfg.stream.
template <typename R>
struct Kernels {
  static cudaError_t raygen(const LaunchCfg& cfg, const CameraView<R>& cam, const ParamsView<R>& par, const Band& band,
                            const PathView<R>& pv);
  // explicit (x,y,sample) list -> rtc_ray (f64) ; used by rtc_camera_rays
  static cudaError_t camera_rays(const LaunchCfg& cfg, const CameraView<R>& cam, const ParamsView<R>& par, int64_t n,
                                 const int32_t* xy, const uint32_t* sample, rtc_ray* out);
  // CreateHorizon for n (pole.xyz, z, theta) tuples (rtc_debug_create_horizon)
  static cudaError_t horizon(const LaunchCfg& cfg, int64_t n, const double* in, double* out);
  // bounce `bounce`: reads queue[q] (nullptr semantics: identity when bounce == 0) and the paths' hpos / dir, writes thit
  static cudaError_t trace(const LaunchCfg& cfg, const SceneView<R>& sc, const PathView<R>& pv, int q, bool identity_queue);
  static cudaError_t shade(const LaunchCfg& cfg, const SceneView<R>& sc, const ParamsView<R>& par, const Band& band,
                           const PathView<R>& pv, int q, int bounce, bool identity_queue);
  static cudaError_t compact(const LaunchCfg& cfg, const PathView<R>& pv, int q, bool identity_queue);
  static cudaError_t accumulate(const LaunchCfg& cfg, const ParamsView<R>& par, const Band& band, const PathView<R>& pv,
                                double* rgb_sum, uint32_t* samples, uint32_t* misses);
  // rtc_trace_closest plumbing
  static cudaError_t import_rays(const LaunchCfg& cfg, const SceneView<R>& sc, int64_t n, const rtc_ray* rays,
                                 const rtc_hit* skip, const int32_t* id_to_slot, const PathView<R>& pv);
  static cudaError_t export_hits(const LaunchCfg& cfg, const SceneView<R>& sc, int64_t n, const PathView<R>& pv,
                                 rtc_hit* out, bool finalize);
  static cudaError_t export_radiance(const LaunchCfg& cfg, const Band& band, const ParamsView<R>& par,
                                     const PathView<R>& pv, double* out_rgb);
  // DebugRaycaster overlay: unjittered camera rays of a band into the path pool / hit primitive ids out of it
  static cudaError_t overlay_rays(const LaunchCfg& cfg, const CameraView<R>& cam, const ParamsView<R>& par, const Band& band,
                                  const PathView<R>& pv);
  static cudaError_t overlay_prims(const LaunchCfg& cfg, const SceneView<R>& sc, const ParamsView<R>& par, const Band& band,
                                   const PathView<R>& pv, int32_t* out);
  static int trace_blocks_per_sm(size_t smem);
};

// BVH.GetIntersectionCount per pixel over the reference-shaped (binary, f64) tree; f64 arithmetic in both modes.
cudaError_t launch_overlay_boxcount(cudaStream_t s, const rtc_bvh_node* nodes, int32_t root, const CameraView<double>& cam,
                                    int32_t width, int32_t height, int32_t* out);

// BVH.Construct on the device (bvh_build.cu): parallel locally-ordered clustering over m bounded primitives. boxes: m x
// (lo[3], hi[3]) f64 on the host, prim_ids: their primitive IDs; nodes_out: 2m-1 nodes (host), leaves first.
cudaError_t build_bvh_ploc(cudaStream_t stream, int32_t m, const double* boxes, const int32_t* prim_ids, int radius,
                           rtc_bvh_node* nodes_out, int32_t* root_out, int32_t* rounds_out);

// Scene.Prepare on the device (prepare_device.cu).
// One device allocation per call, carved up by the two passes below (base / cap set by the caller, used = bytes handed out).
struct PrepareArena {
  char* base = nullptr;
  size_t cap = 0, used = 0;
};
// One primitive as the host stages it for the device flatten, in input (primitive ID) order: the geometry as floats, the
// material record already packed (pack_material), kind | flags << 8 and the transform row.
struct alignas(16) StagedPrim {
  float geom[12];
  uint32_t mat[8];
  uint32_t kind_flags;
  int32_t xform;
  uint32_t pad[2];
};
static_assert(sizeof(StagedPrim) == 96, "StagedPrim layout");
// The host builder's binned-SAH tree (host/bvh_builder.cpp), node for node, over m bounded primitives: d_boxes m x (lo[3],
// hi[3]) f64 leaf boxes, d_prim_ids their primitive IDs (ascending; null = identity), d_nodes room for 2m-1 nodes (root =
// node 0). h_pin: 8 x int32 of pinned host memory. Synchronises `stream`.
void prepare_device_preload();  // loads this unit's kernels (rtc_create, once per device)
size_t build_bvh_sah_scratch_bytes(int32_t m);
cudaError_t build_bvh_sah_device(cudaStream_t stream, PrepareArena* arena, int32_t* h_pin, int32_t m, const double* d_boxes,
                                 const int32_t* d_prim_ids, rtc_bvh_node* d_nodes, int32_t* levels_out);
// The f32-mode flattening of rtc_api.cu (build_device_scene) on the device, byte-identical output. All pointers are device
// pointers.
struct FlattenInput {
  const rtc_bvh_node* nodes;
  int32_t n_nodes, root, n_prims;
  const StagedPrim* staged;
  cudaEvent_t records_ready;  // optional: `staged` is complete once this event has passed (it is read last)
  // called once the number of wide nodes is known: a device buffer of `bytes` for them (the context's scene segment)
  int (*alloc_qnodes)(void* alloc_ctx, size_t bytes, void** out);
  void* alloc_ctx;
};
struct FlattenOutput {
  CNode* qnodes = nullptr;    // from alloc_qnodes; null if there is no bounded primitive
  int32_t n_qnodes = 0, depth = 0, n_bounded = 0;
  std::vector<int32_t> unbounded_prims;  // primitive IDs of the unbounded leaves in left-first order (slots n_bounded ...)
  // caller-allocated, n_prims entries each
  int32_t* slot_prim = nullptr;
  void *prims = nullptr, *mats = nullptr, *sgeom = nullptr;
  int32_t *aux = nullptr, *prim_id = nullptr, *id_to_slot = nullptr;
};
// out[i] = all[ids[i]] for 6-double boxes (the bounded primitives' boxes, compacted)
cudaError_t launch_gather_boxes(cudaStream_t stream, int32_t m, const int32_t* d_ids, const double* d_all, double* d_out);
size_t flatten_scratch_bytes(int32_t n_nodes, int32_t n_prims);
int flatten_device(cudaStream_t stream, PrepareArena* arena, int32_t* h_pin, const FlattenInput& in, FlattenOutput& out,
                   std::string& err);  // RTC_OK or an RTC_ERR_* code

// Queue re-ordering between bounces (reorder.cu, f32 mode): queue_out = queue_in[0 .. *count) sorted by the Morton cell of the
// paths' ray origins on the root node's grid (mode 1) or by cell and direction octant (mode 2); entries beyond *count go last.
size_t reorder_temp_bytes(uint32_t n);
cudaError_t launch_reorder(cudaStream_t stream, const CNode* root, const V4<float>* hpos, const V4<float>* dir, const uint32_t* queue_in,
                           const uint32_t* count, uint32_t n, int mode, uint32_t* keys_in, uint32_t* keys_out, uint32_t* queue_out, void* tmp,
                           size_t tmp_bytes);

// mode-independent
// planes -= base (rtc_reduce_accum after an all-reduce: only the samples rendered since then are contributed again)
cudaError_t launch_subtract_planes(cudaStream_t s, size_t n, double* rgb_sum, uint32_t* samples, uint32_t* misses,
                                   const double* base_rgb, const uint32_t* base_samples, const uint32_t* base_misses);
cudaError_t launch_tonemap(cudaStream_t s, int32_t n, const double* rgb_sum, const uint32_t* samples,
                           const uint32_t* misses, double exposure, double br, double bg, double bb, double ba,
                           uint32_t* argb);

}  // namespace rtc
