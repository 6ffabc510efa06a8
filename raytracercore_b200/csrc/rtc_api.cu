// rtc_api.cu — the C ABI of include/rtcore_b200.h: context, scene hand-over (host f64 arrays -> device SoA +
// flattened BVH), the wavefront loop that sequences the kernels, accumulation-buffer I/O, stats and the
// per-frame NCCL collective. Kernels live in kernels_f32.cu / kernels_f64.cu (rtc_device.cuh).
#include <cuda_fp16.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <limits>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../host/scene.h"
#include "prepare_common.h"
#include "rtc_internal.h"

using namespace rtc;

namespace {
thread_local std::string g_create_error;

struct NcclId {
  char internal[128];
};
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*Reduce)(const void*, void*, size_t, int, int, int, void*, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool load(std::string& err) {
    if (lib) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) {
      err = std::string("cannot load libnccl: ") + dlerror();
      return false;
    }
    GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
    CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
    Reduce = (decltype(Reduce))dlsym(lib, "ncclReduce");
    AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
    Broadcast = (decltype(Broadcast))dlsym(lib, "ncclBroadcast");
    GroupStart = (decltype(GroupStart))dlsym(lib, "ncclGroupStart");
    GroupEnd = (decltype(GroupEnd))dlsym(lib, "ncclGroupEnd");
    GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
    if (!GetUniqueId || !CommInitRank || !CommDestroy || !Reduce || !AllReduce || !Broadcast || !GroupStart || !GroupEnd) {
      err = "libnccl is missing required symbols";
      return false;
    }
    return true;
  }
};
NcclApi g_nccl;
constexpr int kNcclFloat64 = 8, kNcclUint32 = 3, kNcclUint8 = 1, kNcclSum = 0;

struct TimedLaunch {
  cudaEvent_t a, b;
  int kind;
};
}  // namespace

// Host-resident image of the device scene layout (one pinned buffer + segment table): what Scene.Prepare leaves
// behind on the host side, uploadable with plain H2D copies.
struct rtc_baked {
  enum { S_NODES, S_QNODES, S_UNBOUNDED, S_PRIMS, S_MATS, S_XFORMS, S_AUX, S_PRIM_ID, S_ID_TO_SLOT, S_SGEOM, S_COUNT };
  int precision = RTC_F32;
  int32_t n_prims = 0, n_unbounded = 0, n_xforms = 0, bvh_depth = 0;
  uint32_t root_node = 0;
  size_t off[S_COUNT] = {0}, bytes[S_COUNT] = {0};
  char* host = nullptr;
  size_t total = 0;
  bool pinned = false;
  // the asynchronous uploads that still read this image: one event per stream an upload was issued on (recorded behind the
  // last copy); the image is not released before they have passed
  mutable cudaEvent_t in_flight[2] = {nullptr, nullptr};
  void mark_in_flight(int which, cudaStream_t st) const {
    if (!in_flight[which] && cudaEventCreateWithFlags(&in_flight[which], cudaEventDisableTiming) != cudaSuccess) {
      in_flight[which] = nullptr;
      cudaStreamSynchronize(st);  // no event to wait on later: wait now
      return;
    }
    cudaEventRecord(in_flight[which], st);
  }
  ~rtc_baked() {
    for (cudaEvent_t& e : in_flight)
      if (e) {
        cudaEventSynchronize(e);
        cudaEventDestroy(e);
      }
    if (host) {
      if (pinned) cudaFreeHost(host);
      else std::free(host);
    }
  }
  bool alloc(size_t n) {
    total = n;
    if (cudaMallocHost((void**)&host, std::max<size_t>(n, 16)) == cudaSuccess) {
      pinned = true;
      return true;
    }
    cudaGetLastError();
    host = (char*)std::malloc(std::max<size_t>(n, 16));
    pinned = false;
    return host != nullptr;
  }
};

// allocator whose construct() leaves trivially-constructible elements uninitialised: resize() of the two large scene arrays
// must not write 200 MB of zeros that the copy behind it overwrites
template <typename T>
struct NoInitAlloc : std::allocator<T> {
  template <typename U>
  struct rebind {
    using other = NoInitAlloc<U>;
  };
  template <typename U>
  void construct(U* p) noexcept {
    ::new ((void*)p) U;
  }
  template <typename U, typename... A>
  void construct(U* p, A&&... a) {
    ::new ((void*)p) U(std::forward<A>(a)...);
  }
};

struct rtc_ctx {
  int device = 0;
  int precision = RTC_F32;
  std::string err;
  cudaStream_t stream = nullptr;      // the stream all work is issued on
  cudaStream_t own_stream = nullptr;  // created by rtc_create
  int sm_count = 148;

  // host copy of the scene as handed over (f64)
  int32_t n_prims = 0, n_xforms = 0;
  std::vector<uint8_t> kind, flags;
  std::vector<double, NoInitAlloc<double>> geom, material;
  std::vector<double> xforms;
  std::vector<int32_t> xform;
  std::vector<rtc_bvh_node> nodes;
  int32_t root = -1;
  bool scene_set = false, bvh_set = false, camera_set = false, params_set = false;
  rtc_camera cam{};
  rtc_params par{};

  // device scene
  void *d_nodes = nullptr, *d_prims = nullptr, *d_xforms = nullptr, *d_mats = nullptr;
  int32_t *d_aux = nullptr, *d_prim_id = nullptr, *d_id_to_slot = nullptr;
  void* d_sgeom = nullptr;
  void* d_qnodes = nullptr;       // f32 mode: CNode[]
  uint32_t* d_unbounded = nullptr;  // f32 mode: leaf refs of primitives with infinite boxes
  int32_t n_unbounded = 0;
  uint32_t root_node = 0;
  int bvh_depth = 0;
  size_t seg_cap[rtc_baked::S_COUNT] = {0};  // device capacity per scene segment (buffers are reused across uploads)
  size_t seg_bytes[rtc_baked::S_COUNT] = {0};  // bytes of each segment in use by the current device scene
  rtc_baked* baked = nullptr;               // image of the current device scene (null after rtc_prepare_device / rtc_bcast_scene:
                                            // rtc_bake then reads the segments back from the device)
  // rtc_prepare_device: pinned staging ring (host threads convert into it, the copy engine drains it) and a small pinned
  // read-back block for the level-synchronous passes
  char* h_ring[2] = {nullptr, nullptr};
  cudaEvent_t ev_ring[2] = {nullptr, nullptr};
  bool ring_busy[2] = {false, false};
  int32_t* h_pin = nullptr;
  rtc_bvh_node* d_bnodes = nullptr;         // rtc_prepare_device: the reference-shaped tree stays on the device; `nodes` is
  int32_t n_bnodes = 0;                     // filled from it when rtc_get_bvh / the box-count overlay ask

  // path pool
  int64_t max_paths = 1 << 25;
  int64_t pool_cap = 0;
  void *d_dir = nullptr, *d_tint = nullptr, *d_hpos = nullptr, *d_hnrm = nullptr, *d_thit = nullptr, *d_radiance = nullptr,
       *d_skip_pos = nullptr;
  uint32_t* d_queue[2] = {nullptr, nullptr};
  // RTC_OPT_REORDER: a third queue buffer (the sort's output; the three rotate), sort keys and the sort's temporary storage per wave
  int reorder = 0;
  uint32_t *d_rqueue = nullptr, *d_rkeys[2] = {nullptr, nullptr};
  void* d_rtmp[2] = {nullptr, nullptr};
  size_t rtmp_bytes = 0;
  Control* d_ctl = nullptr;  // [2]: one control block per wavefront in flight
  // Two wavefronts in flight (rtc_render): the bands of a frame alternate between the context's stream and wave_stream, each
  // with its own half of the path pool and its own control block, so that the drain phase of one band's persistent trace
  // launch (few rays left, most SMs idle) is filled by the other band's kernels. Bands are disjoint pixel rows: the
  // accumulation stays race-free and ordered per pixel. RTC_OPT_WAVES = 1 serialises (as does RTC_OPT_KERNEL_TIMING).
  int waves = 2;
  cudaStream_t wave_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_accum2 = nullptr;
  int32_t* d_dbg_type = nullptr;
  void* d_dbg_fresnel = nullptr;
  // rtc_trace_closest staging
  rtc_ray* d_rays = nullptr;
  rtc_hit *d_skip = nullptr, *d_hits = nullptr;
  int64_t stage_cap = 0;

  // accumulation planes (SampleSet[,])
  double* d_rgb = nullptr;
  uint32_t *d_samples = nullptr, *d_misses = nullptr;
  int acc_w = 0, acc_h = 0;

  // stats
  rtc_stats stats{};
  bool timing = false, counters = false;
  std::vector<TimedLaunch> pending;
  std::vector<cudaEvent_t> free_events;

  void* nccl_comm = nullptr;
  int nranks = 1, rank = 0;
  // after an all-reduce every rank holds the job's running total; `base` is a copy of it, subtracted again on ranks != 0
  // before the next collective so that only the samples rendered since then are contributed (rtc_reduce_accum)
  bool replicated = false;
  double* d_base_rgb = nullptr;
  uint32_t *d_base_samples = nullptr, *d_base_misses = nullptr;

  // copy engine choreography: the shading half of a pinned scene image (materials, ids) is copied on a second stream
  // behind the geometry half, so traversal starts while it is still in flight; rtc_render_read streams the finished
  // rows of the accumulation planes back while the next band renders.
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_geom = nullptr, ev_shading = nullptr, ev_band = nullptr;
  bool shading_pending = false;  // kernels that read materials / ids must first wait for ev_shading

  // UI read-out beside the render loop (rtc_tonemap_argb, rtc_read_pixel): these two calls may come
  // from a second host thread while the first is inside rtc_render / rtc_sync. They run on their own stream behind the event of
  // the last accumulate launch issued, into persistent device + pinned staging buffers; the render stream never waits for the
  // host, only (for the tens of microseconds of a tonemap kernel) the next accumulate launch waits for a read-out in flight.
  std::mutex ui_mutex;    // serialises the read-out calls (they share the staging buffers)
  std::mutex ev_mutex;    // guards ev_accum / ev_ui_done bookkeeping between the render thread and the read-out thread
  cudaStream_t ui_stream = nullptr;
  cudaEvent_t ev_accum = nullptr, ev_ui_done = nullptr;
  bool accum_recorded = false, accum2_recorded = false, ui_pending = false;
  uint32_t *d_argb = nullptr, *h_argb = nullptr;  // persistent ARGB8 image: device + pinned host staging
  size_t argb_cap = 0;
  double* h_pixel = nullptr;  // pinned: rgb[3] + samples + misses of one pixel
  // scratch for rtc_render_samples / rtc_debug_raycast (grown on demand, reused)
  void* d_scratch = nullptr;
  size_t scratch_cap = 0;
};


namespace {

int fail(rtc_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg;
  return code;
}
#define CU(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e__ = (call);                                                                            \
    if (e__ != cudaSuccess)                                                                              \
      return fail(ctx, RTC_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));               \
  } while (0)

int ensure_copy_stream(rtc_ctx* ctx) {
  if (ctx->copy_stream) return RTC_OK;
  CU(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  CU(cudaEventCreateWithFlags(&ctx->ev_geom, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&ctx->ev_shading, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&ctx->ev_band, cudaEventDisableTiming));
  return RTC_OK;
}

int ensure_ui_stream(rtc_ctx* ctx) {
  if (ctx->ui_stream) return RTC_OK;
  CU(cudaStreamCreateWithFlags(&ctx->ui_stream, cudaStreamNonBlocking));
  CU(cudaEventCreateWithFlags(&ctx->ev_accum, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&ctx->ev_ui_done, cudaEventDisableTiming));
  CU(cudaMallocHost((void**)&ctx->h_pixel, 8 * sizeof(double)));
  return RTC_OK;
}

int ensure_scratch(rtc_ctx* ctx, size_t bytes) {
  if (ctx->scratch_cap >= bytes && ctx->d_scratch) return RTC_OK;
  if (ctx->d_scratch) {
    CU(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->d_scratch);
    ctx->d_scratch = nullptr;
    ctx->scratch_cap = 0;
  }
  CU(cudaMalloc(&ctx->d_scratch, bytes));
  ctx->scratch_cap = bytes;
  return RTC_OK;
}

// Render-thread side of the read-out protocol: called right before a launch that writes the accumulation planes ...
int before_accum_write(rtc_ctx* ctx, int wave = 0) {
  std::lock_guard<std::mutex> g(ctx->ev_mutex);
  if (ctx->ui_pending) {  // a tonemap / pixel read is (or was) in flight on ui_stream: the planes must not change under it
    CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_ui_done, 0));
    if (ctx->wave_stream) CU(cudaStreamWaitEvent(ctx->wave_stream, ctx->ev_ui_done, 0));
    ctx->ui_pending = false;
  }
  (void)wave;
  return RTC_OK;
}
// ... and right after it
int after_accum_write(rtc_ctx* ctx, int wave = 0) {
  if (!ctx->ev_accum) return RTC_OK;  // no read-out has ever been asked for
  std::lock_guard<std::mutex> g(ctx->ev_mutex);
  if (wave == 0) {
    CU(cudaEventRecord(ctx->ev_accum, ctx->stream));
    ctx->accum_recorded = true;
  } else {
    CU(cudaEventRecord(ctx->ev_accum2, ctx->wave_stream));
    ctx->accum2_recorded = true;
  }
  return RTC_OK;
}
// Read-out side: order ui_stream behind every write to the planes issued so far
int ui_begin(rtc_ctx* ctx) {
  int rc = ensure_ui_stream(ctx);
  if (rc) return rc;
  std::lock_guard<std::mutex> g(ctx->ev_mutex);
  if (!ctx->accum_recorded) {  // first read-out: everything issued on the render stream so far
    CU(cudaEventRecord(ctx->ev_accum, ctx->stream));
    ctx->accum_recorded = true;
  }
  CU(cudaStreamWaitEvent(ctx->ui_stream, ctx->ev_accum, 0));
  if (ctx->accum2_recorded) CU(cudaStreamWaitEvent(ctx->ui_stream, ctx->ev_accum2, 0));  // the second wavefront's last accumulate
  return RTC_OK;
}
int ui_end(rtc_ctx* ctx) {
  {
    std::lock_guard<std::mutex> g(ctx->ev_mutex);
    CU(cudaEventRecord(ctx->ev_ui_done, ctx->ui_stream));
    ctx->ui_pending = true;
  }
  CU(cudaEventSynchronize(ctx->ev_ui_done));
  return RTC_OK;
}

// Called before the first launch that reads materials, primitive ids or face normals after an upload.
int wait_shading_upload(rtc_ctx* ctx) {
  if (!ctx->shading_pending) return RTC_OK;
  CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_shading, 0));
  ctx->shading_pending = false;
  return RTC_OK;
}

size_t rsize(const rtc_ctx* c) { return c->precision == RTC_F64 ? sizeof(double) : sizeof(float); }

void free_dev(void*& p) {
  if (p) cudaFree(p);
  p = nullptr;
}
template <typename T>
void free_dev_t(T*& p) {
  if (p) cudaFree(p);
  p = nullptr;
}

void free_scene_device(rtc_ctx* c) {
  free_dev(c->d_nodes);
  free_dev(c->d_prims);
  free_dev(c->d_xforms);
  free_dev(c->d_mats);
  free_dev_t(c->d_aux);
  free_dev_t(c->d_prim_id);
  free_dev_t(c->d_id_to_slot);
  free_dev(c->d_sgeom);
  free_dev(c->d_qnodes);
  free_dev_t(c->d_unbounded);
  free_dev_t(c->d_bnodes);
  c->n_bnodes = 0;
  c->n_unbounded = 0;
  for (size_t& v : c->seg_cap) v = 0;
}

void free_pool(rtc_ctx* c) {
  free_dev(c->d_dir);
  free_dev(c->d_tint);
  free_dev(c->d_hpos);
  free_dev(c->d_hnrm);
  free_dev(c->d_thit);
  for (int i = 0; i < 2; i++) free_dev_t(c->d_queue[i]);
  free_dev_t(c->d_rqueue);
  for (int i = 0; i < 2; i++) {
    free_dev_t(c->d_rkeys[i]);
    free_dev(c->d_rtmp[i]);
  }
  free_dev(c->d_radiance);
  free_dev(c->d_skip_pos);
  free_dev_t(c->d_dbg_type);
  free_dev(c->d_dbg_fresnel);
  c->pool_cap = 0;
}

// The pool holds what the largest wavefront so far needed (at most RTC_OPT_MAX_PATHS paths): it only grows, in 64 Ki steps.
int ensure_pool(rtc_ctx* ctx, int64_t want) {
  want = std::max<int64_t>(want, 1024);
  if (ctx->pool_cap >= want && ctx->d_dir) return RTC_OK;
  want = (want + 65535) & ~(int64_t)65535;
  CU(cudaStreamSynchronize(ctx->stream));
  free_pool(ctx);
  size_t v4 = rsize(ctx) * 4;
  CU(cudaMalloc(&ctx->d_dir, v4 * want));
  CU(cudaMalloc(&ctx->d_tint, v4 * want));
  CU(cudaMalloc(&ctx->d_hpos, v4 * want));
  CU(cudaMalloc(&ctx->d_hnrm, v4 * want));
  CU(cudaMalloc(&ctx->d_thit, (ctx->precision == RTC_F64 ? sizeof(THit<double>) : sizeof(THit<float>)) * want));
  for (int i = 0; i < 2; i++) CU(cudaMalloc((void**)&ctx->d_queue[i], sizeof(uint32_t) * want));
  CU(cudaMalloc(&ctx->d_radiance, v4 * want));
  if (!ctx->d_ctl) {
    CU(cudaMalloc((void**)&ctx->d_ctl, 2 * sizeof(Control)));
    CU(cudaMemset(ctx->d_ctl, 0, 2 * sizeof(Control)));
  }
  ctx->pool_cap = want;
  return RTC_OK;
}

// buffers of the queue re-ordering, sized like the pool
int ensure_reorder(rtc_ctx* ctx) {
  if (ctx->d_rqueue) return RTC_OK;
  const size_t n = (size_t)ctx->pool_cap;
  CU(cudaMalloc((void**)&ctx->d_rqueue, sizeof(uint32_t) * n));
  ctx->rtmp_bytes = reorder_temp_bytes((uint32_t)n);
  for (int i = 0; i < 2; i++) {
    CU(cudaMalloc((void**)&ctx->d_rkeys[i], sizeof(uint32_t) * n));
    CU(cudaMalloc(&ctx->d_rtmp[i], std::max<size_t>(ctx->rtmp_bytes, 16)));
  }
  return RTC_OK;
}

int ensure_wave_stream(rtc_ctx* ctx) {
  if (ctx->wave_stream) return RTC_OK;
  CU(cudaStreamCreateWithFlags(&ctx->wave_stream, cudaStreamNonBlocking));
  CU(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&ctx->ev_accum2, cudaEventDisableTiming));
  return RTC_OK;
}

// wave 1 works in the upper part of every pool array, from element `offset` on, with the second control block
template <typename R>
PathView<R> path_view(rtc_ctx* c, int wave = 0, int64_t offset = 0) {
  PathView<R> pv;
  pv.dir = (V4<R>*)c->d_dir + offset;
  pv.tint = (V4<R>*)c->d_tint + offset;
  pv.hpos = (V4<R>*)c->d_hpos + offset;
  pv.hnrm = (V4<R>*)c->d_hnrm + offset;
  pv.thit = (THit<R>*)c->d_thit + offset;
  for (int i = 0; i < 2; i++) pv.queue[i] = c->d_queue[i] + offset;
  pv.radiance = (V4<R>*)c->d_radiance + offset;
  pv.skip_pos = nullptr;
  pv.ctl = c->d_ctl + wave;
  pv.dbg_type = nullptr;
  pv.dbg_fresnel = nullptr;
  return pv;
}

template <typename R>
SceneView<R> scene_view(rtc_ctx* c) {
  SceneView<R> sv;
  sv.nodes = (const DNode<R>*)c->d_nodes;
  sv.prims = (const DPrim<R>*)c->d_prims;
  sv.xforms = (const DXform<R>*)c->d_xforms;
  sv.mats = (const DMat<R>*)c->d_mats;
  sv.aux = c->d_aux;
  sv.prim_id = c->d_prim_id;
  sv.sgeom = (const V4<R>*)c->d_sgeom;
  sv.root = c->root_node;
  sv.n_prims = c->n_prims;
  sv.qnodes = (const CNode*)c->d_qnodes;
  sv.unbounded = c->d_unbounded;
  sv.n_unbounded = c->n_unbounded;
  sv.q_stack = std::max(2, c->bvh_depth + (c->precision == RTC_F64 ? 2 : 1));  // f64: the binary tree's stack need + 2
  return sv;
}

template <typename R>
CameraView<R> camera_view(const rtc_camera& c) {
  CameraView<R> v;
  v.kind = c.kind;
  for (int i = 0; i < 3; i++) {
    v.position[i] = (R)c.position[i];
    v.look[i] = (R)c.look[i];
    v.side[i] = (R)c.side[i];
    v.up[i] = (R)c.up[i];
  }
  v.w2 = (R)c.w2;
  v.h2 = (R)c.h2;
  v.tan_fov_x2 = (R)c.tan_fov_x2;
  v.tan_fov_y2 = (R)c.tan_fov_y2;
  v.h_mult = (R)c.h_mult;
  v.v_mult = (R)c.v_mult;
  v.image_plane = (R)c.image_plane;
  v.dof_amount = (R)c.dof_amount;
  v.focal_length = (R)c.focal_length;
  return v;
}

template <typename R>
ParamsView<R> params_view(const rtc_params& p) {
  ParamsView<R> v;
  v.width = p.width;
  v.height = p.height;
  v.recursion = p.recursion;
  v.debug_geom = p.debug_geom;
  for (int i = 0; i < 3; i++) v.ambient[i] = (R)p.ambient[i];
  v.air_ior = (R)p.air_ior;
  v.seed_lo = (uint32_t)p.seed;
  v.seed_hi = (uint32_t)(p.seed >> 32);
  return v;
}

// raw bits into the w lane of a record (matches code_of / set_code on the device)
inline void set_ref_bits(float& w, uint32_t bits) { std::memcpy(&w, &bits, 4); }
inline void set_ref_bits(double& w, uint32_t bits) {
  const uint64_t b = bits;
  std::memcpy(&w, &b, 8);
}

// f64 -> R with outward rounding for box bounds
template <typename R>
R round_down(double x) {
  if constexpr (std::is_same<R, double>::value) return x;
  else {
    // round toward -inf, then one more ulp: the f32 slab test evaluates fma(bound, 1/d, -o/d), whose rounding is
    // equivalent to moving the bound by up to an ulp of the coordinate magnitude
    float f = (float)x;
    if ((double)f > x) f = std::nextafterf(f, -std::numeric_limits<float>::infinity());
    return std::nextafterf(f, -std::numeric_limits<float>::infinity());
  }
}
template <typename R>
R round_up(double x) {
  if constexpr (std::is_same<R, double>::value) return x;
  else {
    float f = (float)x;
    if ((double)f < x) f = std::nextafterf(f, std::numeric_limits<float>::infinity());
    return std::nextafterf(f, std::numeric_limits<float>::infinity());
  }
}

// H2D copy of a baked scene image into (reused) device buffers. A pinned image goes in two halves: what the traversal
// reads (tree, primitive records, transforms) on the context's stream, then what only shading and export read
// (materials, ids, face-normal z) on the copy stream, so that ray generation and the first trace launch overlap it.
int upload_baked_image(rtc_ctx* ctx, const rtc_baked* bk) {
  void** dst[rtc_baked::S_COUNT] = {&ctx->d_nodes, &ctx->d_qnodes, (void**)&ctx->d_unbounded, &ctx->d_prims, &ctx->d_mats,
                                    &ctx->d_xforms, (void**)&ctx->d_aux, (void**)&ctx->d_prim_id, (void**)&ctx->d_id_to_slot,
                                    &ctx->d_sgeom};
  const bool shading_seg[rtc_baked::S_COUNT] = {false, false, false, false, true, false, false, true, true, true};
  const bool split = bk->pinned;
  if (split) {
    int rc = ensure_copy_stream(ctx);
    if (rc) return rc;
  }
  if (ctx->shading_pending) {  // an earlier image's shading half is still owed to the main stream: order behind it
    int rc = wait_shading_upload(ctx);
    if (rc) return rc;
  }
  for (int i = 0; i < rtc_baked::S_COUNT; i++) {
    const size_t need = std::max<size_t>(bk->bytes[i], 16);
    if (ctx->seg_cap[i] < need || !*dst[i]) {
      if (*dst[i]) {
        CU(cudaStreamSynchronize(ctx->stream));
        if (ctx->copy_stream) CU(cudaStreamSynchronize(ctx->copy_stream));
        cudaFree(*dst[i]);
        *dst[i] = nullptr;
      }
      CU(cudaMalloc(dst[i], need));
      ctx->seg_cap[i] = need;
    }
  }
  for (int i = 0; i < rtc_baked::S_COUNT; i++)
    if (bk->bytes[i] && !(split && shading_seg[i]))
      CU(cudaMemcpyAsync(*dst[i], bk->host + bk->off[i], bk->bytes[i], cudaMemcpyHostToDevice, ctx->stream));
  if (split) {
    // the shading half starts when the geometry half is through (same PCIe link: side by side they would only slow
    // each other down) and, through the same event, after every earlier kernel that still reads the old materials
    CU(cudaEventRecord(ctx->ev_geom, ctx->stream));
    CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_geom, 0));
    for (int i = 0; i < rtc_baked::S_COUNT; i++)
      if (bk->bytes[i] && shading_seg[i])
        CU(cudaMemcpyAsync(*dst[i], bk->host + bk->off[i], bk->bytes[i], cudaMemcpyHostToDevice, ctx->copy_stream));
    CU(cudaEventRecord(ctx->ev_shading, ctx->copy_stream));
    ctx->shading_pending = true;
    bk->mark_in_flight(1, ctx->copy_stream);
  }
  if (bk->pinned) bk->mark_in_flight(0, ctx->stream);  // (a pageable image is synchronised below)
  if (bk->bytes[rtc_baked::S_QNODES] == 0 && bk->precision == RTC_F32) {
    // no bounded primitive: the kernel keys on a null qnodes pointer
    CU(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->d_qnodes);
    ctx->d_qnodes = nullptr;
    ctx->seg_cap[rtc_baked::S_QNODES] = 0;
  }
  ctx->n_unbounded = bk->n_unbounded;
  ctx->root_node = bk->root_node;
  ctx->bvh_depth = bk->bvh_depth;
  for (int i = 0; i < rtc_baked::S_COUNT; i++) ctx->seg_bytes[i] = bk->bytes[i];
  if (!bk->pinned) CU(cudaStreamSynchronize(ctx->stream));
  return RTC_OK;
}

// Flatten the reference-shaped tree + primitives into the device layout (see rtc_internal.h).
// Chunked parallel loop over [begin, end) on the host's threads (serial below `grain` items per thread).
template <typename F>
void parallel_for(size_t begin, size_t end, size_t grain, F&& fn) {
  const size_t n = end > begin ? end - begin : 0;
  unsigned hw = std::thread::hardware_concurrency();
  size_t nt = std::min<size_t>(hw ? std::min(hw, 16u) : 1, n / std::max<size_t>(grain, 1));
  if (nt <= 1) {
    for (size_t i = begin; i < end; i++) fn(i);
    return;
  }
  std::vector<std::thread> th;
  th.reserve(nt);
  const size_t per = (n + nt - 1) / nt;
  for (size_t t = 0; t < nt; t++) {
    const size_t b = begin + t * per, e = std::min(end, b + per);
    if (b >= e) break;
    th.emplace_back([b, e, &fn]() {
      for (size_t i = b; i < e; i++) fn(i);
    });
  }
  for (auto& t : th) t.join();
}

// Calls post(i) for every node of the binary tree below `root`, children before their parent. The subtrees hanging off the
// top levels run on the host threads (post(i) may only write what belongs to node i and read what belongs to descendants).
template <typename F>
void postorder_parallel(const std::vector<rtc_bvh_node>& nodes, int32_t root, F&& post) {
  std::vector<int32_t> top, frontier{root};
  while (frontier.size() < 128 && top.size() < 4096) {
    std::vector<int32_t> next;
    bool any = false;
    for (int32_t i : frontier) {
      if (nodes[i].prim < 0) {
        top.push_back(i);
        next.push_back(nodes[i].left);
        next.push_back(nodes[i].right);
        any = true;
      } else {
        post(i);  // a leaf this high up: done right away
      }
    }
    frontier.swap(next);
    if (!any) break;
  }
  parallel_for(0, frontier.size(), 1, [&](size_t f) {
    std::vector<std::pair<int32_t, int>> st;
    st.push_back({frontier[f], 0});
    while (!st.empty()) {
      auto& t = st.back();
      const int32_t i = t.first;
      const rtc_bvh_node& nd = nodes[i];
      if (nd.prim < 0 && t.second == 0) {
        t.second = 1;
        st.push_back({nd.right, 0});
        st.push_back({nd.left, 0});
      } else {
        st.pop_back();
        post(i);
      }
    }
  });
  for (size_t k = top.size(); k-- > 0;) post(top[k]);  // (breadth-first order reversed: children first)
}

void drop_device_tree(rtc_ctx* ctx) {
  if (!ctx->d_bnodes) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  free_dev_t(ctx->d_bnodes);
  ctx->n_bnodes = 0;
}

// the reference-shaped tree on the host, fetched from the device after rtc_prepare_device
int ensure_host_nodes(rtc_ctx* ctx) {
  if (!ctx->nodes.empty() || !ctx->d_bnodes) return RTC_OK;
  cudaSetDevice(ctx->device);
  ctx->nodes.resize((size_t)ctx->n_bnodes);
  CU(cudaMemcpyAsync(ctx->nodes.data(), ctx->d_bnodes, (size_t)ctx->n_bnodes * sizeof(rtc_bvh_node), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return RTC_OK;
}

// device buffer of scene segment `seg` with room for `bytes` (buffers are reused across scenes)
int ensure_seg(rtc_ctx* ctx, int seg, size_t bytes) {
  void** dst[rtc_baked::S_COUNT] = {&ctx->d_nodes, &ctx->d_qnodes, (void**)&ctx->d_unbounded, &ctx->d_prims, &ctx->d_mats,
                                    &ctx->d_xforms, (void**)&ctx->d_aux, (void**)&ctx->d_prim_id, (void**)&ctx->d_id_to_slot,
                                    &ctx->d_sgeom};
  const size_t need = std::max<size_t>(bytes, 16);
  if (ctx->seg_cap[seg] < need || !*dst[seg]) {
    if (*dst[seg]) {
      CU(cudaStreamSynchronize(ctx->stream));
      if (ctx->copy_stream) CU(cudaStreamSynchronize(ctx->copy_stream));
      cudaFree(*dst[seg]);
      *dst[seg] = nullptr;
    }
    CU(cudaMalloc(dst[seg], need));
    ctx->seg_cap[seg] = need;
  }
  return RTC_OK;
}

constexpr size_t kRingBytes = (size_t)8 << 20;

int ensure_stage_ring(rtc_ctx* ctx) {
  if (ctx->h_pin) return RTC_OK;
  for (int i = 0; i < 2; i++) {
    CU(cudaMallocHost((void**)&ctx->h_ring[i], kRingBytes));
    CU(cudaEventCreateWithFlags(&ctx->ev_ring[i], cudaEventDisableTiming));
  }
  CU(cudaMallocHost((void**)&ctx->h_pin, 64));
  return RTC_OK;
}

// n items of item_bytes each: fill(i, dst) writes item i; chunks of the pinned ring are filled on the host threads and copied
// to d_dst on `stream` while the next chunk is being filled. One caller at a time.
template <typename F>
int staged_upload(rtc_ctx* ctx, cudaStream_t stream, size_t n, size_t item_bytes, char* d_dst, F&& fill) {
  const size_t per = kRingBytes / item_bytes;
  size_t k = 0;
  for (size_t first = 0; first < n; first += per, k++) {
    const int slot = (int)(k & 1);
    if (ctx->ring_busy[slot]) CU(cudaEventSynchronize(ctx->ev_ring[slot]));
    const size_t cnt = std::min(per, n - first);
    char* base = ctx->h_ring[slot];
    parallel_for(0, cnt, 2048, [&](size_t i) { fill(first + i, base + i * item_bytes); });
    CU(cudaMemcpyAsync(d_dst + first * item_bytes, base, cnt * item_bytes, cudaMemcpyHostToDevice, stream));
    CU(cudaEventRecord(ctx->ev_ring[slot], stream));
    ctx->ring_busy[slot] = true;
  }
  return RTC_OK;
}

// RTC_B200_VERBOSE: wall-clock of the phases of build_device_scene
struct PhaseTimer {
  bool on = std::getenv("RTC_B200_VERBOSE") != nullptr;
  std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
  std::string log;
  void mark(const char* name) {
    if (!on) return;
    auto now = std::chrono::steady_clock::now();
    char buf[96];
    std::snprintf(buf, sizeof(buf), " %s %.0f ms,", name, std::chrono::duration<double, std::milli>(now - t).count());
    log += buf;
    t = now;
  }
  ~PhaseTimer() {
    if (on && !log.empty()) std::fprintf(stderr, "[rtcore_b200] flatten:%s\n", log.c_str());
  }
};

// Material record of one slot from the flattened description (14 doubles, rtcore_b200.h). f32 mode: 12 halfs + 2 floats.
inline void pack_material(DMat<double>& d, const double* m, bool reflective) {
  d.emis_ior = V4<double>{m[0], m[1], m[2], m[12]};
  d.diff_shin = V4<double>{m[3], m[4], m[5], m[13]};
  d.spec = reflective ? V4<double>{m[6], m[7], m[8], 0} : V4<double>{0, 0, 0, 0};
  d.refr = reflective ? V4<double>{m[9], m[10], m[11], 0} : V4<double>{0, 0, 0, 0};
}
inline uint32_t half_bits(double v) {
  if (std::isfinite(v)) v = std::min(std::max(v, -65504.0), 65504.0);  // largest finite half: colours beyond it saturate
  const __half h = __float2half_rn((float)v);
  uint16_t b;
  std::memcpy(&b, &h, 2);
  return b;
}
inline void pack_material(DMat<float>& d, const double* m, bool reflective) {
  double c[12];
  for (int i = 0; i < 12; i++) c[i] = (i >= 6 && !reflective) ? 0.0 : m[i];
  for (int i = 0; i < 6; i++) d.w[i] = half_bits(c[2 * i]) | (half_bits(c[2 * i + 1]) << 16);
  const float ior = (float)m[12], shin = (float)m[13];
  std::memcpy(&d.w[6], &ior, 4);
  std::memcpy(&d.w[7], &shin, 4);
}

template <typename R>
int build_device_scene(rtc_ctx* ctx) {
  PhaseTimer pt;
  const int32_t n = ctx->n_prims;
  const int32_t nn = (int32_t)ctx->nodes.size();
  const std::vector<rtc_bvh_node>& nodes = ctx->nodes;
  if (ctx->root < 0 || ctx->root >= nn) return fail(ctx, RTC_ERR_INVALID, "BVH root out of range");
  if (n > (int32_t)REF_SLOT_MASK) return fail(ctx, RTC_ERR_UNSUPPORTED, "more than 2^26-1 primitives");

  // Pass 1: validate the tree; collect the leaves in left-first order (BVH.cs:314-315).
  std::vector<int32_t> leaf_slot(nn, -1);
  std::vector<int32_t> dfs_leaves;  // binary-tree leaf nodes, left first
  dfs_leaves.reserve(n);
  std::vector<uint8_t> prim_seen(n, 0), node_seen(nn, 0);
  {
    std::vector<int32_t> st;
    st.push_back(ctx->root);
    while (!st.empty()) {
      int32_t i = st.back();
      st.pop_back();
      if (node_seen[i]) return fail(ctx, RTC_ERR_INVALID, "BVH is not a tree (node reached twice)");
      node_seen[i] = 1;
      const rtc_bvh_node& nd = nodes[i];
      if (nd.prim >= 0) {
        if (nd.prim >= n) return fail(ctx, RTC_ERR_INVALID, "BVH leaf references a primitive out of range");
        if (prim_seen[nd.prim]) return fail(ctx, RTC_ERR_INVALID, "primitive referenced by two BVH leaves");
        prim_seen[nd.prim] = 1;
        dfs_leaves.push_back(i);
      } else {
        if (nd.left < 0 || nd.left >= nn || nd.right < 0 || nd.right >= nn)
          return fail(ctx, RTC_ERR_INVALID, "BVH child index out of range");
        st.push_back(nd.right);
        st.push_back(nd.left);
      }
    }
  }
  if ((int32_t)dfs_leaves.size() != n) return fail(ctx, RTC_ERR_INVALID, "BVH does not reference every primitive exactly once");
  std::vector<int32_t> slot_prim(n, -1);

  auto leaf_ref = [&](int32_t node) -> uint32_t {
    int32_t p = nodes[node].prim;
    uint8_t k = ctx->kind[p], f = ctx->flags[p];
    uint32_t dk = k == RTC_KIND_TRIANGLE ? DK_TRI : k == RTC_KIND_PLANE ? DK_PLANE : ((f & RTC_FLAG_TRANSFORMED) && ctx->xform[p] >= 0) ? DK_XSPHERE : DK_SPHERE;
    uint32_t r = REF_LEAF | (dk << REF_KIND_SHIFT) | (uint32_t)leaf_slot[node];
    if (f & RTC_FLAG_MIRROR) r |= REF_MIRROR;
    if (f & RTC_FLAG_TWOSIDED) r |= REF_TWOSIDED;
    if (f & RTC_FLAG_INVERT) r |= REF_INVERT;
    return r;
  };
  auto area = [&](const double* lo, const double* hi) -> double {
    double dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    double a = (dx * dy + dy * dz + dz * dx) * 2;
    return std::isfinite(a) ? a : std::numeric_limits<double>::max();
  };

  std::vector<DNode<R>> dn;
  std::vector<CNode> qn;
  std::vector<uint32_t> unbounded;
  if constexpr (std::is_same<R, double>::value) {
    // ---- f64 mode: slot == left-first leaf order (the reference's tie-break key); W-wide uncompressed nodes ----
    for (int32_t s2 = 0; s2 < n; s2++) {
      leaf_slot[dfs_leaves[s2]] = s2;
      slot_prim[s2] = nodes[dfs_leaves[s2]].prim;
    }
    // Collapse the binary tree into W-wide nodes. A wide node starts from the two children of a binary node and
    // repeatedly replaces its largest-area inner child by that child's two children (in place, so the
    // left-to-right order is kept) until it has W children or only leaves.
    constexpr int W = Width<R>::value;
    dn.reserve((size_t)std::max(1, nn / W + 16));
    auto set_box = [&](DNode<R>& d, int c, const rtc_bvh_node& b) {
      for (int a = 0; a < 3; a++) {
        d.lo[a][c] = round_down<R>(b.bmin[a]);
        d.hi[a][c] = round_up<R>(b.bmax[a]);
      }
    };
    auto set_empty = [&](DNode<R>& d, int c) {
      for (int a = 0; a < 3; a++) {
        d.lo[a][c] = std::numeric_limits<R>::infinity();
        d.hi[a][c] = -std::numeric_limits<R>::infinity();
      }
      d.child[c] = REF_EMPTY;
    };
    struct Work {
      int32_t bnode, parent, pslot, stack_use;
    };
    std::vector<Work> work;
    int32_t max_stack = 0;
    if (nodes[ctx->root].prim >= 0) {
      dn.emplace_back();
      std::memset(&dn[0], 0, sizeof(DNode<R>));
      set_box(dn[0], 0, nodes[ctx->root]);
      dn[0].child[0] = leaf_ref(ctx->root);
      for (int c = 1; c < W; c++) set_empty(dn[0], c);
    } else {
      work.push_back({ctx->root, -1, 0, 0});
      while (!work.empty()) {
        Work wk = work.back();
        work.pop_back();
        int32_t kids[W];
        int nk = 2;
        kids[0] = nodes[wk.bnode].left;
        kids[1] = nodes[wk.bnode].right;
        while (nk < W) {
          int best = -1;
          double best_a = -1;
          for (int c = 0; c < nk; c++)
            if (nodes[kids[c]].prim < 0) {
              double a = area(nodes[kids[c]].bmin, nodes[kids[c]].bmax);
              if (a > best_a) {
                best_a = a;
                best = c;
              }
            }
          if (best < 0) break;
          int32_t e = kids[best];
          for (int c = nk; c > best + 1; c--) kids[c] = kids[c - 1];
          kids[best] = nodes[e].left;
          kids[best + 1] = nodes[e].right;
          nk++;
        }
        const int32_t me = (int32_t)dn.size();
        dn.emplace_back();
        std::memset(&dn[me], 0, sizeof(DNode<R>));
        if (wk.parent >= 0) dn[wk.parent].child[wk.pslot] = (uint32_t)me;
        const int32_t use = wk.stack_use + (nk - 1);
        max_stack = std::max(max_stack, use);
        for (int c = 0; c < W; c++) {
          if (c >= nk) {
            set_empty(dn[me], c);
            continue;
          }
          set_box(dn[me], c, nodes[kids[c]]);
          dn[me].child[c] = nodes[kids[c]].prim >= 0 ? leaf_ref(kids[c]) : REF_EMPTY;  // inner: patched on emission
        }
        for (int c = nk - 1; c >= 0; c--)  // left child emitted next: depth-first layout
          if (nodes[kids[c]].prim < 0) work.push_back({kids[c], me, c, use});
      }
    }
    ctx->root_node = 0;
    if (max_stack + 2 > kTraceStack)
      return fail(ctx, RTC_ERR_UNSUPPORTED, "BVH needs " + std::to_string(max_stack) + " traversal stack entries, limit is " + std::to_string(kTraceStack - 2));
    ctx->bvh_depth = max_stack;
  } else {
    // ---- f32 mode: quantised 8-wide tree over the bounded primitives; unbounded ones (planes) in a side list ----
    // fbox = union of the finite leaf boxes below a binary node, nf = their number (post-order over the tree)
    std::vector<double> fmin((size_t)nn * 3, std::numeric_limits<double>::infinity()), fmax((size_t)nn * 3, -std::numeric_limits<double>::infinity());
    std::vector<int32_t> nf(nn, 0);
    postorder_parallel(nodes, ctx->root, [&](int32_t i) {
      const rtc_bvh_node& nd = nodes[i];
      if (nd.prim >= 0) {
        bool fin = true;
        for (int a = 0; a < 3; a++) fin = fin && std::isfinite(nd.bmin[a]) && std::isfinite(nd.bmax[a]);
        if (fin) {
          nf[i] = 1;
          for (int a = 0; a < 3; a++) {
            fmin[(size_t)i * 3 + a] = nd.bmin[a];
            fmax[(size_t)i * 3 + a] = nd.bmax[a];
          }
        }
      } else {
        nf[i] = nf[nd.left] + nf[nd.right];
        for (int a = 0; a < 3; a++) {
          fmin[(size_t)i * 3 + a] = std::min(fmin[(size_t)nd.left * 3 + a], fmin[(size_t)nd.right * 3 + a]);
          fmax[(size_t)i * 3 + a] = std::max(fmax[(size_t)nd.left * 3 + a], fmax[(size_t)nd.right * 3 + a]);
        }
      }
    });
    // a binary node with bounded leaves on one side only is transparent
    auto resolve = [&](int32_t i) -> int32_t { return prep::resolve(nodes.data(), nf.data(), i); };
    pt.mark("validate + boxes");
    // Optimal collapse (prepare_common.h: dp_node), children first; cut[m][j] is the left side's share when j slots are split.
    std::vector<float> T((size_t)nn * 8, 0.0f);
    std::vector<uint8_t> cut((size_t)nn * 9, 0);
    postorder_parallel(nodes, ctx->root, [&](int32_t i) {
      const rtc_bvh_node& nd = nodes[i];
      if (nd.prim >= 0 || nf[i] == 0) return;
      if (nf[nd.left] == 0 || nf[nd.right] == 0) return;  // transparent: resolve() skips it
      const int32_t l = resolve(nd.left), r = resolve(nd.right);
      prep::dp_node(&T[(size_t)l * 8], &T[(size_t)r * 8], prep::box_area(&fmin[(size_t)i * 3], &fmax[(size_t)i * 3]), &T[(size_t)i * 8],
                    &cut[(size_t)i * 9]);
    });
    const prep::TreeView tv{nodes.data(), nf.data(), fmin.data(), fmax.data(), T.data(), cut.data()};
    pt.mark("collapse DP");
    const int32_t n_bounded = nf[ctx->root];
    int32_t next_slot = 0;
    int32_t max_depth = 0;
    if (n_bounded > 0) {
      struct QWork {
        int32_t bnode;   // resolved binary node (inner with two bounded sides, or the single bounded leaf for the root case)
        int32_t depth;
      };
      std::vector<QWork> queue;  // index in this vector == CNode index (breadth-first emission)
      queue.reserve((size_t)n_bounded / 3 + 16);
      queue.push_back({resolve(ctx->root), 1});
      qn.reserve((size_t)n_bounded / 3 + 16);
      // Level by level. Phase A (parallel over the level's nodes): children, box, grid, slot assignment, quantised bounds --
      // everything but the two base indices. Phase B (serial, cheap): child_base / prim_base in emission order, leaf slots,
      // and the next level's queue entries, children in slot order exactly as a serial breadth-first pass would push them.
      struct Emit {
        int32_t kids[8];
        int8_t child_in_slot[8];
        int nk;
        CNode cn;
      };
      std::vector<Emit> level;
      for (size_t lv_begin = 0; lv_begin < queue.size();) {
        const size_t lv_end = queue.size();
        level.resize(lv_end - lv_begin);
        parallel_for(lv_begin, lv_end, 2048, [&](size_t qi) {
          const QWork wk = queue[qi];
          Emit& em = level[qi - lv_begin];
          if (nodes[wk.bnode].prim >= 0) {
            em.kids[0] = wk.bnode;  // tree of a single bounded primitive
            em.nk = 1;
          } else {
            em.nk = prep::gather_children(tv, wk.bnode, em.kids);
          }
          prep::make_cnode(tv, em.kids, em.nk, em.cn, em.child_in_slot);  // everything but the two base indices
        });
        for (size_t qi = lv_begin; qi < lv_end; qi++) {
          const int32_t depth = queue[qi].depth;
          max_depth = std::max(max_depth, depth);
          Emit& em = level[qi - lv_begin];
          em.cn.child_base = (uint32_t)queue.size();
          em.cn.prim_base = (uint32_t)next_slot;
          for (int s2 = 0; s2 < 8; s2++) {
            const int c = em.child_in_slot[s2];
            if (c < 0) continue;
            const int32_t k = em.kids[c];
            if (nodes[k].prim >= 0) {
              leaf_slot[k] = next_slot;
              slot_prim[next_slot] = nodes[k].prim;
              next_slot++;
            } else {
              queue.push_back({k, depth + 1});
            }
          }
          qn.push_back(em.cn);
        }
        lv_begin = lv_end;
      }
    }
    if (std::getenv("RTC_B200_VERBOSE")) {
      size_t kids_total = 0, leaf_total = 0;
      for (const CNode& c : qn) {
        kids_total += __builtin_popcount(c.e_imask >> 24) + __builtin_popcount(c.lmask);
        leaf_total += __builtin_popcount(c.lmask);
      }
      std::fprintf(stderr, "[rtcore_b200] q8 tree: %zu nodes, %.2f children/node (%.2f leaves/node), depth %d, %d bounded prims\n",
                   qn.size(), qn.empty() ? 0.0 : (double)kids_total / qn.size(), qn.empty() ? 0.0 : (double)leaf_total / qn.size(), max_depth, n_bounded);
    }
    if (max_depth + 1 > kQStackMax)
      return fail(ctx, RTC_ERR_UNSUPPORTED, "8-wide BVH depth " + std::to_string(max_depth) + " exceeds the traversal stack (" + std::to_string(kQStackMax - 1) + ")");
    ctx->bvh_depth = max_depth;
    // unbounded primitives keep their left-first order, after the bounded ones
    for (int32_t li = 0; li < n; li++) {
      int32_t node = dfs_leaves[li];
      if (leaf_slot[node] >= 0) continue;
      leaf_slot[node] = next_slot;
      slot_prim[next_slot] = nodes[node].prim;
      next_slot++;
      unbounded.push_back(leaf_ref(node));
    }
    ctx->root_node = 0;
  }

  pt.mark("node emission");
  std::vector<DPrim<R>> dp(n);
  std::vector<DMat<R>> dm(n);
  std::vector<int32_t> aux(n, -1), prim_id(n), id_to_slot(n);
  std::vector<uint32_t> prim_ref(n);
  std::vector<V4<R>> sgeom(n);
  for (int32_t i = 0; i < nn; i++) {
    if (leaf_slot[i] < 0) continue;
    prim_ref[leaf_slot[i]] = leaf_ref(i);
  }
  (void)node_seen;
  parallel_for(0, (size_t)n, 65536, [&](size_t si) {  // (every slot writes its own records; id_to_slot[p] is a permutation)
    const int32_t s = (int32_t)si;
    int32_t p = slot_prim[s];
    prim_id[s] = p;
    id_to_slot[p] = s;
    const double* g = &ctx->geom[(size_t)p * RTC_GEOM_STRIDE];
    DPrim<R>& d = dp[s];
    std::memset(&d, 0, sizeof(d));
    uint8_t k = ctx->kind[p], f = ctx->flags[p];
    if (k == RTC_KIND_TRIANGLE) {
      d.a.x = (R)g[0]; d.a.y = (R)g[1]; d.a.z = (R)g[2]; d.a.w = (R)g[9];
      d.b.x = (R)g[3]; d.b.y = (R)g[4]; d.b.z = (R)g[5]; d.b.w = (R)g[10];
      d.c.x = (R)g[6]; d.c.y = (R)g[7]; d.c.z = (R)g[8];
      if ((f & RTC_FLAG_VNORMALS) && ctx->xform[p] >= 0) aux[s] = ctx->xform[p] | (int32_t)REF_VNORMALS_AUX;
      // shading geometry: the face normal and a flag word (1 = vertex normals: the record must be re-evaluated)
      sgeom[s] = V4<R>{(R)g[9], (R)g[10], (R)g[11], R(0)};
      set_ref_bits(sgeom[s].w, aux[s] >= 0 ? 1u : 0u);
    } else {
      d.a.x = (R)g[0]; d.a.y = (R)g[1]; d.a.z = (R)g[2]; d.a.w = (R)g[3];
      if (k == RTC_KIND_SPHERE && (f & RTC_FLAG_TRANSFORMED) && ctx->xform[p] >= 0) aux[s] = ctx->xform[p];
      sgeom[s] = V4<R>{(R)g[0], (R)g[1], (R)g[2], k == RTC_KIND_SPHERE ? (R)g[3] : R(0)};  // sphere: centre, radius; plane: normal
    }
    set_ref_bits(d.c.w, prim_ref[s]);  // the leaf reference rides in the record (one round trip per leaf test)
    const double* m = &ctx->material[(size_t)p * RTC_MATERIAL_STRIDE];
    bool reflective = m[13] > 0;  // Primitive.IsReflective (Primitive.cs:106): Specular/Refraction read black otherwise
    pack_material(dm[s], m, reflective);
  });
  std::vector<DXform<R>> dx(std::max(1, ctx->n_xforms));
  std::memset(dx.data(), 0, dx.size() * sizeof(DXform<R>));
  for (int32_t j = 0; j < ctx->n_xforms; j++) {
    const double* x = &ctx->xforms[(size_t)j * RTC_XFORM_STRIDE];
    for (int mtx = 0; mtx < 3; mtx++)
      for (int row = 0; row < 3; row++) {
        const double* r = x + mtx * 16 + row * 4;
        dx[j].r[mtx * 3 + row] = V4<R>{(R)r[0], (R)r[1], (R)r[2], (R)r[3]};
      }
  }
  // vertex-normal triangles keep n0,n1,n2 in the first 9 doubles of their row
  for (int32_t s = 0; s < n; s++) {
    if (aux[s] >= 0 && (aux[s] & (int32_t)REF_VNORMALS_AUX)) {
      int32_t j = aux[s] & 0x3FFFFFFF;
      const double* x = &ctx->xforms[(size_t)j * RTC_XFORM_STRIDE];
      for (int k2 = 0; k2 < 3; k2++) dx[j].r[k2] = V4<R>{(R)x[k2 * 3], (R)x[k2 * 3 + 1], (R)x[k2 * 3 + 2], R(0)};
    }
  }

  pt.mark("primitive + material records");
  rtc_baked* bk = new rtc_baked();
  bk->precision = ctx->precision;
  bk->n_prims = n;
  bk->n_unbounded = (int32_t)unbounded.size();
  bk->n_xforms = ctx->n_xforms;
  bk->bvh_depth = ctx->bvh_depth;
  bk->root_node = ctx->root_node;
  const void* src[rtc_baked::S_COUNT] = {dn.data(), qn.data(), unbounded.data(), dp.data(), dm.data(), dx.data(),
                                         aux.data(), prim_id.data(), id_to_slot.data(), sgeom.data()};
  const size_t sz[rtc_baked::S_COUNT] = {dn.size() * sizeof(DNode<R>), qn.size() * sizeof(CNode), unbounded.size() * sizeof(uint32_t),
                                         dp.size() * sizeof(DPrim<R>), dm.size() * sizeof(DMat<R>), dx.size() * sizeof(DXform<R>),
                                         aux.size() * sizeof(int32_t), prim_id.size() * sizeof(int32_t),
                                         id_to_slot.size() * sizeof(int32_t), sgeom.size() * sizeof(V4<R>)};
  size_t total = 0;
  for (int i = 0; i < rtc_baked::S_COUNT; i++) {
    bk->off[i] = total;
    bk->bytes[i] = sz[i];
    total += (sz[i] + 255) & ~(size_t)255;
  }
  if (!bk->alloc(total)) {
    delete bk;
    return fail(ctx, RTC_ERR_NOMEM, "host allocation for the baked scene failed");
  }
  {  // copy into the pinned image in 4 MB pieces on the host threads
    struct Piece {
      char* dst;
      const char* src;
      size_t len;
    };
    std::vector<Piece> pieces;
    const size_t kPiece = (size_t)4 << 20;
    for (int i = 0; i < rtc_baked::S_COUNT; i++)
      for (size_t o = 0; o < sz[i]; o += kPiece)
        pieces.push_back({(char*)bk->host + bk->off[i] + o, (const char*)src[i] + o, std::min(kPiece, sz[i] - o)});
    parallel_for(0, pieces.size(), 4, [&](size_t k) { std::memcpy(pieces[k].dst, pieces[k].src, pieces[k].len); });
  }
  delete ctx->baked;
  ctx->baked = bk;
  pt.mark("pinned image");
  const int rc_up = upload_baked_image(ctx, bk);
  if (pt.on) cudaStreamSynchronize(ctx->stream);
  pt.mark("upload");
  return rc_up;
}

int ready(rtc_ctx* ctx, bool need_camera) {
  if (!ctx->scene_set) return fail(ctx, RTC_ERR_STATE, "no scene uploaded (rtc_upload_scene)");
  if (!ctx->bvh_set) return fail(ctx, RTC_ERR_STATE, "no BVH (rtc_upload_bvh or rtc_build_bvh)");
  if (need_camera && !ctx->camera_set) return fail(ctx, RTC_ERR_STATE, "no camera (rtc_set_camera)");
  if (need_camera && !ctx->params_set) return fail(ctx, RTC_ERR_STATE, "no render parameters (rtc_set_params)");
  return RTC_OK;
}

cudaEvent_t get_event(rtc_ctx* c) {
  if (!c->free_events.empty()) {
    cudaEvent_t e = c->free_events.back();
    c->free_events.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

void drain_timing(rtc_ctx* c) {
  for (TimedLaunch& t : c->pending) {
    float ms = 0;
    if (cudaEventSynchronize(t.b) == cudaSuccess && cudaEventElapsedTime(&ms, t.a, t.b) == cudaSuccess) c->stats.ms[t.kind] += ms;
    c->free_events.push_back(t.a);
    c->free_events.push_back(t.b);
  }
  c->pending.clear();
}

struct Timed {
  rtc_ctx* c;
  int kind;
  cudaStream_t st;
  cudaEvent_t a = nullptr, b = nullptr;
  Timed(rtc_ctx* ctx, int k, cudaStream_t stream = nullptr) : c(ctx), kind(k), st(stream ? stream : ctx->stream) {
    c->stats.launches[k]++;
    if (c->timing) {
      a = get_event(c);
      b = get_event(c);
      cudaEventRecord(a, st);
    }
  }
  ~Timed() {
    if (c->timing) {
      cudaEventRecord(b, st);
      c->pending.push_back({a, b, kind});
      if (c->pending.size() > 8192) drain_timing(c);
    }
  }
};

// One wavefront over `band`: raygen, (trace, shade + compaction, bookkeeping) x (recursion+1), then accumulate or export.
template <typename R>
int run_band(rtc_ctx* ctx, const Band& band, bool accumulate, double* d_out_rgb, bool debug, int wave = 0, int64_t pool_offset = 0) {
  cudaStream_t stream = wave ? ctx->wave_stream : ctx->stream;
  LaunchCfg cfg{stream, ctx->sm_count, ctx->counters};
  SceneView<R> sv = scene_view<R>(ctx);
  PathView<R> pv = path_view<R>(ctx, wave, pool_offset);
  if (debug) {
    pv.dbg_type = ctx->d_dbg_type;
    pv.dbg_fresnel = (R*)ctx->d_dbg_fresnel;
  }
  CameraView<R> cv = camera_view<R>(ctx->cam);
  ParamsView<R> par = params_view<R>(ctx->par);
  uint32_t* spare = nullptr;  // RTC_OPT_REORDER: the third queue buffer of this wave's part of the pool
  if (ctx->reorder > 0 && ctx->precision == RTC_F32) {
    int rcr = ensure_reorder(ctx);
    if (rcr) return rcr;
    spare = ctx->d_rqueue + pool_offset;
  }
  {
    Timed t(ctx, RTC_K_RAYGEN, stream);
    CU(Kernels<R>::raygen(cfg, cv, par, band, pv));
  }
  ctx->stats.paths += band.n_paths;
  const int bounces = std::max(0, ctx->par.recursion) + 1;
  for (int i = 0; i < bounces; i++) {
    const int q = i & 1;
    const bool ident = (i == 0);
    {
      Timed t(ctx, RTC_K_TRACE, stream);
      CU(Kernels<R>::trace(cfg, sv, pv, q, ident));
    }
    if (i == 0 && wave == 0) {
      int rcw = wait_shading_upload(ctx);  // materials may still be arriving behind the first trace launch
      if (rcw) return rcw;
    }
    {
      Timed t(ctx, RTC_K_SHADE, stream);
      CU(Kernels<R>::shade(cfg, sv, par, band, pv, q, i, ident));
    }
    {
      Timed t(ctx, RTC_K_COMPACT, stream);
      CU(Kernels<R>::compact(cfg, pv, q, ident));
    }
    if constexpr (std::is_same<R, float>::value) {
      if (ctx->reorder > 0 && i + 1 < bounces && ctx->d_qnodes && !debug) {
        // the survivors (queue[q ^ 1]) in Morton order of their new origins into the spare buffer, which then takes the
        // place of queue[q ^ 1]; the unsorted one becomes the spare
        Timed t(ctx, RTC_K_COMPACT, stream);
        CU(launch_reorder(stream, (const CNode*)ctx->d_qnodes, pv.hpos, pv.dir, pv.queue[q ^ 1], &pv.ctl->count[q ^ 1], band.n_paths, ctx->reorder,
                          ctx->d_rkeys[0] + pool_offset, ctx->d_rkeys[1] + pool_offset, spare, ctx->d_rtmp[wave ? 1 : 0], ctx->rtmp_bytes));
        std::swap(pv.queue[q ^ 1], spare);
      }
    }
  }
  if (accumulate) {
    int rcu = before_accum_write(ctx, wave);
    if (rcu) return rcu;
    {
      Timed t(ctx, RTC_K_ACCUMULATE, stream);
      CU(Kernels<R>::accumulate(cfg, par, band, pv, ctx->d_rgb, ctx->d_samples, ctx->d_misses));
    }
    rcu = after_accum_write(ctx, wave);
    if (rcu) return rcu;
  } else if (d_out_rgb) {
    CU(Kernels<R>::export_radiance(cfg, band, par, pv, d_out_rgb));
  }
  return RTC_OK;
}

// Host destinations of rtc_render_read: the rows of a band are copied back on the copy stream as soon as the band's
// last accumulate launch is through, while the next band renders.
struct ReadBack {
  double* rgb;
  uint32_t *samples, *misses;
};

template <typename R>
int render_rect(rtc_ctx* ctx, int x0, int y0, int x1, int y1, uint32_t first_sample, uint32_t n_samples, bool accumulate,
                double* d_out_rgb, const ReadBack* rb = nullptr) {
  const int64_t rw = x1 - x0, rh = y1 - y0;
  const int64_t cap = std::max<int64_t>(ctx->max_paths, rw);
  const int64_t total = rw * rh * (int64_t)n_samples;
  // two wavefronts in flight when the frame gives each of them whole 4-row groups and at least 4 Mi paths (measured on the
  // B200: +2 % on the triangle soups, +1..10 % on the shading-heavy scenes at 16..32 Mi paths per frame, -2..3 % at 2 Mi
  // paths per wavefront, where halving the launches costs more than their overlap returns); per-kernel event timing keeps
  // the launches of one stream back to back, so it uses one wavefront
  const bool dual = ctx->waves > 1 && accumulate && !ctx->timing && rh >= 8 && std::min(total, cap) >= (int64_t)1 << 23 && cap / 2 >= rw * 4;
  int rc = ensure_pool(ctx, std::min<int64_t>(cap, total));
  if (rc) return rc;
  int64_t cap_w = cap, half = 0;
  if (dual) {
    rc = ensure_wave_stream(ctx);
    if (rc) return rc;
    half = ctx->pool_cap / 2;
    cap_w = half;
    CU(cudaEventRecord(ctx->ev_fork, ctx->stream));  // the second stream starts behind everything issued so far
    CU(cudaStreamWaitEvent(ctx->wave_stream, ctx->ev_fork, 0));
    if (ctx->shading_pending) CU(cudaStreamWaitEvent(ctx->wave_stream, ctx->ev_shading, 0));
  }
  int64_t rows_per_band = std::max<int64_t>(1, std::min<int64_t>(rh, cap_w / rw));
  if (dual && rows_per_band * 2 > rh) rows_per_band = (rh + 1) / 2;  // at least two bands
  if (rows_per_band >= 4) rows_per_band = dual ? ((rows_per_band + 3) & ~(int64_t)3) : (rows_per_band & ~(int64_t)3);  // whole 8 x 4 pixel tiles per band (band_pix_xy)
  if (rows_per_band * rw > cap_w) rows_per_band = std::max<int64_t>(4, (cap_w / rw) & ~(int64_t)3);
  int band_index = 0;
  for (int ya = y0; ya < y1; ya += (int)rows_per_band, band_index++) {
    int yb = (int)std::min<int64_t>(y1, ya + rows_per_band);
    int64_t npix = rw * (yb - ya);
    uint32_t s_chunk = (uint32_t)std::max<int64_t>(1, cap_w / npix);
    const int wave = dual ? (band_index & 1) : 0;
    cudaStream_t stream = wave ? ctx->wave_stream : ctx->stream;
    for (uint32_t s = 0; s < n_samples; s += s_chunk) {
      Band b;
      b.x0 = x0; b.x1 = x1; b.y0 = ya; b.y1 = yb;
      b.first_sample = first_sample + s;
      b.n_samples = std::min(s_chunk, n_samples - s);
      b.n_pix = (uint32_t)npix;
      b.n_paths = (uint32_t)(npix * b.n_samples);
      rc = run_band<R>(ctx, b, accumulate, d_out_rgb, false, wave, wave ? half : 0);
      if (rc) break;
    }
    if (rc) break;
    if (rb) {  // full-width bands: rows [ya, yb) are contiguous in the row-major planes
      const size_t off = (size_t)ya * ctx->acc_w, cnt = (size_t)(yb - ya) * ctx->acc_w;
      CU(cudaEventRecord(ctx->ev_band, stream));
      CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_band, 0));
      if (rb->rgb) CU(cudaMemcpyAsync(rb->rgb + off * 3, ctx->d_rgb + off * 3, cnt * 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->copy_stream));
      if (rb->samples) CU(cudaMemcpyAsync(rb->samples + off, ctx->d_samples + off, cnt * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->copy_stream));
      if (rb->misses) CU(cudaMemcpyAsync(rb->misses + off, ctx->d_misses + off, cnt * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->copy_stream));
    }
  }
  if (dual) {  // the context's stream continues behind both wavefronts (also on an error path: nothing is left dangling)
    cudaError_t e1 = cudaEventRecord(ctx->ev_join, ctx->wave_stream);
    cudaError_t e2 = cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0);
    if (!rc) {
      CU(e1);
      CU(e2);
    }
  }
  return rc;
}

int ensure_accum(rtc_ctx* ctx) {
  int w = ctx->par.width, h = ctx->par.height;
  if (ctx->d_rgb && ctx->acc_w == w && ctx->acc_h == h) return RTC_OK;
  if (ctx->d_rgb) CU(cudaStreamSynchronize(ctx->stream));
  free_dev_t(ctx->d_rgb);
  free_dev_t(ctx->d_samples);
  free_dev_t(ctx->d_misses);
  free_dev_t(ctx->d_base_rgb);
  free_dev_t(ctx->d_base_samples);
  free_dev_t(ctx->d_base_misses);
  ctx->replicated = false;
  size_t n = (size_t)w * h;
  CU(cudaMalloc((void**)&ctx->d_rgb, n * 3 * sizeof(double)));
  CU(cudaMalloc((void**)&ctx->d_samples, n * sizeof(uint32_t)));
  CU(cudaMalloc((void**)&ctx->d_misses, n * sizeof(uint32_t)));
  CU(cudaMemsetAsync(ctx->d_rgb, 0, n * 3 * sizeof(double), ctx->stream));
  CU(cudaMemsetAsync(ctx->d_samples, 0, n * sizeof(uint32_t), ctx->stream));
  CU(cudaMemsetAsync(ctx->d_misses, 0, n * sizeof(uint32_t), ctx->stream));
  ctx->acc_w = w;
  ctx->acc_h = h;
  return RTC_OK;
}

template <typename R>
int trace_batch(rtc_ctx* ctx, int64_t n, const rtc_ray* rays, const rtc_hit* skip, rtc_hit* out) {
  const int64_t cap = ctx->max_paths;
  int rc = ensure_pool(ctx, std::min(cap, n));
  if (rc) return rc;
  if (!ctx->d_skip_pos) CU(cudaMalloc(&ctx->d_skip_pos, rsize(ctx) * 4 * ctx->pool_cap));
  if (ctx->stage_cap < std::min(cap, n)) {
    free_dev_t(ctx->d_rays);
    free_dev_t(ctx->d_skip);
    free_dev_t(ctx->d_hits);
    ctx->stage_cap = std::min(cap, n);
    CU(cudaMalloc((void**)&ctx->d_rays, sizeof(rtc_ray) * ctx->stage_cap));
    CU(cudaMalloc((void**)&ctx->d_skip, sizeof(rtc_hit) * ctx->stage_cap));
    CU(cudaMalloc((void**)&ctx->d_hits, sizeof(rtc_hit) * ctx->stage_cap));
  }
  LaunchCfg cfg{ctx->stream, ctx->sm_count, ctx->counters};
  SceneView<R> sv = scene_view<R>(ctx);
  PathView<R> pv = path_view<R>(ctx);
  pv.skip_pos = (V4<R>*)ctx->d_skip_pos;
  for (int64_t off = 0; off < n; off += ctx->stage_cap) {
    int64_t m = std::min(ctx->stage_cap, n - off);
    CU(cudaMemcpyAsync(ctx->d_rays, rays + off, sizeof(rtc_ray) * m, cudaMemcpyHostToDevice, ctx->stream));
    if (skip) CU(cudaMemcpyAsync(ctx->d_skip, skip + off, sizeof(rtc_hit) * m, cudaMemcpyHostToDevice, ctx->stream));
    CU(Kernels<R>::import_rays(cfg, sv, m, ctx->d_rays, skip ? ctx->d_skip : nullptr, ctx->d_id_to_slot, pv));
    {
      Timed t(ctx, RTC_K_TRACE);
      CU(Kernels<R>::trace(cfg, sv, pv, 0, true));
    }
    ctx->stats.rays += (uint64_t)m;
    CU(Kernels<R>::export_hits(cfg, sv, m, pv, ctx->d_hits, true));
    CU(cudaMemcpyAsync(out + off, ctx->d_hits, sizeof(rtc_hit) * m, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
  }
  return RTC_OK;
}

int fetch_device_counters(rtc_ctx* ctx, Control* host) {
  if (!ctx->d_ctl) {
    std::memset(host, 0, sizeof(*host));
    return RTC_OK;
  }
  Control two[2];
  CU(cudaMemcpyAsync(two, ctx->d_ctl, 2 * sizeof(Control), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  *host = two[0];
  host->rays += two[1].rays;
  host->nodes_visited += two[1].nodes_visited;
  host->prims_tested += two[1].prims_tested;
  host->node_steps += two[1].node_steps;
  host->leaf_steps += two[1].leaf_steps;
  return RTC_OK;
}

}  // namespace

// =========================================================================================================
extern "C" {

int rtc_abi_version(void) { return RTC_ABI_VERSION; }

int rtc_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int rtc_create(int device, int precision, rtc_ctx** out) {
  if (!out) return RTC_ERR_INVALID;
  *out = nullptr;
  if (precision != RTC_F32 && precision != RTC_F64) {
    g_create_error = "precision must be RTC_F32 or RTC_F64";
    return RTC_ERR_INVALID;
  }
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    g_create_error = std::string("no CUDA device available (") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                     "); librtcore_b200 has no CPU fallback";
    cudaGetLastError();
    return RTC_ERR_CUDA;
  }
  if (device < 0 || device >= n) {
    g_create_error = "device index out of range";
    return RTC_ERR_INVALID;
  }
  if ((e = cudaSetDevice(device)) != cudaSuccess) {
    g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
    return RTC_ERR_CUDA;
  }
  rtc_ctx* ctx = new rtc_ctx();
  ctx->device = device;
  ctx->precision = precision;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
  if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
    g_create_error = std::string("cudaStreamCreate: ") + cudaGetErrorString(e);
    delete ctx;
    return RTC_ERR_CUDA;
  }
  ctx->own_stream = ctx->stream;
  {  // once per device: load the kernels of Scene.Prepare (lazy module loading would otherwise bill the first prepare for it)
    static std::mutex mu;
    static std::vector<int> done;
    std::lock_guard<std::mutex> g(mu);
    if (std::find(done.begin(), done.end(), device) == done.end()) {
      prepare_device_preload();
      done.push_back(device);
    }
  }
  *out = ctx;
  return RTC_OK;
}

int rtc_set_stream(rtc_ctx* ctx, void* cuda_stream) {
  if (!ctx) return RTC_ERR_INVALID;
  cudaSetDevice(ctx->device);
  CU(cudaStreamSynchronize(ctx->stream));
  drain_timing(ctx);
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return RTC_OK;
}

void rtc_destroy(rtc_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->copy_stream) {
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamDestroy(ctx->copy_stream);
    cudaEventDestroy(ctx->ev_geom);
    cudaEventDestroy(ctx->ev_shading);
    cudaEventDestroy(ctx->ev_band);
  }
  if (ctx->wave_stream) {
    cudaStreamSynchronize(ctx->wave_stream);
    cudaStreamDestroy(ctx->wave_stream);
    cudaEventDestroy(ctx->ev_fork);
    cudaEventDestroy(ctx->ev_join);
    cudaEventDestroy(ctx->ev_accum2);
  }
  if (ctx->ui_stream) {
    cudaStreamSynchronize(ctx->ui_stream);
    cudaStreamDestroy(ctx->ui_stream);
    cudaEventDestroy(ctx->ev_accum);
    cudaEventDestroy(ctx->ev_ui_done);
    cudaFreeHost(ctx->h_pixel);
  }
  for (int i = 0; i < 2; i++) {
    if (ctx->h_ring[i]) cudaFreeHost(ctx->h_ring[i]);
    if (ctx->ev_ring[i]) cudaEventDestroy(ctx->ev_ring[i]);
  }
  if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
  if (ctx->h_argb) cudaFreeHost(ctx->h_argb);
  free_dev_t(ctx->d_argb);
  free_dev(ctx->d_scratch);
  rtc_comm_destroy(ctx);
  drain_timing(ctx);
  for (cudaEvent_t e : ctx->free_events) cudaEventDestroy(e);
  free_scene_device(ctx);
  delete ctx->baked;
  free_pool(ctx);
  free_dev_t(ctx->d_ctl);
  free_dev_t(ctx->d_rays);
  free_dev_t(ctx->d_skip);
  free_dev_t(ctx->d_hits);
  free_dev_t(ctx->d_rgb);
  free_dev_t(ctx->d_samples);
  free_dev_t(ctx->d_misses);
  free_dev_t(ctx->d_base_rgb);
  free_dev_t(ctx->d_base_samples);
  free_dev_t(ctx->d_base_misses);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

const char* rtc_last_error(rtc_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int rtc_set_option(rtc_ctx* ctx, int option, int64_t value) {
  if (!ctx) return RTC_ERR_INVALID;
  switch (option) {
    case RTC_OPT_KERNEL_TIMING: ctx->timing = value != 0; return RTC_OK;
    case RTC_OPT_COUNTERS: ctx->counters = value != 0; return RTC_OK;
    case RTC_OPT_WAVES:
      if (value < 1 || value > 2) return fail(ctx, RTC_ERR_INVALID, "wavefronts in flight must be 1 or 2");
      ctx->waves = (int)value;
      return RTC_OK;
    case RTC_OPT_REORDER:
      if (value < 0 || value > 2) return fail(ctx, RTC_ERR_INVALID, "queue re-ordering mode must be 0, 1 or 2");
      ctx->reorder = (int)value;
      return RTC_OK;
    case RTC_OPT_MAX_PATHS:
      if (value < 1024 || value > (int64_t)1 << 28) return fail(ctx, RTC_ERR_INVALID, "max paths must be in [1024, 2^28]");
      if (value != ctx->max_paths) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        free_pool(ctx);
        free_dev_t(ctx->d_rays);
        free_dev_t(ctx->d_skip);
        free_dev_t(ctx->d_hits);
        ctx->stage_cap = 0;
        ctx->max_paths = value;
      }
      return RTC_OK;
    default: return fail(ctx, RTC_ERR_INVALID, "unknown option");
  }
}

int rtc_upload_scene(rtc_ctx* ctx, const rtc_scene_desc* s) {
  if (!ctx) return RTC_ERR_INVALID;
  if (!s || s->n_prims < 0 || s->n_xforms < 0) return fail(ctx, RTC_ERR_INVALID, "scene description is null or has negative counts");
  if (s->n_prims > 0 && (!s->kind || !s->flags || !s->geom || !s->material)) return fail(ctx, RTC_ERR_INVALID, "scene arrays must not be null");
  if (s->n_xforms > 0 && (!s->xform || !s->xforms)) return fail(ctx, RTC_ERR_INVALID, "xform arrays must not be null when n_xforms > 0");
  const size_t n = (size_t)s->n_prims;
  for (size_t i = 0; i < n; i++) {
    if (s->kind[i] > RTC_KIND_PLANE) return fail(ctx, RTC_ERR_INVALID, "unknown primitive kind at index " + std::to_string(i));
    if (s->xform && s->xform[i] >= s->n_xforms) return fail(ctx, RTC_ERR_INVALID, "xform index out of range at primitive " + std::to_string(i));
  }
  ctx->n_prims = s->n_prims;
  ctx->n_xforms = s->n_xforms;
  ctx->kind.assign(s->kind, s->kind + n);
  ctx->flags.assign(s->flags, s->flags + n);
  ctx->geom.resize(n * RTC_GEOM_STRIDE);  // (no value-initialisation: NoInitAlloc)
  ctx->material.resize(n * RTC_MATERIAL_STRIDE);
  {  // the two large arrays are copied in 4 MB pieces on the host threads
    struct Piece {
      char* dst;
      const char* src;
      size_t len;
    };
    std::vector<Piece> pieces;
    const size_t kPiece = (size_t)4 << 20;
    const size_t gb = n * RTC_GEOM_STRIDE * sizeof(double), mb = n * RTC_MATERIAL_STRIDE * sizeof(double);
    for (size_t o = 0; o < gb; o += kPiece) pieces.push_back({(char*)ctx->geom.data() + o, (const char*)s->geom + o, std::min(kPiece, gb - o)});
    for (size_t o = 0; o < mb; o += kPiece) pieces.push_back({(char*)ctx->material.data() + o, (const char*)s->material + o, std::min(kPiece, mb - o)});
    parallel_for(0, pieces.size(), 2, [&](size_t k) { std::memcpy(pieces[k].dst, pieces[k].src, pieces[k].len); });
  }
  if (s->xform)
    ctx->xform.assign(s->xform, s->xform + n);
  else
    ctx->xform.assign(n, -1);
  if (s->n_xforms > 0)
    ctx->xforms.assign(s->xforms, s->xforms + (size_t)s->n_xforms * RTC_XFORM_STRIDE);
  else
    ctx->xforms.clear();
  ctx->scene_set = true;
  ctx->bvh_set = false;  // Scene.AddPrimitive -> ResetAccelerator (Scene.cs:58-63)
  ctx->nodes.clear();
  ctx->root = -1;
  drop_device_tree(ctx);
  return RTC_OK;
}

static int finish_bvh(rtc_ctx* ctx) {
  cudaSetDevice(ctx->device);
  drop_device_tree(ctx);  // (the host tree in ctx->nodes is the current one)
  int rc = ctx->precision == RTC_F64 ? build_device_scene<double>(ctx) : build_device_scene<float>(ctx);
  ctx->bvh_set = rc == RTC_OK;
  return rc;
}

int rtc_upload_bvh(rtc_ctx* ctx, int32_t n_nodes, const rtc_bvh_node* nodes, int32_t root) {
  if (!ctx) return RTC_ERR_INVALID;
  if (!ctx->scene_set) return fail(ctx, RTC_ERR_STATE, "upload the scene before the BVH");
  if (ctx->n_prims == 0) return fail(ctx, RTC_ERR_INVALID, "scene has no primitives");
  if (!nodes || n_nodes <= 0) return fail(ctx, RTC_ERR_INVALID, "nodes is null or empty");
  ctx->nodes.assign(nodes, nodes + n_nodes);
  ctx->root = root;
  return finish_bvh(ctx);
}

int rtc_build_bvh(rtc_ctx* ctx) {
  if (!ctx) return RTC_ERR_INVALID;
  if (!ctx->scene_set || (int32_t)ctx->kind.size() != ctx->n_prims) return fail(ctx, RTC_ERR_STATE, "upload the scene before building the BVH");
  if (ctx->n_prims == 0) return fail(ctx, RTC_ERR_INVALID, "scene has no primitives");
  rtc_scene_desc d;
  d.n_prims = ctx->n_prims;
  d.n_xforms = ctx->n_xforms;
  d.kind = ctx->kind.data();
  d.flags = ctx->flags.data();
  d.geom = ctx->geom.data();
  d.xform = ctx->xform.data();
  d.xforms = ctx->xforms.empty() ? nullptr : ctx->xforms.data();
  d.material = ctx->material.data();
  const int n = d.n_prims;
  std::vector<double> lo((size_t)n * 3), hi((size_t)n * 3);
  for (int i = 0; i < n; i++) rtcore::DescPrimitiveBounds(d, i, &lo[(size_t)i * 3], &hi[(size_t)i * 3]);
  int threads = (int)std::thread::hardware_concurrency();
  ctx->root = rtcore::BuildBVH(n, lo.data(), hi.data(), ctx->nodes, threads > 0 ? threads : 1);
  return finish_bvh(ctx);
}

int rtc_build_bvh_device(rtc_ctx* ctx, int32_t radius, int32_t* rounds_out) {
  if (!ctx) return RTC_ERR_INVALID;
  if (!ctx->scene_set || (int32_t)ctx->kind.size() != ctx->n_prims) return fail(ctx, RTC_ERR_STATE, "upload the scene before building the BVH");
  if (ctx->n_prims == 0) return fail(ctx, RTC_ERR_INVALID, "scene has no primitives");
  if (radius <= 0) radius = 16;
  rtc_scene_desc d;
  d.n_prims = ctx->n_prims;
  d.n_xforms = ctx->n_xforms;
  d.kind = ctx->kind.data();
  d.flags = ctx->flags.data();
  d.geom = ctx->geom.data();
  d.xform = ctx->xform.data();
  d.xforms = ctx->xforms.empty() ? nullptr : ctx->xforms.data();
  d.material = ctx->material.data();
  const int n = d.n_prims;
  // leaf boxes exactly as AABB.CreateFromBounded (AABB.cs:20-36); primitives with infinite boxes (planes) stay out of the
  // clustering and are chained above its root in ID order, like the host builder does
  std::vector<double> boxes;
  std::vector<int32_t> ids, unbounded;
  boxes.reserve((size_t)n * 6);
  ids.reserve(n);
  std::vector<double> ub;
  for (int i = 0; i < n; i++) {
    double lo[3], hi[3];
    rtcore::DescPrimitiveBounds(d, i, lo, hi);
    bool finite = true;
    for (int k = 0; k < 3; k++) finite = finite && std::isfinite(lo[k]) && std::isfinite(hi[k]);
    if (finite) {
      ids.push_back(i);
      boxes.insert(boxes.end(), lo, lo + 3);
      boxes.insert(boxes.end(), hi, hi + 3);
    } else {
      unbounded.push_back(i);
      ub.insert(ub.end(), lo, lo + 3);
      ub.insert(ub.end(), hi, hi + 3);
    }
  }
  const int32_t m = (int32_t)ids.size();
  std::vector<rtc_bvh_node>& nodes = ctx->nodes;
  nodes.clear();
  nodes.resize((m > 0 ? (size_t)2 * m - 1 : 0) + 2 * unbounded.size());
  int32_t root = -1, rounds = 0, next = 0;
  cudaSetDevice(ctx->device);
  if (m > 0) {
    cudaError_t e = build_bvh_ploc(ctx->stream, m, boxes.data(), ids.data(), radius, nodes.data(), &root, &rounds);
    if (e != cudaSuccess) return fail(ctx, RTC_ERR_CUDA, std::string("build_bvh_ploc: ") + cudaGetErrorString(e));
    next = 2 * m - 1;
  }
  for (size_t j = unbounded.size(); j-- > 0;) {
    rtc_bvh_node& leaf = nodes[next];
    std::memset(&leaf, 0, sizeof(leaf));
    for (int k = 0; k < 3; k++) {
      leaf.bmin[k] = ub[j * 6 + k];
      leaf.bmax[k] = ub[j * 6 + 3 + k];
    }
    leaf.left = leaf.right = -1;
    leaf.prim = unbounded[j];
    if (root < 0) {
      root = next++;
      continue;
    }
    rtc_bvh_node& par = nodes[next + 1];
    std::memset(&par, 0, sizeof(par));
    par.left = next;
    par.right = root;
    par.prim = -1;
    for (int k = 0; k < 3; k++) {
      par.bmin[k] = std::fmin(leaf.bmin[k], nodes[root].bmin[k]);
      par.bmax[k] = std::fmax(leaf.bmax[k], nodes[root].bmax[k]);
    }
    root = next + 1;
    next += 2;
  }
  nodes.resize(next);
  ctx->root = root;
  if (rounds_out) *rounds_out = rounds;
  return finish_bvh(ctx);
}

int rtc_prepare_device(rtc_ctx* ctx, int32_t builder, int32_t radius, rtc_prepare_stats* stats) {
  if (!ctx) return RTC_ERR_INVALID;
  if (!ctx->scene_set || (int32_t)ctx->kind.size() != ctx->n_prims) return fail(ctx, RTC_ERR_STATE, "upload the scene before preparing it");
  if (ctx->n_prims == 0) return fail(ctx, RTC_ERR_INVALID, "scene has no primitives");
  if (builder != RTC_BUILDER_SAH && builder != RTC_BUILDER_PLOC) return fail(ctx, RTC_ERR_INVALID, "unknown builder");
  if (ctx->n_prims > (int32_t)REF_SLOT_MASK) return fail(ctx, RTC_ERR_UNSUPPORTED, "more than 2^26-1 primitives");
  if (radius <= 0) radius = 16;
  using clk = std::chrono::steady_clock;
  const auto t0 = clk::now();
  auto ms_since = [](clk::time_point a) { return std::chrono::duration<double, std::milli>(clk::now() - a).count(); };
  rtc_prepare_stats st;
  std::memset(&st, 0, sizeof(st));
  cudaSetDevice(ctx->device);
  const int32_t n = ctx->n_prims;
  const bool f32 = ctx->precision == RTC_F32;
  int rc = wait_shading_upload(ctx);
  if (rc) return rc;
  if ((rc = ensure_copy_stream(ctx)) || (rc = ensure_stage_ring(ctx))) return rc;
  CU(cudaStreamSynchronize(ctx->stream));
  drop_device_tree(ctx);
  ctx->nodes.clear();
  ctx->root = -1;
  ctx->bvh_set = false;
  rtc_scene_desc d;
  d.n_prims = n;
  d.n_xforms = ctx->n_xforms;
  d.kind = ctx->kind.data();
  d.flags = ctx->flags.data();
  d.geom = ctx->geom.data();
  d.xform = ctx->xform.data();
  d.xforms = ctx->xforms.empty() ? nullptr : ctx->xforms.data();
  d.material = ctx->material.data();

  // One device allocation for everything this call needs beside the scene itself: the boxes, the staged records, the slot
  // table, and the work arrays of the builder and of the flatten (which run one after the other and share their part).
  const size_t nz = (size_t)n;
  const int32_t nn_max = 2 * n;  // (2m - 1 + 2 (n - m) <= 2n)
  const size_t fixed_bytes = nz * 6 * sizeof(double) + nz * sizeof(int32_t) * 2 + (f32 ? nz * sizeof(StagedPrim) : 0) + 16 * 256;
  const size_t work_bytes = std::max(build_bvh_sah_scratch_bytes(n), f32 ? flatten_scratch_bytes(nn_max, n) : (size_t)0);
  PrepareArena arena;
  CU(cudaMalloc((void**)&arena.base, fixed_bytes + work_bytes));
  arena.cap = fixed_bytes + work_bytes;
  struct ArenaFree {
    rtc_ctx* c;
    PrepareArena& a;
    std::thread& t;
    ~ArenaFree() {  // (after the helper thread has ended and everything queued on the two streams has run)
      if (t.joinable()) t.join();
      cudaStreamSynchronize(c->stream);
      if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
      cudaFree(a.base);
    }
  };
  std::thread uploader;
  ArenaFree arena_free{ctx, arena, uploader};
  auto carve = [&](size_t bytes) {
    void* p = arena.base + arena.used;
    arena.used += (std::max<size_t>(bytes, 16) + 255) & ~(size_t)255;
    return p;
  };
  double* d_allboxes = (double*)carve(nz * 6 * sizeof(double));
  int32_t* d_ids = (int32_t*)carve(nz * sizeof(int32_t));
  int32_t* d_slot_prim = (int32_t*)carve(nz * sizeof(int32_t));
  StagedPrim* d_staged = f32 ? (StagedPrim*)carve(nz * sizeof(StagedPrim)) : nullptr;

  // ---- leaf boxes exactly as AABB.CreateFromBounded (AABB.cs:20-36): host threads -> pinned ring -> device ------------------
  // primitives with infinite boxes (planes) stay out of the build and are chained above its root in ID order, like the host
  // builder does; the bounded ones keep their order (ascending ID)
  struct Unb {
    int32_t id;
    double box[6];
  };
  std::vector<Unb> unb;
  std::mutex unb_mutex;
  rc = staged_upload(ctx, ctx->stream, nz, 6 * sizeof(double), (char*)d_allboxes, [&](size_t i, char* dst) {
    double* bx = (double*)dst;
    rtcore::DescPrimitiveBounds(d, (int)i, bx, bx + 3);
    bool finite = true;
    for (int k = 0; k < 6; k++) finite = finite && std::isfinite(bx[k]);
    if (!finite) {
      Unb u;
      u.id = (int32_t)i;
      std::memcpy(u.box, bx, sizeof(u.box));
      std::lock_guard<std::mutex> g(unb_mutex);
      unb.push_back(u);
    }
  });
  if (rc) return rc;
  // the ring is handed to the helper thread below with nothing in flight (no event recorded by one thread is waited for by
  // another), and the host needs the boxes to be through before it sizes the tree anyway
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->ring_busy[0] = ctx->ring_busy[1] = false;
  std::sort(unb.begin(), unb.end(), [](const Unb& x, const Unb& y) { return x.id < y.id; });
  const int32_t nu = (int32_t)unb.size(), m = n - nu;
  const double* d_boxes = d_allboxes;
  const int32_t* d_prim_ids = nullptr;  // identity
  std::vector<int32_t> ids;
  if (nu > 0 && m > 0) {
    ids.reserve(m);
    size_t u = 0;
    for (int32_t i = 0; i < n; i++) {
      if (u < unb.size() && unb[u].id == i) {
        u++;
        continue;
      }
      ids.push_back(i);
    }
    // the bounded primitives' boxes, compacted: out of the work area (scenes with planes only; the builder's pool allocates
    // on its own what no longer fits)
    double* d_boxes_c = (double*)carve((size_t)m * 6 * sizeof(double));
    CU(cudaMemcpyAsync(d_ids, ids.data(), (size_t)m * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    cudaError_t e = launch_gather_boxes(ctx->stream, m, d_ids, d_allboxes, d_boxes_c);
    if (e != cudaSuccess) return fail(ctx, RTC_ERR_CUDA, std::string("gather boxes: ") + cudaGetErrorString(e));
    CU(cudaStreamSynchronize(ctx->stream));  // (ids is read by the copy until here)
    d_boxes = d_boxes_c;
    d_prim_ids = d_ids;
  }
  st.boxes_ms = ms_since(t0);
  const auto t1 = clk::now();

  const int32_t nn = (m > 0 ? 2 * m - 1 : 0) + (m > 0 ? 2 * nu : (nu > 0 ? 2 * nu - 1 : 0));
  rtc_bvh_node* d_bn = nullptr;
  CU(cudaMalloc((void**)&d_bn, (size_t)nn * sizeof(rtc_bvh_node)));
  ctx->d_bnodes = d_bn;  // (owned by the context from here; dropped again on failure)
  ctx->n_bnodes = nn;
  auto bail = [&](int code, const std::string& msg) {
    if (uploader.joinable()) uploader.join();
    cudaStreamSynchronize(ctx->stream);
    drop_device_tree(ctx);
    return fail(ctx, code, msg);
  };

#define BCK(call)                                                                                    \
  do {                                                                                               \
    cudaError_t e__ = (call);                                                                        \
    if (e__ != cudaSuccess) return bail(RTC_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
  } while (0)
  // the records travel on the copy stream, staged by a second host thread (float conversion and material packing on the host
  // threads, through the same pinned ring) while this thread drives the level-synchronous build; the records kernel at the very
  // end is the only reader
  int up_rc = RTC_OK;
  if (f32) {
    uploader = std::thread([&]() {
      cudaSetDevice(ctx->device);
      up_rc = staged_upload(ctx, ctx->copy_stream, nz, sizeof(StagedPrim), (char*)d_staged, [&](size_t i, char* dst) {
        StagedPrim* sp = (StagedPrim*)dst;
        const double* g = &ctx->geom[i * RTC_GEOM_STRIDE];
        for (int k = 0; k < 12; k++) sp->geom[k] = (float)g[k];
        const double* mt = &ctx->material[i * RTC_MATERIAL_STRIDE];
        DMat<float> dm;
        pack_material(dm, mt, mt[13] > 0);  // Primitive.IsReflective (Primitive.cs:106)
        std::memcpy(sp->mat, dm.w, sizeof(dm.w));
        sp->kind_flags = (uint32_t)ctx->kind[i] | ((uint32_t)ctx->flags[i] << 8);
        sp->xform = ctx->xform[i];
        sp->pad[0] = sp->pad[1] = 0;
      });
      if (up_rc == RTC_OK && cudaEventRecord(ctx->ev_geom, ctx->copy_stream) != cudaSuccess) up_rc = RTC_ERR_CUDA;
      if (up_rc == RTC_OK && cudaStreamSynchronize(ctx->copy_stream) != cudaSuccess) up_rc = RTC_ERR_CUDA;
      ctx->ring_busy[0] = ctx->ring_busy[1] = false;  // (drained: the next call starts with an idle ring)
    });
  }

  // ---- the tree ----------------------------------------------------------------------------------------------------------
  int32_t root = -1, levels = 0;
  if (m > 0) {
    if (builder == RTC_BUILDER_SAH) {
      const size_t mark = arena.used;
      cudaError_t e = build_bvh_sah_device(ctx->stream, &arena, ctx->h_pin, m, d_boxes, d_prim_ids, d_bn, &levels);
      arena.used = mark;
      if (e != cudaSuccess) return bail(RTC_ERR_CUDA, std::string("build_bvh_sah_device: ") + cudaGetErrorString(e));
      root = 0;
    } else {
      std::vector<double> boxes((size_t)m * 6);
      std::vector<int32_t> pid(m);
      BCK(cudaMemcpyAsync(boxes.data(), d_boxes, boxes.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
      BCK(cudaStreamSynchronize(ctx->stream));
      for (int32_t i = 0; i < m; i++) pid[i] = ids.empty() ? i : ids[i];
      std::vector<rtc_bvh_node> hn((size_t)2 * m - 1);
      cudaError_t e = build_bvh_ploc(ctx->stream, m, boxes.data(), pid.data(), radius, hn.data(), &root, &levels);
      if (e != cudaSuccess) return bail(RTC_ERR_CUDA, std::string("build_bvh_ploc: ") + cudaGetErrorString(e));
      BCK(cudaMemcpyAsync(d_bn, hn.data(), hn.size() * sizeof(rtc_bvh_node), cudaMemcpyHostToDevice, ctx->stream));
      BCK(cudaStreamSynchronize(ctx->stream));
    }
  }
  if (nu > 0) {  // the chain of unbounded leaves above the root (bvh_builder.cpp: BuildBVH), first plane outermost-left
    std::vector<rtc_bvh_node> chain;
    rtc_bvh_node top;
    std::memset(&top, 0, sizeof(top));
    if (root >= 0) {
      BCK(cudaMemcpyAsync(&top, d_bn + root, sizeof(top), cudaMemcpyDeviceToHost, ctx->stream));
      BCK(cudaStreamSynchronize(ctx->stream));
    }
    int32_t next = m > 0 ? 2 * m - 1 : 0;
    const int32_t chain_begin = next;
    for (int32_t j = nu; j-- > 0;) {
      rtc_bvh_node leaf;
      std::memset(&leaf, 0, sizeof(leaf));
      for (int k = 0; k < 3; k++) {
        leaf.bmin[k] = unb[j].box[k];
        leaf.bmax[k] = unb[j].box[3 + k];
      }
      leaf.left = leaf.right = -1;
      leaf.prim = unb[j].id;
      chain.push_back(leaf);
      if (root < 0) {
        root = next++;
        top = leaf;
        continue;
      }
      rtc_bvh_node par;
      std::memset(&par, 0, sizeof(par));
      par.left = next;
      par.right = root;
      par.prim = -1;
      for (int k = 0; k < 3; k++) {
        par.bmin[k] = std::fmin(leaf.bmin[k], top.bmin[k]);
        par.bmax[k] = std::fmax(leaf.bmax[k], top.bmax[k]);
      }
      chain.push_back(par);
      top = par;
      root = next + 1;
      next += 2;
    }
    BCK(cudaMemcpyAsync(d_bn + chain_begin, chain.data(), chain.size() * sizeof(rtc_bvh_node), cudaMemcpyHostToDevice, ctx->stream));
    BCK(cudaStreamSynchronize(ctx->stream));
  }
  ctx->root = root;
  st.build_levels = levels;
  st.build_ms = ms_since(t1);
  const auto t2 = clk::now();

  if (!f32) {  // parity mode: the f64 layout is made on the host from the tree just built
    rc = ensure_host_nodes(ctx);
    if (rc) return bail(RTC_ERR_CUDA, "tree read-back failed");
    cudaStreamSynchronize(ctx->stream);
    free_dev_t(ctx->d_bnodes);  // (finish_bvh keeps the host tree)
    ctx->n_bnodes = 0;
    rc = finish_bvh(ctx);
    st.flatten_ms = ms_since(t2);
    st.total_ms = ms_since(t0);
    st.n_bounded = m;
    if (stats) *stats = st;
    return rc;
  }

  // ---- the device layout ---------------------------------------------------------------------------------------------------
  using S = rtc_baked;
  if ((rc = ensure_seg(ctx, S::S_PRIMS, nz * sizeof(DPrim<float>))) || (rc = ensure_seg(ctx, S::S_MATS, nz * sizeof(DMat<float>))) ||
      (rc = ensure_seg(ctx, S::S_AUX, nz * sizeof(int32_t))) || (rc = ensure_seg(ctx, S::S_PRIM_ID, nz * sizeof(int32_t))) ||
      (rc = ensure_seg(ctx, S::S_ID_TO_SLOT, nz * sizeof(int32_t))) || (rc = ensure_seg(ctx, S::S_SGEOM, nz * sizeof(V4<float>))) ||
      (rc = ensure_seg(ctx, S::S_NODES, 0)) || (rc = ensure_seg(ctx, S::S_UNBOUNDED, (size_t)nu * sizeof(uint32_t))) ||
      (rc = ensure_seg(ctx, S::S_XFORMS, (size_t)std::max(1, ctx->n_xforms) * sizeof(DXform<float>))))
    return bail(rc, ctx->err);
  FlattenInput fin;
  fin.nodes = d_bn;
  fin.n_nodes = nn;
  fin.root = root;
  fin.n_prims = n;
  fin.staged = d_staged;
  fin.records_ready = nullptr;  // (joined below: the event is recorded by the helper thread)
  fin.alloc_ctx = ctx;
  fin.alloc_qnodes = [](void* c, size_t bytes, void** out) {
    rtc_ctx* cx = (rtc_ctx*)c;
    const int r = ensure_seg(cx, rtc_baked::S_QNODES, bytes);
    *out = cx->d_qnodes;
    return r;
  };
  FlattenOutput fout;
  fout.slot_prim = d_slot_prim;
  fout.prims = ctx->d_prims;
  fout.mats = ctx->d_mats;
  fout.sgeom = ctx->d_sgeom;
  fout.aux = ctx->d_aux;
  fout.prim_id = ctx->d_prim_id;
  fout.id_to_slot = ctx->d_id_to_slot;
  // the helper thread has usually finished long before the tree has; its event orders the records kernel behind the copies
  if (uploader.joinable()) uploader.join();
  if (up_rc != RTC_OK) return bail(up_rc, "staging the scene records failed");
  fin.records_ready = ctx->ev_geom;
  std::string ferr;
  rc = flatten_device(ctx->stream, &arena, ctx->h_pin, fin, fout, ferr);
  if (rc) return bail(rc, ferr);
  if (fout.depth + 1 > kQStackMax)
    return bail(RTC_ERR_UNSUPPORTED, "8-wide BVH depth " + std::to_string(fout.depth) + " exceeds the traversal stack (" + std::to_string(kQStackMax - 1) + ")");
  if (fout.n_qnodes == 0 && ctx->d_qnodes) {  // no bounded primitive: the kernel keys on a null qnodes pointer
    cudaFree(ctx->d_qnodes);
    ctx->d_qnodes = nullptr;
    ctx->seg_cap[S::S_QNODES] = 0;
  }
  // unbounded leaf references and the transform rows: small, made on the host like build_device_scene does
  std::vector<uint32_t> unb_refs;
  for (size_t r = 0; r < fout.unbounded_prims.size(); r++) {
    const int32_t p = fout.unbounded_prims[r];
    unb_refs.push_back(prep::leaf_ref_of(ctx->kind[p], ctx->flags[p], ctx->xform[p], (uint32_t)(fout.n_bounded + (int32_t)r)));
  }
  if (!unb_refs.empty())
    BCK(cudaMemcpyAsync(ctx->d_unbounded, unb_refs.data(), unb_refs.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
  std::vector<DXform<float>> dx(std::max(1, ctx->n_xforms));
  std::memset(dx.data(), 0, dx.size() * sizeof(DXform<float>));
  for (int32_t j = 0; j < ctx->n_xforms; j++) {
    const double* x = &ctx->xforms[(size_t)j * RTC_XFORM_STRIDE];
    for (int mtx = 0; mtx < 3; mtx++)
      for (int row = 0; row < 3; row++) {
        const double* r = x + mtx * 16 + row * 4;
        dx[j].r[mtx * 3 + row] = V4<float>{(float)r[0], (float)r[1], (float)r[2], (float)r[3]};
      }
  }
  if (ctx->n_xforms > 0)
    for (int32_t p = 0; p < n; p++)  // vertex-normal triangles keep n0,n1,n2 in the first 9 doubles of their row
      if (ctx->kind[p] == RTC_KIND_TRIANGLE && (ctx->flags[p] & RTC_FLAG_VNORMALS) && ctx->xform[p] >= 0) {
        const int32_t j = ctx->xform[p];
        const double* x = &ctx->xforms[(size_t)j * RTC_XFORM_STRIDE];
        for (int k2 = 0; k2 < 3; k2++) dx[j].r[k2] = V4<float>{(float)x[k2 * 3], (float)x[k2 * 3 + 1], (float)x[k2 * 3 + 2], 0.0f};
      }
  BCK(cudaMemcpyAsync(ctx->d_xforms, dx.data(), dx.size() * sizeof(DXform<float>), cudaMemcpyHostToDevice, ctx->stream));
  BCK(cudaStreamSynchronize(ctx->stream));
  ctx->n_unbounded = nu;
  ctx->root_node = 0;
  ctx->bvh_depth = fout.depth;
  const size_t sb[S::S_COUNT] = {0, (size_t)fout.n_qnodes * sizeof(CNode), (size_t)nu * sizeof(uint32_t), nz * sizeof(DPrim<float>),
                                 nz * sizeof(DMat<float>), dx.size() * sizeof(DXform<float>), nz * sizeof(int32_t), nz * sizeof(int32_t),
                                 nz * sizeof(int32_t), nz * sizeof(V4<float>)};
  for (int i = 0; i < S::S_COUNT; i++) ctx->seg_bytes[i] = sb[i];
  delete ctx->baked;
  ctx->baked = nullptr;
  ctx->bvh_set = true;
  st.flatten_ms = ms_since(t2);
  st.total_ms = ms_since(t0);
  st.wide_depth = fout.depth;
  st.n_wide_nodes = fout.n_qnodes;
  st.n_bounded = fout.n_bounded;
  if (stats) *stats = st;
  if (std::getenv("RTC_B200_VERBOSE"))
    std::fprintf(stderr, "[rtcore_b200] prepare on the device: %d primitives, boxes %.1f ms, build %.1f ms (%d levels), flatten %.1f ms, total %.1f ms; "
                         "%d wide nodes, depth %d\n", n, st.boxes_ms, st.build_ms, levels, st.flatten_ms, st.total_ms, fout.n_qnodes, fout.depth);
  return RTC_OK;
}
#undef BCK

int rtc_bake(rtc_ctx* ctx, rtc_baked** out) {
  if (!ctx || !out) return RTC_ERR_INVALID;
  *out = nullptr;
  if (!ctx->bvh_set) return fail(ctx, RTC_ERR_STATE, "no scene + BVH to bake (rtc_upload_scene, then rtc_upload_bvh, rtc_build_bvh or rtc_prepare_device)");
  cudaSetDevice(ctx->device);
  if (!ctx->baked) {  // prepared on the device, or received from another rank: the image is read back segment by segment
    int rc = wait_shading_upload(ctx);
    if (rc) return rc;
    rtc_baked* c = new rtc_baked();
    c->precision = ctx->precision;
    c->n_prims = ctx->n_prims;
    c->n_unbounded = ctx->n_unbounded;
    c->n_xforms = ctx->n_xforms;
    c->bvh_depth = ctx->bvh_depth;
    c->root_node = ctx->root_node;
    size_t total = 0;
    for (int i = 0; i < rtc_baked::S_COUNT; i++) {
      c->off[i] = total;
      c->bytes[i] = ctx->seg_bytes[i];
      total += (ctx->seg_bytes[i] + 255) & ~(size_t)255;
    }
    if (!c->alloc(total)) {
      delete c;
      return fail(ctx, RTC_ERR_NOMEM, "host allocation for the baked scene failed");
    }
    const void* src[rtc_baked::S_COUNT] = {ctx->d_nodes, ctx->d_qnodes, ctx->d_unbounded, ctx->d_prims, ctx->d_mats,
                                           ctx->d_xforms, ctx->d_aux, ctx->d_prim_id, ctx->d_id_to_slot, ctx->d_sgeom};
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < rtc_baked::S_COUNT && e == cudaSuccess; i++)
      if (c->bytes[i]) e = cudaMemcpyAsync(c->host + c->off[i], src[i], c->bytes[i], cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      delete c;
      return fail(ctx, RTC_ERR_CUDA, std::string("rtc_bake (read-back): ") + cudaGetErrorString(e));
    }
    *out = c;
    return RTC_OK;
  }
  rtc_baked* c = new rtc_baked();
  const rtc_baked* b = ctx->baked;
  c->precision = b->precision;
  c->n_prims = b->n_prims;
  c->n_unbounded = b->n_unbounded;
  c->n_xforms = b->n_xforms;
  c->bvh_depth = b->bvh_depth;
  c->root_node = b->root_node;
  std::memcpy(c->off, b->off, sizeof(c->off));
  std::memcpy(c->bytes, b->bytes, sizeof(c->bytes));
  if (!c->alloc(b->total)) {
    delete c;
    return fail(ctx, RTC_ERR_NOMEM, "host allocation for the baked scene failed");
  }
  std::memcpy(c->host, b->host, b->total);
  *out = c;
  return RTC_OK;
}

int rtc_upload_baked(rtc_ctx* ctx, const rtc_baked* baked) {
  if (!ctx) return RTC_ERR_INVALID;
  if (!baked) return fail(ctx, RTC_ERR_INVALID, "baked scene is null");
  if (baked->precision != ctx->precision) return fail(ctx, RTC_ERR_INVALID, "baked scene was made for the other arithmetic mode");
  cudaSetDevice(ctx->device);
  int rc = upload_baked_image(ctx, baked);
  if (rc) return rc;
  ctx->n_prims = baked->n_prims;
  ctx->n_xforms = baked->n_xforms;
  if ((int32_t)ctx->kind.size() != baked->n_prims) {  // no host-side description behind this image
    ctx->kind.clear();
    ctx->flags.clear();
    ctx->geom.clear();
    ctx->material.clear();
    ctx->xform.clear();
    ctx->xforms.clear();
    ctx->nodes.clear();
    ctx->root = -1;
    drop_device_tree(ctx);
  }
  ctx->scene_set = true;
  ctx->bvh_set = true;
  return RTC_OK;
}

int64_t rtc_baked_bytes(const rtc_baked* baked) {
  if (!baked) return 0;
  int64_t t = 0;
  for (size_t b : baked->bytes) t += (int64_t)b;
  return t;
}

int rtc_baked_segment(const rtc_baked* baked, int32_t segment, const void** data, int64_t* bytes) {
  static_assert((int)rtc_baked::S_COUNT == (int)RTC_BAKED_SEGMENTS, "segment count in the public header");
  if (!baked || !data || !bytes || segment < 0 || segment >= rtc_baked::S_COUNT) return RTC_ERR_INVALID;
  *data = baked->host + baked->off[segment];
  *bytes = (int64_t)baked->bytes[segment];
  return RTC_OK;
}

void rtc_baked_free(rtc_baked* baked) { delete baked; }

int rtc_get_bvh_size(rtc_ctx* ctx, int32_t* n_nodes, int32_t* root) {
  if (!ctx || !n_nodes || !root) return RTC_ERR_INVALID;
  if (ctx->bvh_set && ensure_host_nodes(ctx)) return RTC_ERR_CUDA;
  if (!ctx->bvh_set || ctx->nodes.empty()) return fail(ctx, RTC_ERR_STATE, "no host-side BVH (none set, or the scene came from rtc_upload_baked)");
  *n_nodes = (int32_t)ctx->nodes.size();
  *root = ctx->root;
  return RTC_OK;
}

int rtc_get_bvh(rtc_ctx* ctx, int32_t capacity, rtc_bvh_node* nodes) {
  if (!ctx || !nodes) return RTC_ERR_INVALID;
  if (!ctx->bvh_set) return fail(ctx, RTC_ERR_STATE, "no BVH");
  if (ensure_host_nodes(ctx)) return RTC_ERR_CUDA;
  if (capacity < (int32_t)ctx->nodes.size()) return fail(ctx, RTC_ERR_INVALID, "capacity too small");
  std::memcpy(nodes, ctx->nodes.data(), ctx->nodes.size() * sizeof(rtc_bvh_node));
  return RTC_OK;
}

int rtc_set_camera(rtc_ctx* ctx, const rtc_camera* camera) {
  if (!ctx) return RTC_ERR_INVALID;
  if (!camera) return fail(ctx, RTC_ERR_INVALID, "camera is null");
  if (camera->kind != RTC_CAMERA_FRUSTUM && camera->kind != RTC_CAMERA_ORTHO) return fail(ctx, RTC_ERR_INVALID, "unknown camera kind");
  ctx->cam = *camera;
  ctx->camera_set = true;
  return RTC_OK;
}

int rtc_set_params(rtc_ctx* ctx, const rtc_params* p) {
  if (!ctx) return RTC_ERR_INVALID;
  if (!p) return fail(ctx, RTC_ERR_INVALID, "params is null");
  if (p->width <= 0 || p->height <= 0 || p->recursion < 0) return fail(ctx, RTC_ERR_INVALID, "width/height must be positive and recursion non-negative");
  if ((int64_t)p->width * p->height > (int64_t)1 << 30) return fail(ctx, RTC_ERR_UNSUPPORTED, "image larger than 2^30 pixels");
  ctx->par = *p;
  ctx->params_set = true;
  cudaSetDevice(ctx->device);
  return ensure_accum(ctx);
}

int rtc_trace_closest(rtc_ctx* ctx, int64_t n, const rtc_ray* rays, const rtc_hit* skip, rtc_hit* out) {
  if (!ctx) return RTC_ERR_INVALID;
  if (n < 0 || (n > 0 && (!rays || !out))) return fail(ctx, RTC_ERR_INVALID, "rays/out must not be null");
  int rc = ready(ctx, false);
  if (rc) return rc;
  if (n == 0) return RTC_OK;
  cudaSetDevice(ctx->device);
  rc = wait_shading_upload(ctx);
  if (rc) return rc;
  return ctx->precision == RTC_F64 ? trace_batch<double>(ctx, n, rays, skip, out) : trace_batch<float>(ctx, n, rays, skip, out);
}

int rtc_camera_rays(rtc_ctx* ctx, int64_t n, const int32_t* xy, const uint32_t* sample, rtc_ray* out) {
  if (!ctx) return RTC_ERR_INVALID;
  if (n < 0 || (n > 0 && (!xy || !sample || !out))) return fail(ctx, RTC_ERR_INVALID, "xy/sample/out must not be null");
  if (!ctx->camera_set || !ctx->params_set) return fail(ctx, RTC_ERR_STATE, "camera and params must be set");
  if (n == 0) return RTC_OK;
  cudaSetDevice(ctx->device);
  int32_t* dxy = nullptr;
  uint32_t* ds = nullptr;
  rtc_ray* dout = nullptr;
  CU(cudaMalloc((void**)&dxy, sizeof(int32_t) * 2 * n));
  CU(cudaMalloc((void**)&ds, sizeof(uint32_t) * n));
  CU(cudaMalloc((void**)&dout, sizeof(rtc_ray) * n));
  CU(cudaMemcpyAsync(dxy, xy, sizeof(int32_t) * 2 * n, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ds, sample, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, ctx->stream));
  LaunchCfg cfg{ctx->stream, ctx->sm_count, false};
  cudaError_t e;
  if (ctx->precision == RTC_F64)
    e = Kernels<double>::camera_rays(cfg, camera_view<double>(ctx->cam), params_view<double>(ctx->par), n, dxy, ds, dout);
  else
    e = Kernels<float>::camera_rays(cfg, camera_view<float>(ctx->cam), params_view<float>(ctx->par), n, dxy, ds, dout);
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, dout, sizeof(rtc_ray) * n, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(dxy);
  cudaFree(ds);
  cudaFree(dout);
  if (e != cudaSuccess) return fail(ctx, RTC_ERR_CUDA, std::string("camera_rays: ") + cudaGetErrorString(e));
  return RTC_OK;
}

int rtc_debug_create_horizon(rtc_ctx* ctx, int64_t n, const double* pole_z_theta, double* out) {
  if (!ctx) return RTC_ERR_INVALID;
  if (n < 0 || (n > 0 && (!pole_z_theta || !out))) return fail(ctx, RTC_ERR_INVALID, "pole_z_theta/out must not be null");
  if (n == 0) return RTC_OK;
  cudaSetDevice(ctx->device);
  int rc = ensure_scratch(ctx, (size_t)n * 8 * sizeof(double));
  if (rc) return rc;
  double* din = (double*)ctx->d_scratch;
  double* dout = din + 5 * n;
  CU(cudaMemcpyAsync(din, pole_z_theta, sizeof(double) * 5 * n, cudaMemcpyHostToDevice, ctx->stream));
  LaunchCfg cfg{ctx->stream, ctx->sm_count, false};
  cudaError_t e = ctx->precision == RTC_F64 ? Kernels<double>::horizon(cfg, n, din, dout) : Kernels<float>::horizon(cfg, n, din, dout);
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, dout, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return fail(ctx, RTC_ERR_CUDA, std::string("create_horizon: ") + cudaGetErrorString(e));
  return RTC_OK;
}

int rtc_render(rtc_ctx* ctx, int32_t x0, int32_t y0, int32_t x1, int32_t y1, uint32_t first_sample, uint32_t n_samples) {
  if (!ctx) return RTC_ERR_INVALID;
  int rc = ready(ctx, true);
  if (rc) return rc;
  if (x0 < 0 || y0 < 0 || x1 > ctx->par.width || y1 > ctx->par.height || x0 >= x1 || y0 >= y1)
    return fail(ctx, RTC_ERR_INVALID, "render rectangle is empty or outside the image");
  if (n_samples == 0) return RTC_OK;
  cudaSetDevice(ctx->device);
  rc = ensure_accum(ctx);
  if (rc) return rc;
  return ctx->precision == RTC_F64 ? render_rect<double>(ctx, x0, y0, x1, y1, first_sample, n_samples, true, nullptr)
                                   : render_rect<float>(ctx, x0, y0, x1, y1, first_sample, n_samples, true, nullptr);
}

int rtc_render_read(rtc_ctx* ctx, uint32_t first_sample, uint32_t n_samples, double* rgb_sum, uint32_t* samples, uint32_t* misses) {
  if (!ctx) return RTC_ERR_INVALID;
  int rc = ready(ctx, true);
  if (rc) return rc;
  cudaSetDevice(ctx->device);
  rc = ensure_accum(ctx);
  if (rc) return rc;
  rc = ensure_copy_stream(ctx);
  if (rc) return rc;
  ReadBack rb{rgb_sum, samples, misses};
  const int w = ctx->par.width, h = ctx->par.height;
  if (n_samples == 0) {  // nothing to render: a plain read of the whole planes
    return rtc_read_accum(ctx, rgb_sum, samples, misses);
  }
  rc = ctx->precision == RTC_F64 ? render_rect<double>(ctx, 0, 0, w, h, first_sample, n_samples, true, nullptr, &rb)
                                 : render_rect<float>(ctx, 0, 0, w, h, first_sample, n_samples, true, nullptr, &rb);
  // both streams are drained even on error: the caller owns the host buffers again when this returns
  cudaError_t e1 = cudaStreamSynchronize(ctx->copy_stream), e2 = cudaStreamSynchronize(ctx->stream);
  drain_timing(ctx);
  if (rc) return rc;
  CU(e1);
  CU(e2);
  return RTC_OK;
}

int rtc_sync(rtc_ctx* ctx) {
  if (!ctx) return RTC_ERR_INVALID;
  cudaSetDevice(ctx->device);
  CU(cudaStreamSynchronize(ctx->stream));
  if (ctx->copy_stream) CU(cudaStreamSynchronize(ctx->copy_stream));
  drain_timing(ctx);
  return RTC_OK;
}

int rtc_clear_accum(rtc_ctx* ctx) {
  if (!ctx) return RTC_ERR_INVALID;
  if (!ctx->params_set) return fail(ctx, RTC_ERR_STATE, "no render parameters (rtc_set_params)");
  cudaSetDevice(ctx->device);
  int rc = ensure_accum(ctx);
  if (rc) return rc;
  size_t n = (size_t)ctx->acc_w * ctx->acc_h;
  rc = before_accum_write(ctx);
  if (rc) return rc;
  CU(cudaMemsetAsync(ctx->d_rgb, 0, n * 3 * sizeof(double), ctx->stream));
  CU(cudaMemsetAsync(ctx->d_samples, 0, n * sizeof(uint32_t), ctx->stream));
  CU(cudaMemsetAsync(ctx->d_misses, 0, n * sizeof(uint32_t), ctx->stream));
  ctx->replicated = false;
  return after_accum_write(ctx);
}

int rtc_read_accum(rtc_ctx* ctx, double* rgb_sum, uint32_t* samples, uint32_t* misses) {
  if (!ctx) return RTC_ERR_INVALID;
  if (!ctx->d_rgb) return fail(ctx, RTC_ERR_STATE, "no accumulation buffer (rtc_set_params)");
  cudaSetDevice(ctx->device);
  size_t n = (size_t)ctx->acc_w * ctx->acc_h;
  if (rgb_sum) CU(cudaMemcpyAsync(rgb_sum, ctx->d_rgb, n * 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (samples) CU(cudaMemcpyAsync(samples, ctx->d_samples, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  if (misses) CU(cudaMemcpyAsync(misses, ctx->d_misses, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  drain_timing(ctx);
  return RTC_OK;
}

int rtc_write_accum(rtc_ctx* ctx, const double* rgb_sum, const uint32_t* samples, const uint32_t* misses) {
  if (!ctx) return RTC_ERR_INVALID;
  if (!rgb_sum || !samples || !misses) return fail(ctx, RTC_ERR_INVALID, "accumulation planes must not be null");
  if (!ctx->d_rgb) return fail(ctx, RTC_ERR_STATE, "no accumulation buffer (rtc_set_params)");
  cudaSetDevice(ctx->device);
  size_t n = (size_t)ctx->acc_w * ctx->acc_h;
  int rcw = before_accum_write(ctx);
  if (rcw) return rcw;
  CU(cudaMemcpyAsync(ctx->d_rgb, rgb_sum, n * 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->d_samples, samples, n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->d_misses, misses, n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->replicated = false;  // the planes are this rank's own contribution again
  return after_accum_write(ctx);
}

int rtc_accum_device_ptrs(rtc_ctx* ctx, void** rgb_sum, void** samples, void** misses) {
  if (!ctx) return RTC_ERR_INVALID;
  if (!ctx->d_rgb) return fail(ctx, RTC_ERR_STATE, "no accumulation buffer (rtc_set_params)");
  if (rgb_sum) *rgb_sum = ctx->d_rgb;
  if (samples) *samples = ctx->d_samples;
  if (misses) *misses = ctx->d_misses;
  return RTC_OK;
}

int rtc_tonemap_argb(rtc_ctx* ctx, double exposure, const double back_rgb[3], double back_a, uint32_t* argb) {
  if (!ctx) return RTC_ERR_INVALID;
  if (!back_rgb || !argb) return fail(ctx, RTC_ERR_INVALID, "back_rgb/argb must not be null");
  if (!ctx->d_rgb) return fail(ctx, RTC_ERR_STATE, "no accumulation buffer (rtc_set_params)");
  cudaSetDevice(ctx->device);
  std::lock_guard<std::mutex> ui(ctx->ui_mutex);
  const size_t n = (size_t)ctx->acc_w * ctx->acc_h;
  if (ctx->argb_cap < n) {  // persistent image buffers, sized once per image size
    if (ctx->ui_stream) CU(cudaStreamSynchronize(ctx->ui_stream));
    free_dev_t(ctx->d_argb);
    if (ctx->h_argb) cudaFreeHost(ctx->h_argb);
    ctx->h_argb = nullptr;
    ctx->argb_cap = 0;
    CU(cudaMalloc((void**)&ctx->d_argb, sizeof(uint32_t) * n));
    CU(cudaMallocHost((void**)&ctx->h_argb, sizeof(uint32_t) * n));
    ctx->argb_cap = n;
  }
  int rc = ui_begin(ctx);
  if (rc) return rc;
  CU(launch_tonemap(ctx->ui_stream, (int32_t)n, ctx->d_rgb, ctx->d_samples, ctx->d_misses, exposure, back_rgb[0], back_rgb[1], back_rgb[2], back_a,
                    ctx->d_argb));
  CU(cudaMemcpyAsync(ctx->h_argb, ctx->d_argb, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, ctx->ui_stream));
  rc = ui_end(ctx);
  if (rc) return rc;
  std::memcpy(argb, ctx->h_argb, sizeof(uint32_t) * n);
  return RTC_OK;
}

int rtc_read_pixel(rtc_ctx* ctx, int32_t x, int32_t y, double rgb_sum[3], uint32_t* samples, uint32_t* misses) {
  if (!ctx) return RTC_ERR_INVALID;
  if (!rgb_sum || !samples || !misses) return fail(ctx, RTC_ERR_INVALID, "rgb_sum/samples/misses must not be null");
  if (!ctx->d_rgb) return fail(ctx, RTC_ERR_STATE, "no accumulation buffer (rtc_set_params)");
  if (x < 0 || y < 0 || x >= ctx->acc_w || y >= ctx->acc_h) return fail(ctx, RTC_ERR_INVALID, "pixel outside the image");
  cudaSetDevice(ctx->device);
  std::lock_guard<std::mutex> ui(ctx->ui_mutex);
  int rc = ui_begin(ctx);
  if (rc) return rc;
  const size_t i = (size_t)y * ctx->acc_w + x;
  uint32_t* hu = reinterpret_cast<uint32_t*>(ctx->h_pixel + 3);
  CU(cudaMemcpyAsync(ctx->h_pixel, ctx->d_rgb + i * 3, 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->ui_stream));
  CU(cudaMemcpyAsync(hu, ctx->d_samples + i, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->ui_stream));
  CU(cudaMemcpyAsync(hu + 1, ctx->d_misses + i, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->ui_stream));
  rc = ui_end(ctx);
  if (rc) return rc;
  rgb_sum[0] = ctx->h_pixel[0];
  rgb_sum[1] = ctx->h_pixel[1];
  rgb_sum[2] = ctx->h_pixel[2];
  *samples = hu[0];
  *misses = hu[1];
  return RTC_OK;
}

int rtc_render_samples(rtc_ctx* ctx, uint32_t sample, double* out_rgb) {
  if (!ctx) return RTC_ERR_INVALID;
  if (!out_rgb) return fail(ctx, RTC_ERR_INVALID, "out_rgb must not be null");
  int rc = ready(ctx, true);
  if (rc) return rc;
  cudaSetDevice(ctx->device);
  size_t n = (size_t)ctx->par.width * ctx->par.height;
  rc = ensure_scratch(ctx, n * 3 * sizeof(double));
  if (rc) return rc;
  double* d = (double*)ctx->d_scratch;
  rc = ctx->precision == RTC_F64 ? render_rect<double>(ctx, 0, 0, ctx->par.width, ctx->par.height, sample, 1, false, d)
                                 : render_rect<float>(ctx, 0, 0, ctx->par.width, ctx->par.height, sample, 1, false, d);
  cudaError_t e = cudaSuccess;
  if (rc == RTC_OK) e = cudaMemcpyAsync(out_rgb, d, n * 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (rc) return rc;
  if (e != cudaSuccess) return fail(ctx, RTC_ERR_CUDA, std::string("render_samples: ") + cudaGetErrorString(e));
  return RTC_OK;
}

int rtc_debug_trace(rtc_ctx* ctx, int32_t x, int32_t y, uint32_t sample, int32_t capacity, rtc_debug_ray* out, int32_t* n_out) {
  if (!ctx) return RTC_ERR_INVALID;
  if (!out || !n_out || capacity <= 0) return fail(ctx, RTC_ERR_INVALID, "out/n must not be null and capacity positive");
  int rc = ready(ctx, true);
  if (rc) return rc;
  if (x < 0 || y < 0 || x >= ctx->par.width || y >= ctx->par.height) return fail(ctx, RTC_ERR_INVALID, "pixel outside the image");
  cudaSetDevice(ctx->device);
  rc = wait_shading_upload(ctx);
  if (rc) return rc;
  rc = ensure_pool(ctx, 1024);
  if (rc) return rc;
  if (!ctx->d_dbg_type) {
    CU(cudaMalloc((void**)&ctx->d_dbg_type, sizeof(int32_t) * 32));
    CU(cudaMalloc(&ctx->d_dbg_fresnel, sizeof(double) * 32));
  }
  // Replay the one path bounce by bounce: a band of one pixel, with recursion clamped to i so that every prefix
  // of the path ends in a recorded state. Cheap (<= recursion+1 single-thread wavefronts) and exact.
  const int full = ctx->par.recursion;
  *n_out = 0;
  rtc_params saved = ctx->par;
  std::vector<rtc_hit> hits(1);
  for (int i = 0; i <= full && i < capacity; i++) {
    Band b;
    b.x0 = x; b.x1 = x + 1; b.y0 = y; b.y1 = y + 1;
    b.first_sample = sample; b.n_samples = 1; b.n_pix = 1; b.n_paths = 1;
    // run bounces 0..i with the real recursion limit, but stop after bounce i
    LaunchCfg cfg{ctx->stream, ctx->sm_count, false};
    int32_t type = 0;
    double fres = 0;
    float fres32 = 0;
    uint32_t alive = 0;
    auto run = [&](auto tag) -> int {
      using R = decltype(tag);
      SceneView<R> sv = scene_view<R>(ctx);
      PathView<R> pv = path_view<R>(ctx);
      pv.dbg_type = ctx->d_dbg_type;
      pv.dbg_fresnel = (R*)ctx->d_dbg_fresnel;
      CU(Kernels<R>::raygen(cfg, camera_view<R>(ctx->cam), params_view<R>(ctx->par), b, pv));
      for (int k = 0; k <= i; k++) {
        CU(Kernels<R>::trace(cfg, sv, pv, k & 1, k == 0));
        CU(Kernels<R>::shade(cfg, sv, params_view<R>(ctx->par), b, pv, k & 1, k, k == 0));
        if (k < i) CU(Kernels<R>::compact(cfg, pv, k & 1, k == 0));
      }
      if (!ctx->d_hits) {
        CU(cudaMalloc((void**)&ctx->d_hits, sizeof(rtc_hit) * 1024));
      }
      CU(Kernels<R>::export_hits(cfg, sv, 1, pv, ctx->d_hits, false));
      CU(cudaMemcpyAsync(hits.data(), ctx->d_hits, sizeof(rtc_hit), cudaMemcpyDeviceToHost, ctx->stream));
      CU(cudaMemcpyAsync(&type, ctx->d_dbg_type, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
      if (sizeof(R) == 8)
        CU(cudaMemcpyAsync(&fres, ctx->d_dbg_fresnel, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
      else
        CU(cudaMemcpyAsync(&fres32, ctx->d_dbg_fresnel, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
      // k_shade appended the path to the next bounce's queue iff it is still alive
      CU(cudaMemcpyAsync(&alive, &ctx->d_ctl->count[(i & 1) ^ 1], sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
      CU(cudaStreamSynchronize(ctx->stream));
      if (sizeof(R) == 4) fres = (double)fres32;
      return RTC_OK;
    };
    rc = ctx->precision == RTC_F64 ? run(double()) : run(float());
    if (rc) {
      ctx->par = saved;
      return rc;
    }
    out[i].hit = hits[0];
    out[i].type = type;
    out[i].pad = 0;
    out[i].fresnel_ratio = fres;
    *n_out = i + 1;
    if (alive == 0) break;  // the path ended at bounce i
  }
  ctx->par = saved;
  return RTC_OK;
}

int rtc_debug_raycast(rtc_ctx* ctx, int32_t mode, int32_t* out) {
  if (!ctx) return RTC_ERR_INVALID;
  if (!out) return fail(ctx, RTC_ERR_INVALID, "out must not be null");
  if (mode != RTC_OVERLAY_PRIMITIVES && mode != RTC_OVERLAY_BOUNDING_VOLUMES) return fail(ctx, RTC_ERR_INVALID, "unknown overlay mode");
  int rc = ready(ctx, true);
  if (rc) return rc;
  cudaSetDevice(ctx->device);
  rc = wait_shading_upload(ctx);
  if (rc) return rc;
  const int w = ctx->par.width, h = ctx->par.height;
  const size_t n = (size_t)w * h;
  if (mode == RTC_OVERLAY_BOUNDING_VOLUMES && ensure_host_nodes(ctx)) return RTC_ERR_CUDA;
  if (mode == RTC_OVERLAY_BOUNDING_VOLUMES && ctx->nodes.empty())
    return fail(ctx, RTC_ERR_STATE, "no host-side BVH (the scene came from rtc_upload_baked)");
  // persistent scratch: [w*h ids | the reference-shaped tree for the box-count mode]
  const size_t ids_bytes = (n * sizeof(int32_t) + 255) & ~(size_t)255;
  rc = ensure_scratch(ctx, ids_bytes + (mode == RTC_OVERLAY_BOUNDING_VOLUMES ? ctx->nodes.size() * sizeof(rtc_bvh_node) : 0));
  if (rc) return rc;
  int32_t* d_out = (int32_t*)ctx->d_scratch;
  cudaError_t e = cudaSuccess;
  rtc_bvh_node* d_nodes = (rtc_bvh_node*)((char*)ctx->d_scratch + ids_bytes);
  if (mode == RTC_OVERLAY_BOUNDING_VOLUMES) {
    e = cudaMemcpyAsync(d_nodes, ctx->nodes.data(), ctx->nodes.size() * sizeof(rtc_bvh_node), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = launch_overlay_boxcount(ctx->stream, d_nodes, ctx->root, camera_view<double>(ctx->cam), w, h, d_out);
  } else {
    auto run = [&](auto tag) -> int {
      using R = decltype(tag);
      const int64_t cap = std::max<int64_t>(ctx->max_paths, w);
      int rc2 = ensure_pool(ctx, std::min<int64_t>(cap, (int64_t)w * h));
      if (rc2) return rc2;
      LaunchCfg cfg{ctx->stream, ctx->sm_count, false};
      SceneView<R> sv = scene_view<R>(ctx);
      PathView<R> pv = path_view<R>(ctx);
      const int64_t rows = std::max<int64_t>(1, std::min<int64_t>(h, cap / w));
      for (int ya = 0; ya < h; ya += (int)rows) {
        Band b;
        b.x0 = 0; b.x1 = w; b.y0 = ya; b.y1 = (int)std::min<int64_t>(h, ya + rows);
        b.first_sample = 0; b.n_samples = 1;
        b.n_pix = (uint32_t)(w * (b.y1 - b.y0));
        b.n_paths = b.n_pix;
        CU(Kernels<R>::overlay_rays(cfg, camera_view<R>(ctx->cam), params_view<R>(ctx->par), b, pv));
        CU(Kernels<R>::trace(cfg, sv, pv, 0, true));
        CU(Kernels<R>::overlay_prims(cfg, sv, params_view<R>(ctx->par), b, pv, d_out));
      }
      return RTC_OK;
    };
    rc = ctx->precision == RTC_F64 ? run(double()) : run(float());
  }
  if (rc == RTC_OK && e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  else cudaStreamSynchronize(ctx->stream);
  if (rc) return rc;
  if (e != cudaSuccess) return fail(ctx, RTC_ERR_CUDA, std::string("debug_raycast: ") + cudaGetErrorString(e));
  return RTC_OK;
}

int rtc_debug_raycast_selection(rtc_ctx* ctx, int32_t n_sel, const int32_t* prim_ids, int32_t* out) {
  if (!ctx) return RTC_ERR_INVALID;
  if (!out || n_sel < 0 || (n_sel > 0 && !prim_ids)) return fail(ctx, RTC_ERR_INVALID, "prim_ids/out must not be null");
  if (!ctx->camera_set || !ctx->params_set) return fail(ctx, RTC_ERR_STATE, "camera and params must be set");
  if (!ctx->scene_set || (int32_t)ctx->kind.size() != ctx->n_prims)
    return fail(ctx, RTC_ERR_STATE, "no host-side scene description (the scene came from rtc_upload_baked / rtc_bcast_scene)");
  const size_t npix = (size_t)ctx->par.width * ctx->par.height;
  for (int32_t i = 0; i < n_sel; i++)
    if (prim_ids[i] < 0 || prim_ids[i] >= ctx->n_prims) return fail(ctx, RTC_ERR_INVALID, "selected primitive out of range");
  if (n_sel == 0) {
    std::fill(out, out + npix, -1);
    return RTC_OK;
  }
  // The selected primitives as a scene of their own, prepared on the device (a few milliseconds even for a whole mesh) and
  // raycast like DisplayMode.Primitives: the nearest hit among the selection alone, whatever hides it in the full scene.
  const size_t k = (size_t)n_sel;
  std::vector<uint8_t> kind(k), flags(k);
  std::vector<int32_t> xform(k);
  std::vector<double> geom(k * RTC_GEOM_STRIDE), material(k * RTC_MATERIAL_STRIDE);
  for (size_t i = 0; i < k; i++) {
    const size_t p = (size_t)prim_ids[i];
    kind[i] = ctx->kind[p];
    flags[i] = ctx->flags[p];
    xform[i] = ctx->xform[p];
    std::memcpy(&geom[i * RTC_GEOM_STRIDE], &ctx->geom[p * RTC_GEOM_STRIDE], RTC_GEOM_STRIDE * sizeof(double));
    std::memcpy(&material[i * RTC_MATERIAL_STRIDE], &ctx->material[p * RTC_MATERIAL_STRIDE], RTC_MATERIAL_STRIDE * sizeof(double));
  }
  rtc_scene_desc d;
  d.n_prims = n_sel;
  d.n_xforms = ctx->n_xforms;
  d.kind = kind.data();
  d.flags = flags.data();
  d.geom = geom.data();
  d.xform = xform.data();
  d.xforms = ctx->xforms.empty() ? nullptr : ctx->xforms.data();
  d.material = material.data();
  rtc_ctx* sub = nullptr;
  int rc = rtc_create(ctx->device, ctx->precision, &sub);
  if (rc) return fail(ctx, rc, std::string("selection context: ") + rtc_last_error(nullptr));
  if ((rc = rtc_upload_scene(sub, &d)) == RTC_OK && (rc = rtc_prepare_device(sub, RTC_BUILDER_SAH, 0, nullptr)) == RTC_OK &&
      (rc = rtc_set_params(sub, &ctx->par)) == RTC_OK && (rc = rtc_set_camera(sub, &ctx->cam)) == RTC_OK)
    rc = rtc_debug_raycast(sub, RTC_OVERLAY_PRIMITIVES, out);
  if (rc) fail(ctx, rc, std::string("selection overlay: ") + rtc_last_error(sub));
  rtc_destroy(sub);
  if (rc) return rc;
  for (size_t i = 0; i < npix; i++)
    if (out[i] >= 0) out[i] = prim_ids[out[i]];  // the sub-scene numbers its primitives in list order
  return RTC_OK;
}

int rtc_get_stats(rtc_ctx* ctx, rtc_stats* stats) {
  if (!ctx || !stats) return RTC_ERR_INVALID;
  cudaSetDevice(ctx->device);
  Control h;
  int rc = fetch_device_counters(ctx, &h);
  if (rc) return rc;
  drain_timing(ctx);
  *stats = ctx->stats;
  stats->rays += h.rays;
  stats->nodes_visited = h.nodes_visited;
  stats->prims_tested = h.prims_tested;
  stats->node_steps = h.node_steps;
  stats->leaf_steps = h.leaf_steps;
  return RTC_OK;
}

int rtc_reset_stats(rtc_ctx* ctx) {
  if (!ctx) return RTC_ERR_INVALID;
  cudaSetDevice(ctx->device);
  CU(cudaStreamSynchronize(ctx->stream));
  drain_timing(ctx);
  std::memset(&ctx->stats, 0, sizeof(ctx->stats));
  if (ctx->d_ctl) CU(cudaMemset(ctx->d_ctl, 0, 2 * sizeof(Control)));
  return RTC_OK;
}

// ---- multi-GPU ---------------------------------------------------------------------------------------
int rtc_comm_unique_id(void* id128) {
  if (!id128) return RTC_ERR_INVALID;
  if (!g_nccl.load(g_create_error)) return RTC_ERR_NCCL;
  NcclId id;
  int r = g_nccl.GetUniqueId(&id);
  if (r != 0) {
    g_create_error = "ncclGetUniqueId failed";
    return RTC_ERR_NCCL;
  }
  std::memcpy(id128, &id, sizeof(id));
  return RTC_OK;
}

int rtc_comm_init(rtc_ctx* ctx, int32_t nranks, int32_t rank, const void* id128) {
  if (!ctx) return RTC_ERR_INVALID;
  if (!id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail(ctx, RTC_ERR_INVALID, "bad rank / nranks / id");
  if (!g_nccl.load(ctx->err)) return RTC_ERR_NCCL;
  cudaSetDevice(ctx->device);
  rtc_comm_destroy(ctx);
  NcclId id;
  std::memcpy(&id, id128, sizeof(id));
  int r = g_nccl.CommInitRank(&ctx->nccl_comm, nranks, id, rank);
  if (r != 0) return fail(ctx, RTC_ERR_NCCL, std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error"));
  ctx->nranks = nranks;
  ctx->rank = rank;
  return RTC_OK;
}

int rtc_reduce_accum(rtc_ctx* ctx, int32_t root) {
  if (!ctx) return RTC_ERR_INVALID;
  if (!ctx->nccl_comm) return fail(ctx, RTC_ERR_STATE, "rtc_comm_init has not been called");
  if (!ctx->d_rgb) return fail(ctx, RTC_ERR_STATE, "no accumulation buffer (rtc_set_params)");
  if (root >= ctx->nranks) return fail(ctx, RTC_ERR_INVALID, "root out of range");
  cudaSetDevice(ctx->device);
  const size_t n = (size_t)ctx->acc_w * ctx->acc_h;
  int rcw = before_accum_write(ctx);
  if (rcw) return rcw;
  // The planes of a rank hold what it has rendered and not yet handed over. After an all-reduce they hold the job's total
  // on every rank: all but rank 0 take it out again first, so that the total is contributed exactly once.
  if (ctx->replicated && ctx->rank != 0)
    CU(launch_subtract_planes(ctx->stream, n, ctx->d_rgb, ctx->d_samples, ctx->d_misses, ctx->d_base_rgb, ctx->d_base_samples,
                              ctx->d_base_misses));
  ctx->replicated = false;
  int r = g_nccl.GroupStart();
  if (r == 0) {
    if (root < 0) {
      r = g_nccl.AllReduce(ctx->d_rgb, ctx->d_rgb, n * 3, kNcclFloat64, kNcclSum, ctx->nccl_comm, ctx->stream);
      if (r == 0) r = g_nccl.AllReduce(ctx->d_samples, ctx->d_samples, n, kNcclUint32, kNcclSum, ctx->nccl_comm, ctx->stream);
      if (r == 0) r = g_nccl.AllReduce(ctx->d_misses, ctx->d_misses, n, kNcclUint32, kNcclSum, ctx->nccl_comm, ctx->stream);
    } else {
      r = g_nccl.Reduce(ctx->d_rgb, ctx->d_rgb, n * 3, kNcclFloat64, kNcclSum, root, ctx->nccl_comm, ctx->stream);
      if (r == 0) r = g_nccl.Reduce(ctx->d_samples, ctx->d_samples, n, kNcclUint32, kNcclSum, root, ctx->nccl_comm, ctx->stream);
      if (r == 0) r = g_nccl.Reduce(ctx->d_misses, ctx->d_misses, n, kNcclUint32, kNcclSum, root, ctx->nccl_comm, ctx->stream);
    }
    int r2 = g_nccl.GroupEnd();
    if (r == 0) r = r2;
  }
  if (r != 0) return fail(ctx, RTC_ERR_NCCL, std::string("nccl reduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error"));
  if (root >= 0) {
    if (ctx->rank != root) {  // the contribution has moved to the root, which keeps the running total
      CU(cudaMemsetAsync(ctx->d_rgb, 0, n * 3 * sizeof(double), ctx->stream));
      CU(cudaMemsetAsync(ctx->d_samples, 0, n * sizeof(uint32_t), ctx->stream));
      CU(cudaMemsetAsync(ctx->d_misses, 0, n * sizeof(uint32_t), ctx->stream));
    }
  } else {
    if (ctx->rank != 0) {
      if (!ctx->d_base_rgb) {
        CU(cudaMalloc((void**)&ctx->d_base_rgb, n * 3 * sizeof(double)));
        CU(cudaMalloc((void**)&ctx->d_base_samples, n * sizeof(uint32_t)));
        CU(cudaMalloc((void**)&ctx->d_base_misses, n * sizeof(uint32_t)));
      }
      CU(cudaMemcpyAsync(ctx->d_base_rgb, ctx->d_rgb, n * 3 * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
      CU(cudaMemcpyAsync(ctx->d_base_samples, ctx->d_samples, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
      CU(cudaMemcpyAsync(ctx->d_base_misses, ctx->d_misses, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    ctx->replicated = true;
  }
  return after_accum_write(ctx);
}

// What a receiving rank needs to know about the root's device scene before the segments arrive.
struct SceneHeader {
  int32_t precision, n_prims, n_unbounded, n_xforms, bvh_depth;
  uint32_t root_node;
  uint64_t bytes[rtc_baked::S_COUNT];
};

int rtc_bcast_scene(rtc_ctx* ctx, int32_t root) {
  if (!ctx) return RTC_ERR_INVALID;
  if (!ctx->nccl_comm) return fail(ctx, RTC_ERR_STATE, "rtc_comm_init has not been called");
  if (root < 0 || root >= ctx->nranks) return fail(ctx, RTC_ERR_INVALID, "root out of range");
  cudaSetDevice(ctx->device);
  const bool is_root = ctx->rank == root;
  if (is_root) {
    int rc = ready(ctx, false);
    if (rc) return rc;
    rc = wait_shading_upload(ctx);  // the shading half of a staged upload may still be in flight on the copy stream
    if (rc) return rc;
  }
  auto nccl_fail = [&](int r, const char* what) {
    return fail(ctx, RTC_ERR_NCCL, std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error"));
  };
  // 1. the header: precision, counts and the byte length of every segment
  SceneHeader* h_hdr = nullptr;
  void* d_hdr = nullptr;
  CU(cudaMallocHost((void**)&h_hdr, sizeof(SceneHeader)));
  cudaError_t ce = cudaMalloc(&d_hdr, sizeof(SceneHeader));
  if (ce != cudaSuccess) {
    cudaFreeHost(h_hdr);
    return fail(ctx, RTC_ERR_CUDA, cudaGetErrorString(ce));
  }
  std::memset(h_hdr, 0, sizeof(SceneHeader));
  if (is_root) {
    h_hdr->precision = ctx->precision;
    h_hdr->n_prims = ctx->n_prims;
    h_hdr->n_unbounded = ctx->n_unbounded;
    h_hdr->n_xforms = ctx->n_xforms;
    h_hdr->bvh_depth = ctx->bvh_depth;
    h_hdr->root_node = ctx->root_node;
    for (int i = 0; i < rtc_baked::S_COUNT; i++) h_hdr->bytes[i] = ctx->seg_bytes[i];
    cudaMemcpyAsync(d_hdr, h_hdr, sizeof(SceneHeader), cudaMemcpyHostToDevice, ctx->stream);
  }
  int r = g_nccl.Broadcast(d_hdr, d_hdr, sizeof(SceneHeader), kNcclUint8, root, ctx->nccl_comm, ctx->stream);
  if (r == 0) {
    cudaMemcpyAsync(h_hdr, d_hdr, sizeof(SceneHeader), cudaMemcpyDeviceToHost, ctx->stream);
    ce = cudaStreamSynchronize(ctx->stream);
  }
  const SceneHeader hdr = *h_hdr;
  cudaFreeHost(h_hdr);
  cudaFree(d_hdr);
  if (r != 0) return nccl_fail(r, "ncclBroadcast (scene header)");
  if (ce != cudaSuccess) return fail(ctx, RTC_ERR_CUDA, cudaGetErrorString(ce));
  // every rank returns the same verdict on a mode mismatch (no rank is left waiting in the second collective)
  if (hdr.precision != ctx->precision) return fail(ctx, RTC_ERR_INVALID, "the root's scene was made for the other arithmetic mode");
  // 2. the segments, into (reused) device buffers
  void** dst[rtc_baked::S_COUNT] = {&ctx->d_nodes, &ctx->d_qnodes, (void**)&ctx->d_unbounded, &ctx->d_prims, &ctx->d_mats,
                                    &ctx->d_xforms, (void**)&ctx->d_aux, (void**)&ctx->d_prim_id, (void**)&ctx->d_id_to_slot,
                                    &ctx->d_sgeom};
  if (!is_root) {
    if (ctx->shading_pending) {
      int rc = wait_shading_upload(ctx);
      if (rc) return rc;
    }
    for (int i = 0; i < rtc_baked::S_COUNT; i++) {
      const size_t need = std::max<size_t>((size_t)hdr.bytes[i], 16);
      if (ctx->seg_cap[i] < need || !*dst[i]) {
        if (*dst[i]) {
          CU(cudaStreamSynchronize(ctx->stream));
          cudaFree(*dst[i]);
          *dst[i] = nullptr;
        }
        CU(cudaMalloc(dst[i], need));
        ctx->seg_cap[i] = need;
      }
    }
  }
  r = g_nccl.GroupStart();
  if (r == 0) {
    for (int i = 0; i < rtc_baked::S_COUNT && r == 0; i++)
      if (hdr.bytes[i]) r = g_nccl.Broadcast(*dst[i], *dst[i], (size_t)hdr.bytes[i], kNcclUint8, root, ctx->nccl_comm, ctx->stream);
    int r2 = g_nccl.GroupEnd();
    if (r == 0) r = r2;
  }
  if (r != 0) return nccl_fail(r, "ncclBroadcast (scene segments)");
  if (!is_root) {
    if (hdr.bytes[rtc_baked::S_QNODES] == 0 && hdr.precision == RTC_F32 && ctx->d_qnodes) {  // the kernel keys on a null pointer
      CU(cudaStreamSynchronize(ctx->stream));
      cudaFree(ctx->d_qnodes);
      ctx->d_qnodes = nullptr;
      ctx->seg_cap[rtc_baked::S_QNODES] = 0;
    }
    ctx->n_prims = hdr.n_prims;
    ctx->n_xforms = hdr.n_xforms;
    ctx->n_unbounded = hdr.n_unbounded;
    ctx->bvh_depth = hdr.bvh_depth;
    ctx->root_node = hdr.root_node;
    for (int i = 0; i < rtc_baked::S_COUNT; i++) ctx->seg_bytes[i] = (size_t)hdr.bytes[i];
    // no host-side description behind this scene (as after rtc_upload_baked of a foreign image)
    ctx->kind.clear();
    ctx->flags.clear();
    ctx->geom.clear();
    ctx->material.clear();
    ctx->xform.clear();
    ctx->xforms.clear();
    ctx->nodes.clear();
    ctx->root = -1;
    drop_device_tree(ctx);
    delete ctx->baked;
    ctx->baked = nullptr;
    ctx->scene_set = true;
    ctx->bvh_set = true;
  }
  return RTC_OK;
}

int rtc_comm_destroy(rtc_ctx* ctx) {
  if (!ctx) return RTC_ERR_INVALID;
  if (ctx->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->nccl_comm);
  ctx->nccl_comm = nullptr;
  ctx->nranks = 1;
  ctx->rank = 0;
  return RTC_OK;
}

}  // extern "C"
