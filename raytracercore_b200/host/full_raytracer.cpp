// full_raytracer.cpp — FullRaytracer-shaped progressive renderer over the kernel ABI.
//
// Mirrors the public surface of the reference's Raytracing/FullRaytracer.cs that MainWindow / RayInspector call
// (ctor :66, Start :243, Stop :409, Pause :377, Resume :403, IsRunning/IsPaused/IsStopping :375/:401/:416,
// GetSampleSet :131, GetBitmap :179, Exposure :36). The worker threads, tile cursor and ConcurrentQueue of the
// reference (:271-344) are replaced by GPU sample passes over the whole image; the 100 ms status loop becomes one
// status callback per pass.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>

#include "../../include/rtcore_host.h"
#include "scene.h"

using namespace rtcore;

namespace {

std::string FormatTimeSpan(double seconds) {  // Util.FormatTimeSpan, Util.cs:20-32
  long long ms_total = (long long)(seconds * 1000.0);
  long long days = ms_total / 86400000LL;
  int hours = (int)((ms_total / 3600000LL) % 24), minutes = (int)((ms_total / 60000LL) % 60), secs = (int)((ms_total / 1000LL) % 60),
      ms = (int)(ms_total % 1000LL);
  std::string result;
  char buf[64];
  if (days > 0) {
    std::snprintf(buf, sizeof(buf), "%lld days ", days);
    result += buf;
  }
  if (!result.empty() || hours > 0) {
    std::snprintf(buf, sizeof(buf), "%d:", hours);
    result += buf;
  }
  if (!result.empty() || minutes > 0) {
    std::snprintf(buf, sizeof(buf), "%02d:", minutes);
    result += buf;
  }
  std::snprintf(buf, sizeof(buf), "%02d.%03d", secs, ms);
  return result + buf;
}

std::string GroupDigits(unsigned long long v) {  // {value:N0}
  std::string s = std::to_string(v), out;
  int c = 0;
  for (size_t i = s.size(); i-- > 0;) {
    out.insert(out.begin(), s[i]);
    if (++c % 3 == 0 && i > 0) out.insert(out.begin(), ',');
  }
  return out;
}

}  // namespace

struct rtcs_raytracer {
  std::shared_ptr<Scene> scene;  // kept alive for as long as the raytracer exists
  rtc_ctx* ctx = nullptr;
  uint64_t seed = 0;
  rtcs_status_fn status = nullptr;
  void* user = nullptr;
  std::atomic<double> Exposure{1};
  // Running / Stopping change only under state_mutex: Start's "wait until the previous run is over, then claim the
  // renderer and clear the stop request" (FullRaytracer.cs:245-250) is one critical section, so two Starts cannot both
  // proceed and a Stop that arrives after Start has claimed the renderer is never lost. Readers (IsRunning ...) are lock-free.
  std::mutex state_mutex;
  std::condition_variable state_cv;
  std::atomic<bool> Stopping{false}, Running{false}, Paused{false};
  std::mutex ctx_mutex;  // render-side calls of the kernel ABI are single-caller per handle (the read-out calls are exempt)
  std::mutex pause_mutex;
  std::condition_variable pause_cv;
  std::mutex err_mutex;
  std::string err;
  std::atomic<bool> have_image{false};

  void UpdateStatus(const std::string& text, double progress) {  // FullRaytracer.cs:91-94
    if (status) status(user, text.c_str(), progress);
  }
  void SetError(const std::string& m) {
    std::lock_guard<std::mutex> g(err_mutex);
    err = m;
  }
  void Finish() {  // the end of Start(): `Running = false` (:373), waking a waiting Start / destroy
    {
      std::lock_guard<std::mutex> g(state_mutex);
      Running = false;
    }
    state_cv.notify_all();
  }
  int Fail(int rc, const char* what = nullptr) {
    const std::string m = what ? std::string(what) : std::string(rtc_last_error(ctx));
    SetError(m);
    UpdateStatus("Error: " + m, 0);  // the reference would surface a .NET exception on the render thread
    Finish();
    return rc;
  }
};

extern "C" {

rtcs_raytracer* rtcs_raytracer_create(rtcs_scene* scene, int32_t device, int32_t precision, uint64_t seed,
                                      rtcs_status_fn status, void* user, char* err, int32_t err_cap) {
  auto set_err = [&](const char* m) {
    if (err && err_cap > 0) std::snprintf(err, (size_t)err_cap, "%s", m);
  };
  set_err("");
  if (!scene) {
    set_err("scene is null");
    return nullptr;
  }
  rtc_ctx* ctx = nullptr;
  int rc = rtc_create(device, precision, &ctx);
  if (rc != RTC_OK) {
    set_err(rtc_last_error(nullptr));
    return nullptr;
  }
  rtcs_raytracer* r = new rtcs_raytracer();
  r->scene = scene->scene;
  r->ctx = ctx;
  r->seed = seed;
  r->status = status;
  r->user = user;
  return r;
}

void rtcs_raytracer_destroy(rtcs_raytracer* r) {
  if (!r) return;
  rtcs_raytracer_stop(r);
  {
    std::unique_lock<std::mutex> lk(r->state_mutex);
    r->state_cv.wait(lk, [&] { return !r->Running.load(); });
  }
  rtc_destroy(r->ctx);
  delete r;
}

int rtcs_raytracer_start(rtcs_raytracer* r, uint32_t samples_per_pass, uint32_t max_samples) {
  if (!r) return RTC_ERR_INVALID;
  if (samples_per_pass == 0) {
    // automatic: passes of about 8 Mi paths -- what keeps the two wavefronts of rtc_render busy while a pass
    // still ends, and Stop / Pause / the status line are served, every few tens of milliseconds
    const double pixels = std::max(1.0, (double)r->scene->Width * (double)r->scene->Height);
    samples_per_pass = (uint32_t)std::min(64.0, std::max(1.0, std::ceil((double)(8u << 20) / pixels)));
  }
  {
    std::unique_lock<std::mutex> lk(r->state_mutex);
    r->state_cv.wait(lk, [&] { return !r->Running.load(); });  // `while (Running) ;` FullRaytracer.cs:245
    r->Running = true;                                         // :249-250, as one step with the wait
    r->Stopping = false;
  }
  Scene& sc = *r->scene;
  r->UpdateStatus("Preparing scene...", 0);  // :252
  int rc;
  {
    std::lock_guard<std::mutex> g(r->ctx_mutex);
    // Scene.Prepare (:253): hand primitives + accelerator to the device
    rc = rtc_upload_scene(r->ctx, &sc.Desc());
    if (rc) return r->Fail(rc);
    if (sc.HasAccelerator()) {  // built on the host earlier (Scene.Accelerator: the inspector asked for it) and cached there
      int root = -1;
      const std::vector<rtc_bvh_node>& nodes = sc.Accelerator(&root);
      rc = rtc_upload_bvh(r->ctx, (int32_t)nodes.size(), nodes.data(), root);
    } else {  // tree and device layout made on the GPU: the same tree, node for node, in a twentieth of the time
      rc = rtc_prepare_device(r->ctx, RTC_BUILDER_SAH, 0, nullptr);
    }
    if (rc) return r->Fail(rc);
    // :255-269 new SampleSet[w,h]; Camera.InitRender(w,h)
    rtc_params par = sc.Params(r->seed);
    rc = rtc_set_params(r->ctx, &par);
    if (rc) return r->Fail(rc);
    rc = rtc_clear_accum(r->ctx);
    if (rc) return r->Fail(rc);
    if (sc.Cameras.empty()) return r->Fail(RTC_ERR_STATE, "scene has no camera");
    rtc_camera cam = sc.Cameras[sc.CurrentCamera].InitRender(sc.Width, sc.Height);
    rc = rtc_set_camera(r->ctx, &cam);
    if (rc) return r->Fail(rc);
    r->have_image = true;
  }
  r->UpdateStatus("Beginning render...", 0);  // :290
  double total_seconds = 0;
  unsigned long long total_passes = 0;
  uint32_t done = 0;
  const double w = sc.Width, h = sc.Height;
  while (!r->Stopping) {
    if (r->Paused) {  // workers park at a pass boundary (:223-224)
      std::unique_lock<std::mutex> lk(r->pause_mutex);
      r->pause_cv.wait(lk, [&] { return !r->Paused || r->Stopping; });
      continue;
    }
    uint32_t n = samples_per_pass;
    if (max_samples > 0) {
      if (done >= max_samples) break;
      n = std::min(n, max_samples - done);
    }
    auto t0 = std::chrono::steady_clock::now();
    {
      std::lock_guard<std::mutex> g(r->ctx_mutex);
      rc = rtc_render(r->ctx, 0, 0, sc.Width, sc.Height, done, n);
      if (rc == RTC_OK) rc = rtc_sync(r->ctx);
    }
    if (rc) return r->Fail(rc);
    total_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    done += n;
    total_passes++;
    // :346-357 status text and progress
    double perPixel = ((double)done * w * h) / (w * h);
    double samplesPerSecond = perPixel / total_seconds;
    double progress = perPixel / (perPixel + 1000);
    char buf[256];
    std::snprintf(buf, sizeof(buf), "Tiles: %s Elapsed: %s %.2f/px %.3f/px/sec", GroupDigits(total_passes).c_str(),
                  FormatTimeSpan(total_seconds).c_str(), perPixel, samplesPerSecond);
    r->UpdateStatus(buf, progress);
  }
  r->Finish();
  return RTC_OK;
}

void rtcs_raytracer_stop(rtcs_raytracer* r) {  // :409-414
  if (!r) return;
  {
    std::lock_guard<std::mutex> g(r->state_mutex);
    r->Stopping = true;
  }
  rtcs_raytracer_resume(r);
}
void rtcs_raytracer_pause(rtcs_raytracer* r) {
  if (r) r->Paused = true;
}
void rtcs_raytracer_resume(rtcs_raytracer* r) {
  if (!r) return;
  {
    std::lock_guard<std::mutex> lk(r->pause_mutex);
    r->Paused = false;
  }
  r->pause_cv.notify_all();
}
int rtcs_raytracer_is_running(rtcs_raytracer* r) { return r && r->Running ? 1 : 0; }
int rtcs_raytracer_is_paused(rtcs_raytracer* r) { return r && r->Paused ? 1 : 0; }
int rtcs_raytracer_is_stopping(rtcs_raytracer* r) { return r && r->Stopping ? 1 : 0; }
void rtcs_raytracer_set_exposure(rtcs_raytracer* r, double exposure) {
  if (r) r->Exposure = exposure;
}

int rtcs_raytracer_get_sample_set(rtcs_raytracer* r, int32_t x, int32_t y, double rgb[3], uint32_t* samples, uint32_t* misses) {
  if (!r || !rgb || !samples || !misses) return RTC_ERR_INVALID;
  rgb[0] = rgb[1] = rgb[2] = 0;
  *samples = *misses = 0;
  if (!r->have_image) return RTC_OK;  // `new SampleSet()` when nothing was rendered yet (:143-144)
  Scene& sc = *r->scene;
  // :137-138 clamps to [0, Width] / [0, Height]; the inclusive upper bound would index out of range in the
  // reference, so the mirror clamps to the last pixel.
  x = std::min(std::max(x, 0), sc.Width - 1);
  y = std::min(std::max(y, 0), sc.Height - 1);
  // one pixel on the read-out stream: no lock against the render loop, no full-image copy (MainWindow.cs:360-370 calls this
  // on every mouse move)
  int rc = rtc_read_pixel(r->ctx, x, y, rgb, samples, misses);
  if (rc) r->SetError(rtc_last_error(r->ctx));
  return rc;
}

int rtcs_raytracer_get_bitmap(rtcs_raytracer* r, uint32_t* argb) {
  if (!r || !argb) return RTC_ERR_INVALID;
  if (!r->have_image) {
    r->SetError("nothing rendered yet");
    return RTC_ERR_STATE;  // GetBitmap returns null when SampleSets == null (:184-185)
  }
  Scene& sc = *r->scene;
  double back[3] = {sc.BackgroundRGB.R, sc.BackgroundRGB.G, sc.BackgroundRGB.B};
  // read-out call: runs beside the render loop without taking its lock (rtcore_b200.h, rtc_tonemap_argb)
  int rc = rtc_tonemap_argb(r->ctx, r->Exposure.load(), back, sc.BackgroundAlpha, argb);
  if (rc) r->SetError(rtc_last_error(r->ctx));
  return rc;
}

rtc_ctx* rtcs_raytracer_ctx(rtcs_raytracer* r) { return r ? r->ctx : nullptr; }
const char* rtcs_raytracer_last_error(rtcs_raytracer* r) {
  if (!r) return "";
  static thread_local std::string copy;  // the caller's own copy: another thread may replace r->err meanwhile
  std::lock_guard<std::mutex> g(r->err_mutex);
  copy = r->err;
  return copy.c_str();
}

}  // extern "C"
