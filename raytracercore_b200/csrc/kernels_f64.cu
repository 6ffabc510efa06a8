// kernels_f64.cu — RTC_F64 (parity) instantiation of the wavefront kernels, plus the f64 tonemap kernel.
// Built with -fmad=false: the only fused operations are the explicit fma() calls that mirror the reference's
// Fma.* intrinsics.
#include "rtc_device.cuh"

namespace rtc {
template struct Kernels<double>;

// Util.Clamp via SSE MaxScalar/MinScalar (Util.cs:126-134): NaN -> minimum
__device__ __forceinline__ double clamp_sse(double v, double lo, double hi) {
  double m = v > lo ? v : lo;
  return m < hi ? m : hi;
}
__device__ __forceinline__ uint32_t color_code(double r, double g, double b, double a) {  // SampleSet.cs:47-53
  return ((uint32_t)(int)(clamp_sse(a, 0, 1) * 255) << 24) | ((uint32_t)(int)(clamp_sse(r, 0, 1) * 255) << 16) |
         ((uint32_t)(int)(clamp_sse(g, 0, 1) * 255) << 8) | ((uint32_t)(int)(clamp_sse(b, 0, 1) * 255));
}

// SampleSet.GetOutput over the whole image (SampleSet.cs:55-113, FullRaytracer.GetBitmap FullRaytracer.cs:179-205)
__global__ void k_tonemap(int32_t n, const double* __restrict__ rgb_sum, const uint32_t* __restrict__ samples,
                          const uint32_t* __restrict__ misses, double exposure, double br, double bg, double bb, double ba,
                          uint32_t* __restrict__ argb) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t S = samples[i], M = misses[i];
  if (S == 0) {  // :57-58
    argb[i] = color_code(br * exposure, bg * exposure, bb * exposure, ba);
    return;
  }
  double total = (double)(uint32_t)(S + M);  // :85
  double mult = exposure / (double)S;        // :86
  double r = rgb_sum[(size_t)i * 3] * mult, g = rgb_sum[(size_t)i * 3 + 1] * mult, b = rgb_sum[(size_t)i * 3 + 2] * mult, a = 1;
  double back_alpha_amt = (double)M / total;  // :93
  double back_amt = back_alpha_amt * ba;
  r += (br - r) * back_amt;
  g += (bg - g) * back_amt;
  b += (bb - b) * back_amt;
  a += (ba - a) * back_alpha_amt;
  const double gamma = 1 / 2.2;  // :101
  r = pow(r, gamma);
  g = pow(g, gamma);
  b = pow(b, gamma);
  argb[i] = color_code(r, g, b, a);
}

// BVH<T>.GetIntersectionCount (Acceleration/BVH.cs:352-363): 1 + count(left) + count(right) for every node whose
// Volume.Intersect(ray).far >= 0; leaves count 1. One thread per pixel, explicit stack over the reference-shaped tree.
__global__ void k_overlay_boxcount(const rtc_bvh_node* __restrict__ nodes, int32_t root, CameraView<double> cam, int32_t width,
                                   int32_t height, int32_t* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= width * height) return;
  const int x = i % width, y = i / width;
  V3<double> o, d;
  camera_get_ray(cam, (double)x, (double)y, o, d);
  o = o + (d * cam.image_plane);
  const V3<double> inv = mk3(1.0 / d.x, 1.0 / d.y, 1.0 / d.z);
  int32_t stack[kTraceStack];
  int sp = 0, count = 0;
  if (root >= 0) stack[sp++] = root;
  while (sp > 0) {
    const rtc_bvh_node& nd = nodes[stack[--sp]];
    double nr;
    if (!box_test<double>(nd.bmin[0], nd.bmax[0], nd.bmin[1], nd.bmax[1], nd.bmin[2], nd.bmax[2], o, d, inv, nr)) continue;
    count++;
    if (nd.prim < 0 && sp + 2 <= kTraceStack) {
      stack[sp++] = nd.right;
      stack[sp++] = nd.left;
    }
  }
  out[i] = count;
}

cudaError_t launch_overlay_boxcount(cudaStream_t s, const rtc_bvh_node* nodes, int32_t root, const CameraView<double>& cam,
                                    int32_t width, int32_t height, int32_t* out) {
  int n = width * height;
  if (n <= 0) return cudaSuccess;
  k_overlay_boxcount<<<(n + 127) / 128, 128, 0, s>>>(nodes, root, cam, width, height, out);
  return cudaGetLastError();
}

__global__ void k_subtract_planes(size_t n, double* __restrict__ rgb_sum, uint32_t* __restrict__ samples, uint32_t* __restrict__ misses,
                                  const double* __restrict__ base_rgb, const uint32_t* __restrict__ base_samples,
                                  const uint32_t* __restrict__ base_misses) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  rgb_sum[i * 3] -= base_rgb[i * 3];
  rgb_sum[i * 3 + 1] -= base_rgb[i * 3 + 1];
  rgb_sum[i * 3 + 2] -= base_rgb[i * 3 + 2];
  samples[i] -= base_samples[i];
  misses[i] -= base_misses[i];
}

cudaError_t launch_subtract_planes(cudaStream_t s, size_t n, double* rgb_sum, uint32_t* samples, uint32_t* misses,
                                   const double* base_rgb, const uint32_t* base_samples, const uint32_t* base_misses) {
  if (n == 0) return cudaSuccess;
  k_subtract_planes<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, rgb_sum, samples, misses, base_rgb, base_samples, base_misses);
  return cudaGetLastError();
}

cudaError_t launch_tonemap(cudaStream_t s, int32_t n, const double* rgb_sum, const uint32_t* samples, const uint32_t* misses,
                           double exposure, double br, double bg, double bb, double ba, uint32_t* argb) {
  if (n <= 0) return cudaSuccess;
  k_tonemap<<<(n + 255) / 256, 256, 0, s>>>(n, rgb_sum, samples, misses, exposure, br, bg, bb, ba, argb);
  return cudaGetLastError();
}
}  // namespace rtc
