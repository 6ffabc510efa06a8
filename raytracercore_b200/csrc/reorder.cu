// reorder.cu — queue re-ordering between bounces for scenes beyond the L2.
//
// After a bounce the live queue lists the surviving paths in the order the shading warps compacted them: neighbours in the
// queue left neighbouring pixels, but their new origins (the hit points) and directions are scattered over the scene. While the
// tree and the primitive records fit the 126 MB L2 that costs little; for a 10 M-triangle scene (1.3 GB) every warp of the
// next trace launch then pulls 32 unrelated root-to-leaf paths through DRAM. Sorting the queue by the Morton cell of the ray
// origin (optionally refined by the direction octant) makes the lanes of a warp, and the warps resident together, walk the
// same part of the tree. Only the queue of path indices is permuted: path state stays in place, results are per path and
// the accumulation is ordered per pixel, so the image is bit-identical with and without it (tests/test_gpu_render.py).
//
// Library use: cub::DeviceRadixSort (16-bit keys: two digit passes); the key kernel is hand-written.
#include <cub/cub.cuh>

#include "rtc_internal.h"

namespace rtc {
namespace {

__device__ __forceinline__ uint32_t spread3(uint32_t v, int bits) {  // bit i -> bit 3 i
  uint32_t r = 0;
  for (int i = 0; i < bits; i++) r |= ((v >> i) & 1u) << (3 * i);
  return r;
}

// key of queue entry i: Morton cell of the ray origin on the root node's grid (mode 1: 5 bits per axis; mode 2: 4 bits per
// axis, then the direction octant); entries beyond the live count sort behind everything (0x8000)
__global__ void k_reorder_keys(const CNode* __restrict__ root, const V4<float>* __restrict__ hpos, const V4<float>* __restrict__ dir,
                               const uint32_t* __restrict__ queue, const uint32_t* __restrict__ count, uint32_t n, int mode,
                               uint32_t* __restrict__ keys) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t key = 0x8000u;
  if (i < *count) {
    const uint32_t path = queue[i];
    const V4<float> o = hpos[path];
    const uint32_t em = root->e_imask;
    // the root grid spans 256 steps of 2^(e - 127) from its origin; position in [0, 1) along each axis
    const float sx = __uint_as_float((254u - 8u - (em & 0xFFu)) << 23), sy = __uint_as_float((254u - 8u - ((em >> 8) & 0xFFu)) << 23),
                sz = __uint_as_float((254u - 8u - ((em >> 16) & 0xFFu)) << 23);
    const int bits = mode == 2 ? 4 : 5;
    const float scale = (float)(1 << bits);
    const int top = (1 << bits) - 1;
    const int qx = min(max((int)((o.x - root->px) * sx * scale), 0), top);
    const int qy = min(max((int)((o.y - root->py) * sy * scale), 0), top);
    const int qz = min(max((int)((o.z - root->pz) * sz * scale), 0), top);
    key = (spread3((uint32_t)qx, bits) << 2) | (spread3((uint32_t)qy, bits) << 1) | spread3((uint32_t)qz, bits);
    if (mode == 2) {
      const V4<float> d = dir[path];
      key = (key << 3) | (d.x < 0 ? 1u : 0u) | (d.y < 0 ? 2u : 0u) | (d.z < 0 ? 4u : 0u);
    }
  }
  keys[i] = key;
}

}  // namespace

size_t reorder_temp_bytes(uint32_t n) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                  (int)n, 0, 16);
  return bytes;
}

cudaError_t launch_reorder(cudaStream_t stream, const CNode* root, const V4<float>* hpos, const V4<float>* dir, const uint32_t* queue_in,
                           const uint32_t* count, uint32_t n, int mode, uint32_t* keys_in, uint32_t* keys_out, uint32_t* queue_out, void* tmp,
                           size_t tmp_bytes) {
  if (n == 0) return cudaSuccess;
  k_reorder_keys<<<(n + 255) / 256, 256, 0, stream>>>(root, hpos, dir, queue_in, count, n, mode, keys_in);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  return cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, (const uint32_t*)keys_in, keys_out, queue_in, queue_out, (int)n, 0, 16, stream);
}

}  // namespace rtc
