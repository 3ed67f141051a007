// Class-aware batched NMS for the test-time consumer of the merged BAGS scores -- EXPERIMENTAL:
// written at the end of round 1, compiled for sm_100a, not yet run on a GPU (BAGS_NMS_NATIVE=1 opts in).
//
// The reference loops over the 1230 foreground classes in Python and calls its NMS op once per class
// (mmdet/core/post_processing/bbox_nms.py:34-54; IoU with "+1" widths and "suppress if IoU > thr" in
// mmdet/ops/nms/src/nms_kernel.cu:13-21,60; candidates pre-sorted by score, :76-78).  Here all classes are
// processed by ONE launch: the caller passes the candidates (score > score_thr) sorted by (class, score descending)
// with per-class segment offsets; one CTA per class builds the suppression bit matrix of its segment in shared
// memory and one warp runs the greedy scan over it.
//
//   boxes   [M, 4] fp32 (x1, y1, x2, y2), candidates in segment order
//   seg_off [S + 1] int32 segment boundaries (segment s = candidates [seg_off[s], seg_off[s+1]))
//   keep    [M] uint8 out: 1 = survives NMS inside its class
// Segments longer than kNmsMaxSeg are rejected by the host wrapper (test-time RoI count is 1000 per image,
// configs/bags/*: rpn max_num = 1000).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bags {

constexpr int kNmsMaxSeg = 1024;
constexpr int kNmsThreads = 256;

__device__ __forceinline__ float iou_plus1(const float4 a, const float4 b) {
  const float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z);
  const float top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
  const float width = fmaxf(right - left + 1.f, 0.f), height = fmaxf(bottom - top + 1.f, 0.f);
  const float inter = width * height;
  const float sa = (a.z - a.x + 1.f) * (a.w - a.y + 1.f);
  const float sb = (b.z - b.x + 1.f) * (b.w - b.y + 1.f);
  return inter / (sa + sb - inter);
}

__global__ void __launch_bounds__(kNmsThreads)
class_nms_kernel(const float4* __restrict__ boxes, const int* __restrict__ seg_off, float iou_thr,
                 uint8_t* __restrict__ keep) {
  extern __shared__ __align__(16) uint8_t nms_smem[];
  const int s0 = seg_off[blockIdx.x], n = seg_off[blockIdx.x + 1] - s0;
  if (n <= 0) return;
  const int words = (n + 31) >> 5;                                  // mask words per row
  float4* sbox = reinterpret_cast<float4*>(nms_smem);                // [n]
  uint32_t* mask = reinterpret_cast<uint32_t*>(sbox + ((n + 1) & ~1));   // [n][words]: bit j of row i = "i suppresses j"
  for (int i = threadIdx.x; i < n; i += kNmsThreads) sbox[i] = boxes[s0 + i];
  __syncthreads();
  // one (row, word) pair per iteration; only j > i matters (the scan visits boxes in score order)
  for (int t = threadIdx.x; t < n * words; t += kNmsThreads) {
    const int i = t / words, w = t - i * words;
    uint32_t bits = 0u;
    if (w >= (i >> 5)) {
      const float4 bi = sbox[i];
      const int j0 = w << 5;
#pragma unroll 4
      for (int b = 0; b < 32; ++b) {
        const int j = j0 + b;
        if (j > i && j < n && iou_plus1(bi, sbox[j]) > iou_thr) bits |= (1u << b);
      }
    }
    mask[t] = bits;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    uint32_t removed = 0u;            // lane w holds bits [32w, 32w+32) of the "already suppressed" set
    for (int i = 0; i < n; ++i) {
      const uint32_t word = __shfl_sync(0xffffffffu, removed, i >> 5);
      const bool alive = ((word >> (i & 31)) & 1u) == 0u;            // warp-uniform
      if (alive) {
        if (lane < words) removed |= mask[i * words + lane];
        if (lane == 0) keep[s0 + i] = 1;
      } else if (lane == 0) {
        keep[s0 + i] = 0;
      }
    }
  }
}

}  // namespace bags
