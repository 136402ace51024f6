// Shared by the translation units that run the per-image NMS kernel (nms_image.cu defines it; detect_select.cu and
// rpn_select.cu launch it through launch_nms_image).
#pragma once
#include <math.h>

#include "common.cuh"

namespace fsg {

constexpr int kMaxLevels = 8;
constexpr int kNmsThreads = 1024;
constexpr int kNmsBigSeg = 512;                     // class segments above this size are suppressed by the whole CTA
constexpr int kNmsCap = 8192;                       // candidates per image the NMS kernel holds

// ------------------------------------------------------------------------------------------
// NMS: one CTA per image (or per stand-alone call)
// ------------------------------------------------------------------------------------------
struct NmsArgs {
  const float4* boxes;      // per image: slots_per_image entries
  const float* scores;
  const int64_t* classes;   // may be NULL (single class)
  int64_t slots_per_image;
  const int* lvl_count;     // (N, L) or NULL -> fixed_count
  int L;
  int topk;                 // slot stride per level
  int fixed_count;
  float thr;                // largest float <= the double threshold (strict > compare, see fsg_nms)
  int max_out;              // truncate to this many (DETECTIONS_PER_IMAGE); <= 0: all
  int sorted_runs;          // the L levels of an image are runs sorted by (score descending, position ascending):
                            // nms_runs_kernel ranks by merging instead of sorting (needs lvl_count and classes)
  int split;                // CTAs per image; CTA c owns the classes with class % split == c
  int part_cap;             // survivors each CTA hands to the merge (max_out, or the candidate count)
  uint64_t* part_keys;      // (N, split, part_cap) scratch
  int* part_cnt;            // (N, split)
  unsigned* done;           // (N)
  unsigned* alive;          // (N, kNmsCap / 32) survivor bits by merged rank, zero-initialised (sorted_runs only)
  uint16_t* rank2cand;      // (N, kNmsCap) candidate (concatenation index) of every merged rank (sorted_runs only)
  // outputs
  int64_t* keep;            // (N, keep_stride) candidate indices in concatenation order
  int64_t keep_stride;
  int32_t* num_keep;        // (N)
  float4* out_boxes;        // (N, max_out) or NULL
  float* out_scores;
  int64_t* out_classes;
  const float4* post;       // (N) [scale_x, scale_y, clip_w, clip_h] or NULL: detector_postprocess fused in
  // optional compact export of the candidates
  float4* exp_boxes;        // (N, L*topk)
  float* exp_scores;
  int64_t* exp_classes;
  int32_t* exp_count;
};

// Boxes.scale (boxes.py:205-210: fp32 * fp32(scale)) then Boxes.clip (boxes.py:122-136: clamp(min=0, max=size));
// pp = [scale_x, scale_y, clip_w, clip_h]
inline int nms_split_for(int N) {
  int s = 1;
  while (s * 2 * N <= 148 && s < 8) s <<= 1;   // fill the 148 SMs: one CTA per SM, up to 8 per image
  return s;
}
struct NmsWs {
  size_t off_done, off_alive, off_cnt, off_keys, off_r2c, total;   // [off_done, off_cnt) is zeroed before every call
};
inline NmsWs nms_ws_layout(int N, int split, int part_cap) {
  NmsWs w;
  size_t o = 0;
  w.off_done = o;  o += align_up(sizeof(unsigned) * (size_t)N, 16);
  w.off_alive = o; o += align_up(sizeof(unsigned) * (size_t)N * (kNmsCap / 32), 16);
  w.off_cnt = o;   o += align_up(sizeof(int) * (size_t)N * split, 16);
  w.off_keys = o;  o += align_up(sizeof(uint64_t) * (size_t)N * split * part_cap, 16);
  w.off_r2c = o;   o += align_up(sizeof(uint16_t) * (size_t)N * kNmsCap, 16);
  w.total = o;
  return w;
}

inline float threshold_floor(double thr) {
  // fp32 IoU `ovr > (double)thr`  <=>  `ovr > f` with f the largest float <= thr
  float f = (float)thr;
  if ((double)f > thr) f = nextafterf(f, -INFINITY);
  return f;
}

constexpr size_t kNmsSmem = (size_t)kNmsCap * 27;

// enqueue nms_image_kernel for N images (grid: a.split x N); returns an fsg_status
// pdl: launch under programmatic dependent launch (the kernel waits on the device for its predecessor)
int launch_nms_image(const NmsArgs& a, int N, cudaStream_t s, bool pdl = false);

}  // namespace fsg
