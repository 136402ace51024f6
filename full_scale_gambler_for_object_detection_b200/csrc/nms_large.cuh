// General-n NMS (n > 8192 boxes): see nms_large.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace fsg {

constexpr int64_t kNmsLargeMax = 262144;  // bit matrix n * ceil(n/64) * 8 B = 8.6 GB at this size

struct NmsLargeWs {
  size_t off_box, off_cls, off_idx, off_mask, total;
};
NmsLargeWs nms_large_ws_layout(int64_t n);
int nms_large(const float* boxes, const float* scores, const int64_t* class_ids, int64_t n, float thr,
              int64_t* keep, int32_t* num_keep, void* workspace, size_t workspace_bytes, cudaStream_t s);

}  // namespace fsg
