// Launch interface of the non-GEMM kernels (layout, pooling, RPN selection, NMS, ROIPool,
// detection tail).  All launches are stream-ordered and never synchronise with the host.
#pragma once
#include "common.cuh"
#include "conv.cuh"

namespace vltk {

// ---- elementwise.cu -----------------------------------------------------------------------
int nchw3_to_nhwc4(const float* x, void* y, DType dt, int N, int H, int W, cudaStream_t st);
// 7x7 s2 p3 patches of the NHWC4 f32 image -> bf16 rows of 192 (147 + zero pad), bf16 mode stem
int stem_im2col(const float* x_nhwc4, void* a, int N, int H, int W, int OH, int OW, cudaStream_t st);
// the same matrix straight from the NCHW f32 image (shared-memory staged, no NHWC4 intermediate)
int stem_im2col_nchw(const float* x_nchw, void* a, int N, int H, int W, int OH, int OW, cudaStream_t st);
int maxpool3x3s2_ceil(const void* x, void* y, DType dt, int N, int H, int W, int C, int OH, int OW,
                      cudaStream_t st);
int mean_rows(const void* x, float* y, DType dt, int R, int P, int C, cudaStream_t st);
int cast_f32(const float* x, void* y, DType dt, int64_t n, cudaStream_t st);
int preprocess_image(const uint8_t* raw, int rh, int rw, int nh, int nw, int Hm, int Wm,
                     const float* mean, const float* stdv, float pad_value, float* out_nchw,
                     void* out_nhwc4, DType dt, cudaStream_t st);

// ---- h2ops.cu: the same stages on split-fp16 activations (DT_H2, exact_tc mode) -----------------
int maxpool3x3s2_ceil_f32_to_h2(const float* x, void* y, int N, int H, int W, int C, int OH, int OW, cudaStream_t st);
int roi_pool_h2(const void* feat, int N, int H, int W, int C, const float* rois, const int* count, int R, int P,
                float scale, void* out, cudaStream_t st);
int mean_rows_h2(const void* x, float* y, int R, int P, int C, cudaStream_t st);
// out[r] = split(relu?(x[r] + add?[r])) as [rows][hi(C) | lo'(C)]
int split_f32_h2(const float* x, const float* add, int relu, void* out, int64_t rows, int C, cudaStream_t st);
int widen_h2(const void* x, float* y, int64_t rows, int C, cudaStream_t st);

// ---- rpn.cu -------------------------------------------------------------------------------
struct RpnSelectArgs {
  const float* head;      // [N, HW, ldh] f32 RPN head output, one row per res4 pixel:
                          //   [delta_off + a*4 + coord] anchor deltas, [logit_off + a] objectness
  int ldh, delta_off, logit_off;
  int N, H4, W4, A;
  int stride;             // anchor stride (16)
  const float* cell;      // [A,4] cell anchors
  const int* sizes_hw;    // [N,2] device, resized (h,w) per image
  int pre_topk;           // <= 8192
  float min_size;
  float wx, wy, ww, wh;   // RPN bbox weights
  // outputs (sorted by logit desc, ties lower index first)
  float* boxes;           // [N, K, 4] decoded + clipped
  float* scores;          // [N, K]
  int* anchor_idx;        // [N, K] flattened anchor index (y*W+x)*A+a (for tests)
  uint8_t* valid;         // [N, K] non-empty flag
  int K;                  // = min(pre_topk, H4*W4*A)
  // `ignorey` branch (frcnn.py:328-366), active when both pointers are set: J caller-given y-ranges per image,
  // divided by scales_yx[n][1]; proposals spanning a range are dropped, the others clipped to its nearer end.
  const float* ignorey;   // [N, J, 2] device or nullptr
  const float* scales_yx; // [N, 2] device or nullptr
  int J;
};
int rpn_select(const RpnSelectArgs& a, cudaStream_t st);

struct NmsArgs {
  const float* boxes;     // [N, K, 4] in score order
  const float* scores;    // [N, K] (copied to outputs)
  const uint8_t* valid;   // [N, K] or nullptr (all valid)
  int N, K;
  float thresh;
  int max_keep;
  unsigned long long* mask;  // workspace [N, K, ceil(K/64)]
  int* done;                 // workspace [N] or nullptr: enables the exact prefix-first schedule for K > 1024
  // outputs
  float* out_boxes;       // [N, max_keep, 4]  (zero-filled past count)
  float* out_scores;      // [N, max_keep]
  int* out_idx;           // [N, max_keep] positions in the sorted input list (-1 past count)
  int* out_count;         // [N]
};
size_t nms_mask_bytes(int N, int K);
// stable argsort(-scores) of one box list (K <= 8192) + gather; and keep-index remapping
int sort_boxes_desc(const float* boxes, const float* scores, int K, float* sboxes, float* sscores, int* order,
                    cudaStream_t st);
int remap_indices(const int* idx, const int* order, int n, int* out, cudaStream_t st);
int nms_sorted(const NmsArgs& a, cudaStream_t st);

// ---- roipool.cu ---------------------------------------------------------------------------
// torchvision RoIPool(P, scale) on NHWC features; ROI r of image n = rois[n, r]; r >= count[n]
// rows are zero-filled.  out: [N*R, P, P, C]
int roi_pool(const void* feat, DType dt, int N, int H, int W, int C, const float* rois,
             const int* count, int R, int P, float scale, void* out, cudaStream_t st);
// stage-entry variant: explicit batch index per ROI (rois [R,4], bidx [R]); f32 only
int roi_pool_indexed(const void* feat, DType dt, int H, int W, int C, const float* boxes, const int* bidx,
                     int R, int P, float scale, void* out, cudaStream_t st);

// ---- tail.cu ------------------------------------------------------------------------------
struct TailArgs {
  int N, R;                    // images, proposal slots per image
  const float* cls_logits;     // [N*R, ldc]  (num_classes+1 valid)
  int ldc;
  const float* bbox_deltas;    // [N*R, ldb]  (num_classes*4 valid), or nullptr when the weights below are given
  int ldb;
  // bbox_pred evaluated ONLY for each ROI's winning class (4 of its 6400 rows; frcnn.py:1244-1250 computes all,
  // 116-131 keeps one): deltas = W[4c..4c+3, :] . feats + b with W = hi + lo bf16 planes [rows][D] (fp32-faithful)
  const bf16* bbox_w_hi; const bf16* bbox_w_lo; const float* bbox_bias;
  const float* bbox_w_f32;     // exact_tc: W as fp32 [rows][D] instead of the two bf16 planes
  const float* attr_logits;    // [N*R, lda]  (num_attrs+1 valid)
  int lda;
  const float* feats;          // [N*R, D]
  int D;
  const float* proposals;      // [N, R, 4]
  const int* count;            // [N]
  const int* sizes_hw;         // [N,2]
  const float* scales_yx;      // [N,2] or nullptr
  int num_classes, num_attrs;
  float wx, wy, ww, wh;
  const float* nms_thresh;     // host array
  int n_thresh;
  int min_det, max_det;
  float pad_value;
  // outputs, dense [N, max_det, ...]
  float* boxes; float* norm_boxes; long long* obj_ids; float* obj_probs;
  long long* attr_ids; float* attr_probs; float* roi_features; int* preds_per_image;
  int* keep_idx;               // [N, max_det] index into the image's proposal list (-1 pad)
  float* stats;                // scratch [N*R, 8] f32: per-ROI box / prob / attr prob / ids (roi_stats_kernel)
};
int roi_tail(const TailArgs& a, cudaStream_t st);

// per-row argmax over the first `n` columns (first max wins) -> int32
int row_argmax(const float* x, int ld, int rows, int n, int* out, cudaStream_t st);
// out[r, :] = table[idx[r], :]
int gather_rows(const float* table, int ld, const int* idx, int rows, int cols, float* out, int ldo,
                cudaStream_t st);

}  // namespace vltk
