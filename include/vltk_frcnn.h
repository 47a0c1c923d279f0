/*
 * C ABI of libvltk_frcnn.so — the B200-native drop-in for vltk's Faster R-CNN R101-C4
 * Visual-Genome region-feature extraction path.
 *
 * The reference has no native boundary for this path (it is pure Python over torch /
 * torchvision); the entry points below are what a binding of that path
 * replaces.  Each one cites the reference interface it stands in for (paths relative to
 * the reference tree).  INTEGRATION.md shows the ctypes stub a vltk maintainer would add.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success and a
 * negative code on failure (vltk_frcnn_last_error() describes it); nothing throws; calls
 * are stream-ordered on the cudaStream_t passed as `void* stream` (NULL = default stream);
 * a handle is bound to one device and is not thread-safe (the reference caller is a
 * single-threaded loop, vltk/abc/extraction.py:142-199).  There is NO CPU fallback.
 */
#ifndef VLTK_FRCNN_H_
#define VLTK_FRCNN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vltk_frcnn vltk_frcnn_t;

/* Arithmetic modes of the dense layers (activations NHWC in all of them). */
enum {
  VLTK_MODE_FP32 = 0, /* fp32 storage, fp32 FMA on the CUDA cores: index-exact parity mode      */
  VLTK_MODE_BF16 = 1, /* bf16 storage, tcgen05 tensor-core implicit GEMM, fp32 accumulate       */
  VLTK_MODE_EXACT_TC = 2 /* fp32-FAITHFUL on tcgen05: activations stored as two fp16 planes (x = hi + lo*2^-11),
                          * weights as three, 3 kind::f16 passes per K chunk, chunk sums promoted to an fp32
                          * register accumulator (csrc/conv_tcx.cu).  Index-exact parity mode on the tensor
                          * pipe; requires |activation| <= 65504.                                          */
};

/* Architecture + selection knobs.  Mirrors the cfg.* keys FRCNN.__init__ reads
 * (vltk/modeling/frcnn.py:1744-1755 and the constructors it calls). */
typedef struct {
  int stem_out_channels;    /* RESNETS.STEM_OUT_CHANNELS   (64)   */
  int res2_out_channels;    /* RESNETS.RES2_OUT_CHANNELS   (256)  */
  int blocks[3];            /* res2,res3,res4 block counts (3,4,23) */
  int res5_blocks;          /* 3 */
  int num_anchors;          /* len(sizes)*len(ratios)      (15)   */
  int anchor_stride;        /* 16 */
  int rpn_hidden;           /* PROPOSAL_GENERATOR.HIDDEN_CHANNELS (512) */
  float rpn_nms_thresh;     /* RPN.NMS_THRESH              (0.7)  */
  int rpn_pre_nms_topk;     /* RPN.PRE_NMS_TOPK_TEST       (6000, <= 8192) */
  int rpn_post_nms_topk;    /* RPN.POST_NMS_TOPK_TEST      (300,  <= 512)  */
  float rpn_min_size;       /* PROPOSAL_GENERATOR.MIN_SIZE (0)    */
  float rpn_bbox_weights[4];/* RPN.BBOX_REG_WEIGHTS        (1,1,1,1) */
  int pooler_resolution;    /* ROI_BOX_HEAD.POOLER_RESOLUTION (14) */
  int num_classes;          /* ROI_HEADS.NUM_CLASSES       (1600) */
  int num_attrs;            /* ROI_BOX_HEAD.NUM_ATTRS      (400)  */
  float roi_bbox_weights[4];/* ROI_BOX_HEAD.BBOX_REG_WEIGHTS (10,10,5,5) */
  int mode;                 /* VLTK_MODE_* */
} vltk_frcnn_config;

/* Per-call detection knobs == the mutable ROIOutputs attributes callers poke
 * (frcnn.py:1233-1240; tests/frcnn_test.py:16-19). */
typedef struct {
  float nms_thresh[4];      /* roi_outputs.nms_thresh (list, tried in order) */
  int n_nms_thresh;         /* 1..4 */
  int min_detections;
  int max_detections;
  float pad_value;          /* forward(..., pad_value=) */
  /* forward(..., ignorey=) — frcnn.py:328-366: HOST [N, n_ignorey, 2] f32 y-ranges (raw-image coordinates) or
   * NULL.  Honoured only when scales_yx is given too (the reference's own condition); ranges are divided by
   * scales_yx[n][1], RPN proposals spanning a range are dropped and the others clipped to its nearer end.
   * n_ignorey <= 16. */
  const float* ignorey;
  int n_ignorey;
} vltk_frcnn_knobs;

/* Dense, caller-allocated DEVICE outputs of one forward call: the model-dict of
 * FRCNN.inference (frcnn.py:1996-2004) in its padding="max_detections" layout
 * (v1.0.0 contract). */
typedef struct {
  float* boxes;             /* [N, max_det, 4] x1,y1,x2,y2 scaled by scales_yx          */
  float* normalized_boxes;  /* [N, max_det, 4] boxes / (sizes*scales_yx)                 */
  int64_t* obj_ids;         /* [N, max_det]                                              */
  float* obj_probs;         /* [N, max_det]                                              */
  int64_t* attr_ids;        /* [N, max_det]                                              */
  float* attr_probs;        /* [N, max_det]                                              */
  float* roi_features;      /* [N, max_det, 2048]                                        */
  int32_t* preds_per_image; /* [N]                                                       */
  int32_t* keep_idx;        /* [N, max_det] index into the image's proposal list, -1 pad */
} vltk_frcnn_out;

const char* vltk_frcnn_last_error(void);
const char* vltk_frcnn_version(void);

/* FRCNN(cfg) — vltk/modeling/frcnn.py:1744-1755. */
int vltk_frcnn_create(const vltk_frcnn_config* cfg, int device, vltk_frcnn_t** out);
void vltk_frcnn_destroy(vltk_frcnn_t* h);

/* model.load_state_dict(state_dict) — frcnn.py:1881 (keys/shapes: SURVEY.md Appendix C).
 * `data` is a HOST fp32 array of `numel` elements in the reference's own layout
 * ([Cout,Cin,kH,kW] convs, [out,in] linears); BN buffers under "<conv>.norm.*". */
int vltk_frcnn_load_tensor(vltk_frcnn_t* h, const char* name, const float* data, int64_t numel);
/* Folds frozen BN (eps 1e-5) into per-channel scale/shift, repacks weights for the kernels
 * and uploads them.  Must be called once after all tensors are loaded. */
int vltk_frcnn_finalize(vltk_frcnn_t* h);

/* Scratch the engine needs for a batch of N padded HxW images. */
size_t vltk_frcnn_workspace_bytes(vltk_frcnn_t* h, int n, int height, int width);

/* FRCNN.forward(images, image_shapes, scales_yx=..., padding="max_detections",
 * max_detections=...) — frcnn.py:1924-2004.
 *   images_nchw : DEVICE [N,3,H,W] f32, normalised + padded (Preprocess output)
 *   sizes_hw    : HOST   [N,2] int32, resized (h,w) per image (clip + pooling use these)
 *   scales_yx   : HOST   [N,2] f32 raw/resized, or NULL
 *   workspace   : DEVICE scratch of >= vltk_frcnn_workspace_bytes(N,H,W)           */
int vltk_frcnn_forward(vltk_frcnn_t* h, const float* images_nchw, const int32_t* sizes_hw,
                       const float* scales_yx, int n, int height, int width,
                       const vltk_frcnn_knobs* knobs, const vltk_frcnn_out* out,
                       void* workspace, size_t workspace_bytes, void* stream);

/* Preprocess.__call__ for one image — vltk/legacy/processing.py:112-150: raw BGR u8
 * [raw_h, raw_w, 3] (DEVICE) -> bilinear shortest-edge resize to (new_h,new_w) -> (x-mean)/std
 * -> written at batch slot `index` of the zero-padded canvas images_nchw [N,3,H,W] (DEVICE). */
int vltk_frcnn_preprocess(const uint8_t* raw_bgr, int raw_h, int raw_w, int new_h, int new_w,
                          const float mean[3], const float std[3], float pad_value,
                          float* images_nchw, int index, int height, int width, void* stream);

/* ---- stage entry points (teacher-forced parity tests and the config-4 microbenchmarks) ---- */

/* One conv / linear layer on NHWC activations with the fused epilogue
 * y = act((x (*) w) * scale + shift + residual)   — frcnn.py:794-822, 963-979.
 * weight: DEVICE f32 in the reference layout [Cout,Cin,KH,KW]; scale/shift/residual may be NULL.
 * x/y/residual dtype follows `mode` (f32 or bf16); `use_tensor_cores` selects tcgen05 (bf16 only). */
int vltk_conv2d_nhwc(const void* x, const float* weight, const float* scale, const float* shift,
                     const void* residual, void* y, int n, int h, int w, int cin, int cout,
                     int kh, int kw, int stride, int pad, int dil, int relu, int mode,
                     int use_tensor_cores, void* stream);

/* The res5 tail in bf16 mode (frcnn.py:1389, 1401): conv + BN + residual + ReLU on tcgen05 whose output tile
 * is reduced instead of stored — pooled[g, c] = mean over the g-th group of `pool_rows` consecutive output
 * pixels (one ROI = 14*14 = 196 rows).  x/residual bf16 NHWC (DEVICE), weight f32 [Cout,Cin,k,k], pooled f32
 * [N*OH*OW/pool_rows, Cout].  cin % 64 == 0, cout % 256 == 0, pool_rows >= 128 and dividing N*OH*OW. */
int vltk_conv2d_meanpool_nhwc(const void* x, const float* weight, const float* scale, const float* shift,
                              const void* residual, float* pooled, int n, int h, int w, int cin, int cout,
                              int k, int stride, int pad, int dil, int relu, int pool_rows, void* stream);

/* A projection bottleneck's tail in bf16 mode (frcnn.py:918-925, 971-979): conv3 and the 1x1 shortcut as ONE
 * K-concatenated tcgen05 GEMM, y = act(x . w^T + x2[::stride2, ::stride2] . w2^T + shift), frozen-BN scales
 * already folded into w / w2 by the caller.  x [N,h,w,cin], x2 [N,h2,w2,cin2] bf16 NHWC (DEVICE) with
 * (h2-1)/stride2+1 == h (same for w); weight [cout,cin], weight2 [cout,cin2] DEVICE f32 (rounded to bf16);
 * shift [cout] or NULL; y [N,h,w,cout] bf16.  cin, cin2, cout % 64 == 0. */
/* The same fused tail in exact_tc mode (csrc/conv_tcx.cu on CTA pairs, ROI-aligned tiles): x [N,h,w,cin], residual
 * [N,h,w,cout] and weight [Cout,Cin] are DEVICE fp32 (split into fp16 planes inside), a 1x1 convolution;
 * 128 < pool_rows <= 256. */
int vltk_conv2d_meanpool_exact_nhwc(const float* x, const float* weight, const float* scale, const float* shift,
                                    const float* residual, float* pooled, int n, int h, int w, int cin, int cout,
                                    int relu, int pool_rows, void* stream);

int vltk_conv2d_dual_nhwc(const void* x, const float* weight, const void* x2, const float* weight2,
                          const float* shift, void* y, int n, int h, int w, int cin, int h2, int w2,
                          int cin2, int stride2, int cout, int relu, void* stream);

/* Kernel selection for the tcgen05 convolutions (process-wide; A/B measurements and the kernel-equivalence tests).
 * Layers with a 256-wide cout tile, bf16 output and at least `min_pixels` output pixels run on CTA pairs
 * (tcgen05 cta_group::2: two SMs share one W tile); 0 = never.  `residual_layers` = 0 keeps the layers with a
 * shortcut / fused-mean epilogue on the single-CTA kernel; 1 runs them on pairs (4 operand stages + 4-slab shortcut
 * ring); 2 runs the shortcut layers on pairs with 3 stages + a 6-slab ring.  A negative argument leaves that setting
 * unchanged.
 * Defaults: 32768 and 0 (environment: VLTK_CTA2, VLTK_CTA2_RES).  Both kernels accumulate in the same order and
 * produce bit-identical outputs.  Always returns 0. */
int vltk_conv_tc_set_cta_pairs(int min_pixels, int residual_layers);
/* The same switch for the exact_tc kernels (csrc/conv_tcx.cu): layers with a 256-wide cout tile and at least
 * `min_pixels` output pixels run on CTA pairs (five 32 KB operand stages instead of three 48 KB ones); 0 = never,
 * negative = unchanged.  Default 32768 (environment: VLTK_TCX_CTA2).  Bit-identical outputs.  Always returns 0. */
int vltk_conv_tcx_set_cta_pairs(int min_pixels);

/* One PART of the model through the engine's own layers, weights and arithmetic mode, on caller-provided DEVICE fp32
 * tensors (converted to / from the mode's activation type inside): the teacher-forced stage tests feed the oracle's
 * stage input and compare the stage output.
 *   part 0   BasicStem (frcnn.py:872-879): x = images NCHW [n,3,hh,ww] -> y = pooled NHWC [n,Hp,Wp,64]
 *   part 2-4 res2 / res3 / res4 (frcnn.py:963-979, 1076-1090): x NHWC [n,hh,ww,cin] -> y NHWC [n,oh,ow,cout];
 *            only blocks [block_begin, block_end) of the stage run (block_end < 0: to the end)
 *   part 5   RPNHead (frcnn.py:1561-1572): x = res4 NHWC [n,hh,ww,1024] -> y = fp32 rows [n*hh*ww, ld]:
 *            columns [0,4A) anchor deltas (a*4+coord), [4A,5A) objectness
 * out_dims receives {oh, ow, channels (ld for part 5)}; y_cap = capacity of y in floats.  Synchronises. */
int vltk_frcnn_run_part(vltk_frcnn_t* h, int part, int block_begin, int block_end, const float* x, int n, int hh, int ww,
                        float* y, int64_t y_cap, int32_t* out_dims, void* stream);

/* Pipeline trace of the single-CTA tcgen05 kernel (diagnosis only; tools/tc_trace.py).  In a library built with
 * VLTK_TRACE=1 csrc/build.sh, the following conv launches make CTA `cta` append (tag, clock64) records per role
 * (0 TMA producer, 1 MMA issuer, 2 residual producer, 3/4 the two epilogue groups) to dev_buf, a DEVICE array of
 * 5 * cap_per_role * 2 int64; dev_buf = NULL switches tracing off.  The shipped library has the hooks compiled out and
 * returns -1 (vltk_frcnn_last_error says so). */
int vltk_conv_tc_set_trace(void* dev_buf, int cap_per_role, int cta);

/* nn.Linear on the tensor pipe with fp32-faithful arithmetic (frcnn.py:1729-1737 in bf16 mode):
 * y[m,n] = act(x[m,k] . weight[n,k]^T + bias), all DEVICE f32; operands are split into bf16
 * hi+lo planes and accumulated as hi*hi + lo*hi + hi*lo in one fp32 TMEM tile.  k,n % 64 == 0. */
int vltk_linear_tc3(const float* x, const float* weight, const float* bias, float* y, int m, int k,
                    int n, int relu, void* stream);

/* find_top_rpn_proposals + RPN.inference — frcnn.py:264-390, 1615-1638 — on the RPN head's
 * NCHW outputs (DEVICE): logits [N,A,H4,W4], deltas [N,4A,H4,W4]; cell anchors [A,4] (HOST).
 * Outputs (DEVICE): proposals [N,post,4], proposal_logits [N,post], counts [N]. */
int vltk_rpn_proposals(const float* logits_nchw, const float* deltas_nchw, const float* cell_anchors,
                       const int32_t* sizes_hw, int n, int a, int h4, int w4, int stride,
                       int pre_topk, int post_topk, float nms_thresh, float min_size,
                       const float weights[4], float* proposals, float* proposal_logits,
                       int32_t* counts, void* stream);

/* torchvision.ops.nms on DEVICE boxes [K,4] / scores [K] (frcnn.py:132, 383):
 * keep [max_keep] int32 indices into boxes (score order), count [1]. */
int vltk_nms(const float* boxes, const float* scores, int k, float thresh, int max_keep,
             int32_t* keep, int32_t* count, void* stream);

/* torchvision.ops.RoIPool(P, scale) (frcnn.py:1179,1198) on an NCHW f32 DEVICE map [N,C,H,W]
 * with rois [R,5]=(batch,x1,y1,x2,y2) (DEVICE); out [R,C,P,P] f32 (DEVICE). */
int vltk_roi_pool_nchw(const float* feat, int n, int c, int h, int w, const float* rois, int r,
                       int p, float scale, float* out, void* stream);

/* ROIOutputs.inference — frcnn.py:1262-1294 — from predictor outputs (all DEVICE f32):
 * obj_logits [N*R, C+1], attr_logits [N*R, A+1], box_deltas [N*R, 4C], feats [N*R, D],
 * proposals [N,R,4], counts [N]; sizes/scales HOST as in vltk_frcnn_forward. */
int vltk_roi_outputs(const float* obj_logits, const float* attr_logits, const float* box_deltas,
                     const float* feats, const float* proposals, const int32_t* counts,
                     const int32_t* sizes_hw, const float* scales_yx, int n, int r,
                     int num_classes, int num_attrs, int d, const float weights[4],
                     const vltk_frcnn_knobs* knobs, const vltk_frcnn_out* out, void* stream);

/* Debug taps: after a forward call, copies an intermediate to HOST memory (tests only).
 * name in {"res4" [N,H4,W4,1024], "rpn_head" [N,H4*W4,80], "proposals" [N,post,4],
 * "proposal_count" [N] (int32 payload), "topk_anchor_idx" [N,K] / "proposal_pos" [N,post] (int32
 * payload: anchor index of each sorted candidate / position of each kept proposal in that list), "feats" [N*post,2048], "cls_logits", "attr_logits",
 * "bbox_deltas"}; returns the number of floats written, or <0.  bf16 taps are widened. */
int64_t vltk_frcnn_debug_read(vltk_frcnn_t* h, const char* name, float* host_dst, int64_t capacity);

/* Number of kernels the engine launched since creation (bench.py's gpu_launches). */
int64_t vltk_frcnn_launch_count(vltk_frcnn_t* h);

/* Per-launch CUDA-event timing of the dense kernels (bench.py's roofline leg).  While enabled,
 * every conv/GEMM launch is bracketed by two events on the launch stream.  profile_read
 * synchronises, then fills agg[6] = {tcgen05 ms, tcgen05 FLOPs, tcgen05 launches, SIMT ms,
 * SIMT FLOPs, SIMT launches} since the last read — a kind's ms is the UNION of its launches' event intervals
 * (the two backbone half-batch streams overlap; overlapped time is counted once) — optionally writes one CSV
 * line per launch ("kind,M,K,Cout,ms", each launch's own duration) into csv (NUL-terminated, truncated at
 * cap), and clears the log. */
int vltk_frcnn_profile_enable(vltk_frcnn_t* h, int enable);
int vltk_frcnn_profile_read(vltk_frcnn_t* h, double* agg, char* csv, size_t cap);

/* Reader side (vltk/abc/adapter.py:186-199 `get(img_id)` over `img_to_row_map`, consumed by
 * dataset/visnlangdataset.py:370-405): a whole feature column resident in HBM, batches gathered by row index.
 * out[r, :cols] = table[idx[r], :cols] for r < rows; table [n_rows, ld] f32, idx [rows] int32, out [rows, cols]
 * f32, all DEVICE, 16-byte aligned; cols, ld multiples of 4.  Rows whose index is outside [0, n_rows) are left
 * untouched. */
int vltk_gather_rows_f32(const float* table, int64_t n_rows, int64_t ld, const int32_t* idx, int rows,
                         int cols, float* out, void* stream);

/* ---- JPEG front end: replaces the decode inside `cv2.imread` at vltk/compat.py:573-579
 * (`img_tensorize`, reached from legacy/processing.py:119-129).  The host parses the markers and runs the
 * (inherently serial) Huffman entropy decoder; dequantisation, the inverse DCT, chroma upsampling and colour
 * conversion run on the GPU and reproduce libjpeg-turbo's default pipeline (JDCT_ISLOW, fancy upsampling) bit
 * for bit.  Supported: baseline / extended-sequential (one interleaved scan; entropy decoding on the device) and
 * progressive (entropy decoding of its scans on the host) Huffman JPEG, 8-bit, grayscale or 3 components with 4:4:4,
 * 4:2:2 or 4:2:0 sampling, restart intervals.  Anything else returns -3 (the caller
 * decides what to do with such a file; nothing is decoded on the CPU behind its back). */
typedef struct {
  int width, height, ncomp;
  int comp_id[3], hs[3], vs[3];
  int hmax, vmax, mcus_x, mcus_y;
  int blocks_w[3], blocks_h[3];          /* padded 8x8-block grid of each component                       */
  int comp_w[3], comp_h[3];              /* real (downsampled) sample extent of each component            */
  int64_t coef_offset[3], coef_count;    /* int16 elements: component c starts at coef_offset[c]          */
  int64_t plane_offset[3], plane_bytes;  /* u8 sample planes (row stride blocks_w*8): device scratch      */
  uint16_t qt[3][64];                    /* quantisation tables, natural (row-major) order                */
  int restart_interval;
  int orientation;                       /* EXIF orientation tag (0 = absent); cv2.imread applies it      */
  int progressive;
  int color_transform;                   /* Adobe APP14 transform flag, -1 = no Adobe marker              */
} vltk_jpeg_info;

/* Parses the headers up to the first scan.  0, or -2 (corrupt) / -3 (valid but unsupported). */
int vltk_jpeg_parse(const uint8_t* data, size_t len, vltk_jpeg_info* info);
/* Entropy-decodes the scan into `coef` (HOST, >= info.coef_count int16): per component, blocks in raster order
 * over its padded block grid, 64 coefficients per block in natural order, still quantised. */
int vltk_jpeg_decode_coefficients(const uint8_t* data, size_t len, int16_t* coef, int64_t coef_capacity);
/* The same for n images on up to n_threads host threads (images are independent); status[i] per image. */
int vltk_jpeg_decode_coefficients_batch(int n, const uint8_t* const* datas, const size_t* lens,
                                        int16_t* const* coefs, const int64_t* capacities, int n_threads,
                                        int* status);
/* GPU entropy decoding: the host only parses the markers and removes the byte stuffing; the Huffman decoding of
 * each scan runs on the device (one CTA per image, self-synchronising subsequences — jpeg.cu).
 *   blob_bound   : bytes of pinned host staging needed for n files of the given lengths
 *   prepare_batch: fills infos[n], the upload blob (HOST), coef_offsets[n] / coef_total (int16 elements of one
 *                  batch coefficient buffer, each image 16-byte aligned) and on_gpu[n] (1 = decoded by
 *                  entropy_decode; reserved for stream kinds that must take vltk_jpeg_decode_coefficients instead)
 *   entropy_decode: blob (DEVICE copy, 8-byte aligned) -> coef (DEVICE, zero-filled here); iterations (DEVICE
 *                  int32[n] or NULL) receives the number of synchronisation iterations per image */
size_t vltk_jpeg_gpu_blob_bound(int n, const size_t* lens);
int vltk_jpeg_gpu_prepare_batch(int n, const uint8_t* const* datas, const size_t* lens, vltk_jpeg_info* infos,
                                uint8_t* blob, size_t cap, size_t* used, int64_t* coef_offsets,
                                int64_t* coef_total, int* on_gpu);
int vltk_jpeg_gpu_entropy_decode(int n, const uint8_t* blob, int16_t* coef, int64_t coef_total,
                                 int32_t* iterations, void* stream);
/* coef (DEVICE, 16-byte aligned) -> bgr (DEVICE u8 [height, width, 3], the array cv2.imread returns, EXIF
 * orientation NOT applied); planes = DEVICE scratch of info.plane_bytes.  Stream-ordered, no host sync. */
int vltk_jpeg_reconstruct(const int16_t* coef, const vltk_jpeg_info* info, uint8_t* planes, uint8_t* bgr,
                          void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VLTK_FRCNN_H_ */
