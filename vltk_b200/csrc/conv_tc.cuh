// tcgen05 implicit-GEMM convolution (conv_tc.cu) + weight packing helpers (pack.cu).
#pragma once
#include <cuda.h>

#include <map>
#include <tuple>

#include "conv.cuh"

namespace vltk {

// Host-side cache of encoded TMA descriptors: encoding costs ~1 us of driver work per map and
// the engine replays the same ~220 (activation, weight) maps every forward.
struct TensorMapCache {
  typedef std::tuple<const void*, int, int, int, int, int, int, int, int, int, int> Key;
  std::map<Key, CUtensorMap> maps;
};

// y = act((x (*) w) * scale + shift + residual) on the tensor pipe.
//   x, y, residual: bf16 NHWC;  w_nk: bf16 [cout_pad][KH*KW*Cin] (cin fastest), cout_pad % 64 == 0
//   requires Cin % 64 == 0; scale/shift must hold cout_pad entries.
// Optional split-precision operands: x = x_hi + x_lo, w = w_hi + w_lo (bf16 pairs).  When given, the
// kernel accumulates x_hi*w_hi + x_lo*w_hi + x_hi*w_lo in one TMEM tile (fp32-faithful, ~2^-17 rel).
// With x_lo == nullptr the activations are taken as exact bf16 and only the weights are split: x*w_hi + x*w_lo.
struct TcSplit {
  const void* x_lo = nullptr;   // same geometry as p.x
  const bf16* w_lo = nullptr;   // same geometry as w_nk
};
// Optional fused row-group mean (res5 tail, frcnn.py:1401): rows are ROIs of `rows` consecutive pixels; the
// layer's output tile is NOT stored — out[M/rows][Cout] (f32) receives the mean of each group, computed from
// the fp32 epilogue values.  `partial` is scratch of conv_tc_pool_partial_bytes(M, Cout).  Requires a residual,
// Cout % 256 == 0, K > 256, rows >= 128 (a 128-row tile then spans at most two ROIs).
struct TcPool {
  float* out = nullptr;
  float* partial = nullptr;
  int rows = 0;
};
size_t conv_tc_pool_partial_bytes(int64_t M, int cout);
// Optional second operand pair accumulated into the SAME output tile (K-concatenation of two 1x1 convs):
//   y = act((x (*) w + x2 (*) w2) * scale + shift (+ residual))
// x2 is a bf16 NHWC tensor [N, H2, W2, Cin2] sampled with stride2 (a strided 1x1, pad 0) onto the same
// [N, OH, OW] output grid; w2 is bf16 [cout_pad][Cin2].  This is how a bottleneck's projection shortcut
// (frcnn.py:918-925, 971-976) is folded into its conv3: with both frozen-BN scales pre-multiplied into the
// weights, conv3(t2) + shortcut(x) is ONE GEMM over K = mid + cin and the shortcut tensor is never written
// to or re-read from HBM.  Requires a 1x1 primary conv.
struct TcConcat {
  const void* x2 = nullptr;
  int ldx2 = 0, H2 = 0, W2 = 0, Cin2 = 0, stride2 = 1;
  const bf16* w2 = nullptr;
};
// out_dtype may be DT_F32 (no residual) or DT_BF16.
int conv_tc_launch(const ConvProblem& p, const bf16* w_nk, int cout_pad, TensorMapCache* cache,
                   cudaStream_t st, const TcSplit* split = nullptr, const TcPool* pool = nullptr,
                   const TcConcat* concat = nullptr);

// fp32-faithful convolution on the tensor pipe (conv_tcx.cu): split-fp16 activations (DT_H2) x three fp16 weight
// planes, three kind::f16 passes per K chunk, chunk sums promoted to an fp32 register accumulator.
//   w3: fp16 [cout_pad][3*K] = rows of (WA | WB | WC), built by pack_weight_h3; p.scale must already carry the
//   per-row 2^-s of the packing.  out_dtype DT_H2 (optional DT_H2 residual) or DT_F32 (cout_pad % 128 == 0).
// Optional `concat` (x2 / ldx2 / H2 / W2 / Cin2 / stride2; w2 unused): K-concatenation as in conv_tc_launch — x2 is a
// DT_H2 tensor, and every plane of a w3 row is the concatenation [K primary | Cin2] (planes 3 * (K + Cin2) per row).
// Optional `pool`: the fused row-group mean of conv_tc_launch (needs a residual, Cout % 256 == 0, 128 < rows <= 256; runs on
// CTA pairs with ROI-aligned tiles); p.y is not written.
int conv_tcx_launch(const ConvProblem& p, const void* w3, int cout_pad, TensorMapCache* cache, cudaStream_t st,
                    const TcConcat* concat = nullptr, const TcPool* pool = nullptr);
// out[roi][c] = (partial[2 roi][c] + partial[2 roi + 1][c]) / rows  (second pass of the pooled epilogues)
int tc_pool_finish(const float* partial, float* out, int rois, int rows, int C, cudaStream_t st);
// conv_tcx layers with a 256-wide cout tile and at least `min_pixels` output pixels run on CTA pairs (0 = never)
void conv_tcx_set_cta_pairs(int min_pixels);

// CTA-pair (tcgen05 cta_group::2) dispatch: layers with BN = 256, bf16 output and at least `min_pixels` output pixels
// run on conv_tc3_kernel (0 = never); `residual_layers` = 0 keeps residual / pooled layers on v2.  Negative = unchanged.
void conv_tc_set_cta_pairs(int min_pixels, int residual_layers);

// Diagnosis builds (-DVLTK_TC_TRACE): the next conv_tc2 launches record CTA `cta`'s pipeline events into dev_buf
// (5 roles x cap_per_role x 2 int64).  dev_buf = nullptr switches it off.  Returns -1 in a library built without tracing.
int conv_tc_set_trace(void* dev_buf, int cap_per_role, int cta);

// ---- descriptor encoders (conv_tc.cu), shared with conv_tcx.cu.  128B-swizzled boxes.
int tc_num_sms();
int tc_encode_tiled(CUtensorMap* out, CUtensorMapDataType dt, const void* ptr, uint64_t cols, uint64_t rows,
                    uint64_t row_stride_bytes, uint32_t box_cols, uint32_t box_rows, bool promote256);
// NHWC tensor [N,H,W,C] (pixel stride ld_elems elements of esz bytes) gathered as 128-pixel x 64-channel im2col boxes
int tc_encode_im2col(CUtensorMap* out, CUtensorMapDataType dt, const void* ptr, int N, int H, int W, int C, int ld_elems,
                     int esz, int KH, int KW, int stride, int pad, int dil);

// ---- pack.cu: reference-layout weights [cout][cin][taps] (DEVICE f32) -> kernel layouts
int pack_weight_kn(const float* w, float* w_kn, int cout, int cin, int taps, int ldw, bool round_bf16,
                   cudaStream_t st);
int pack_weight_nk(const float* w, bf16* w_nk, int cout, int cin, int taps, cudaStream_t st);
int pad_vector(const float* src, float* dst, int n, int n_pad, float fill, cudaStream_t st);
// hi = bf16(v), lo = bf16(v - hi) with v = relu?(x + add?) elementwise (add may be nullptr)
int split_f32(const float* x, const float* add, int relu, bf16* hi, bf16* lo, int64_t n, cudaStream_t st);

}  // namespace vltk
