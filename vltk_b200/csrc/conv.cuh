// Convolution / GEMM launch interface shared by the SIMT fp32 path (conv_simt.cu) and the
// tcgen05 path (conv_tc.cu).  Every dense layer of the detector is one of these:
//   y[m, n] = act( (sum_k A[m, k] * W[n, k]) * scale[n] + shift[n] + residual[m, n] )
// with m = output pixel (NHWC row), k = (kh, kw, cin), n = output channel.
// Reference layers: vltk/modeling/frcnn.py:794-822 (Conv2d+BN+ReLU), :963-979 (bottleneck),
// :1561-1572 (RPN head), :1726-1740 (predictor linears).
#pragma once
#include "common.cuh"

namespace vltk {

// DT_H2 ("split fp16", the exact_tc mode): an fp32 value x stored as two fp16 planes, x = hi + lo' * 2^-11 with
// hi = fp16(x), lo' = fp16((x - hi) * 2^11) (|x - (hi + lo' 2^-11)| <= 2^-24 |x|).  A pixel row of C channels is
// 2C halfs: [hi(0..C) | lo'(0..C)], so `ld` counts halfs and is 2C for a dense tensor.
enum DType { DT_F32 = 0, DT_BF16 = 1, DT_H2 = 2 };

struct ConvProblem {
  // activations, NHWC
  const void* x;     // [N, H, W, Cin], channel stride 1, pixel stride ldx
  int ldx;           // elements between consecutive pixels of x (>= Cin)
  void* y;           // [N, OH, OW, Cout], pixel stride ldy
  int ldy;
  const void* residual;  // nullptr or [N, OH, OW, Cout] with pixel stride ldr (same dtype as y)
  int ldr;
  int N, H, W, Cin;
  int OH, OW, Cout;
  int KH, KW, stride, pad, dil;
  // epilogue
  const float* scale;  // [Cout] or nullptr (== 1)
  const float* shift;  // [Cout] or nullptr (== 0)
  int relu;
  DType in_dtype, out_dtype;
};

// weights for the SIMT kernel: fp32 [K_pad][ldw], K = KH*KW*Cin (cin fastest), ldw = round_up(Cout,4)
int conv_simt_launch(const ConvProblem& p, const float* w_kn, int ldw, cudaStream_t st);

}  // namespace vltk
