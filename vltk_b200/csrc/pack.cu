// Weight repacking kernels: reference layout [cout][cin][taps] -> GEMM operand layouts.
#include "conv_tc.cuh"

namespace vltk {

namespace {

__global__ void pack_kn_kernel(const float* __restrict__ w, float* __restrict__ out, int cout, int cin,
                               int taps, int ldw, int round_bf16) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t tot = (int64_t)cout * cin * taps;
  if (i >= tot) return;
  int t = (int)(i % taps);
  int64_t r = i / taps;
  int c = (int)(r % cin);
  int o = (int)(r / cin);
  float v = w[i];
  if (round_bf16) v = __bfloat162float(__float2bfloat16_rn(v));
  out[((int64_t)t * cin + c) * ldw + o] = v;
}

__global__ void pack_nk_kernel(const float* __restrict__ w, bf16* __restrict__ out, int cout, int cin, int taps) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t tot = (int64_t)cout * cin * taps;
  if (i >= tot) return;
  int t = (int)(i % taps);
  int64_t r = i / taps;
  int c = (int)(r % cin);
  int o = (int)(r / cin);
  out[(int64_t)o * taps * cin + (int64_t)t * cin + c] = __float2bfloat16_rn(w[i]);
}

__global__ void pad_vector_kernel(const float* __restrict__ src, float* __restrict__ dst, int n, int n_pad, float fill) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pad) dst[i] = (src && i < n) ? src[i] : fill;
}

}  // namespace

int pack_weight_kn(const float* w, float* w_kn, int cout, int cin, int taps, int ldw, bool round_bf16, cudaStream_t st) {
  int64_t tot = (int64_t)cout * cin * taps;
  if (tot == 0) return 0;
  pack_kn_kernel<<<(unsigned)ceil_div64(tot, 256), 256, 0, st>>>(w, w_kn, cout, cin, taps, ldw, round_bf16 ? 1 : 0);
  VLTK_LAUNCH_CHECK();
  return 0;
}

int pack_weight_nk(const float* w, bf16* w_nk, int cout, int cin, int taps, cudaStream_t st) {
  int64_t tot = (int64_t)cout * cin * taps;
  if (tot == 0) return 0;
  pack_nk_kernel<<<(unsigned)ceil_div64(tot, 256), 256, 0, st>>>(w, w_nk, cout, cin, taps);
  VLTK_LAUNCH_CHECK();
  return 0;
}

int pad_vector(const float* src, float* dst, int n, int n_pad, float fill, cudaStream_t st) {
  if (n_pad == 0) return 0;
  pad_vector_kernel<<<ceil_div(n_pad, 256), 256, 0, st>>>(src, dst, n, n_pad, fill);
  VLTK_LAUNCH_CHECK();
  return 0;
}

}  // namespace vltk
