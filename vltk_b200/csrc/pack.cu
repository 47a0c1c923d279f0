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

__global__ void split_f32_kernel(const float* __restrict__ x, const float* __restrict__ add, int relu,
                                 bf16* __restrict__ hi, bf16* __restrict__ lo, int64_t n4) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 v = load4(x + i * 4);
  if (add) {
    float4 a = load4(add + i * 4);
    v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
  }
  if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
  float4 h = make_float4(__bfloat162float(__float2bfloat16_rn(v.x)), __bfloat162float(__float2bfloat16_rn(v.y)),
                         __bfloat162float(__float2bfloat16_rn(v.z)), __bfloat162float(__float2bfloat16_rn(v.w)));
  store4(hi + i * 4, h);  // exact: h is already bf16-representable
  store4(lo + i * 4, make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w));
}

}  // namespace

int split_f32(const float* x, const float* add, int relu, bf16* hi, bf16* lo, int64_t n, cudaStream_t st) {
  VLTK_CHECK(n % 4 == 0, "split_f32: n must be a multiple of 4");
  if (n == 0) return 0;
  split_f32_kernel<<<(unsigned)ceil_div64(n / 4, 256), 256, 0, st>>>(x, add, relu, hi, lo, n / 4);
  VLTK_LAUNCH_CHECK();
  return 0;
}

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
