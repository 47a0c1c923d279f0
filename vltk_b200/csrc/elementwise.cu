// HBM-bound layout / pooling kernels of the extraction path (all coalesced along the
// channel dimension of NHWC activations, 8- or 16-byte vector accesses).
#include "kernels.cuh"

namespace vltk {

// ------------------------------------------------------------------------------------------
// K1a: model input [N,3,H,W] f32 (reference contract, frcnn.py:1924) -> NHWC with C padded to 4
template <typename TO>
__global__ void nchw3_to_nhwc4_kernel(const float* __restrict__ x, TO* __restrict__ y, int N, int HW) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)N * HW) return;
  int n = (int)(i / HW);
  int p = (int)(i - (int64_t)n * HW);
  const float* b = x + (int64_t)n * 3 * HW + p;
  store4(y + i * 4, make_float4(b[0], b[HW], b[2 * (int64_t)HW], 0.f));
}

int nchw3_to_nhwc4(const float* x, void* y, DType dt, int N, int H, int W, cudaStream_t st) {
  int64_t tot = (int64_t)N * H * W;
  if (tot == 0) return 0;
  unsigned grid = (unsigned)ceil_div64(tot, 256);
  if (dt == DT_F32) nchw3_to_nhwc4_kernel<float><<<grid, 256, 0, st>>>(x, (float*)y, N, H * W);
  else nchw3_to_nhwc4_kernel<bf16><<<grid, 256, 0, st>>>(x, (bf16*)y, N, H * W);
  VLTK_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// K2a (bf16 mode): im2col of the 7x7 s2 p3 stem (frcnn.py:860-868) so that the 3-channel conv becomes
// a plain [M,192] x [64,192]^T tensor-core GEMM.  Row m = output pixel; k = (kh*7+kw)*3 + c for
// k < 147, zero for k in [147,192).  One thread per 16-byte chunk (8 consecutive k) of a row.
__global__ void stem_im2col_kernel(const float* __restrict__ x, bf16* __restrict__ a, int H, int W, int OH, int OW) {
  // grid (ceil(OW*24/256), OH, N): no runtime-divisor or 64-bit index arithmetic per thread
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= OW * 24) return;
  const int ow = t / 24, chunk = t - ow * 24;
  const int oh = blockIdx.y, n = blockIdx.z;
  const float* xb = x + (int64_t)n * H * W * 4;
  const int ih0 = oh * 2 - 3, iw0 = ow * 2 - 3;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = chunk * 8 + j;
    float f = 0.f;
    if (k < 147) {
      const int tap = k / 3, c = k - tap * 3;
      const int kh = tap / 7, kw = tap - kh * 7;
      const int ih = ih0 + kh, iw = iw0 + kw;
      if (ih >= 0 && ih < H && iw >= 0 && iw < W) f = xb[((int64_t)ih * W + iw) * 4 + c];
    }
    v[j] = f;
  }
  uint4 o;
  __nv_bfloat162* ob = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
  for (int j = 0; j < 4; ++j) ob[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
  const int64_t m = ((int64_t)n * OH + oh) * OW + ow;
  *reinterpret_cast<uint4*>(a + m * 192 + chunk * 8) = o;
}

// Same A matrix straight from the NCHW f32 image (no NHWC4 intermediate): one CTA = 64 consecutive output pixels of
// one output row.  The 7 input rows x 3 channels x 133 columns it needs are loaded once (coalesced along W, zero
// outside the image), rounded to bf16 into shared memory, and every thread then assembles 16-byte chunks of A rows
// (8 consecutive k = (kh*7+kw)*3+c) from shared memory through a constant offset table.
constexpr int IM_OW = 64, IM_COLS = IM_OW * 2 + 5 + 3;   // 136: 133 columns + padding
__constant__ short c_stem_lut[192];                     // k -> (kh*3+c)*IM_COLS + kw, -1 for the zero tail k >= 147

__global__ void __launch_bounds__(256)
stem_im2col_nchw_kernel(const float* __restrict__ x, bf16* __restrict__ a, int H, int W, int OH, int OW) {
  __shared__ bf16 s[21 * IM_COLS + 8];
  __shared__ short lut[192];             // per-thread-divergent lookups: shared memory, not the constant cache
  if (threadIdx.x < 192) lut[threadIdx.x] = c_stem_lut[threadIdx.x];
  const int ow0 = blockIdx.x * IM_OW, oh = blockIdx.y, n = blockIdx.z;
  const int ih0 = oh * 2 - 3, iw0 = ow0 * 2 - 3;
  const float* xb = x + (int64_t)n * 3 * H * W;
  for (int e = threadIdx.x; e < 21 * IM_COLS; e += 256) {
    const int rc = e / IM_COLS, col = e - rc * IM_COLS;   // rc = kh*3 + c
    const int kh = rc / 3, c = rc - kh * 3;
    const int ih = ih0 + kh, iw = iw0 + col;
    float v = 0.f;
    if (ih >= 0 && ih < H && iw >= 0 && iw < W) v = xb[((int64_t)c * H + ih) * W + iw];
    s[e] = __float2bfloat16_rn(v);
  }
  if (threadIdx.x < 8) s[21 * IM_COLS + threadIdx.x] = __float2bfloat16_rn(0.f);
  __syncthreads();
  const int nout = min(IM_OW, OW - ow0);
  const int64_t m0 = ((int64_t)n * OH + oh) * OW + ow0;
  const unsigned short* su = reinterpret_cast<const unsigned short*>(s);
  for (int t = threadIdx.x; t < nout * 24; t += 256) {
    const int owl = t / 24, chunk = t - owl * 24;
    unsigned int w4[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int l0 = lut[chunk * 8 + 2 * j], l1 = lut[chunk * 8 + 2 * j + 1];
      const unsigned int v0 = l0 >= 0 ? su[l0 + 2 * owl] : 0u, v1 = l1 >= 0 ? su[l1 + 2 * owl] : 0u;
      w4[j] = v0 | (v1 << 16);
    }
    *reinterpret_cast<uint4*>(a + (m0 + owl) * 192 + chunk * 8) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
  }
}

int stem_im2col_nchw(const float* x_nchw, void* a, int N, int H, int W, int OH, int OW, cudaStream_t st) {
  if ((int64_t)N * OH * OW == 0) return 0;
  VLTK_CHECK(OH <= 65535 && N <= 65535, "stem_im2col: image too tall / batch too large for the grid");
  static DeviceOnce once;
  if (once.first()) {
    short lut[192];
    for (int k = 0; k < 192; ++k) {
      if (k >= 147) { lut[k] = -1; continue; }
      const int tap = k / 3, c = k - tap * 3, kh = tap / 7, kw = tap - kh * 7;
      lut[k] = (short)((kh * 3 + c) * IM_COLS + kw);
    }
    VLTK_CUDA(cudaMemcpyToSymbol(c_stem_lut, lut, sizeof(lut)));
  }
  dim3 grid(ceil_div(OW, IM_OW), OH, N);
  stem_im2col_nchw_kernel<<<grid, 256, 0, st>>>(x_nchw, (bf16*)a, H, W, OH, OW);
  VLTK_LAUNCH_CHECK();
  return 0;
}

int stem_im2col(const float* x_nhwc4, void* a, int N, int H, int W, int OH, int OW, cudaStream_t st) {
  if ((int64_t)N * OH * OW == 0) return 0;
  VLTK_CHECK(OH <= 65535 && N <= 65535, "stem_im2col: image too tall / batch too large for the grid");
  dim3 grid(ceil_div(OW * 24, 256), OH, N);
  stem_im2col_kernel<<<grid, 256, 0, st>>>(x_nhwc4, (bf16*)a, H, W, OH, OW);
  VLTK_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// K2b: max_pool2d(k=3, s=2, p=0, ceil_mode=True) (frcnn.py:875-876) on NHWC, 4 channels/thread
template <typename T>
__global__ void maxpool3x3s2_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W,
                                    int C, int OH, int OW) {
  // grid (ceil(OW*C/4 / 256), OH, N)
  const int c4 = C / 4;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= OW * c4) return;
  const int ow = t / c4, c = (t - ow * c4) * 4;
  const int oh = blockIdx.y, n = blockIdx.z;
  float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  const int h0 = oh * 2, w0 = ow * 2;
#pragma unroll
  for (int dh = 0; dh < 3; ++dh) {
    const int h = h0 + dh;
    if (h >= H) break;
#pragma unroll
    for (int dw = 0; dw < 3; ++dw) {
      const int w = w0 + dw;
      if (w >= W) break;
      float4 v = load4(x + (((int64_t)n * H + h) * W + w) * C + c);
      m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
    }
  }
  store4(y + (((int64_t)n * OH + oh) * OW + ow) * C + c, m);
}

int maxpool3x3s2_ceil(const void* x, void* y, DType dt, int N, int H, int W, int C, int OH, int OW,
                      cudaStream_t st) {
  VLTK_CHECK(C % 4 == 0, "maxpool: C=%d must be a multiple of 4", C);
  if ((int64_t)N * OH * OW == 0) return 0;
  VLTK_CHECK(OH <= 65535 && N <= 65535, "maxpool: image too tall / batch too large for the grid");
  dim3 grid(ceil_div(OW * (C / 4), 256), OH, N);
  if (dt == DT_F32) maxpool3x3s2_kernel<float><<<grid, 256, 0, st>>>((const float*)x, (float*)y, N, H, W, C, OH, OW);
  else maxpool3x3s2_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)x, (bf16*)y, N, H, W, C, OH, OW);
  VLTK_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// K12: mean over the 14x14 positions of each ROI (frcnn.py:1401): [R, P, C] -> [R, C] f32
template <typename T>
__global__ void mean_rows_kernel(const T* __restrict__ x, float* __restrict__ y, int R, int P, int C) {
  const int c4 = C / 4;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)R * c4) return;
  int c = (int)(i % c4) * 4;
  int r = (int)(i / c4);
  const T* b = x + (int64_t)r * P * C + c;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int p = 0; p < P; ++p) {
    float4 v = load4(b + (int64_t)p * C);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  // torch.mean = sum / count (true division)
  store4(y + (int64_t)r * C + c, make_float4(s.x / (float)P, s.y / (float)P, s.z / (float)P, s.w / (float)P));
}

int mean_rows(const void* x, float* y, DType dt, int R, int P, int C, cudaStream_t st) {
  VLTK_CHECK(C % 4 == 0, "mean_rows: C=%d must be a multiple of 4", C);
  int64_t tot = (int64_t)R * (C / 4);
  if (tot == 0) return 0;
  unsigned grid = (unsigned)ceil_div64(tot, 128);
  if (dt == DT_F32) mean_rows_kernel<float><<<grid, 128, 0, st>>>((const float*)x, y, R, P, C);
  else mean_rows_kernel<bf16><<<grid, 128, 0, st>>>((const bf16*)x, y, R, P, C);
  VLTK_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// f32 -> T cast of a dense buffer (predictor inputs / test plumbing)
template <typename T>
__global__ void cast_kernel(const float* __restrict__ x, T* __restrict__ y, int64_t n4) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n4) store4(y + i * 4, load4(x + i * 4));
}

int cast_f32(const float* x, void* y, DType dt, int64_t n, cudaStream_t st) {
  VLTK_CHECK(n % 4 == 0, "cast: n must be a multiple of 4");
  if (n == 0) return 0;
  unsigned grid = (unsigned)ceil_div64(n / 4, 256);
  if (dt == DT_F32) cast_kernel<float><<<grid, 256, 0, st>>>(x, (float*)y, n / 4);
  else cast_kernel<bf16><<<grid, 256, 0, st>>>(x, (bf16*)y, n / 4);
  VLTK_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// K1: fused Preprocess (legacy/processing.py:112-150): raw BGR u8 [h,w,3] -> bilinear resize
// (align_corners=False, no antialias; ATen upsample_bilinear2d index math) -> (x-mean)/std ->
// zero-pad to the batch max.  Writes the reference-contract NCHW f32 batch and/or the
// engine's NHWC4 layout in the same pass.  One thread per output pixel of the padded canvas.
template <typename TO>
__global__ void preprocess_kernel(const uint8_t* __restrict__ raw, int rh, int rw, int nh, int nw,
                                  int Hm, int Wm, float3 mean, float3 stdv,
                                  float pad_value, float* __restrict__ out_nchw,
                                  TO* __restrict__ out_nhwc4) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)Hm * Wm) return;
  int oy = (int)(i / Wm), ox = (int)(i - (int64_t)oy * Wm);
  float3 v = make_float3(pad_value, pad_value, pad_value);
  if (oy < nh && ox < nw) {
    float b, g, r;
    if (nh == rh && nw == rw) {  // F.interpolate to the same size is the identity
      const uint8_t* p = raw + ((int64_t)oy * rw + ox) * 3;
      b = (float)p[0]; g = (float)p[1]; r = (float)p[2];
    } else {
      const float sh = (float)rh / (float)nh, sw = (float)rw / (float)nw;
      float fy = __fsub_rn(__fmul_rn(sh, (float)oy + 0.5f), 0.5f);
      float fx = __fsub_rn(__fmul_rn(sw, (float)ox + 0.5f), 0.5f);
      fy = fy < 0.f ? 0.f : fy;
      fx = fx < 0.f ? 0.f : fx;
      int y0 = (int)fy, x0 = (int)fx;
      int y1 = y0 + (y0 < rh - 1 ? 1 : 0), x1 = x0 + (x0 < rw - 1 ? 1 : 0);
      float ly1 = fy - (float)y0, lx1 = fx - (float)x0;
      float ly0 = 1.f - ly1, lx0 = 1.f - lx1;
      const uint8_t* p00 = raw + ((int64_t)y0 * rw + x0) * 3;
      const uint8_t* p01 = raw + ((int64_t)y0 * rw + x1) * 3;
      const uint8_t* p10 = raw + ((int64_t)y1 * rw + x0) * 3;
      const uint8_t* p11 = raw + ((int64_t)y1 * rw + x1) * 3;
      auto lerp2 = [&](int c) {
        float top = __fadd_rn(__fmul_rn(lx0, (float)p00[c]), __fmul_rn(lx1, (float)p01[c]));
        float bot = __fadd_rn(__fmul_rn(lx0, (float)p10[c]), __fmul_rn(lx1, (float)p11[c]));
        return __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot));
      };
      b = lerp2(0); g = lerp2(1); r = lerp2(2);
    }
    v.x = (b - mean.x) / stdv.x;
    v.y = (g - mean.y) / stdv.y;
    v.z = (r - mean.z) / stdv.z;
  }
  if (out_nchw) {
    int64_t hw = (int64_t)Hm * Wm;
    out_nchw[i] = v.x;
    out_nchw[hw + i] = v.y;
    out_nchw[2 * hw + i] = v.z;
  }
  if (out_nhwc4) store4(out_nhwc4 + i * 4, make_float4(v.x, v.y, v.z, 0.f));
}

int preprocess_image(const uint8_t* raw, int rh, int rw, int nh, int nw, int Hm, int Wm,
                     const float* mean, const float* stdv, float pad_value, float* out_nchw,
                     void* out_nhwc4, DType dt, cudaStream_t st) {
  int64_t tot = (int64_t)Hm * Wm;
  if (tot == 0) return 0;
  unsigned grid = (unsigned)ceil_div64(tot, 256);
  float3 m = make_float3(mean[0], mean[1], mean[2]);
  float3 s = make_float3(stdv[0], stdv[1], stdv[2]);
  if (dt == DT_F32)
    preprocess_kernel<float><<<grid, 256, 0, st>>>(raw, rh, rw, nh, nw, Hm, Wm, m, s, pad_value, out_nchw, (float*)out_nhwc4);
  else
    preprocess_kernel<bf16><<<grid, 256, 0, st>>>(raw, rh, rw, nh, nw, Hm, Wm, m, s, pad_value, out_nchw, (bf16*)out_nhwc4);
  VLTK_LAUNCH_CHECK();
  return 0;
}

}  // namespace vltk
