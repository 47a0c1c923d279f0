// Non-GEMM kernels of the exact_tc mode on split-fp16 activations (DT_H2, conv.cuh): an fp32 value x is stored as
// x = hi + lo' * 2^-11 (two fp16 planes per pixel row: [hi(0..C) | lo'(0..C)]).  Each kernel reconstructs fp32 values,
// does the reference's fp32 arithmetic and, where it writes activations, splits again:
//   * stem max-pool (frcnn.py:875-876)    fp32 NHWC in -> DT_H2 out
//   * RoIPool (frcnn.py:1179, 1195-1198)  DT_H2 -> DT_H2; a max SELECTS an input, so the (hi, lo') pair is copied
//   * 14x14 mean (frcnn.py:1401)          DT_H2 -> fp32
//   * predictor inputs (frcnn.py:1729-1737)  fp32 (+ add) (ReLU) -> DT_H2
#include <cuda_fp16.h>

#include "kernels.cuh"

namespace vltk {

namespace {

constexpr float LO_SCALE = 2048.f, LO_INV = 1.f / 2048.f;

__device__ __forceinline__ void split1(float x, __half& hi, __half& lo) {
  const float c = fminf(fmaxf(x, -65504.f), 65504.f);
  hi = __float2half_rn(c);
  lo = __float2half_rn(fminf(fmaxf((x - __half2float(hi)) * LO_SCALE, -65504.f), 65504.f));
}
__device__ __forceinline__ float join1(__half hi, __half lo) { return fmaf(__half2float(lo), LO_INV, __half2float(hi)); }

struct H4 { __half v[4]; };   // 4 halfs = 8 bytes
__device__ __forceinline__ H4 ld_h4(const __half* p) { H4 r; *reinterpret_cast<uint2*>(r.v) = *reinterpret_cast<const uint2*>(p); return r; }
__device__ __forceinline__ void st_h4(__half* p, const H4& r) { *reinterpret_cast<uint2*>(p) = *reinterpret_cast<const uint2*>(r.v); }

__global__ void maxpool_f32_h2_kernel(const float* __restrict__ x, __half* __restrict__ y, int N, int H, int W, int C,
                                      int OH, int OW) {
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
      const float4 v = load4(x + (((int64_t)n * H + h) * W + w) * C + c);
      m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
    }
  }
  H4 hi, lo;
  split1(m.x, hi.v[0], lo.v[0]); split1(m.y, hi.v[1], lo.v[1]); split1(m.z, hi.v[2], lo.v[2]); split1(m.w, hi.v[3], lo.v[3]);
  __half* o = y + (((int64_t)n * OH + oh) * OW + ow) * 2 * C + c;
  st_h4(o, hi);
  st_h4(o + C, lo);
}

// One CTA per (roi, output row ph); threadIdx.x walks 8-channel vectors (16 B per plane), threadIdx.y splits the P bins.
__global__ void __launch_bounds__(256)
roi_pool_h2_kernel(const __half* __restrict__ feat, int H, int W, int C, const float* __restrict__ rois,
                   const int* __restrict__ count, int R, int P, float scale, __half* __restrict__ out) {
  const int cv = C / 8;
  const int ph = blockIdx.x, roi = blockIdx.y;
  const int n = roi / R, r = roi - n * R;
  __half* orow = out + ((int64_t)roi * P + ph) * P * 2 * C;
  const uint4 zero = make_uint4(0, 0, 0, 0);
  if (r >= count[n]) {
    for (int pw = threadIdx.y; pw < P; pw += blockDim.y)
      for (int c = threadIdx.x; c < 2 * cv; c += blockDim.x) *reinterpret_cast<uint4*>(orow + (int64_t)pw * 2 * C + c * 8) = zero;
    return;
  }
  const float4 b = reinterpret_cast<const float4*>(rois)[roi];
  const int sw = (int)roundf(b.x * scale), sh = (int)roundf(b.y * scale);
  const int ew = (int)roundf(b.z * scale), eh = (int)roundf(b.w * scale);
  const int rw = max(ew - sw + 1, 1), rh = max(eh - sh + 1, 1);
  const float bin_h = (float)rh / (float)P, bin_w = (float)rw / (float)P;
  int hs = (int)floorf((float)ph * bin_h) + sh, he = (int)ceilf((float)(ph + 1) * bin_h) + sh;
  hs = min(max(hs, 0), H); he = min(max(he, 0), H);
  const __half* f = feat + (int64_t)n * H * W * 2 * C;
  for (int c = threadIdx.x; c < cv; c += blockDim.x) {
    const __half* fc = f + c * 8;
    for (int pw = threadIdx.y; pw < P; pw += blockDim.y) {
      int ws = (int)floorf((float)pw * bin_w) + sw, we = (int)ceilf((float)(pw + 1) * bin_w) + sw;
      ws = min(max(ws, 0), W); we = min(max(we, 0), W);
      const bool empty = he <= hs || we <= ws;
      // x = hi + lo' 2^-11 with hi = fp16(x): x is ordered like the pair (hi, lo') lexicographically (hi is monotone in x,
      // and for equal hi the remainder decides), so the maximum is selected with packed fp16 compares — no conversion to
      // fp32.  A strict ">" keeps the FIRST maximum like the scalar loop (and torchvision); NaNs never win, as before.
      uint4 bh = empty ? zero : make_uint4(0xFC00FC00u, 0xFC00FC00u, 0xFC00FC00u, 0xFC00FC00u);   // -inf | 0
      uint4 bl = zero;
      for (int h = hs; h < he; ++h) {
        const __half* row = fc + (int64_t)h * W * 2 * C;
        for (int w = ws; w < we; ++w) {
          const uint4 vh = *reinterpret_cast<const uint4*>(row + (int64_t)w * 2 * C);
          const uint4 vl = *reinterpret_cast<const uint4*>(row + (int64_t)w * 2 * C + C);
          const uint32_t* ph_ = reinterpret_cast<const uint32_t*>(&vh);
          const uint32_t* pl_ = reinterpret_cast<const uint32_t*>(&vl);
          uint32_t* sh_ = reinterpret_cast<uint32_t*>(&bh);
          uint32_t* sl_ = reinterpret_cast<uint32_t*>(&bl);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const __half2 h2 = *reinterpret_cast<const __half2*>(&ph_[i]), l2 = *reinterpret_cast<const __half2*>(&pl_[i]);
            const __half2 mh = *reinterpret_cast<const __half2*>(&sh_[i]), ml = *reinterpret_cast<const __half2*>(&sl_[i]);
            const uint32_t take = __hgt2_mask(h2, mh) | (__heq2_mask(h2, mh) & __hgt2_mask(l2, ml));
            sh_[i] = (ph_[i] & take) | (sh_[i] & ~take);
            sl_[i] = (pl_[i] & take) | (sl_[i] & ~take);
          }
        }
      }
      __half* o = orow + (int64_t)pw * 2 * C + c * 8;
      *reinterpret_cast<uint4*>(o) = bh;
      *reinterpret_cast<uint4*>(o + C) = bl;
    }
  }
}

__global__ void mean_rows_h2_kernel(const __half* __restrict__ x, float* __restrict__ y, int R, int P, int C) {
  const int c4 = C / 4;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)R * c4) return;
  const int c = (int)(i % c4) * 4, r = (int)(i / c4);
  const __half* b = x + (int64_t)r * P * 2 * C + c;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int p = 0; p < P; ++p) {
    const H4 hi = ld_h4(b + (int64_t)p * 2 * C), lo = ld_h4(b + (int64_t)p * 2 * C + C);
    s.x += join1(hi.v[0], lo.v[0]); s.y += join1(hi.v[1], lo.v[1]); s.z += join1(hi.v[2], lo.v[2]); s.w += join1(hi.v[3], lo.v[3]);
  }
  store4(y + (int64_t)r * C + c, make_float4(s.x / (float)P, s.y / (float)P, s.z / (float)P, s.w / (float)P));
}

__global__ void split_f32_h2_kernel(const float* __restrict__ x, const float* __restrict__ add, int relu,
                                    __half* __restrict__ out, int64_t rows, int C) {
  const int c4 = C / 4;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * c4) return;
  const int c = (int)(i % c4) * 4;
  const int64_t r = i / c4;
  float4 v = load4(x + r * C + c);
  if (add) { const float4 a = load4(add + r * C + c); v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w; }
  if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
  H4 hi, lo;
  split1(v.x, hi.v[0], lo.v[0]); split1(v.y, hi.v[1], lo.v[1]); split1(v.z, hi.v[2], lo.v[2]); split1(v.w, hi.v[3], lo.v[3]);
  st_h4(out + r * 2 * C + c, hi);
  st_h4(out + r * 2 * C + C + c, lo);
}

__global__ void widen_h2_kernel(const __half* __restrict__ x, float* __restrict__ y, int64_t rows, int C) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * C) return;
  const int64_t r = i / C;
  const int c = (int)(i - r * C);
  y[i] = join1(x[r * 2 * C + c], x[r * 2 * C + C + c]);
}

}  // namespace

int maxpool3x3s2_ceil_f32_to_h2(const float* x, void* y, int N, int H, int W, int C, int OH, int OW, cudaStream_t st) {
  VLTK_CHECK(C % 4 == 0, "maxpool: C=%d must be a multiple of 4", C);
  if ((int64_t)N * OH * OW == 0) return 0;
  VLTK_CHECK(OH <= 65535 && N <= 65535, "maxpool: image too tall / batch too large for the grid");
  dim3 grid(ceil_div(OW * (C / 4), 256), OH, N);
  maxpool_f32_h2_kernel<<<grid, 256, 0, st>>>(x, (__half*)y, N, H, W, C, OH, OW);
  VLTK_LAUNCH_CHECK();
  return 0;
}

int roi_pool_h2(const void* feat, int N, int H, int W, int C, const float* rois, const int* count, int R, int P,
                float scale, void* out, cudaStream_t st) {
  VLTK_CHECK(C % 8 == 0, "roi_pool: C=%d must be a multiple of 8", C);
  if (N * R == 0) return 0;
  const int tx = min(128, round_up(C / 8, 32));
  roi_pool_h2_kernel<<<dim3(P, N * R), dim3(tx, 256 / tx), 0, st>>>((const __half*)feat, H, W, C, rois, count, R, P, scale, (__half*)out);
  VLTK_LAUNCH_CHECK();
  return 0;
}

int mean_rows_h2(const void* x, float* y, int R, int P, int C, cudaStream_t st) {
  VLTK_CHECK(C % 4 == 0, "mean_rows: C=%d must be a multiple of 4", C);
  const int64_t tot = (int64_t)R * (C / 4);
  if (tot == 0) return 0;
  mean_rows_h2_kernel<<<(unsigned)ceil_div64(tot, 128), 128, 0, st>>>((const __half*)x, y, R, P, C);
  VLTK_LAUNCH_CHECK();
  return 0;
}

int split_f32_h2(const float* x, const float* add, int relu, void* out, int64_t rows, int C, cudaStream_t st) {
  VLTK_CHECK(C % 4 == 0, "split_f32_h2: C must be a multiple of 4");
  const int64_t tot = rows * (C / 4);
  if (tot == 0) return 0;
  split_f32_h2_kernel<<<(unsigned)ceil_div64(tot, 256), 256, 0, st>>>(x, add, relu, (__half*)out, rows, C);
  VLTK_LAUNCH_CHECK();
  return 0;
}

int widen_h2(const void* x, float* y, int64_t rows, int C, cudaStream_t st) {
  const int64_t tot = rows * C;
  if (tot == 0) return 0;
  widen_h2_kernel<<<(unsigned)ceil_div64(tot, 256), 256, 0, st>>>((const __half*)x, y, rows, C);
  VLTK_LAUNCH_CHECK();
  return 0;
}

}  // namespace vltk
