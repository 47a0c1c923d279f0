// K10: torchvision.ops.RoIPool((P,P), scale) restated for NHWC features
// (reference call site frcnn.py:1179, 1195-1198; ROI format frcnn.py:426-441).
// HBM/L2-bound: the map (<=78 MB) is L2-resident, the output write is the algorithmic traffic.
#include "kernels.cuh"

namespace vltk {

namespace {

// 16-byte vector of channels: 4 x f32 or 8 x bf16
template <typename T> struct Vec;
template <> struct Vec<float> {
  static constexpr int N = 4;
  float v[4];
  __device__ static Vec load(const float* p) { Vec r; *reinterpret_cast<float4*>(r.v) = *reinterpret_cast<const float4*>(p); return r; }
  __device__ void store(float* p) const { *reinterpret_cast<float4*>(p) = *reinterpret_cast<const float4*>(v); }
  __device__ static Vec fill(float x) { Vec r; r.v[0] = r.v[1] = r.v[2] = r.v[3] = x; return r; }
  __device__ void max_with(const Vec& o) {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = fmaxf(v[i], o.v[i]);
  }
};
template <> struct Vec<bf16> {
  static constexpr int N = 8;
  __nv_bfloat162 v[4];
  __device__ static Vec load(const bf16* p) { Vec r; *reinterpret_cast<uint4*>(r.v) = *reinterpret_cast<const uint4*>(p); return r; }
  __device__ void store(bf16* p) const { *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(v); }
  __device__ static Vec fill(float x) { Vec r; r.v[0] = r.v[1] = r.v[2] = r.v[3] = __float2bfloat162_rn(x); return r; }
  __device__ void max_with(const Vec& o) {       // max of bf16 values is exact in any precision
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __hmax2(v[i], o.v[i]);
  }
};

// One CTA per (roi, output row ph).  threadIdx.x walks the 16-byte channel vectors and threadIdx.y splits the
// P bins of the row, so each thread pools ~P/2 bins whose loads are mutually independent (several L2 requests
// in flight per thread instead of one exposed latency per tiny CTA), the ROI geometry is computed once per
// thread, and every access is a coalesced channel run.
template <typename T>
__global__ void __launch_bounds__(256)
roi_pool_kernel(const T* __restrict__ feat, int H, int W, int C, const float* __restrict__ rois,
                const int* __restrict__ count, const int* __restrict__ bidx, int R, int P, float scale,
                T* __restrict__ out) {
  using V = Vec<T>;
  const int cv = C / V::N;
  const int ph = blockIdx.x, roi = blockIdx.y;
  // engine path: ROI slot r of image n, masked by count[n]; stage path: explicit batch index
  const int n = bidx ? bidx[roi] : roi / R, r = roi - (roi / R) * R;
  T* orow = out + ((int64_t)roi * P + ph) * P * C;
  if (!bidx && r >= count[n]) {
    for (int pw = threadIdx.y; pw < P; pw += blockDim.y)
      for (int c = threadIdx.x; c < cv; c += blockDim.x) V::fill(0.f).store(orow + (int64_t)pw * C + c * V::N);
    return;
  }
  const float4 b = reinterpret_cast<const float4*>(rois)[roi];
  // round half away from zero of the float product (C `round`)
  const int sw = (int)roundf(b.x * scale), sh = (int)roundf(b.y * scale);
  const int ew = (int)roundf(b.z * scale), eh = (int)roundf(b.w * scale);
  const int rw = max(ew - sw + 1, 1), rh = max(eh - sh + 1, 1);
  const float bin_h = (float)rh / (float)P, bin_w = (float)rw / (float)P;
  int hs = (int)floorf((float)ph * bin_h) + sh, he = (int)ceilf((float)(ph + 1) * bin_h) + sh;
  hs = min(max(hs, 0), H); he = min(max(he, 0), H);
  const T* f = feat + (int64_t)n * H * W * C;
  for (int c = threadIdx.x; c < cv; c += blockDim.x) {
    const T* fc = f + c * V::N;
#pragma unroll 4
    for (int pw = threadIdx.y; pw < P; pw += blockDim.y) {
      int ws = (int)floorf((float)pw * bin_w) + sw, we = (int)ceilf((float)(pw + 1) * bin_w) + sw;
      ws = min(max(ws, 0), W); we = min(max(we, 0), W);
      V m = (he <= hs || we <= ws) ? V::fill(0.f) : V::fill(-INFINITY);   // empty bin -> 0
      for (int h = hs; h < he; ++h) {
        const T* row = fc + (int64_t)h * W * C;
        for (int w = ws; w < we; ++w) m.max_with(V::load(row + (int64_t)w * C));
      }
      m.store(orow + (int64_t)pw * C + c * V::N);
    }
  }
}

}  // namespace

int roi_pool(const void* feat, DType dt, int N, int H, int W, int C, const float* rois,
             const int* count, int R, int P, float scale, void* out, cudaStream_t st) {
  VLTK_CHECK(C % 8 == 0, "roi_pool: C=%d must be a multiple of 8", C);
  if (N * R == 0) return 0;
  if (dt == DT_F32) {
    const int tx = min(128, round_up(C / 4, 32));
    roi_pool_kernel<float><<<dim3(P, N * R), dim3(tx, 256 / tx), 0, st>>>((const float*)feat, H, W, C, rois, count, nullptr, R, P, scale, (float*)out);
  } else {
    const int tx = min(128, round_up(C / 8, 32));
    roi_pool_kernel<bf16><<<dim3(P, N * R), dim3(tx, 256 / tx), 0, st>>>((const bf16*)feat, H, W, C, rois, count, nullptr, R, P, scale, (bf16*)out);
  }
  VLTK_LAUNCH_CHECK();
  return 0;
}

int roi_pool_indexed(const void* feat, DType dt, int H, int W, int C, const float* boxes, const int* bidx,
                     int R, int P, float scale, void* out, cudaStream_t st) {
  VLTK_CHECK(C % 4 == 0, "roi_pool: C=%d must be a multiple of 4", C);
  VLTK_CHECK(dt == DT_F32, "roi_pool_indexed: f32 only");
  if (R == 0) return 0;
  const int tx = min(128, round_up(C / 4, 32));
  roi_pool_kernel<float><<<dim3(P, R), dim3(tx, 256 / tx), 0, st>>>((const float*)feat, H, W, C, boxes, nullptr, bidx, R, P, scale, (float*)out);
  VLTK_LAUNCH_CHECK();
  return 0;
}

}  // namespace vltk
