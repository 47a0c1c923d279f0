// K10: torchvision.ops.RoIPool((P,P), scale) restated for NHWC features
// (reference call site frcnn.py:1179, 1195-1198; ROI format frcnn.py:426-441).
// One CTA per (roi, output row); threads sweep the channel dimension with 8/16-byte loads,
// so every feature-map access is a fully coalesced channel run.  HBM/L2-bound: the map
// (<=78 MB) is L2-resident, the output write is the algorithmic traffic.
#include "kernels.cuh"

namespace vltk {

namespace {

template <typename T>
__global__ void __launch_bounds__(256)
roi_pool_kernel(const T* __restrict__ feat, int H, int W, int C, const float* __restrict__ rois,
                const int* __restrict__ count, const int* __restrict__ bidx, int R, int P, float scale,
                T* __restrict__ out) {
  const int roi = blockIdx.x, ph = blockIdx.y;
  // engine path: ROI slot r of image n, masked by count[n]; stage path: explicit batch index
  const int n = bidx ? bidx[roi] : roi / R, r = roi - (roi / R) * R;
  const int c4n = C / 4;
  T* o = out + ((int64_t)roi * P + ph) * P * C;
  if (!bidx && r >= count[n]) {
    for (int i = threadIdx.x; i < P * c4n; i += blockDim.x) store4(o + (int64_t)i * 4, make_float4(0.f, 0.f, 0.f, 0.f));
    return;
  }
  const float4 b = reinterpret_cast<const float4*>(rois)[roi];
  // round half away from zero of the float product (C `round`)
  const int sw = (int)roundf(b.x * scale), sh = (int)roundf(b.y * scale);
  const int ew = (int)roundf(b.z * scale), eh = (int)roundf(b.w * scale);
  const int rw = max(ew - sw + 1, 1), rh = max(eh - sh + 1, 1);
  const float bin_h = (float)rh / (float)P, bin_w = (float)rw / (float)P;
  int hs = (int)floorf((float)ph * bin_h) + sh;
  int he = (int)ceilf((float)(ph + 1) * bin_h) + sh;
  hs = min(max(hs, 0), H);
  he = min(max(he, 0), H);
  const T* f = feat + (int64_t)n * H * W * C;
  for (int pw = 0; pw < P; ++pw) {
    int ws = (int)floorf((float)pw * bin_w) + sw;
    int we = (int)ceilf((float)(pw + 1) * bin_w) + sw;
    ws = min(max(ws, 0), W);
    we = min(max(we, 0), W);
    const bool empty = (he <= hs) || (we <= ws);
    for (int c4 = threadIdx.x; c4 < c4n; c4 += blockDim.x) {
      float4 m = empty ? make_float4(0.f, 0.f, 0.f, 0.f)
                       : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      for (int h = hs; h < he; ++h)
        for (int w = ws; w < we; ++w) {
          float4 v = load4(f + ((int64_t)h * W + w) * C + c4 * 4);
          m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
        }
      store4(o + (int64_t)pw * C + c4 * 4, m);
    }
  }
}

}  // namespace

int roi_pool(const void* feat, DType dt, int N, int H, int W, int C, const float* rois,
             const int* count, int R, int P, float scale, void* out, cudaStream_t st) {
  VLTK_CHECK(C % 4 == 0, "roi_pool: C=%d must be a multiple of 4", C);
  if (N * R == 0) return 0;
  dim3 grid(N * R, P);
  int threads = min(256, round_up(C / 4, 32));
  if (dt == DT_F32)
    roi_pool_kernel<float><<<grid, threads, 0, st>>>((const float*)feat, H, W, C, rois, count, nullptr, R, P, scale, (float*)out);
  else
    roi_pool_kernel<bf16><<<grid, threads, 0, st>>>((const bf16*)feat, H, W, C, rois, count, nullptr, R, P, scale, (bf16*)out);
  VLTK_LAUNCH_CHECK();
  return 0;
}

int roi_pool_indexed(const void* feat, DType dt, int H, int W, int C, const float* boxes, const int* bidx,
                     int R, int P, float scale, void* out, cudaStream_t st) {
  VLTK_CHECK(C % 4 == 0, "roi_pool: C=%d must be a multiple of 4", C);
  VLTK_CHECK(dt == DT_F32, "roi_pool_indexed: f32 only");
  if (R == 0) return 0;
  dim3 grid(R, P);
  int threads = min(256, round_up(C / 4, 32));
  roi_pool_kernel<float><<<grid, threads, 0, st>>>((const float*)feat, H, W, C, boxes, nullptr, bidx, R, P, scale, (float*)out);
  VLTK_LAUNCH_CHECK();
  return 0;
}

}  // namespace vltk
