// K14: the detection tail, one CTA per image, everything in fp32:
//   softmax over all classes -> drop background -> per-ROI best class + prob
//   attribute softmax over the first num_attrs logits -> best attribute + prob
//   decode ONLY the winning class's deltas (weights 10,10,5,5), clip to the resized image
//   class-agnostic greedy NMS in score order, first max_det survivors, threshold list retry
//   scale by scales_yx, gather the pooled features, zero-pad to [max_det, ...]
// Reference: ROIOutputs.inference / do_nms (frcnn.py:1242-1294, 116-143); the padded
// layout + normalized_boxes follow the v1.0.0 contract.
// Compiled with -fmad=false (see rpn.cu).
#include "kernels.cuh"

namespace vltk {

namespace {

constexpr int TAIL_THREADS = 512;
constexpr int TAIL_MAX_R = 512;
constexpr float SCALE_CLAMP = 4.135166556742356f;

__device__ __forceinline__ void warp_argmax(float& v, int& i) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, v, o);
    int oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
  }
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Phase 1, for every ROI slot of the batch in parallel (one warp per ROI, ~N*R warps over the whole GPU instead of
// 16 warps per image): class / attribute posteriors and the winning class's decoded + clipped box.
//   stats[row] = {x1, y1, x2, y2, prob, attr_prob, bitcast(cls), bitcast(attr)}   (32 B per ROI)
__global__ void __launch_bounds__(256)
roi_stats_kernel(TailArgs a, float* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= (int64_t)a.N * a.R) return;
  const int n = (int)(row / a.R), r = (int)(row - (int64_t)n * a.R);
  if (r >= min(a.count[n], a.R)) return;                 // empty slot: never read by the tail
  const float img_h = (float)a.sizes_hw[2 * n], img_w = (float)a.sizes_hw[2 * n + 1];
  const int NC = a.num_classes, NA = a.num_attrs;
  const float* __restrict__ l = a.cls_logits + row * a.ldc;
  float best = -INFINITY; int bi = 0x7fffffff; float mx = -INFINITY;
  for (int i = lane; i <= NC; i += 32) {
    float v = l[i];
    mx = fmaxf(mx, v);
    if (i < NC && v > best) { best = v; bi = i; }
  }
  mx = warp_max(mx);
  warp_argmax(best, bi);
  if (bi == 0x7fffffff) bi = 0;            // all-NaN / -inf row (the reference asserts isfinite, frcnn.py:148): stay in bounds
  float s = 0.f;
  for (int i = lane; i <= NC; i += 32) s += expf(l[i] - mx);
  s = warp_sum(s);
  const float prob = expf(best - mx) / s;

  const float* __restrict__ al = a.attr_logits + row * a.lda;
  float ab = -INFINITY; int ai = 0x7fffffff;
  for (int i = lane; i < NA; i += 32) {
    float v = al[i];
    if (v > ab) { ab = v; ai = i; }
  }
  warp_argmax(ab, ai);
  if (ai == 0x7fffffff) ai = 0;
  float as = 0.f;
  for (int i = lane; i < NA; i += 32) as += expf(al[i] - ab);
  as = warp_sum(as);

  float4 d;
  if (a.bbox_deltas) d = *reinterpret_cast<const float4*>(a.bbox_deltas + row * a.ldb + 4 * bi);
  else {                                   // the winning class's four regression outputs, straight from the weights
    const float* __restrict__ f = a.feats + row * a.D;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = lane * 4; k < a.D; k += 128) {
      const float4 x = *reinterpret_cast<const float4*>(f + k);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t o = (int64_t)(4 * bi + j) * a.D + k;
        if (a.bbox_w_f32) {                  // exact_tc: the fp32 weights themselves
          const float4 w = *reinterpret_cast<const float4*>(a.bbox_w_f32 + o);
          acc[j] += x.x * w.x + x.y * w.y + x.z * w.z + x.w * w.w;
        } else {
          const float4 wh = load4(a.bbox_w_hi + o), wl = load4(a.bbox_w_lo + o);
          acc[j] += x.x * (wh.x + wl.x) + x.y * (wh.y + wl.y) + x.z * (wh.z + wl.z) + x.w * (wh.w + wl.w);
        }
      }
    }
    d = make_float4(warp_sum(acc[0]) + a.bbox_bias[4 * bi], warp_sum(acc[1]) + a.bbox_bias[4 * bi + 1],
                    warp_sum(acc[2]) + a.bbox_bias[4 * bi + 2], warp_sum(acc[3]) + a.bbox_bias[4 * bi + 3]);
  }
  if (lane == 0) {
    const float4 p = reinterpret_cast<const float4*>(a.proposals)[row];
    const float w = p.z - p.x, h = p.w - p.y;
    const float cx = p.x + 0.5f * w, cy = p.y + 0.5f * h;
    const float dx = d.x / a.wx, dy = d.y / a.wy;
    const float dw = fminf(d.z / a.ww, SCALE_CLAMP), dh = fminf(d.w / a.wh, SCALE_CLAMP);
    const float pcx = dx * w + cx, pcy = dy * h + cy;
    const float pw = expf(dw) * w, ph = expf(dh) * h;
    float x1 = pcx - 0.5f * pw, y1 = pcy - 0.5f * ph, x2 = pcx + 0.5f * pw, y2 = pcy + 0.5f * ph;
    x1 = fminf(fmaxf(x1, 0.f), img_w); y1 = fminf(fmaxf(y1, 0.f), img_h);
    x2 = fminf(fmaxf(x2, 0.f), img_w); y2 = fminf(fmaxf(y2, 0.f), img_h);
    float4* o = reinterpret_cast<float4*>(stats + row * 8);
    o[0] = make_float4(x1, y1, x2, y2);
    o[1] = make_float4(prob, 1.0f / as /* exp(max-max)/sum */, __int_as_float(bi), __int_as_float(ai));
  }
}

__global__ void __launch_bounds__(TAIL_THREADS)
roi_tail_kernel(TailArgs a, const float* __restrict__ stats, float t0, float t1, float t2, float t3, int SORT_N, int RW) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // carve dynamic smem
  float4* s_box = reinterpret_cast<float4*>(smem_raw);                       // [R]
  unsigned long long* s_keys = reinterpret_cast<unsigned long long*>(s_box + a.R);  // [SORT_N]
  unsigned long long* s_mask = s_keys + SORT_N;                               // [R][RW]
  unsigned long long* s_removed = s_mask + (size_t)a.R * RW;                  // [RW]
  float* s_score = reinterpret_cast<float*>(s_removed + RW);                  // [R]
  float* s_attr_p = s_score + a.R;                                            // [R]
  int* s_cls = reinterpret_cast<int*>(s_attr_p + a.R);                        // [R]
  int* s_attr = s_cls + a.R;                                                  // [R]
  int* s_order = s_attr + a.R;                                                // [R]
  int* s_kept = s_order + a.R;                                                // [max_det]
  __shared__ int s_nkept, s_done;

  const int tid = threadIdx.x;
  const int n = blockIdx.x, R = a.R;
  const int cnt = min(a.count[n], R);
  const float img_h = (float)a.sizes_hw[2 * n], img_w = (float)a.sizes_hw[2 * n + 1];

  // ---- 1. per-ROI posteriors + decoded boxes were computed by roi_stats_kernel
  for (int r = tid; r < cnt; r += TAIL_THREADS) {
    const float4* st4 = reinterpret_cast<const float4*>(stats + ((int64_t)n * R + r) * 8);
    const float4 b = st4[0], q = st4[1];
    s_box[r] = b;
    s_score[r] = q.x;
    s_attr_p[r] = q.y;
    s_cls[r] = __float_as_int(q.z);
    s_attr[r] = __float_as_int(q.w);
  }
  // ---- 2. stable descending sort of the scores (ties: lower ROI index first)
  for (int i = tid; i < SORT_N; i += TAIL_THREADS) s_keys[i] = 0ull;
  __syncthreads();
  for (int r = tid; r < cnt; r += TAIL_THREADS)
    s_keys[r] = ((unsigned long long)float_to_key(s_score[r]) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)r);
  __syncthreads();
  for (int k = 2; k <= SORT_N; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < SORT_N / 2; t += TAIL_THREADS) {
        int i = ((t / j) * 2 * j) + (t % j), l2 = i + j;
        unsigned long long x = s_keys[i], y = s_keys[l2];
        bool desc = (i & k) == 0;
        if (desc ? (x < y) : (x > y)) { s_keys[i] = y; s_keys[l2] = x; }
      }
      __syncthreads();
    }
  for (int i = tid; i < cnt; i += TAIL_THREADS) s_order[i] = (int)(0xFFFFFFFFu - (uint32_t)(s_keys[i] & 0xFFFFFFFFull));
  if (tid == 0) s_done = 0;
  __syncthreads();

  // ---- 3. class-agnostic NMS, thresholds tried in order (frcnn.py:1274-1278)
  const float thr_list[4] = {t0, t1, t2, t3};
  for (int ti = 0; ti < a.n_thresh; ++ti) {
    if (s_done) break;
    const float thr = thr_list[ti];
    for (int e = tid; e < cnt * RW; e += TAIL_THREADS) {
      const int i = e / RW, wd = e - i * RW;
      const float4 bi4 = s_box[s_order[i]];
      const float ai_ = (bi4.z - bi4.x) * (bi4.w - bi4.y);
      unsigned long long bits = 0ull;
      const int j0 = wd * 64;
      for (int b = 0; b < 64; ++b) {
        const int j = j0 + b;
        if (j <= i || j >= cnt) continue;
        const float4 bj = s_box[s_order[j]];
        const float aj = (bj.z - bj.x) * (bj.w - bj.y);
        float xx1 = fmaxf(bi4.x, bj.x), yy1 = fmaxf(bi4.y, bj.y);
        float xx2 = fminf(bi4.z, bj.z), yy2 = fminf(bi4.w, bj.w);
        float w = fmaxf(0.f, xx2 - xx1), h = fmaxf(0.f, yy2 - yy1);
        float inter = w * h;
        float ovr = inter / (ai_ + aj - inter);
        if (ovr > thr) bits |= 1ull << b;
      }
      s_mask[e] = bits;
    }
    if (tid < RW) s_removed[tid] = 0ull;
    __syncthreads();
    if (tid == 0) {
      int nk = 0;
      for (int i = 0; i < cnt && nk < a.max_det; ++i) {
        if ((s_removed[i >> 6] >> (i & 63)) & 1ull) continue;
        s_kept[nk++] = i;
        for (int wd = i >> 6; wd < RW; ++wd) s_removed[wd] |= s_mask[i * RW + wd];
      }
      s_nkept = nk;
      if (nk >= a.min_det && nk <= a.max_det) s_done = 1;
    }
    __syncthreads();
  }

  // ---- 4. outputs, dense [max_det, ...] with pad_value past the count
  const int nk = (a.n_thresh > 0) ? s_nkept : 0;
  const int MD = a.max_det;
  float sy = 1.f, sx = 1.f;
  if (a.scales_yx) { sy = a.scales_yx[2 * n]; sx = a.scales_yx[2 * n + 1]; }
  const float raw_h = img_h * sy, raw_w = img_w * sx;
  if (tid == 0) a.preds_per_image[n] = nk;
  for (int t = tid; t < MD; t += TAIL_THREADS) {
    const int64_t o = (int64_t)n * MD + t;
    float4 bx = make_float4(a.pad_value, a.pad_value, a.pad_value, a.pad_value), nb = bx;
    long long oid = (long long)a.pad_value, aid = (long long)a.pad_value;
    float op = a.pad_value, ap = a.pad_value;
    int ki = -1;
    if (t < nk) {
      const int r = s_order[s_kept[t]];
      const float4 b = s_box[r];
      bx = a.scales_yx ? make_float4(b.x * sx, b.y * sy, b.z * sx, b.w * sy) : b;
      nb = make_float4(bx.x / raw_w, bx.y / raw_h, bx.z / raw_w, bx.w / raw_h);
      oid = s_cls[r]; aid = s_attr[r]; op = s_score[r]; ap = s_attr_p[r]; ki = r;
    }
    reinterpret_cast<float4*>(a.boxes)[o] = bx;
    reinterpret_cast<float4*>(a.norm_boxes)[o] = nb;
    a.obj_ids[o] = oid; a.attr_ids[o] = aid; a.obj_probs[o] = op; a.attr_probs[o] = ap;
    a.keep_idx[o] = ki;
  }
  const int d4 = a.D / 4;
  for (int e = tid; e < MD * d4; e += TAIL_THREADS) {
    const int t = e / d4, c = e - t * d4;
    float4 v = make_float4(a.pad_value, a.pad_value, a.pad_value, a.pad_value);
    if (t < nk) {
      const int r = s_order[s_kept[t]];
      v = reinterpret_cast<const float4*>(a.feats + ((int64_t)n * R + r) * a.D)[c];
    }
    reinterpret_cast<float4*>(a.roi_features + ((int64_t)n * MD + t) * a.D)[c] = v;
  }
}

__global__ void row_argmax_kernel(const float* __restrict__ x, int ld, int rows, int n, int* __restrict__ out) {
  const int row = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* l = x + (int64_t)row * ld;
  float best = -INFINITY; int bi = 0x7fffffff;
  for (int i = lane; i < n; i += 32) {
    float v = l[i];
    if (v > best) { best = v; bi = i; }
  }
  warp_argmax(best, bi);
  if (lane == 0) out[row] = bi == 0x7fffffff ? 0 : bi;   // never an out-of-range table index (non-finite row)
}

__global__ void gather_rows_kernel(const float* __restrict__ table, int ld, const int* __restrict__ idx,
                                   int rows, int cols4, float* __restrict__ out, int ldo) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)rows * cols4) return;
  int r = (int)(i / cols4), c = (int)(i - (int64_t)r * cols4);
  reinterpret_cast<float4*>(out + (int64_t)r * ldo)[c] =
      reinterpret_cast<const float4*>(table + (int64_t)idx[r] * ld)[c];
}

}  // namespace

int roi_tail(const TailArgs& a, cudaStream_t st) {
  VLTK_CHECK(a.R >= 1 && a.R <= TAIL_MAX_R, "roi_tail: R=%d out of range (1..%d)", a.R, TAIL_MAX_R);
  VLTK_CHECK(a.n_thresh >= 1 && a.n_thresh <= 4, "roi_tail: 1..4 NMS thresholds supported, got %d", a.n_thresh);
  VLTK_CHECK(a.D % 4 == 0 && a.ldb % 4 == 0, "roi_tail: D and ldb must be multiples of 4");
  VLTK_CHECK(a.max_det >= 1 && a.max_det <= a.R, "roi_tail: max_det=%d must be in 1..R", a.max_det);
  if (a.N == 0) return 0;
  int sort_n = 2;
  while (sort_n < a.R) sort_n <<= 1;
  const int RW = ceil_div(a.R, 64);
  size_t smem = (size_t)a.R * 16 + (size_t)sort_n * 8 + (size_t)a.R * RW * 8 + (size_t)RW * 8 +
                (size_t)a.R * 4 * 5 + (size_t)a.max_det * 4;
  static DeviceOnce once;
  if (once.first()) {
    VLTK_CUDA(cudaFuncSetAttribute(roi_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  }
  VLTK_CHECK(smem <= 100 * 1024, "roi_tail: smem %zu too large", smem);
  float t[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i = 0; i < a.n_thresh; ++i) t[i] = a.nms_thresh[i];
  VLTK_CHECK(a.stats != nullptr, "roi_tail: stats scratch (N*R*8 floats) missing");
  const int64_t rows = (int64_t)a.N * a.R;
  roi_stats_kernel<<<(unsigned)ceil_div64(rows, 8), 256, 0, st>>>(a, a.stats);
  VLTK_LAUNCH_CHECK();
  roi_tail_kernel<<<a.N, TAIL_THREADS, smem, st>>>(a, a.stats, t[0], t[1], t[2], t[3], sort_n, RW);
  VLTK_LAUNCH_CHECK();
  return 0;
}

int row_argmax(const float* x, int ld, int rows, int n, int* out, cudaStream_t st) {
  if (rows == 0) return 0;
  row_argmax_kernel<<<ceil_div(rows, 8), 256, 0, st>>>(x, ld, rows, n, out);
  VLTK_LAUNCH_CHECK();
  return 0;
}

int gather_rows(const float* table, int ld, const int* idx, int rows, int cols, float* out, int ldo,
                cudaStream_t st) {
  VLTK_CHECK(cols % 4 == 0 && ld % 4 == 0 && ldo % 4 == 0, "gather_rows: cols/ld must be multiples of 4");
  int64_t tot = (int64_t)rows * (cols / 4);
  if (tot == 0) return 0;
  gather_rows_kernel<<<(unsigned)ceil_div64(tot, 256), 256, 0, st>>>(table, ld, idx, rows, cols / 4, out, ldo);
  VLTK_LAUNCH_CHECK();
  return 0;
}

}  // namespace vltk
