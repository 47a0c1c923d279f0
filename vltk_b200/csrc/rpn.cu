// RPN proposal selection (K5-K9 of SURVEY.md §2.1): anchors are generated on the fly, the
// top-k objectness logits are radix-selected and sorted inside one CTA per image, only the
// selected anchors are decoded/clipped, and greedy NMS runs as an IoU bitmask + an on-device
// sequential reduce with early exit.  No host synchronisation anywhere (the reference syncs
// at frcnn.py:375).  Reference semantics: frcnn.py:176-197, 264-390, 548-584, 748-781,
// 1463-1510; torchvision.ops.nms (stable score order, suppress iff IoU > thr).
//
// Compiled with -fmad=false: box arithmetic must round like the reference's unfused
// torch ops (mul then add), otherwise near-threshold IoU decisions can flip.
#include "kernels.cuh"

namespace vltk {

namespace {

constexpr int SEL_THREADS = 1024;
constexpr float SCALE_CLAMP = 4.135166556742356f;  // log(1000/16), frcnn.py:510

__device__ __forceinline__ uint32_t logit_key(const float* __restrict__ logits, int ldh, int A, int i) {
  int pix = i / A, an = i - pix * A;
  return float_to_key(logits[(int64_t)pix * ldh + an]);
}

// One CTA per image.  keys: dynamic smem, SORT_N u64.
__global__ void __launch_bounds__(SEL_THREADS)
rpn_select_kernel(RpnSelectArgs a, int SORT_N) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);
  __shared__ int hist[256];
  __shared__ int warp_sums[32];
  __shared__ uint32_t s_prefix;
  __shared__ int s_remaining, s_count;

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n = blockIdx.x;
  const int HW = a.H4 * a.W4;
  const int n_el = HW * a.A;
  const int K = a.K;
  const float* __restrict__ head = a.head + (int64_t)n * HW * a.ldh;
  const float* __restrict__ logits = head + a.logit_off;

  // ---- 1. radix-select the K-th largest key (4 x 8-bit passes, MSB first)
  uint32_t T = 0;
  int need_eq = 0x7fffffff;
  if (n_el > K) {
    if (tid == 0) { s_prefix = 0; s_remaining = K; }
    uint32_t pmask = 0;
    for (int pass = 3; pass >= 0; --pass) {
      if (tid < 256) hist[tid] = 0;
      __syncthreads();
      const uint32_t prefix = s_prefix;
      const int sh = 8 * pass;
      for (int i = tid; i < n_el; i += SEL_THREADS) {
        uint32_t k = logit_key(logits, a.ldh, a.A, i);
        if ((k & pmask) == prefix) atomicAdd(&hist[(k >> sh) & 255], 1);
      }
      __syncthreads();
      if (tid == 0) {
        int rem = s_remaining, cum = 0, b = 255;
        for (; b > 0; --b) {
          if (cum + hist[b] >= rem) break;
          cum += hist[b];
        }
        s_remaining = rem - cum;
        s_prefix = prefix | ((uint32_t)b << sh);
      }
      pmask |= 0xFFu << sh;
      __syncthreads();
    }
    T = s_prefix;
    need_eq = s_remaining;
  }

  // ---- 2. compact: key > T always; key == T lowest index first (ordered block scan)
  if (tid == 0) s_count = 0;
  for (int i = tid; i < SORT_N; i += SEL_THREADS) keys[i] = 0ull;
  __syncthreads();
  int eq_seen = 0;  // identical in every thread
  for (int base = 0; base < n_el; base += SEL_THREADS) {
    const int i = base + tid;
    uint32_t k = 0;
    bool gt = false, eq = false;
    if (i < n_el) {
      k = logit_key(logits, a.ldh, a.A, i);
      gt = k > T;
      eq = k == T;
    }
    unsigned bal = __ballot_sync(0xffffffffu, eq);
    if (lane == 0) warp_sums[wid] = __popc(bal);
    __syncthreads();
    int before = eq_seen, tot = 0;
    for (int w = 0; w < 32; ++w) {
      int v = warp_sums[w];
      if (w < wid) before += v;
      tot += v;
    }
    eq_seen += tot;
    int rank = before + __popc(bal & ((1u << lane) - 1u));
    if (gt || (eq && rank < need_eq)) {
      int pos = atomicAdd(&s_count, 1);
      keys[pos] = ((unsigned long long)k << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)i);
    }
    __syncthreads();  // warp_sums is reused by the next tile
  }

  // ---- 3. bitonic sort, descending (equal logits: lower anchor index first)
  for (int k = 2; k <= SORT_N; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < SORT_N / 2; t += SEL_THREADS) {
        int i = ((t / j) * 2 * j) + (t % j);
        int l = i + j;
        unsigned long long x = keys[i], y = keys[l];
        bool desc = (i & k) == 0;
        if (desc ? (x < y) : (x > y)) { keys[i] = y; keys[l] = x; }
      }
      __syncthreads();
    }
  }

  // ---- 4. decode + clip the selected anchors (frcnn.py:548-584, 147-160)
  const float img_h = (float)a.sizes_hw[2 * n], img_w = (float)a.sizes_hw[2 * n + 1];
  for (int r = tid; r < K; r += SEL_THREADS) {
    const uint32_t idx = 0xFFFFFFFFu - (uint32_t)(keys[r] & 0xFFFFFFFFull);
    const int pix = idx / a.A, an = idx - pix * a.A;
    const int py = pix / a.W4, px = pix - py * a.W4;
    const float sx = (float)(px * a.stride), sy = (float)(py * a.stride);
    const float ax1 = sx + a.cell[4 * an + 0], ay1 = sy + a.cell[4 * an + 1];
    const float ax2 = sx + a.cell[4 * an + 2], ay2 = sy + a.cell[4 * an + 3];
    const float* hp = head + (int64_t)pix * a.ldh;
    const float logit = hp[a.logit_off + an];
    const float4 d = *reinterpret_cast<const float4*>(hp + a.delta_off + 4 * an);
    const float w = ax2 - ax1, h = ay2 - ay1;
    const float cx = ax1 + 0.5f * w, cy = ay1 + 0.5f * h;
    const float dx = d.x / a.wx, dy = d.y / a.wy;
    const float dw = fminf(d.z / a.ww, SCALE_CLAMP), dh = fminf(d.w / a.wh, SCALE_CLAMP);
    const float pcx = dx * w + cx, pcy = dy * h + cy;
    const float pw = expf(dw) * w, ph = expf(dh) * h;
    float x1 = pcx - 0.5f * pw, y1 = pcy - 0.5f * ph;
    float x2 = pcx + 0.5f * pw, y2 = pcy + 0.5f * ph;
    bool alive = true;
    if (a.ignorey && a.scales_yx) {       // frcnn.py:328-366, on the decoded, not yet clipped box
      const float sxs = a.scales_yx[2 * n + 1];            // the reference divides by the X scale (:331)
      for (int j = 0; j < a.J && alive; ++j) {
        const float r0 = a.ignorey[((int64_t)n * a.J + j) * 2 + 0] / sxs;
        const float r1 = a.ignorey[((int64_t)n * a.J + j) * 2 + 1] / sxs;
        if (r1 <= y2 && r0 >= y1) { alive = false; break; }   // spans the whole range: dropped (:333-336)
        const bool past = y1 > r1 && y2 > r0;                 // (:342-343; `box_ignore_below` can never be true)
        if (!past) {
          const float d_top = fabsf(r1 - y2), d_bot = fabsf(r0 - y1);
          if (d_top < d_bot) y2 = (float)(int)r0;             // clip_top    (:351-357, 366)
          else if (d_bot < d_top) y1 = (float)(int)r1;        // clip_bottom (:358-365)
        }
      }
    }
    x1 = fminf(fmaxf(x1, 0.f), img_w); y1 = fminf(fmaxf(y1, 0.f), img_h);
    x2 = fminf(fmaxf(x2, 0.f), img_w); y2 = fminf(fmaxf(y2, 0.f), img_h);
    const int64_t o = (int64_t)n * K + r;
    *reinterpret_cast<float4*>(a.boxes + o * 4) = make_float4(x1, y1, x2, y2);
    a.scores[o] = logit;
    a.anchor_idx[o] = (int)idx;
    a.valid[o] = (alive && (x2 - x1) > a.min_size && (y2 - y1) > a.min_size) ? 1 : 0;
  }
}

// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool iou_gt(const float4& a, float area_a, const float4& b, float area_b,
                                       float thr) {
  float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
  float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
  float w = fmaxf(0.f, xx2 - xx1), h = fmaxf(0.f, yy2 - yy1);
  float inter = w * h;
  float ovr = inter / (area_a + area_b - inter);
  return ovr > thr;  // NaN (0/0) never suppresses
}

// grid (KBe, KBe, N), 64 threads: row block y vs column block x (upper triangle only) over the first Ke
// candidates; the mask row stride stays KB words.  Images whose `done` flag is set are skipped.
__global__ void __launch_bounds__(64)
nms_mask_kernel(const float* __restrict__ boxes, int K, int Ke, int KB, float thr,
                unsigned long long* __restrict__ mask, const int* __restrict__ done) {
  const int cb = blockIdx.x, rb = blockIdx.y, n = blockIdx.z;
  if (cb < rb) return;
  if (done && done[n]) return;
  __shared__ float4 cbox[64];
  __shared__ float carea[64];
  const int tid = threadIdx.x;
  const float4* b4 = reinterpret_cast<const float4*>(boxes) + (int64_t)n * K;
  const int j0 = cb * 64;
  if (j0 + tid < Ke) {
    float4 b = b4[j0 + tid];
    cbox[tid] = b;
    carea[tid] = (b.z - b.x) * (b.w - b.y);
  }
  __syncthreads();
  const int i = rb * 64 + tid;
  if (i >= Ke) return;
  const float4 me = b4[i];
  const float area = (me.z - me.x) * (me.w - me.y);
  unsigned long long bits = 0ull;
  const int lim = min(64, Ke - j0);
  for (int b = 0; b < lim; ++b) {
    if (j0 + b > i && iou_gt(me, area, cbox[b], carea[b], thr)) bits |= 1ull << b;
  }
  mask[((int64_t)n * K + i) * KB + cb] = bits;
}

// One CTA per image; thread t owns the 64-bit "removed" word of chunk t.  Scans the first Ke candidates.
//   phase 1 (Ke < K, done_out != nullptr): greedy NMS decisions for candidate j depend only on candidates
//     before j, so the prefix result is exact; if it already holds max_keep survivors the image is DONE
//     (done_out[n] = 1 and outputs are written), otherwise done_out[n] = 0 and nothing is written;
//   phase 2 / single phase (Ke == K): skipped for images flagged in done_in, else the full scan.
__global__ void __launch_bounds__(128)
nms_scan_kernel(NmsArgs a, int KB, int Ke, const int* __restrict__ done_in, int* __restrict__ done_out) {
  __shared__ unsigned long long diag[64];
  __shared__ unsigned long long s_kept_bits;
  __shared__ int s_count;
  extern __shared__ int kept_list[];  // max_keep ints
  const int tid = threadIdx.x, n = blockIdx.x, K = a.K;
  if (done_in && done_in[n]) return;
  const int KBe = (Ke + 63) / 64;
  const unsigned long long* __restrict__ mask = a.mask + (int64_t)n * K * KB;

  unsigned long long removed = 0ull;
  if (tid < KBe) {
    for (int b = 0; b < 64; ++b) {
      int j = tid * 64 + b;
      bool ok = j < Ke && (a.valid == nullptr || a.valid[(int64_t)n * K + j]);
      if (!ok) removed |= 1ull << b;
    }
  }
  if (tid == 0) s_count = 0;
  __syncthreads();

  for (int c = 0; c < KBe; ++c) {
    if (tid < 64) {
      int i = c * 64 + tid;
      diag[tid] = (i < Ke) ? mask[(int64_t)i * KB + c] : 0ull;
    }
    __syncthreads();
    if (tid == c) {  // owner of this chunk resolves it serially
      unsigned long long cur = removed, kb = 0ull;
      int cnt = s_count;
      for (int b = 0; b < 64 && cnt < a.max_keep; ++b) {
        if (!((cur >> b) & 1ull)) {
          kb |= 1ull << b;
          cur |= diag[b];
          kept_list[cnt++] = c * 64 + b;
        }
      }
      s_kept_bits = kb;
      s_count = cnt;
    }
    __syncthreads();
    if (s_count >= a.max_keep) break;
    unsigned long long kb = s_kept_bits;
    if (tid > c && tid < KBe) {
      // OR the kept rows' masks into this thread's word, 8 INDEPENDENT loads at a time (one load per
      // iteration of a data-dependent loop serialised ~64 L2 latencies per chunk and was the whole cost)
      while (kb) {
        unsigned long long v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[j] = 0ull;
          if (kb) {
            const int b = __ffsll((long long)kb) - 1;
            kb &= kb - 1;
            v[j] = mask[(int64_t)(c * 64 + b) * KB + tid];
          }
        }
        removed |= ((v[0] | v[1]) | (v[2] | v[3])) | ((v[4] | v[5]) | (v[6] | v[7]));
      }
    }
    __syncthreads();
  }
  __syncthreads();
  const int cnt = s_count;
  if (done_out) {
    const bool fin = cnt >= a.max_keep || Ke >= K;
    if (tid == 0) done_out[n] = fin ? 1 : 0;
    if (!fin) return;                 // uniform: the full phase will produce this image
  }
  if (tid == 0) a.out_count[n] = cnt;
  for (int t = tid; t < a.max_keep; t += blockDim.x) {
    const int64_t o = (int64_t)n * a.max_keep + t;
    if (t < cnt) {
      int p = kept_list[t];
      reinterpret_cast<float4*>(a.out_boxes)[o] = reinterpret_cast<const float4*>(a.boxes)[(int64_t)n * K + p];
      if (a.out_scores) a.out_scores[o] = a.scores[(int64_t)n * K + p];
      a.out_idx[o] = p;
    } else {
      reinterpret_cast<float4*>(a.out_boxes)[o] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a.out_scores) a.out_scores[o] = 0.f;
      a.out_idx[o] = -1;
    }
  }
}

// single CTA: order = stable argsort(-scores); gathers boxes/scores into sorted order
__global__ void __launch_bounds__(SEL_THREADS)
sort_boxes_kernel(const float* __restrict__ boxes, const float* __restrict__ scores, int K, int SORT_N,
                  float* __restrict__ sboxes, float* __restrict__ sscores, int* __restrict__ order) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);
  const int tid = threadIdx.x;
  for (int i = tid; i < SORT_N; i += SEL_THREADS)
    keys[i] = i < K ? (((unsigned long long)float_to_key(scores[i]) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)i)) : 0ull;
  __syncthreads();
  for (int k = 2; k <= SORT_N; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < SORT_N / 2; t += SEL_THREADS) {
        int i = ((t / j) * 2 * j) + (t % j), l = i + j;
        unsigned long long x = keys[i], y = keys[l];
        bool desc = (i & k) == 0;
        if (desc ? (x < y) : (x > y)) { keys[i] = y; keys[l] = x; }
      }
      __syncthreads();
    }
  for (int r = tid; r < K; r += SEL_THREADS) {
    int idx = (int)(0xFFFFFFFFu - (uint32_t)(keys[r] & 0xFFFFFFFFull));
    order[r] = idx;
    sscores[r] = scores[idx];
    reinterpret_cast<float4*>(sboxes)[r] = reinterpret_cast<const float4*>(boxes)[idx];
  }
}

__global__ void remap_kernel(const int* __restrict__ idx, const int* __restrict__ order, int n, int* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = idx[i] >= 0 ? order[idx[i]] : -1;
}

}  // namespace

int sort_boxes_desc(const float* boxes, const float* scores, int K, float* sboxes, float* sscores, int* order,
                    cudaStream_t st) {
  VLTK_CHECK(K <= 8192, "sort_boxes_desc: K=%d too large", K);
  if (K == 0) return 0;
  int sort_n = 2;
  while (sort_n < K) sort_n <<= 1;
  static DeviceOnce once;
  if (once.first()) {
    VLTK_CUDA(cudaFuncSetAttribute(sort_boxes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 8));
  }
  sort_boxes_kernel<<<1, SEL_THREADS, (size_t)sort_n * 8, st>>>(boxes, scores, K, sort_n, sboxes, sscores, order);
  VLTK_LAUNCH_CHECK();
  return 0;
}

int remap_indices(const int* idx, const int* order, int n, int* out, cudaStream_t st) {
  if (n == 0) return 0;
  remap_kernel<<<ceil_div(n, 256), 256, 0, st>>>(idx, order, n, out);
  VLTK_LAUNCH_CHECK();
  return 0;
}

int rpn_select(const RpnSelectArgs& a, cudaStream_t st) {
  VLTK_CHECK(a.K >= 1 && a.K <= 8192, "rpn_select: K=%d out of range (1..8192)", a.K);
  VLTK_CHECK(a.ldh % 4 == 0 && a.delta_off % 4 == 0,
             "rpn_select: ldh=%d and delta_off=%d must be multiples of 4", a.ldh, a.delta_off);
  int sort_n = 2;
  while (sort_n < a.K) sort_n <<= 1;
  size_t smem = (size_t)sort_n * sizeof(unsigned long long);
  static DeviceOnce once;
  if (once.first()) {
    VLTK_CUDA(cudaFuncSetAttribute(rpn_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 8));
  }
  if (a.N == 0) return 0;
  rpn_select_kernel<<<a.N, SEL_THREADS, smem, st>>>(a, sort_n);
  VLTK_LAUNCH_CHECK();
  return 0;
}

size_t nms_mask_bytes(int N, int K) { return (size_t)N * K * ceil_div(K, 64) * sizeof(unsigned long long); }

int nms_sorted(const NmsArgs& a, cudaStream_t st) {
  const int KB = ceil_div(a.K, 64);
  VLTK_CHECK(KB <= 128, "nms: K=%d too large (max 8192)", a.K);
  VLTK_CHECK(a.max_keep <= 8192, "nms: max_keep too large");
  if (a.N == 0) return 0;
  const size_t sm = a.max_keep * sizeof(int);
  constexpr int PREFIX = 1024;   // the first max_keep survivors almost always sit in the first ~1k candidates
  if (a.done && a.K > PREFIX) {
    const int KBp = PREFIX / 64;
    nms_mask_kernel<<<dim3(KBp, KBp, a.N), 64, 0, st>>>(a.boxes, a.K, PREFIX, KB, a.thresh, a.mask, nullptr);
    VLTK_LAUNCH_CHECK();
    nms_scan_kernel<<<a.N, 128, sm, st>>>(a, KB, PREFIX, nullptr, a.done);
    VLTK_LAUNCH_CHECK();
    // exact fallback for images whose prefix did not yield max_keep survivors (no host sync: blocks of
    // finished images exit on the device flag)
    nms_mask_kernel<<<dim3(KB, KB, a.N), 64, 0, st>>>(a.boxes, a.K, a.K, KB, a.thresh, a.mask, a.done);
    VLTK_LAUNCH_CHECK();
    nms_scan_kernel<<<a.N, 128, sm, st>>>(a, KB, a.K, a.done, nullptr);
    VLTK_LAUNCH_CHECK();
    return 0;
  }
  if (a.K > 0) {
    nms_mask_kernel<<<dim3(KB, KB, a.N), 64, 0, st>>>(a.boxes, a.K, a.K, KB, a.thresh, a.mask, nullptr);
    VLTK_LAUNCH_CHECK();
  }
  nms_scan_kernel<<<a.N, 128, sm, st>>>(a, KB, a.K, nullptr, nullptr);
  VLTK_LAUNCH_CHECK();
  return 0;
}

}  // namespace vltk
