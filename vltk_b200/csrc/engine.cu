// C ABI + host-side orchestration of the extraction path (include/vltk_frcnn.h).
// The whole forward is enqueued on one stream with no host synchronisation: dynamic counts
// (non-empty proposals, NMS survivors, detections) stay on the device.
#include <cuda_fp16.h>
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "../../include/vltk_frcnn.h"
#include "conv.cuh"
#include "conv_tc.cuh"
#include "kernels.cuh"

namespace vltk {

static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

namespace {

struct LayerW {           // one conv / linear layer, packed for the kernels
  int cin = 0, cout = 0, k = 1, stride = 1, pad = 0, dil = 1, relu = 0;
  int cin_pad = 0;        // cin as stored in the activation (stem: 3 -> 4)
  int ldw = 0;            // round_up(cout, 4)
  float* w_kn = nullptr;  // SIMT: f32 [K_pad][ldw]
  bf16* w_nk = nullptr;   // tcgen05: bf16 [cout_pad][K]  (bf16 mode, Cin % 64 == 0 only)
  bf16* w_lo = nullptr;   // split-precision layers: w = w_nk (hi) + w_lo
  void* w_h3 = nullptr;   // exact_tc: fp16 [cout_pad][3K] rows of (WA | WB | WC), see conv_tcx.cu
  float* scale_x = nullptr;  // exact_tc: scale[o] * 2^-s[o] (the row scaling of w_h3 folded back), cout_pad entries
  int cout_pad = 0;
  float* scale = nullptr; // [ldw] or nullptr
  float* shift = nullptr; // [ldw]
};

struct Block {
  LayerW c1, c2, c3, sc;
  bool has_sc = false;
  // projection blocks on the tensor pipe: conv3 and shortcut as ONE K-concatenated GEMM (conv_tc.cuh TcConcat).
  // Both frozen-BN scales are folded into the bf16 weights, the shifts are summed.
  LayerW c3f;               // w_nk = bf16(scale3 * W3), scale = nullptr, shift = shift3 + shift_sc
  bf16* scf_w = nullptr;    // bf16(scale_sc * W_sc)  [cout][cin]
  // exact_tc: the same fusion on conv_tcx — w_h3 = planes of the fp32 rows [scale3 * W3 | scale_sc * W_sc], scale_x = the
  // rows' 2^-s, shift = shift3 + shift_sc
  LayerW c3x;
};

struct Tap { const void* p = nullptr; int64_t n = 0; DType dt = DT_F32; int c = 0; };   // c: channels per row (DT_H2)

}  // namespace
}  // namespace vltk

using namespace vltk;

struct vltk_frcnn {
  vltk_frcnn_config cfg;
  int device = 0;
  bool finalized = false;
  bool use_tc = false;
  bool use_tcx = false;    // exact_tc mode: conv_tcx.cu on split-fp16 activations
  bool pack_x = true;      // pack_layer builds exact_tc planes (switched off for the predictor linears, packed separately)
  DType act = DT_F32;
  std::map<std::string, std::vector<float>> host;
  std::vector<void*> owned;  // device allocations
  LayerW stem;
  LayerW stem_tc;          // bf16 mode: stem as a [M,192] x [64,192]^T tensor-core GEMM over im2col rows
  std::vector<std::vector<Block>> stages;  // res2, res3, res4
  std::vector<Block> res5;
  LayerW rpn_conv, rpn_head, cls_score, bbox_pred, fc_attr, attr_score;
  float* rpn_head_shift_simt = nullptr;   // ldw-sized bias of the CUDA-core RPN head (the tensor-pipe copy is padded to 64)
  float* bbox_w_f32 = nullptr;  // exact_tc: bbox_pred.weight [4 NC][D] in fp32 (winner-only evaluation in roi_stats_kernel)
  float* attr_table = nullptr;  // [C+1][512] = emb @ fc_attr.W[:, D:]^T  (bias stays in fc_attr.shift)
  float* cell = nullptr;        // [A,4]
  std::map<std::string, Tap> taps;
  int64_t launches = 0;
  TensorMapCache tmaps;
  // optional per-launch event timing (bench roofline leg)
  bool profiling = false;
  struct ProfRec { int kind; double flops; int64_t M; int K, Cout; cudaEvent_t e0, e1; };
  std::vector<ProfRec> prof;
  std::vector<cudaEvent_t> event_pool;
  // res2-res4 of a batch run as two image halves on two streams (see forward): fork/join plumbing.  One side stream
  // per CALLER stream, so forwards that overlap on different caller streams (different workspaces) keep their second
  // halves independent instead of queueing them on one shared stream.
  struct Side { cudaStream_t caller; cudaStream_t side; cudaEvent_t ev_fork, ev_join; };
  std::vector<Side> sides;
};

namespace vltk_eng {

int dev_alloc(vltk_frcnn* h, void** p, size_t bytes) {
  VLTK_CUDA(cudaMalloc(p, bytes ? bytes : 16));
  h->owned.push_back(*p);
  return 0;
}

int upload(vltk_frcnn* h, const std::vector<float>& v, float** p) {
  if (dev_alloc(h, (void**)p, v.size() * 4)) return -1;
  VLTK_CUDA(cudaMemcpy(*p, v.data(), v.size() * 4, cudaMemcpyHostToDevice));
  return 0;
}

inline float bf16_round(float f) { return __bfloat162float(__float2bfloat16_rn(f)); }

const std::vector<float>* find(vltk_frcnn* h, const std::string& k, int64_t numel) {
  auto it = h->host.find(k);
  if (it == h->host.end()) { set_error("state_dict is missing '%s'", k.c_str()); return nullptr; }
  if ((int64_t)it->second.size() != numel) {
    set_error("'%s' has %lld elements, expected %lld", k.c_str(), (long long)it->second.size(), (long long)numel);
    return nullptr;
  }
  return &it->second;
}

// exact_tc weight planes (conv_tcx.cu): row o of wk ([cout][K], kernel K order) -> fp16 (WA | WB | WC) with a per-row
// power-of-two scale 2^s chosen so that max|w| 2^s lies in [2^12, 2^13) (every plane then sits in fp16's normal
// range): WA = fp16(2^s w), WB = WA 2^-11, WC = fp16(2^s w - WA).  scale_x[o] = base_scale[o] 2^-s undoes the scaling
// in the epilogue (exact: powers of two).  Rows past `cout` are zero with scale 1.
void build_h3(const float* wk, int cout, int K, int cout_pad, const float* base_scale, std::vector<__half>& pl, std::vector<float>& sx) {
  pl.assign((size_t)cout_pad * 3 * K, __float2half_rn(0.f));
  sx.assign(cout_pad, 1.f);
  for (int o = 0; o < cout; ++o) {
    const float* w = wk + (size_t)o * K;
    float mx = 0.f;
    for (int k = 0; k < K; ++k) mx = std::max(mx, fabsf(w[k]));
    int s = 0;
    if (mx > 0.f && std::isfinite(mx)) {
      int e;
      frexpf(mx, &e);                 // mx = m 2^e, m in [0.5, 1)
      s = std::min(std::max(13 - e, -60), 60);
    }
    const float up = ldexpf(1.f, s), dn = ldexpf(1.f, -s);
    __half* row = pl.data() + (size_t)o * 3 * K;
    for (int k = 0; k < K; ++k) {
      const float v = w[k] * up;
      const __half a = __float2half_rn(v);
      const float af = __half2float(a);
      row[k] = a;
      row[K + k] = __float2half_rn(af * (1.f / 2048.f));
      row[2 * K + k] = __float2half_rn(v - af);
    }
    sx[o] = (base_scale ? base_scale[o] : 1.f) * dn;
  }
}

int pack_h3(vltk_frcnn* h, LayerW& L, const std::vector<float>& wk, int cout, int K, int cout_pad, const float* base_scale) {
  L.cout_pad = cout_pad;
  std::vector<__half> pl;
  std::vector<float> sx;
  build_h3(wk.data(), cout, K, cout_pad, base_scale, pl, sx);
  if (dev_alloc(h, &L.w_h3, pl.size() * 2)) return -1;
  VLTK_CUDA(cudaMemcpy(L.w_h3, pl.data(), pl.size() * 2, cudaMemcpyHostToDevice));
  if (upload(h, sx, &L.scale_x)) return -1;
  return 0;
}

// exact_tc linear (fp32 out): W[cout][k_full] uses columns [0, k_used); rows padded to a multiple of 128
int pack_h3_linear(vltk_frcnn* h, LayerW& L, const std::vector<float>& w, int cout, int k_full, int k_used,
                   const std::vector<float>& bias) {
  std::vector<float> wk((size_t)cout * k_used);
  for (int o = 0; o < cout; ++o)
    for (int k = 0; k < k_used; ++k) wk[(size_t)o * k_used + k] = w[(size_t)o * k_full + k];
  if (pack_h3(h, L, wk, cout, k_used, round_up(cout, 128), nullptr)) return -1;
  std::vector<float> sh(L.cout_pad, 0.f);
  for (int o = 0; o < cout; ++o) sh[o] = bias[o];
  if (upload(h, sh, &L.shift)) return -1;
  return 0;
}

// Packs a reference-layout weight [cout][cin][k][k] into the kernel layouts.
//   round_bf16: weights are rounded to bf16 values (bf16 mode), so the SIMT cross-check and the
//   tensor-core kernel see identical operands.
int pack_layer(vltk_frcnn* h, LayerW& L, const std::string& name, int cin, int cout, int k, int stride,
               int pad, int dil, int relu, bool bn, bool bias, bool round_bf16, bool want_tc) {
  L.cin = cin; L.cout = cout; L.k = k; L.stride = stride; L.pad = pad; L.dil = dil; L.relu = relu;
  L.cin_pad = round_up(cin, 4);
  L.ldw = round_up(cout, 4);
  const std::vector<float>* w = find(h, name + ".weight", (int64_t)cout * cin * k * k);
  if (!w) return -2;
  const int K = k * k * L.cin_pad, K_pad = round_up(K, 16);
  std::vector<float> kn((size_t)K_pad * L.ldw, 0.f);
  for (int o = 0; o < cout; ++o)
    for (int c = 0; c < cin; ++c)
      for (int t = 0; t < k * k; ++t) {
        float v = (*w)[((size_t)o * cin + c) * k * k + t];
        if (round_bf16) v = bf16_round(v);
        kn[((size_t)t * L.cin_pad + c) * L.ldw + o] = v;
      }
  if (upload(h, kn, &L.w_kn)) return -1;
  if (want_tc && cin % 64 == 0) {
    L.cout_pad = round_up(cout, 64);
    std::vector<bf16> nk((size_t)L.cout_pad * K, __float2bfloat16_rn(0.f));
    for (int o = 0; o < cout; ++o)
      for (int c = 0; c < cin; ++c)
        for (int t = 0; t < k * k; ++t)
          nk[(size_t)o * K + (size_t)t * cin + c] = __float2bfloat16_rn((*w)[((size_t)o * cin + c) * k * k + t]);
    if (dev_alloc(h, (void**)&L.w_nk, nk.size() * 2)) return -1;
    VLTK_CUDA(cudaMemcpy(L.w_nk, nk.data(), nk.size() * 2, cudaMemcpyHostToDevice));
  }
  std::vector<float> sc(std::max(L.ldw, L.cout_pad), 1.f), sh(std::max(L.ldw, L.cout_pad), 0.f);
  if (bn) {
    const std::string n = name + ".norm";
    const auto* g = find(h, n + ".weight", cout); const auto* b = find(h, n + ".bias", cout);
    const auto* m = find(h, n + ".running_mean", cout); const auto* v = find(h, n + ".running_var", cout);
    if (!g || !b || !m || !v) return -2;
    for (int o = 0; o < cout; ++o) {  // frozen BN, eps 1e-5 (frcnn.py:163-173): y = x*alpha + beta
      float invstd = 1.0f / sqrtf((*v)[o] + 1e-5f);
      sc[o] = (*g)[o] * invstd;
      sh[o] = (*b)[o] - (*m)[o] * sc[o];
    }
    if (upload(h, sc, &L.scale)) return -1;
  }
  if (bias) {
    const auto* b = find(h, name + ".bias", cout);
    if (!b) return -2;
    for (int o = 0; o < cout; ++o) sh[o] = (*b)[o];
  }
  if (bn || bias) { if (upload(h, sh, &L.shift)) return -1; }
  if (h->use_tcx && h->pack_x && cin % 64 == 0) {
    const int K = k * k * cin;
    std::vector<float> wk((size_t)cout * K);
    for (int o = 0; o < cout; ++o)
      for (int c = 0; c < cin; ++c)
        for (int t = 0; t < k * k; ++t) wk[(size_t)o * K + (size_t)t * cin + c] = (*w)[((size_t)o * cin + c) * k * k + t];
    if (pack_h3(h, L, wk, cout, K, round_up(cout, 64), bn ? sc.data() : nullptr)) return -1;
  }
  return 0;
}

// Predictor linears on the tensor pipe keep fp32-faithful logits: W[out][:k_used] is split into
// bf16 hi + lo planes [cout_pad][k_used] (rows past `cout` are zero), bias padded to cout_pad.
int pack_split_linear(vltk_frcnn* h, LayerW& L, const std::vector<float>& w, int cout, int k_full, int k_used,
                      const std::vector<float>& bias) {
  L.cout_pad = round_up(cout, 64);
  std::vector<bf16> hi((size_t)L.cout_pad * k_used, __float2bfloat16_rn(0.f)), lo(hi);
  for (int o = 0; o < cout; ++o)
    for (int k = 0; k < k_used; ++k) {
      float v = w[(size_t)o * k_full + k];
      bf16 b = __float2bfloat16_rn(v);
      hi[(size_t)o * k_used + k] = b;
      lo[(size_t)o * k_used + k] = __float2bfloat16_rn(v - __bfloat162float(b));
    }
  if (dev_alloc(h, (void**)&L.w_nk, hi.size() * 2) || dev_alloc(h, (void**)&L.w_lo, lo.size() * 2)) return -1;
  VLTK_CUDA(cudaMemcpy(L.w_nk, hi.data(), hi.size() * 2, cudaMemcpyHostToDevice));
  VLTK_CUDA(cudaMemcpy(L.w_lo, lo.data(), lo.size() * 2, cudaMemcpyHostToDevice));
  std::vector<float> sh(L.cout_pad, 0.f);
  for (int o = 0; o < cout; ++o) sh[o] = bias[o];
  if (upload(h, sh, &L.shift)) return -1;   // replaces the SIMT-sized copy; both hold the same bias
  return 0;
}

// frozen BN (eps 1e-5, frcnn.py:163-173) as y = x*sc + sh
int bn_fold(vltk_frcnn* h, const std::string& n, int cout, std::vector<float>& sc, std::vector<float>& sh) {
  const auto* g = find(h, n + ".weight", cout); const auto* b = find(h, n + ".bias", cout);
  const auto* m = find(h, n + ".running_mean", cout); const auto* v = find(h, n + ".running_var", cout);
  if (!g || !b || !m || !v) return -2;
  for (int o = 0; o < cout; ++o) {
    const float invstd = 1.0f / sqrtf((*v)[o] + 1e-5f);
    sc[o] = (*g)[o] * invstd;
    sh[o] = (*b)[o] - (*m)[o] * sc[o];
  }
  return 0;
}

int pack_block(vltk_frcnn* h, Block& B, const std::string& p, int cin, int mid, int cout, int stride,
               int dil, bool rb, bool tc) {
  B.has_sc = cin != cout;
  if (B.has_sc && pack_layer(h, B.sc, p + ".shortcut", cin, cout, 1, stride, 0, 1, 0, true, false, rb, tc)) return -1;
  if (pack_layer(h, B.c1, p + ".conv1", cin, mid, 1, stride, 0, 1, 1, true, false, rb, tc)) return -1;
  if (pack_layer(h, B.c2, p + ".conv2", mid, mid, 3, 1, dil, dil, 1, true, false, rb, tc)) return -1;
  // conv3: BN only; the ReLU comes after the residual add and is applied by the same epilogue
  if (pack_layer(h, B.c3, p + ".conv3", mid, cout, 1, 1, 0, 1, 1, true, false, rb, tc)) return -1;
  if (tc && B.has_sc && B.c3.w_nk && B.sc.w_nk && cout % 64 == 0) {
    const auto* w3 = find(h, p + ".conv3.weight", (int64_t)cout * mid);
    const auto* ws = find(h, p + ".shortcut.weight", (int64_t)cout * cin);
    if (!w3 || !ws) return -2;
    std::vector<float> s3(cout), b3(cout), ss(cout), bs(cout);
    if (bn_fold(h, p + ".conv3.norm", cout, s3, b3) || bn_fold(h, p + ".shortcut.norm", cout, ss, bs)) return -2;
    std::vector<bf16> f3((size_t)cout * mid), fs((size_t)cout * cin);
    for (int o = 0; o < cout; ++o) {
      for (int c = 0; c < mid; ++c) f3[(size_t)o * mid + c] = __float2bfloat16_rn(s3[o] * (*w3)[(size_t)o * mid + c]);
      for (int c = 0; c < cin; ++c) fs[(size_t)o * cin + c] = __float2bfloat16_rn(ss[o] * (*ws)[(size_t)o * cin + c]);
      b3[o] += bs[o];
    }
    B.c3f = B.c3;
    B.c3f.scale = nullptr;
    if (dev_alloc(h, (void**)&B.c3f.w_nk, f3.size() * 2) || dev_alloc(h, (void**)&B.scf_w, fs.size() * 2)) return -1;
    VLTK_CUDA(cudaMemcpy(B.c3f.w_nk, f3.data(), f3.size() * 2, cudaMemcpyHostToDevice));
    VLTK_CUDA(cudaMemcpy(B.scf_w, fs.data(), fs.size() * 2, cudaMemcpyHostToDevice));
    if (upload(h, b3, &B.c3f.shift)) return -1;
  }
  if (h->use_tcx && B.has_sc && B.c3.w_h3 && B.sc.w_h3 && cin % 64 == 0 && mid % 64 == 0) {
    const auto* w3 = find(h, p + ".conv3.weight", (int64_t)cout * mid);
    const auto* ws = find(h, p + ".shortcut.weight", (int64_t)cout * cin);
    if (!w3 || !ws) return -2;
    std::vector<float> s3(cout), b3(cout), ss(cout), bs(cout);
    if (bn_fold(h, p + ".conv3.norm", cout, s3, b3) || bn_fold(h, p + ".shortcut.norm", cout, ss, bs)) return -2;
    const int K = mid + cin;
    std::vector<float> wk((size_t)cout * K);
    for (int o = 0; o < cout; ++o) {
      for (int c = 0; c < mid; ++c) wk[(size_t)o * K + c] = s3[o] * (*w3)[(size_t)o * mid + c];
      for (int c = 0; c < cin; ++c) wk[(size_t)o * K + mid + c] = ss[o] * (*ws)[(size_t)o * cin + c];
      b3[o] += bs[o];
    }
    B.c3x = B.c3;
    B.c3x.scale = nullptr;
    if (pack_h3(h, B.c3x, wk, cout, K, round_up(cout, 64), nullptr)) return -1;
    std::vector<float> sh(std::max(B.c3.ldw, B.c3x.cout_pad), 0.f);
    for (int o = 0; o < cout; ++o) sh[o] = b3[o];
    if (upload(h, sh, &B.c3x.shift)) return -1;
  }
  return 0;
}

struct Bump {  // workspace carve-up, 256-byte aligned
  char* base; size_t off = 0, cap;
  Bump(void* b, size_t c) : base((char*)b), cap(c) {}
  void* take(size_t bytes) {
    size_t o = (off + 255) & ~(size_t)255;
    off = o + bytes;
    return base ? base + o : nullptr;
  }
};

size_t esz(DType d) { return d == DT_BF16 ? 2 : 4; }   // per channel element (DT_H2: two fp16 planes)

struct Shapes {
  int N, H, W, Hs, Ws, Hp, Wp, h2, w2, h3, w3, h4, w4, R, P, K;
};

Shapes make_shapes(const vltk_frcnn_config& c, int N, int H, int W) {
  Shapes s;
  s.N = N; s.H = H; s.W = W;
  s.Hs = (H + 6 - 7) / 2 + 1; s.Ws = (W + 6 - 7) / 2 + 1;
  auto pool = [](int n) { int o = (n - 3 + 1) / 2 + 1; if ((o - 1) * 2 >= n) --o; return o; };
  s.Hp = pool(s.Hs); s.Wp = pool(s.Ws);
  s.h2 = s.Hp; s.w2 = s.Wp;
  s.h3 = (s.h2 - 1) / 2 + 1; s.w3 = (s.w2 - 1) / 2 + 1;
  s.h4 = (s.h3 - 1) / 2 + 1; s.w4 = (s.w3 - 1) / 2 + 1;
  s.R = c.rpn_post_nms_topk; s.P = c.pooler_resolution;
  s.K = std::min(c.rpn_pre_nms_topk, s.h4 * s.w4 * c.num_anchors);
  return s;
}

// Runs one layer.  x: [N,H,W,cin_pad]
int run_conv(vltk_frcnn* h, const LayerW& L, const void* x, DType xdt, int N, int H, int W, void* y,
             DType ydt, int ldy, const void* residual, int ldr, int relu, cudaStream_t st, int* oh_out = nullptr,
             int* ow_out = nullptr, const TcPool* pool = nullptr, const TcConcat* cc = nullptr) {
  ConvProblem p;
  memset(&p, 0, sizeof(p));
  p.x = x; p.ldx = L.cin_pad; p.y = y; p.ldy = ldy; p.residual = residual; p.ldr = ldr;
  if (xdt == DT_H2) p.ldx = 2 * L.cin;                 // two planes per pixel row (conv.cuh)
  if (ydt == DT_H2) { p.ldy = 2 * ldy; p.ldr = 2 * ldr; }
  p.N = N; p.H = H; p.W = W; p.Cin = L.cin_pad;
  p.KH = p.KW = L.k; p.stride = L.stride; p.pad = L.pad; p.dil = L.dil;
  p.OH = (H + 2 * L.pad - (L.dil * (L.k - 1) + 1)) / L.stride + 1;
  p.OW = (W + 2 * L.pad - (L.dil * (L.k - 1) + 1)) / L.stride + 1;
  p.Cout = L.cout; p.scale = L.scale; p.shift = L.shift; p.relu = relu;
  p.in_dtype = xdt; p.out_dtype = ydt;
  if (oh_out) *oh_out = p.OH;
  if (ow_out) *ow_out = p.OW;
  h->launches++;
  const bool tc = h->use_tc && L.w_nk && xdt == DT_BF16 && ydt == DT_BF16;
  const bool tcx = h->use_tcx && L.w_h3 && xdt == DT_H2;
  vltk_frcnn::ProfRec rec;
  if (h->profiling) {
    auto get_event = [&]() {
      cudaEvent_t e;
      if (!h->event_pool.empty()) { e = h->event_pool.back(); h->event_pool.pop_back(); }
      else cudaEventCreate(&e);
      return e;
    };
    rec.kind = (tc || tcx) ? 0 : 1;
    rec.M = (int64_t)N * p.OH * p.OW; rec.K = L.k * L.k * L.cin; rec.Cout = L.cout;
    if (cc) rec.K += cc->Cin2;                           // K-concatenated second GEMM (projection shortcut)
    rec.flops = 2.0 * (double)rec.M * rec.K * rec.Cout;  // algorithmic: unpadded cin/cout
    rec.e0 = get_event(); rec.e1 = get_event();
    cudaEventRecord(rec.e0, st);
  }
  if ((pool || cc) && !tc && !tcx) { set_error("internal: fused mean-pool / concat need the tensor-core path"); return -2; }
  int rc;
  if (tcx) {
    VLTK_CHECK(ydt == DT_H2, "internal: run_conv on split-fp16 input writes split-fp16");
    p.scale = L.scale_x;
    rc = conv_tcx_launch(p, L.w_h3, L.cout_pad, &h->tmaps, st, cc, pool);
  } else {
    VLTK_CHECK(xdt != DT_H2 && ydt != DT_H2, "internal: layer %dx%d k%d has no exact_tc weights", L.cin, L.cout, L.k);
    rc = tc ? conv_tc_launch(p, L.w_nk, L.cout_pad, &h->tmaps, st, nullptr, pool, cc) : conv_simt_launch(p, L.w_kn, L.ldw, st);
  }
  if (h->profiling) {
    cudaEventRecord(rec.e1, st);
    h->prof.push_back(rec);
  }
  return rc;
}

// Scratch of the stage entry points (tests / microbenchmarks): released on EVERY exit path, after the stream
// has drained — an early error return must not leak device memory or free buffers a queued kernel still uses.
struct Scratch {
  cudaStream_t st;
  std::vector<void*> ptrs;
  explicit Scratch(cudaStream_t s) : st(s) {}
  Scratch(const Scratch&) = delete;
  Scratch& operator=(const Scratch&) = delete;
  ~Scratch() {
    cudaStreamSynchronize(st);
    for (void* q : ptrs) cudaFree(q);
  }
  template <typename T>
  cudaError_t get(T** out, size_t bytes) {
    cudaError_t e = cudaMalloc((void**)out, bytes ? bytes : 16);
    if (e == cudaSuccess) ptrs.push_back((void*)*out);
    return e;
  }
};

// Brackets a non-GEMM launch with events when profiling (kind >= 2; `bytes` = algorithmic bytes moved)
struct StageTimer {
  vltk_frcnn* h; cudaStream_t st; bool on;
  vltk_frcnn::ProfRec rec;
  StageTimer(vltk_frcnn* h_, int kind, double bytes, cudaStream_t s) : h(h_), st(s), on(h_->profiling) {
    if (!on) return;
    rec.kind = kind; rec.flops = 0.0; rec.M = (int64_t)bytes; rec.K = 0; rec.Cout = 0;
    auto get = [&]() { cudaEvent_t e; if (!h->event_pool.empty()) { e = h->event_pool.back(); h->event_pool.pop_back(); } else cudaEventCreate(&e); return e; };
    rec.e0 = get(); rec.e1 = get();
    cudaEventRecord(rec.e0, st);
  }
  ~StageTimer() {
    if (!on) return;
    cudaEventRecord(rec.e1, st);
    h->prof.push_back(rec);
  }
};
enum { K_TC = 0, K_SIMT = 1, K_ROIPOOL = 2, K_SELECT = 3, K_NMS = 4, K_MEAN = 5, K_TAIL = 6, K_MAXPOOL = 7, K_LAYOUT = 8, K_GLUE = 9, K_NUM = 10 };
static const char* kKindName[K_NUM] = {"tcgen05", "simt", "roi_pool", "rpn_select", "rpn_nms", "mean_rows", "roi_tail", "maxpool", "layout", "predictor_glue"};

// bottleneck (frcnn.py:963-979): x -> conv1 -> conv2 -> conv3 (+shortcut(x) | x) -> relu
int run_block(vltk_frcnn* h, const Block& B, const void* x, int N, int H, int W, void* out, void* t1,
              void* t2, void* sbuf, cudaStream_t st, int* oh, int* ow, const TcPool* pool = nullptr) {
  const DType d = h->act;
  int h1, w1;
  if (run_conv(h, B.c1, x, d, N, H, W, t1, d, B.c1.ldw, nullptr, 0, 1, st, &h1, &w1)) return -1;
  if (run_conv(h, B.c2, t1, d, N, h1, w1, t2, d, B.c2.ldw, nullptr, 0, 1, st)) return -1;
  // projection block on the tensor pipe: out = relu(t2 * W3' + x (*) Wsc' + shift) in ONE launch; the shortcut
  // tensor never exists (saves its HBM write + re-read and a launch).  VLTK_FUSE_SC=0 restores the two-launch form.
  static const bool fuse_sc = [] {
    const char* e = getenv("VLTK_FUSE_SC"); const char* v1 = getenv("VLTK_TC_V1");
    return !(e && e[0] == '0') && !(v1 && v1[0] == '1');
  }();
  if (fuse_sc && B.has_sc && B.scf_w && h->use_tc && d == DT_BF16 && !pool) {
    TcConcat cc;
    cc.x2 = x; cc.ldx2 = B.sc.cin_pad; cc.H2 = H; cc.W2 = W; cc.Cin2 = B.sc.cin; cc.stride2 = B.sc.stride; cc.w2 = B.scf_w;
    if (run_conv(h, B.c3f, t2, d, N, h1, w1, out, d, B.c3.ldw, nullptr, 0, 1, st, nullptr, nullptr, nullptr, &cc)) return -1;
    *oh = h1; *ow = w1;
    return 0;
  }
  if (fuse_sc && B.has_sc && B.c3x.w_h3 && h->use_tcx && d == DT_H2 && !pool) {
    TcConcat cc;
    cc.x2 = x; cc.ldx2 = 2 * B.sc.cin; cc.H2 = H; cc.W2 = W; cc.Cin2 = B.sc.cin; cc.stride2 = B.sc.stride;
    if (run_conv(h, B.c3x, t2, d, N, h1, w1, out, d, B.c3.ldw, nullptr, 0, 1, st, nullptr, nullptr, nullptr, &cc)) return -1;
    *oh = h1; *ow = w1;
    return 0;
  }
  const void* res = x;
  int ldr = B.c3.ldw;
  if (B.has_sc) {
    if (run_conv(h, B.sc, x, d, N, H, W, sbuf, d, B.sc.ldw, nullptr, 0, 0, st)) return -1;
    res = sbuf;
  }
  if (run_conv(h, B.c3, t2, d, N, h1, w1, out, d, B.c3.ldw, res, ldr, 1, st, nullptr, nullptr, pool)) return -1;
  *oh = h1; *ow = w1;
  return 0;
}

// NCHW f32 <-> NHWC helpers for the stage entry points
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, int C, int HW, int ldy, int coff, int64_t tot) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= tot) return;
  int c = (int)(i % C);
  int64_t t = i / C;
  int p = (int)(t % HW);
  int n = (int)(t / HW);
  y[((int64_t)n * HW + p) * ldy + coff + c] = x[((int64_t)n * C + c) * HW + p];
}
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ x, float* __restrict__ y, int C, int HW, int64_t tot) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= tot) return;
  int p = (int)(i % HW);
  int64_t t = i / HW;
  int c = (int)(t % C);
  int n = (int)(t / C);
  y[i] = x[((int64_t)n * HW + p) * C + c];
}
__global__ void rois5_split_kernel(const float* __restrict__ r5, int R, float* __restrict__ boxes, int* __restrict__ bidx) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R) return;
  bidx[i] = (int)r5[5 * i];
  reinterpret_cast<float4*>(boxes)[i] = make_float4(r5[5 * i + 1], r5[5 * i + 2], r5[5 * i + 3], r5[5 * i + 4]);
}
__global__ void widen_bf16_kernel(const bf16* __restrict__ x, float* __restrict__ y, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = __bfloat162float(x[i]);
}

void tap(vltk_frcnn* h, const char* name, const void* p, int64_t n, DType dt, int c = 0) {
  Tap t; t.p = p; t.n = n; t.dt = dt; t.c = c;
  h->taps[name] = t;
}

}  // namespace vltk_eng
using namespace vltk_eng;

// =========================================================================================
extern "C" {

const char* vltk_frcnn_last_error(void) { return get_error(); }
const char* vltk_frcnn_version(void) { return "vltk_b200 frcnn 0.1 (sm_100a)"; }

int vltk_frcnn_create(const vltk_frcnn_config* cfg, int device, vltk_frcnn_t** out) {
  VLTK_CHECK(cfg && out, "create: null argument");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    set_error("no CUDA device available (%s): this library has no CPU fallback", cudaGetErrorString(e));
    return -3;
  }
  VLTK_CHECK(device >= 0 && device < ndev, "create: device %d out of range", device);
  VLTK_CHECK(cfg->rpn_pre_nms_topk >= 1 && cfg->rpn_pre_nms_topk <= 8192, "rpn_pre_nms_topk must be in 1..8192");
  VLTK_CHECK(cfg->rpn_post_nms_topk >= 1 && cfg->rpn_post_nms_topk <= 512, "rpn_post_nms_topk must be in 1..512");
  VLTK_CHECK(cfg->mode == VLTK_MODE_FP32 || cfg->mode == VLTK_MODE_BF16 || cfg->mode == VLTK_MODE_EXACT_TC, "unknown mode %d", cfg->mode);
  VLTK_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  VLTK_CUDA(cudaGetDeviceProperties(&prop, device));
  VLTK_CHECK(prop.major == 10, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
  vltk_frcnn* h = new vltk_frcnn();
  h->cfg = *cfg;
  h->device = device;
  h->act = cfg->mode == VLTK_MODE_BF16 ? DT_BF16 : (cfg->mode == VLTK_MODE_EXACT_TC ? DT_H2 : DT_F32);
  h->use_tcx = cfg->mode == VLTK_MODE_EXACT_TC;
  const char* notc = getenv("VLTK_NO_TC");
  h->use_tc = cfg->mode == VLTK_MODE_BF16 && !(notc && notc[0] == '1');
  *out = h;
  return 0;
}

void vltk_frcnn_destroy(vltk_frcnn_t* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  for (void* p : h->owned) cudaFree(p);
  for (auto& r : h->prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  for (auto e : h->event_pool) cudaEventDestroy(e);
  for (auto& sd : h->sides) { cudaStreamDestroy(sd.side); cudaEventDestroy(sd.ev_fork); cudaEventDestroy(sd.ev_join); }
  delete h;
}

int vltk_frcnn_load_tensor(vltk_frcnn_t* h, const char* name, const float* data, int64_t numel) {
  VLTK_CHECK(h && name && data && numel >= 0, "load_tensor: bad argument");
  VLTK_CHECK(!h->finalized, "load_tensor: weights already finalized");
  h->host[name].assign(data, data + numel);
  return 0;
}

int vltk_frcnn_finalize(vltk_frcnn_t* h) {
  VLTK_CHECK(h && !h->finalized, "finalize: bad handle or already finalized");
  VLTK_CUDA(cudaSetDevice(h->device));
  const vltk_frcnn_config& c = h->cfg;
  const bool rb = h->act == DT_BF16, tc = h->use_tc;
  // stem consumes the fp32 NHWC4 image directly; its weights stay fp32 in both modes
  if (pack_layer(h, h->stem, "backbone.stem.conv1", 3, c.stem_out_channels, 7, 2, 3, 1, 1, true, false, false, false)) return -1;
  if (tc && c.stem_out_channels % 64 == 0) {
    const int so = c.stem_out_channels;
    const auto* w = find(h, "backbone.stem.conv1.weight", (int64_t)so * 3 * 49);
    if (!w) return -2;
    LayerW& L = h->stem_tc;
    L.cin = L.cin_pad = 192; L.cout = so; L.cout_pad = so; L.k = 1; L.ldw = so;
    std::vector<bf16> nk((size_t)so * 192, __float2bfloat16_rn(0.f));
    for (int o = 0; o < so; ++o)
      for (int ci = 0; ci < 3; ++ci)
        for (int t = 0; t < 49; ++t)   // k = tap*3 + c, matching stem_im2col
          nk[(size_t)o * 192 + t * 3 + ci] = __float2bfloat16_rn((*w)[((size_t)o * 3 + ci) * 49 + t]);
    if (dev_alloc(h, (void**)&L.w_nk, nk.size() * 2)) return -1;
    VLTK_CUDA(cudaMemcpy(L.w_nk, nk.data(), nk.size() * 2, cudaMemcpyHostToDevice));
    L.scale = h->stem.scale; L.shift = h->stem.shift;   // same folded BN (64 entries, so == ldw)
  }
  int cin = c.stem_out_channels, cout = c.res2_out_channels, mid = c.res2_out_channels / 4;
  h->stages.resize(3);
  for (int s = 0; s < 3; ++s) {
    h->stages[s].resize(c.blocks[s]);
    for (int b = 0; b < c.blocks[s]; ++b) {
      char p[64];
      snprintf(p, sizeof(p), "backbone.res%d.%d", s + 2, b);
      if (pack_block(h, h->stages[s][b], p, b == 0 ? cin : cout, mid, cout, (b == 0 && s > 0) ? 2 : 1, 1, rb, tc)) return -1;
    }
    cin = cout; cout *= 2; mid *= 2;
  }
  h->res5.resize(c.res5_blocks);
  for (int b = 0; b < c.res5_blocks; ++b) {  // VG head: stride 1, conv2 dilation 2 (frcnn.py:1345-1355)
    char p[64];
    snprintf(p, sizeof(p), "roi_heads.res5.%d", b);
    if (pack_block(h, h->res5[b], p, b == 0 ? cin : cout, mid, cout, 1, 2, rb, tc)) return -1;
  }
  const int c4 = cin, D = cout, A = c.num_anchors, hid = c.rpn_hidden;
  if (pack_layer(h, h->rpn_conv, "proposal_generator.rpn_head.conv", c4, hid, 3, 1, 1, 1, 1, false, true, rb, tc)) return -1;
  {  // fused 1x1 head: columns [0,4A) anchor deltas, [4A,5A) objectness (frcnn.py:1569-1571)
    const auto* wd = find(h, "proposal_generator.rpn_head.anchor_deltas.weight", (int64_t)4 * A * hid);
    const auto* bd = find(h, "proposal_generator.rpn_head.anchor_deltas.bias", 4 * A);
    const auto* wo = find(h, "proposal_generator.rpn_head.objectness_logits.weight", (int64_t)A * hid);
    const auto* bo = find(h, "proposal_generator.rpn_head.objectness_logits.bias", A);
    if (!wd || !bd || !wo || !bo) return -2;
    LayerW& L = h->rpn_head;
    L.cin = L.cin_pad = hid; L.cout = 5 * A; L.k = 1; L.ldw = round_up(5 * A, 4);
    std::vector<float> kn((size_t)round_up(hid, 16) * L.ldw, 0.f), sh(L.ldw, 0.f);
    for (int k = 0; k < hid; ++k) {
      for (int o = 0; o < 4 * A; ++o) kn[(size_t)k * L.ldw + o] = (*wd)[(size_t)o * hid + k];
      for (int o = 0; o < A; ++o) kn[(size_t)k * L.ldw + 4 * A + o] = (*wo)[(size_t)o * hid + k];
    }
    for (int o = 0; o < 4 * A; ++o) sh[o] = (*bd)[o];
    for (int o = 0; o < A; ++o) sh[4 * A + o] = (*bo)[o];
    if (upload(h, kn, &L.w_kn) || upload(h, sh, &L.shift)) return -1;
    if ((tc || h->use_tcx) && hid % 64 == 0) {   // tensor pipe, fp32 out: bf16 mode exact bf16 activations x (w_hi + w_lo), rows padded to 64; exact_tc three fp16 planes, rows padded to 128
      std::vector<float> wrow((size_t)5 * A * hid), brow(5 * A);
      for (int o = 0; o < 4 * A; ++o) { brow[o] = (*bd)[o]; for (int k = 0; k < hid; ++k) wrow[(size_t)o * hid + k] = (*wd)[(size_t)o * hid + k]; }
      for (int o = 0; o < A; ++o) { brow[4 * A + o] = (*bo)[o]; for (int k = 0; k < hid; ++k) wrow[(size_t)(4 * A + o) * hid + k] = (*wo)[(size_t)o * hid + k]; }
      float* simt_shift = L.shift;
      if (h->use_tcx ? pack_h3_linear(h, L, wrow, 5 * A, hid, hid, brow)
                     : pack_split_linear(h, L, wrow, 5 * A, hid, hid, brow)) return -1;   // sets the planes / cout_pad and a padded shift
      h->rpn_head_shift_simt = simt_shift;
    }
  }
  {
    const auto* ca = find(h, "proposal_generator.anchor_generator.cell_anchors.0", 4 * A);
    if (!ca) return -2;
    if (upload(h, *ca, &h->cell)) return -1;
  }
  // predictor (frcnn.py:1726-1740): always fp32 — its argmaxes decide ids
  const int NC = c.num_classes, NA = c.num_attrs, E = D / 8, HA = D / 4;
  h->pack_x = false;
  if (pack_layer(h, h->cls_score, "roi_heads.box_predictor.cls_score", D, NC + 1, 1, 1, 0, 1, 0, false, true, false, false)) return -1;
  if (pack_layer(h, h->bbox_pred, "roi_heads.box_predictor.bbox_pred", D, NC * 4, 1, 1, 0, 1, 0, false, true, false, false)) return -1;
  if (pack_layer(h, h->attr_score, "roi_heads.box_predictor.attr_score", HA, NA + 1, 1, 1, 0, 1, 0, false, true, false, false)) return -1;
  {  // fc_attr(cat[x, emb[c]]) = W[:, :D] x + (W[:, D:] emb[c]) + b: the concat becomes a class-indexed bias
    const auto* w = find(h, "roi_heads.box_predictor.fc_attr.weight", (int64_t)HA * (D + E));
    const auto* b = find(h, "roi_heads.box_predictor.fc_attr.bias", HA);
    const auto* emb = find(h, "roi_heads.box_predictor.cls_embedding.weight", (int64_t)(NC + 1) * E);
    if (!w || !b || !emb) return -2;
    LayerW& L = h->fc_attr;
    L.cin = L.cin_pad = D; L.cout = HA; L.k = 1; L.ldw = HA; L.relu = 1;
    std::vector<float> kn((size_t)D * HA), sh(b->begin(), b->end());
    for (int o = 0; o < HA; ++o)
      for (int k = 0; k < D; ++k) kn[(size_t)k * HA + o] = (*w)[(size_t)o * (D + E) + k];
    std::vector<float> tab((size_t)(NC + 1) * HA);
    for (int cidx = 0; cidx <= NC; ++cidx)
      for (int o = 0; o < HA; ++o) {
        double s = 0.0;
        for (int e = 0; e < E; ++e) s += (double)(*w)[(size_t)o * (D + E) + D + e] * (double)(*emb)[(size_t)cidx * E + e];
        tab[(size_t)cidx * HA + o] = (float)s;
      }
    if (upload(h, kn, &L.w_kn) || upload(h, sh, &L.shift) || upload(h, tab, &h->attr_table)) return -1;
    if (tc && pack_split_linear(h, L, *w, HA, D + E, D, *b)) return -1;
    if (h->use_tcx && pack_h3_linear(h, L, *w, HA, D + E, D, *b)) return -1;
  }
  if (h->use_tcx) {   // cls_score / attr_score on the tensor pipe (fp32-faithful); bbox_pred stays on the CUDA cores
    const auto* wc = find(h, "roi_heads.box_predictor.cls_score.weight", (int64_t)(NC + 1) * D);
    const auto* bc = find(h, "roi_heads.box_predictor.cls_score.bias", NC + 1);
    const auto* wa = find(h, "roi_heads.box_predictor.attr_score.weight", (int64_t)(NA + 1) * HA);
    const auto* ba = find(h, "roi_heads.box_predictor.attr_score.bias", NA + 1);
    if (!wc || !bc || !wa || !ba) return -2;
    if (pack_h3_linear(h, h->cls_score, *wc, NC + 1, D, D, *bc)) return -1;
    if (pack_h3_linear(h, h->attr_score, *wa, NA + 1, HA, HA, *ba)) return -1;
    // bbox_pred is only ever needed for each ROI's winning class (frcnn.py:116-131): keep W [4 NC][D] in fp32 for roi_stats_kernel
    const auto* wb = find(h, "roi_heads.box_predictor.bbox_pred.weight", (int64_t)NC * 4 * D);
    if (!wb) return -2;
    if (upload(h, *wb, &h->bbox_w_f32)) return -1;
  }
  if (tc) {
    const auto* wc = find(h, "roi_heads.box_predictor.cls_score.weight", (int64_t)(NC + 1) * D);
    const auto* bc = find(h, "roi_heads.box_predictor.cls_score.bias", NC + 1);
    const auto* wb = find(h, "roi_heads.box_predictor.bbox_pred.weight", (int64_t)NC * 4 * D);
    const auto* bb = find(h, "roi_heads.box_predictor.bbox_pred.bias", NC * 4);
    const auto* wa = find(h, "roi_heads.box_predictor.attr_score.weight", (int64_t)(NA + 1) * HA);
    const auto* ba = find(h, "roi_heads.box_predictor.attr_score.bias", NA + 1);
    if (!wc || !bc || !wb || !bb || !wa || !ba) return -2;
    if (pack_split_linear(h, h->cls_score, *wc, NC + 1, D, D, *bc)) return -1;
    if (pack_split_linear(h, h->bbox_pred, *wb, NC * 4, D, D, *bb)) return -1;
    if (pack_split_linear(h, h->attr_score, *wa, NA + 1, HA, HA, *ba)) return -1;
  }
  h->host.clear();
  h->finalized = true;
  return 0;
}

static size_t plan(vltk_frcnn* h, const Shapes& s, void* base, size_t cap, void** ptrs);

enum {
  B_IN4, B_STEM, B_POOL, B_A, B_B, B_T1, B_T2, B_S, B_RPNH, B_HEAD, B_SIZES, B_SCALES, B_SBOX, B_SSCORE,
  B_SIDX, B_SVALID, B_MASK, B_PROP, B_PSCORE, B_PIDX, B_COUNT, B_POOLED, B_R5A, B_R5B, B_R5T1, B_R5T2,
  B_R5S, B_FEATS, B_CLS, B_BBOX, B_ARGMAX, B_TG, B_AH, B_ATTR, B_FHI, B_FLO, B_AHHI, B_AHLO, B_STEMA, B_PARTIAL, B_NMSDONE, B_ROISTAT, B_RES4, B_IGNOREY, B_NUM
};

static size_t plan(vltk_frcnn* h, const Shapes& s, void* base, size_t cap, void** p) {
  const vltk_frcnn_config& c = h->cfg;
  const size_t e = esz(h->act);
  Bump b(base, cap);
  const int64_t N = s.N;
  const int c2 = c.res2_out_channels, D = c2 * 8, NR = (int)(N * s.R), PP = s.P * s.P;
  p[B_IN4] = b.take((size_t)N * s.H * s.W * 4 * 4);
  p[B_STEM] = b.take((size_t)N * s.Hs * s.Ws * c.stem_out_channels * e);
  p[B_POOL] = b.take((size_t)N * s.Hp * s.Wp * c.stem_out_channels * e);
  // backbone ping-pong: largest block output / mid tensor over res2..res4
  size_t big = std::max({(size_t)N * s.h2 * s.w2 * c2, (size_t)N * s.h3 * s.w3 * c2 * 2, (size_t)N * s.h4 * s.w4 * c2 * 4});
  size_t midsz = std::max({(size_t)N * s.h2 * s.w2 * (c2 / 4), (size_t)N * s.h3 * s.w3 * (c2 / 2), (size_t)N * s.h4 * s.w4 * c2});
  p[B_A] = b.take(big * e); p[B_B] = b.take(big * e); p[B_S] = b.take(big * e);
  p[B_T1] = b.take(midsz * e); p[B_T2] = b.take(midsz * e);
  p[B_RPNH] = b.take((size_t)N * s.h4 * s.w4 * c.rpn_hidden * e);
  p[B_HEAD] = b.take((size_t)N * s.h4 * s.w4 * std::max(h->rpn_head.ldw, h->rpn_head.cout_pad) * 4);
  p[B_SIZES] = b.take((size_t)N * 2 * 4); p[B_SCALES] = b.take((size_t)N * 2 * 4);
  p[B_SBOX] = b.take((size_t)N * s.K * 16); p[B_SSCORE] = b.take((size_t)N * s.K * 4);
  p[B_SIDX] = b.take((size_t)N * s.K * 4); p[B_SVALID] = b.take((size_t)N * s.K);
  p[B_MASK] = b.take(nms_mask_bytes((int)N, s.K));
  p[B_PROP] = b.take((size_t)NR * 16); p[B_PSCORE] = b.take((size_t)NR * 4);
  p[B_PIDX] = b.take((size_t)NR * 4); p[B_COUNT] = b.take((size_t)N * 4);
  p[B_POOLED] = b.take((size_t)NR * PP * c2 * 4 * e);
  p[B_R5A] = b.take((size_t)NR * PP * D * e); p[B_R5B] = b.take((size_t)NR * PP * D * e);
  p[B_R5S] = b.take((size_t)NR * PP * D * e);
  p[B_R5T1] = b.take((size_t)NR * PP * (D / 4) * e); p[B_R5T2] = b.take((size_t)NR * PP * (D / 4) * e);
  p[B_FEATS] = b.take((size_t)NR * D * 4);
  auto ld = [](const LayerW& L) { return (size_t)std::max(L.ldw, L.cout_pad); };  // split layers pad to 64
  p[B_CLS] = b.take((size_t)NR * ld(h->cls_score) * 4);
  p[B_BBOX] = b.take((size_t)NR * ld(h->bbox_pred) * 4);
  p[B_ARGMAX] = b.take((size_t)NR * 4);
  p[B_TG] = b.take((size_t)NR * (D / 4) * 4); p[B_AH] = b.take((size_t)NR * (D / 4) * 4);
  p[B_ATTR] = b.take((size_t)NR * ld(h->attr_score) * 4);
  // predictor inputs: bf16 hi / lo planes (bf16 mode) or one split-fp16 tensor in B_FHI / B_AHHI (exact_tc)
  const size_t pe = h->use_tcx ? 4 : 2;
  p[B_FHI] = b.take((size_t)NR * D * pe); p[B_FLO] = b.take((size_t)NR * D * 2);
  p[B_AHHI] = b.take((size_t)NR * (D / 4) * pe); p[B_AHLO] = b.take((size_t)NR * (D / 4) * 2);
  p[B_STEMA] = b.take(h->stem_tc.w_nk ? (size_t)N * s.Hs * s.Ws * 192 * 2 : 0);
  p[B_PARTIAL] = b.take((h->use_tc || h->use_tcx) ? conv_tc_pool_partial_bytes((int64_t)NR * PP, D) : 0);
  p[B_NMSDONE] = b.take((size_t)N * 4);
  p[B_ROISTAT] = b.take((size_t)NR * 32);
  p[B_RES4] = b.take((size_t)N * s.h4 * s.w4 * c2 * 4 * e);
  p[B_IGNOREY] = b.take((size_t)N * 16 * 2 * 4);
  return b.off + 256;
}

// BasicStem (frcnn.py:872-879): conv7x7 s2 p3 + BN + ReLU + 3x3 s2 ceil-mode max-pool, images NCHW f32 -> pooled NHWC in
// the engine's activation type.  Buffers: b_in4 [N,H,W,4] f32, b_stema im2col rows (bf16 mode), b_stem, b_pool.
static int run_stem(vltk_frcnn* h, const float* images, int n, int height, int width, const Shapes& s, void* b_in4,
                    void* b_stema, void* b_stem, void* b_pool, cudaStream_t st) {
  const vltk_frcnn_config& c = h->cfg;
  const DType d = h->act;
  const bool stem_on_tc = h->use_tc && h->stem_tc.w_nk;
  static const bool im2col_direct = [] { const char* e = getenv("VLTK_STEM_NHWC4"); return !(e && e[0] == '1'); }();
  if (!(stem_on_tc && im2col_direct)) {
    StageTimer t(h, K_LAYOUT, (double)n * height * width * (12 + 16), st);
    if (nchw3_to_nhwc4(images, b_in4, DT_F32, n, height, width, st)) return -1;
    h->launches++;
  }
  if (stem_on_tc) {
    const int64_t Ms = (int64_t)n * s.Hs * s.Ws;
    { StageTimer t(h, K_LAYOUT, (double)n * height * width * 12.0 + (double)Ms * 384.0, st);
      if (im2col_direct ? stem_im2col_nchw(images, b_stema, n, height, width, s.Hs, s.Ws, st)
                        : stem_im2col((const float*)b_in4, b_stema, n, height, width, s.Hs, s.Ws, st)) return -1; }
    h->launches++;
    ConvProblem q;
    memset(&q, 0, sizeof(q));
    const LayerW& L = h->stem_tc;
    q.x = b_stema; q.ldx = 192; q.y = b_stem; q.ldy = L.cout; q.N = (int)Ms; q.H = q.W = q.OH = q.OW = 1;
    q.Cin = 192; q.Cout = L.cout; q.KH = q.KW = 1; q.stride = 1; q.dil = 1; q.scale = L.scale; q.shift = L.shift;
    q.relu = 1; q.in_dtype = DT_BF16; q.out_dtype = DT_BF16;
    vltk_frcnn::ProfRec rec;
    if (h->profiling) {
      rec.kind = 0; rec.M = Ms; rec.K = 147; rec.Cout = L.cout; rec.flops = 2.0 * (double)Ms * 147 * L.cout;
      cudaEventCreate(&rec.e0); cudaEventCreate(&rec.e1); cudaEventRecord(rec.e0, st);
    }
    h->launches++;
    if (conv_tc_launch(q, L.w_nk, L.cout_pad, &h->tmaps, st)) return -1;
    if (h->profiling) { cudaEventRecord(rec.e1, st); h->prof.push_back(rec); }
  } else if (run_conv(h, h->stem, b_in4, DT_F32, n, height, width, b_stem, d == DT_H2 ? DT_F32 : d, h->stem.ldw, nullptr, 0, 1, st)) return -1;
  { StageTimer t(h, K_MAXPOOL, ((double)n * s.Hs * s.Ws + (double)n * s.Hp * s.Wp) * c.stem_out_channels * esz(d), st);
    // exact_tc: the 3-channel stem runs in fp32 on the CUDA cores (0.14 % of the FLOPs); the pool splits its output
    if (d == DT_H2 ? maxpool3x3s2_ceil_f32_to_h2((const float*)b_stem, b_pool, n, s.Hs, s.Ws, c.stem_out_channels, s.Hp, s.Wp, st)
                   : maxpool3x3s2_ceil(b_stem, b_pool, d, n, s.Hs, s.Ws, c.stem_out_channels, s.Hp, s.Wp, st)) return -1; }
  h->launches++;
  return 0;
}

// RPNHead (frcnn.py:1561-1572): relu(conv3x3 + b) then the fused 1x1 head -> fp32 rows of *ldh_out columns:
// [0,4A) anchor deltas, [4A,5A) objectness.
static int run_rpn_head(vltk_frcnn* h, const void* res4, int n, int h4, int w4, void* b_hidden, void* b_head, int* ldh_out,
                        cudaStream_t st) {
  const DType d = h->act;
  if (run_conv(h, h->rpn_conv, res4, d, n, h4, w4, b_hidden, d, h->rpn_conv.ldw, nullptr, 0, 1, st)) return -1;
  int ldh = h->rpn_head.ldw;
  if (h->use_tc && h->rpn_head.w_lo && d == DT_BF16) {
    // 1x1 head on the tensor pipe, fp32-faithful: the RPN conv output IS bf16, so x*(w_hi + w_lo) in one fp32 TMEM tile
    const LayerW& L = h->rpn_head;
    const int64_t Mh = (int64_t)n * h4 * w4;
    ldh = L.cout_pad;
    ConvProblem q;
    memset(&q, 0, sizeof(q));
    q.x = b_hidden; q.ldx = L.cin_pad; q.y = b_head; q.ldy = ldh; q.N = (int)Mh; q.H = q.W = q.OH = q.OW = 1;
    q.Cin = L.cin_pad; q.Cout = L.cout_pad; q.KH = q.KW = 1; q.stride = 1; q.dil = 1; q.shift = L.shift; q.relu = 0;
    q.in_dtype = DT_BF16; q.out_dtype = DT_F32;
    TcSplit sp; sp.x_lo = nullptr; sp.w_lo = L.w_lo;
    vltk_frcnn::ProfRec rec;
    if (h->profiling) {
      rec.kind = 0; rec.M = Mh; rec.K = L.cin; rec.Cout = L.cout; rec.flops = 2.0 * (double)Mh * L.cin * L.cout;
      cudaEventCreate(&rec.e0); cudaEventCreate(&rec.e1); cudaEventRecord(rec.e0, st);
    }
    h->launches++;
    if (conv_tc_launch(q, L.w_nk, L.cout_pad, &h->tmaps, st, &sp)) return -1;
    if (h->profiling) { cudaEventRecord(rec.e1, st); h->prof.push_back(rec); }
  } else if (h->use_tcx && h->rpn_head.w_h3 && d == DT_H2) {
    // 1x1 head on the tensor pipe, fp32-faithful, fp32 rows of cout_pad (= 128) columns
    const LayerW& L = h->rpn_head;
    const int64_t Mh = (int64_t)n * h4 * w4;
    ldh = L.cout_pad;
    ConvProblem q;
    memset(&q, 0, sizeof(q));
    q.x = b_hidden; q.ldx = 2 * L.cin_pad; q.y = b_head; q.ldy = ldh; q.N = (int)Mh; q.H = q.W = q.OH = q.OW = 1;
    q.Cin = L.cin_pad; q.Cout = L.cout_pad; q.KH = q.KW = 1; q.stride = 1; q.dil = 1; q.scale = L.scale_x; q.shift = L.shift; q.relu = 0;
    q.in_dtype = DT_H2; q.out_dtype = DT_F32;
    vltk_frcnn::ProfRec rec;
    if (h->profiling) {
      rec.kind = 0; rec.M = Mh; rec.K = L.cin; rec.Cout = L.cout; rec.flops = 2.0 * (double)Mh * L.cin * L.cout;
      cudaEventCreate(&rec.e0); cudaEventCreate(&rec.e1); cudaEventRecord(rec.e0, st);
    }
    h->launches++;
    if (conv_tcx_launch(q, L.w_h3, L.cout_pad, &h->tmaps, st)) return -1;
    if (h->profiling) { cudaEventRecord(rec.e1, st); h->prof.push_back(rec); }
  } else {
    LayerW L = h->rpn_head;
    if (h->rpn_head_shift_simt) L.shift = h->rpn_head_shift_simt;
    if (run_conv(h, L, b_hidden, d, n, h4, w4, b_head, DT_F32, L.ldw, nullptr, 0, 0, st)) return -1;
  }
  *ldh_out = ldh;
  return 0;
}

size_t vltk_frcnn_workspace_bytes(vltk_frcnn_t* h, int n, int height, int width) {
  if (!h || !h->finalized || n < 0) return 0;
  void* p[B_NUM];
  return plan(h, make_shapes(h->cfg, n, height, width), nullptr, 0, p);
}

int vltk_frcnn_forward(vltk_frcnn_t* h, const float* images, const int32_t* sizes_hw, const float* scales_yx,
                       int n, int height, int width, const vltk_frcnn_knobs* knobs, const vltk_frcnn_out* out,
                       void* workspace, size_t workspace_bytes, void* stream) {
  VLTK_CHECK(h && h->finalized, "forward: weights not finalized");
  VLTK_CHECK(images && sizes_hw && knobs && out && workspace, "forward: null argument");
  VLTK_CHECK(n >= 1, "forward: empty batch");
  VLTK_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const vltk_frcnn_config& c = h->cfg;
  const Shapes s = make_shapes(c, n, height, width);
  VLTK_CHECK(s.h4 >= 1 && s.w4 >= 1, "forward: image %dx%d too small", height, width);
  for (int i = 0; i < n; ++i)
    VLTK_CHECK(sizes_hw[2 * i] >= 1 && sizes_hw[2 * i] <= height && sizes_hw[2 * i + 1] >= 1 && sizes_hw[2 * i + 1] <= width,
               "forward: image_shapes[%d]=(%d,%d) outside the padded %dx%d batch", i, sizes_hw[2 * i], sizes_hw[2 * i + 1], height, width);
  VLTK_CHECK(knobs->max_detections >= 1 && knobs->max_detections <= s.R, "forward: max_detections=%d must be in 1..%d", knobs->max_detections, s.R);
  void* p[B_NUM];
  size_t need = plan(h, s, workspace, workspace_bytes, p);
  VLTK_CHECK(need <= workspace_bytes, "forward: workspace %zu < required %zu", workspace_bytes, need);
  const DType d = h->act;
  h->taps.clear();

  VLTK_CUDA(cudaMemcpyAsync(p[B_SIZES], sizes_hw, (size_t)n * 8, cudaMemcpyHostToDevice, st));
  if (scales_yx) VLTK_CUDA(cudaMemcpyAsync(p[B_SCALES], scales_yx, (size_t)n * 8, cudaMemcpyHostToDevice, st));
  const bool use_ignorey = knobs->ignorey && knobs->n_ignorey > 0 && scales_yx;   // frcnn.py:328
  if (use_ignorey) {
    VLTK_CHECK(knobs->n_ignorey <= 16, "forward: at most 16 ignorey ranges per image (got %d)", knobs->n_ignorey);
    VLTK_CUDA(cudaMemcpyAsync(p[B_IGNOREY], knobs->ignorey, (size_t)n * knobs->n_ignorey * 8, cudaMemcpyHostToDevice, st));
  }

  // ---- backbone (frcnn.py:1076-1090)
  if (run_stem(h, images, n, height, width, s, p[B_IN4], p[B_STEMA], p[B_STEM], p[B_POOL], st)) return -1;
  const void* x = p[B_POOL];
  int ch = s.Hp, cw = s.Wp;
  void* pp[2] = {p[B_A], p[B_B]};
  int flip = 0;
  // res2-res4 are ~90 SMALL launches (res4: 19 152 pixel rows = 150 row tiles on 148 SMs, 20-40 us each, half of it
  // launch ramp and a second wave of 2 tiles).  Images are independent up to the RPN, so on the tensor pipe the
  // batch is split into two image halves that run the same layers on two streams: each half's kernels need ~75
  // CTAs, the two streams' kernels pack the 148 SMs together and overlap each other's ramps and tails.  Every
  // output row is computed from its own input rows only, so the result is bit-identical to the unsplit order.
  // VLTK_SPLIT_BACKBONE=0 restores the single-stream order.
  static const bool want_split = [] { const char* e = getenv("VLTK_SPLIT_BACKBONE"); return !(e && e[0] == '0'); }();
  const bool split = want_split && (h->use_tc || h->use_tcx) && n >= 2;
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  if (split) {
    for (auto& sd : h->sides)
      if (sd.caller == st) { side = sd.side; ev_fork = sd.ev_fork; ev_join = sd.ev_join; }
    if (!side) {
      if (h->sides.size() >= 8) {                     // callers that keep creating streams: recycle the oldest entry
        cudaStreamSynchronize(h->sides[0].side);
        cudaStreamDestroy(h->sides[0].side); cudaEventDestroy(h->sides[0].ev_fork); cudaEventDestroy(h->sides[0].ev_join);
        h->sides.erase(h->sides.begin());
      }
      vltk_frcnn::Side sd;
      sd.caller = st;
      VLTK_CUDA(cudaStreamCreateWithFlags(&sd.side, cudaStreamNonBlocking));
      VLTK_CUDA(cudaEventCreateWithFlags(&sd.ev_fork, cudaEventDisableTiming));
      VLTK_CUDA(cudaEventCreateWithFlags(&sd.ev_join, cudaEventDisableTiming));
      h->sides.push_back(sd);
      side = sd.side; ev_fork = sd.ev_fork; ev_join = sd.ev_join;
    }
  }
  const int nA = split ? (n + 1) / 2 : n, nB = n - nA;
  if (split) {
    VLTK_CUDA(cudaEventRecord(ev_fork, st));
    VLTK_CUDA(cudaStreamWaitEvent(side, ev_fork, 0));
  }
  {
    // Half B lives at a FIXED offset (= half A's capacity) in every shared scratch buffer, so the two halves' regions
    // stay disjoint even when one stream is a stage ahead of the other (tensor shapes — and with them the natural
    // batch offsets — change from stage to stage).  Only the first block's input (maxpool output) and the last
    // block's output (the res4 map, its own buffer) are addressed as one contiguous [N,H,W,C] tensor.
    const size_t e = esz(d);
    const size_t c2 = c.res2_out_channels;
    const size_t f_big = std::max({(size_t)s.h2 * s.w2 * c2, (size_t)s.h3 * s.w3 * c2 * 2, (size_t)s.h4 * s.w4 * c2 * 4});
    const size_t f_mid = std::max({(size_t)s.h2 * s.w2 * (c2 / 4), (size_t)s.h3 * s.w3 * (c2 / 2), (size_t)s.h4 * s.w4 * c2});
    const size_t offB_big = (size_t)nA * f_big * e, offB_mid = (size_t)nA * f_mid * e;
    size_t nblk = 0, bi = 0;
    for (auto& stage : h->stages) nblk += stage.size();
    int cin_x = c.stem_out_channels;      // channels of the current block input
    for (auto& stage : h->stages)
      for (auto& blk : stage) {
        const bool first = bi == 0, last = bi + 1 == nblk;
        const int h1 = (ch - 1) / blk.c1.stride + 1, w1 = (cw - 1) / blk.c1.stride + 1;
        char* outA = last ? (char*)p[B_RES4] : (char*)pp[flip];
        int oh = 0, ow = 0;
        if (run_block(h, blk, x, nA, ch, cw, outA, p[B_T1], p[B_T2], p[B_S], st, &oh, &ow)) return -1;
        if (nB > 0) {
          const char* xB = (const char*)x + (first ? (size_t)nA * ch * cw * cin_x * e : offB_big);
          char* outB = last ? (char*)p[B_RES4] + (size_t)nA * h1 * w1 * blk.c3.ldw * e : (char*)pp[flip] + offB_big;
          int oh2, ow2;
          if (run_block(h, blk, xB, nB, ch, cw, outB, (char*)p[B_T1] + offB_mid, (char*)p[B_T2] + offB_mid,
                        (char*)p[B_S] + offB_big, side, &oh2, &ow2)) return -1;
        }
        x = outA; flip ^= 1; ch = oh; cw = ow; cin_x = blk.c3.ldw; ++bi;
      }
  }
  if (split) {
    VLTK_CUDA(cudaEventRecord(ev_join, side));
    VLTK_CUDA(cudaStreamWaitEvent(st, ev_join, 0));
  }
  VLTK_CHECK(ch == s.h4 && cw == s.w4, "internal: res4 shape mismatch");
  const void* res4 = x;
  const int C4 = c.res2_out_channels * 4, A = c.num_anchors;
  tap(h, "res4", res4, (int64_t)n * s.h4 * s.w4 * C4, d, C4);

  // ---- RPN head + proposal selection (frcnn.py:1561-1572, 264-390)
  int ldh = 0;
  if (run_rpn_head(h, res4, n, s.h4, s.w4, p[B_RPNH], p[B_HEAD], &ldh, st)) return -1;
  tap(h, "rpn_head", p[B_HEAD], (int64_t)n * s.h4 * s.w4 * ldh, DT_F32);
  RpnSelectArgs ra;
  memset(&ra, 0, sizeof(ra));
  ra.head = (const float*)p[B_HEAD]; ra.ldh = ldh; ra.delta_off = 0; ra.logit_off = 4 * A;
  ra.N = n; ra.H4 = s.h4; ra.W4 = s.w4; ra.A = A; ra.stride = c.anchor_stride; ra.cell = h->cell;
  ra.sizes_hw = (const int*)p[B_SIZES]; ra.pre_topk = c.rpn_pre_nms_topk; ra.min_size = c.rpn_min_size;
  ra.wx = c.rpn_bbox_weights[0]; ra.wy = c.rpn_bbox_weights[1]; ra.ww = c.rpn_bbox_weights[2]; ra.wh = c.rpn_bbox_weights[3];
  ra.boxes = (float*)p[B_SBOX]; ra.scores = (float*)p[B_SSCORE]; ra.anchor_idx = (int*)p[B_SIDX];
  ra.valid = (uint8_t*)p[B_SVALID]; ra.K = s.K;
  if (use_ignorey) { ra.ignorey = (const float*)p[B_IGNOREY]; ra.scales_yx = (const float*)p[B_SCALES]; ra.J = knobs->n_ignorey; }
  { StageTimer t(h, K_SELECT, (double)n * s.h4 * s.w4 * A * 20.0 + (double)n * s.K * 20.0, st);
    if (rpn_select(ra, st)) return -1; }
  NmsArgs na;
  memset(&na, 0, sizeof(na));
  na.boxes = ra.boxes; na.scores = ra.scores; na.valid = ra.valid; na.N = n; na.K = s.K;
  na.thresh = c.rpn_nms_thresh; na.max_keep = s.R; na.mask = (unsigned long long*)p[B_MASK];
  na.out_boxes = (float*)p[B_PROP]; na.out_scores = (float*)p[B_PSCORE]; na.out_idx = (int*)p[B_PIDX];
  na.out_count = (int*)p[B_COUNT]; na.done = (int*)p[B_NMSDONE];
  { StageTimer t(h, K_NMS, (double)n * s.K * 20.0 + (double)n * s.R * 20.0, st);
    if (nms_sorted(na, st)) return -1; }
  h->launches += (s.K > 1024) ? 5 : 3;   // select + (prefix mask/scan + guarded full mask/scan | mask/scan)
  tap(h, "topk_anchor_idx", p[B_SIDX], (int64_t)n * s.K, DT_F32);  // int32 payload, read raw
  tap(h, "proposals", p[B_PROP], (int64_t)n * s.R * 4, DT_F32);
  tap(h, "proposal_logits", p[B_PSCORE], (int64_t)n * s.R, DT_F32);
  tap(h, "proposal_count", p[B_COUNT], n, DT_F32);
  tap(h, "proposal_pos", p[B_PIDX], (int64_t)n * s.R, DT_F32);  // int32: position in the sorted top-k list

  // ---- ROI head: RoIPool -> res5 -> mean (frcnn.py:1387-1403)
  const int NR = n * s.R, PP = s.P * s.P, D = c.res2_out_channels * 8;
  { StageTimer t(h, K_ROIPOOL, ((double)n * s.h4 * s.w4 * C4 + (double)n * s.R * s.P * s.P * C4) * esz(d), st);
    if (d == DT_H2 ? roi_pool_h2(res4, n, s.h4, s.w4, C4, (const float*)p[B_PROP], (const int*)p[B_COUNT], s.R, s.P,
                                 1.0f / (float)c.anchor_stride, p[B_POOLED], st)
                   : roi_pool(res4, d, n, s.h4, s.w4, C4, (const float*)p[B_PROP], (const int*)p[B_COUNT], s.R, s.P,
                              1.0f / (float)c.anchor_stride, p[B_POOLED], st)) return -1; }
  h->launches++;
  tap(h, "pooled", p[B_POOLED], (int64_t)NR * PP * C4, d, C4);
  x = p[B_POOLED];
  void* r5[2] = {p[B_R5A], p[B_R5B]};
  flip = 0;
  // The last block's output is only ever consumed by the 14x14 mean (frcnn.py:1401).  On the tensor pipe its conv3
  // epilogue reduces the fp32 tile per ROI instead of storing 2 GB that is read straight back (conv_tc.cu POOL,
  // tested in test_fused_meanpool_epilogue_matches_conv_then_mean).  Row tiles are ROI-aligned (two per ROI: rows
  // [0,128) and [128,196)), so an ROI's sums do not depend on its position in the batch and every image stays
  // bit-independent of its batch neighbours (test_full_batch8_equals_smaller_batches).  VLTK_FUSE_MEAN=0 restores the
  // separate mean_rows pass over the stored bf16 tensor.
  static const bool want_fuse = [] { const char* e = getenv("VLTK_FUSE_MEAN"); return !(e && e[0] == '0'); }();
  const LayerW& tail = h->res5.back().c3;
  const bool fuse_mean = want_fuse && PP > 128 && PP <= 256 && D % 256 == 0 && tail.cin > 256 && !h->res5.back().has_sc &&
                         ((h->use_tc && tail.w_nk) || (h->use_tcx && tail.w_h3));
  TcPool pool;
  pool.out = (float*)p[B_FEATS]; pool.partial = (float*)p[B_PARTIAL]; pool.rows = PP;
  for (size_t bi = 0; bi < h->res5.size(); ++bi) {
    int oh, ow;
    const bool last = bi + 1 == h->res5.size();
    if (run_block(h, h->res5[bi], x, NR, s.P, s.P, r5[flip], p[B_R5T1], p[B_R5T2], p[B_R5S], st, &oh, &ow,
                  (last && fuse_mean) ? &pool : nullptr)) return -1;
    x = r5[flip]; flip ^= 1;
  }
  if (!fuse_mean) {
    StageTimer t(h, K_MEAN, (double)NR * PP * D * esz(d) + (double)NR * D * 4, st);
    if (d == DT_H2 ? mean_rows_h2(x, (float*)p[B_FEATS], NR, PP, D, st) : mean_rows(x, (float*)p[B_FEATS], d, NR, PP, D, st)) return -1;
  }
  h->launches++;
  tap(h, "feats", p[B_FEATS], (int64_t)NR * D, DT_F32);

  // ---- predictor (frcnn.py:1726-1740): fp32 on the CUDA cores, or fp32-faithful split-bf16 (hi*hi +
  //      lo*hi + hi*lo, fp32 accumulate and fp32 logits) on the tensor pipe
  const bool ptc = h->use_tc && h->cls_score.w_lo;
  const bool ptx = h->use_tcx && h->cls_score.w_h3 && h->fc_attr.w_h3 && h->attr_score.w_h3;
  const int ldc = (ptc || ptx) ? h->cls_score.cout_pad : h->cls_score.ldw;
  const int ldb = ptc ? h->bbox_pred.cout_pad : h->bbox_pred.ldw;
  const int lda = (ptc || ptx) ? h->attr_score.cout_pad : h->attr_score.ldw;
  if (ptx) {
    // exact_tc: fp32-faithful GEMMs on split-fp16 inputs (conv_tcx.cu); bbox_pred in fp32 on the CUDA cores
    auto h2_gemm = [&](const LayerW& L, const void* xh2, int K, void* y, int relu) -> int {
      ConvProblem q;
      memset(&q, 0, sizeof(q));
      q.x = xh2; q.ldx = 2 * K; q.y = y; q.ldy = L.cout_pad; q.N = NR; q.H = q.W = q.OH = q.OW = 1; q.Cin = K;
      q.Cout = L.cout_pad; q.KH = q.KW = 1; q.stride = 1; q.dil = 1; q.scale = L.scale_x; q.shift = L.shift; q.relu = relu;
      q.in_dtype = DT_H2; q.out_dtype = DT_F32;
      h->launches++;
      vltk_frcnn::ProfRec rec;
      if (h->profiling) {
        rec.kind = 0; rec.M = NR; rec.K = K; rec.Cout = L.cout; rec.flops = 2.0 * NR * (double)K * L.cout;
        cudaEventCreate(&rec.e0); cudaEventCreate(&rec.e1); cudaEventRecord(rec.e0, st);
      }
      int rc = conv_tcx_launch(q, L.w_h3, L.cout_pad, &h->tmaps, st);
      if (h->profiling) { cudaEventRecord(rec.e1, st); h->prof.push_back(rec); }
      return rc;
    };
    { StageTimer t(h, K_GLUE, (double)NR * D * 8.0, st);
      if (split_f32_h2((const float*)p[B_FEATS], nullptr, 0, p[B_FHI], NR, D, st)) return -1; }
    if (h2_gemm(h->cls_score, p[B_FHI], D, p[B_CLS], 0)) return -1;
    // bbox_pred: only the winning class's 4 rows are ever used (frcnn.py:116-131) -> evaluated inside roi_stats_kernel (fp32)
    if (row_argmax((const float*)p[B_CLS], ldc, NR, c.num_classes + 1, (int*)p[B_ARGMAX], st)) return -1;
    if (gather_rows(h->attr_table, D / 4, (const int*)p[B_ARGMAX], NR, D / 4, (float*)p[B_TG], D / 4, st)) return -1;
    if (h2_gemm(h->fc_attr, p[B_FHI], D, p[B_AH], 0)) return -1;      // W[:, :D] x + b
    if (split_f32_h2((const float*)p[B_AH], (const float*)p[B_TG], 1, p[B_AHHI], NR, D / 4, st)) return -1;   // + T[argmax], ReLU, split
    if (h2_gemm(h->attr_score, p[B_AHHI], D / 4, p[B_ATTR], 0)) return -1;
    h->launches += 4;
  } else if (ptc) {
    auto split_gemm = [&](const LayerW& L, const void* xhi, const void* xlo, int K, void* y, int relu) -> int {
      ConvProblem q;
      memset(&q, 0, sizeof(q));
      q.x = xhi; q.ldx = K; q.y = y; q.ldy = L.cout_pad; q.N = NR; q.H = q.W = q.OH = q.OW = 1; q.Cin = K;
      q.Cout = L.cout_pad; q.KH = q.KW = 1; q.stride = 1; q.dil = 1; q.shift = L.shift; q.relu = relu;
      q.in_dtype = DT_BF16; q.out_dtype = DT_F32;
      TcSplit sp; sp.x_lo = xlo; sp.w_lo = L.w_lo;
      h->launches++;
      vltk_frcnn::ProfRec rec;
      if (h->profiling) {
        rec.kind = 0; rec.M = NR; rec.K = K; rec.Cout = L.cout; rec.flops = 2.0 * NR * (double)K * L.cout;  // algorithmic (1 pass)
        cudaEventCreate(&rec.e0); cudaEventCreate(&rec.e1); cudaEventRecord(rec.e0, st);
      }
      int rc = conv_tc_launch(q, L.w_nk, L.cout_pad, &h->tmaps, st, &sp);
      if (h->profiling) { cudaEventRecord(rec.e1, st); h->prof.push_back(rec); }
      return rc;
    };
    { StageTimer t(h, K_GLUE, (double)NR * D * 8.0, st);
      if (split_f32((const float*)p[B_FEATS], nullptr, 0, (bf16*)p[B_FHI], (bf16*)p[B_FLO], (int64_t)NR * D, st)) return -1; }
    if (split_gemm(h->cls_score, p[B_FHI], p[B_FLO], D, p[B_CLS], 0)) return -1;
    // bbox_pred: only the winning class's 4 rows are ever used (frcnn.py:116-131) -> evaluated inside roi_stats_kernel
    if (row_argmax((const float*)p[B_CLS], ldc, NR, c.num_classes + 1, (int*)p[B_ARGMAX], st)) return -1;
    if (gather_rows(h->attr_table, D / 4, (const int*)p[B_ARGMAX], NR, D / 4, (float*)p[B_TG], D / 4, st)) return -1;
    if (split_gemm(h->fc_attr, p[B_FHI], p[B_FLO], D, p[B_AH], 0)) return -1;   // W[:, :D] x + b
    // + T[argmax class], ReLU, and the hi/lo split of the hidden vector, in one pass
    if (split_f32((const float*)p[B_AH], (const float*)p[B_TG], 1, (bf16*)p[B_AHHI], (bf16*)p[B_AHLO], (int64_t)NR * (D / 4), st)) return -1;
    if (split_gemm(h->attr_score, p[B_AHHI], p[B_AHLO], D / 4, p[B_ATTR], 0)) return -1;
    h->launches += 4;
  } else {
    if (run_conv(h, h->cls_score, p[B_FEATS], DT_F32, NR, 1, 1, p[B_CLS], DT_F32, ldc, nullptr, 0, 0, st)) return -1;
    if (run_conv(h, h->bbox_pred, p[B_FEATS], DT_F32, NR, 1, 1, p[B_BBOX], DT_F32, ldb, nullptr, 0, 0, st)) return -1;
    if (row_argmax((const float*)p[B_CLS], ldc, NR, c.num_classes + 1, (int*)p[B_ARGMAX], st)) return -1;
    if (gather_rows(h->attr_table, D / 4, (const int*)p[B_ARGMAX], NR, D / 4, (float*)p[B_TG], D / 4, st)) return -1;
    if (run_conv(h, h->fc_attr, p[B_FEATS], DT_F32, NR, 1, 1, p[B_AH], DT_F32, D / 4, p[B_TG], D / 4, 1, st)) return -1;
    if (run_conv(h, h->attr_score, p[B_AH], DT_F32, NR, 1, 1, p[B_ATTR], DT_F32, lda, nullptr, 0, 0, st)) return -1;
    h->launches += 2;
  }
  tap(h, "cls_logits", p[B_CLS], (int64_t)NR * ldc, DT_F32);
  if (!ptc && !ptx) tap(h, "bbox_deltas", p[B_BBOX], (int64_t)NR * ldb, DT_F32);
  tap(h, "attr_logits", p[B_ATTR], (int64_t)NR * lda, DT_F32);

  // ---- detection tail (frcnn.py:1262-1294)
  TailArgs ta;
  memset(&ta, 0, sizeof(ta));
  ta.N = n; ta.R = s.R; ta.cls_logits = (const float*)p[B_CLS]; ta.ldc = ldc;
  ta.bbox_deltas = (ptc || ptx) ? nullptr : (const float*)p[B_BBOX]; ta.ldb = ldb;
  if (ptc) { ta.bbox_w_hi = h->bbox_pred.w_nk; ta.bbox_w_lo = h->bbox_pred.w_lo; ta.bbox_bias = h->bbox_pred.shift; }
  if (ptx) { ta.bbox_w_f32 = h->bbox_w_f32; ta.bbox_bias = h->bbox_pred.shift; }
  ta.attr_logits = (const float*)p[B_ATTR]; ta.lda = lda;
  ta.feats = (const float*)p[B_FEATS]; ta.D = D; ta.proposals = (const float*)p[B_PROP];
  ta.count = (const int*)p[B_COUNT]; ta.sizes_hw = (const int*)p[B_SIZES];
  ta.scales_yx = scales_yx ? (const float*)p[B_SCALES] : nullptr;
  ta.num_classes = c.num_classes; ta.num_attrs = c.num_attrs;
  ta.wx = c.roi_bbox_weights[0]; ta.wy = c.roi_bbox_weights[1]; ta.ww = c.roi_bbox_weights[2]; ta.wh = c.roi_bbox_weights[3];
  ta.nms_thresh = knobs->nms_thresh; ta.n_thresh = knobs->n_nms_thresh;
  ta.min_det = knobs->min_detections; ta.max_det = knobs->max_detections; ta.pad_value = knobs->pad_value;
  ta.boxes = out->boxes; ta.norm_boxes = out->normalized_boxes; ta.obj_ids = (long long*)out->obj_ids;
  ta.obj_probs = out->obj_probs; ta.attr_ids = (long long*)out->attr_ids; ta.attr_probs = out->attr_probs;
  ta.roi_features = out->roi_features; ta.preds_per_image = out->preds_per_image; ta.keep_idx = out->keep_idx;
  ta.stats = (float*)p[B_ROISTAT];
  { StageTimer t(h, K_TAIL, (double)NR * (ldc + lda + 4 + 4) * 4.0 + (double)n * knobs->max_detections * (D + 12) * 4.0, st);
    if (roi_tail(ta, st)) return -1; }
  h->launches += 2;   // roi_stats + roi_tail
  return 0;
}

int vltk_frcnn_preprocess(const uint8_t* raw, int rh, int rw, int nh, int nw, const float mean[3], const float stdv[3],
                          float pad_value, float* images, int index, int height, int width, void* stream) {
  VLTK_CHECK(raw && images && mean && stdv, "preprocess: null argument");
  VLTK_CHECK(nh <= height && nw <= width && nh >= 1 && nw >= 1, "preprocess: resized %dx%d exceeds canvas %dx%d", nh, nw, height, width);
  return preprocess_image(raw, rh, rw, nh, nw, height, width, mean, stdv, pad_value,
                          images + (int64_t)index * 3 * height * width, nullptr, DT_F32, (cudaStream_t)stream);
}

int64_t vltk_frcnn_debug_read(vltk_frcnn_t* h, const char* name, float* dst, int64_t cap) {
  VLTK_CHECK(h && name && dst, "debug_read: null argument");
  auto it = h->taps.find(name);
  VLTK_CHECK(it != h->taps.end(), "debug_read: no tap '%s' (run forward first)", name);
  const Tap& t = it->second;
  VLTK_CHECK(t.n <= cap, "debug_read: '%s' has %lld elements, capacity %lld", name, (long long)t.n, (long long)cap);
  VLTK_CUDA(cudaSetDevice(h->device));
  VLTK_CUDA(cudaDeviceSynchronize());
  if (t.dt == DT_F32) {
    VLTK_CUDA(cudaMemcpy(dst, t.p, (size_t)t.n * 4, cudaMemcpyDeviceToHost));
  } else {
    float* tmp = nullptr;
    VLTK_CUDA(cudaMalloc(&tmp, (size_t)t.n * 4));
    if (t.dt == DT_H2) widen_h2(t.p, tmp, t.n / t.c, t.c, 0);
    else widen_bf16_kernel<<<(unsigned)ceil_div64(t.n, 256), 256>>>((const bf16*)t.p, tmp, t.n);
    cudaError_t e = cudaMemcpy(dst, tmp, (size_t)t.n * 4, cudaMemcpyDeviceToHost);
    cudaFree(tmp);
    VLTK_CUDA(e);
  }
  return t.n;
}

int64_t vltk_frcnn_launch_count(vltk_frcnn_t* h) { return h ? h->launches : -1; }

int vltk_frcnn_profile_enable(vltk_frcnn_t* h, int enable) {
  VLTK_CHECK(h, "profile_enable: null handle");
  h->profiling = enable != 0;
  return 0;
}

int vltk_frcnn_profile_read(vltk_frcnn_t* h, double* agg, char* csv, size_t cap) {
  VLTK_CHECK(h && agg, "profile_read: null argument");
  VLTK_CUDA(cudaSetDevice(h->device));
  VLTK_CUDA(cudaDeviceSynchronize());
  for (int i = 0; i < 6; ++i) agg[i] = 0.0;
  size_t off = 0;
  if (csv && cap) csv[0] = 0;
  // Launches of the two backbone half-batch streams overlap in time: a kind's busy time is the UNION of its
  // launches' [start, end] intervals (each measured against the first recorded event), not their sum.
  std::vector<std::pair<float, float>> iv[2];
  for (auto& r : h->prof) {
    float ms = 0.f, t0 = 0.f;
    cudaEventElapsedTime(&ms, r.e0, r.e1);
    cudaEventElapsedTime(&t0, h->prof.front().e0, r.e0);
    if (r.kind < 2) { iv[r.kind].push_back({t0, t0 + ms}); agg[3 * r.kind + 1] += r.flops; agg[3 * r.kind + 2] += 1.0; }
    if (csv && off + 96 < cap)
      off += snprintf(csv + off, cap - off, "%s,%lld,%d,%d,%.5f\n", kKindName[r.kind < K_NUM ? r.kind : 0], (long long)r.M, r.K, r.Cout, ms);
  }
  for (int k = 0; k < 2; ++k) {
    std::sort(iv[k].begin(), iv[k].end());
    float lo = 0.f, hi = -1.f;
    for (auto& x : iv[k]) {
      if (hi < lo || x.first > hi) { if (hi >= lo) agg[3 * k] += hi - lo; lo = x.first; hi = x.second; }
      else hi = std::max(hi, x.second);
    }
    if (hi >= lo) agg[3 * k] += hi - lo;
  }
  for (auto& r : h->prof) { h->event_pool.push_back(r.e0); h->event_pool.push_back(r.e1); }
  h->prof.clear();
  return 0;
}

// ---------------------------------------------------------------------------- stage entries
int vltk_conv2d_nhwc(const void* x, const float* weight, const float* scale, const float* shift, const void* residual,
                     void* y, int n, int hh, int ww, int cin, int cout, int kh, int kw, int stride, int pad, int dil,
                     int relu, int mode, int use_tc, void* stream) {
  VLTK_CHECK(x && weight && y, "conv2d: null argument");
  VLTK_CHECK(kh == kw, "conv2d: square kernels only");
  VLTK_CHECK(cin % 4 == 0 && cout % 4 == 0, "conv2d: cin/cout must be multiples of 4 for the stage entry");
  cudaStream_t st = (cudaStream_t)stream;
  Scratch scratch(st);
  const DType d = mode == VLTK_MODE_BF16 ? DT_BF16 : DT_F32;
  ConvProblem p;
  memset(&p, 0, sizeof(p));
  p.x = x; p.ldx = cin; p.y = y; p.ldy = cout; p.residual = residual; p.ldr = cout;
  p.N = n; p.H = hh; p.W = ww; p.Cin = cin; p.KH = kh; p.KW = kw; p.stride = stride; p.pad = pad; p.dil = dil;
  p.OH = (hh + 2 * pad - (dil * (kh - 1) + 1)) / stride + 1;
  p.OW = (ww + 2 * pad - (dil * (kw - 1) + 1)) / stride + 1;
  p.Cout = cout; p.scale = scale; p.shift = shift; p.relu = relu; p.in_dtype = d; p.out_dtype = d;
  const int K = kh * kw * cin, K_pad = round_up(K, 16);
  int rc = 0;
  if (mode == VLTK_MODE_EXACT_TC) {
    // fp32 NHWC tensors in and out; split / widened around the split-fp16 tensor-core kernel.  use_tc = 2: fp32 rows out
    // straight from the kernel's fp32 epilogue (the predictor / RPN-head form; cout % 128 == 0, no residual).
    VLTK_CHECK(cin % 64 == 0 && cout % 64 == 0, "conv2d(exact_tc): cin and cout must be multiples of 64");
    const bool f32out = use_tc == 2;
    const int64_t Min = (int64_t)n * hh * ww, Mout = (int64_t)n * p.OH * p.OW;
    std::vector<float> hw((size_t)cout * K), wk((size_t)cout * K), hs(cout, 1.f);
    VLTK_CUDA(cudaStreamSynchronize(st));
    VLTK_CUDA(cudaMemcpy(hw.data(), weight, hw.size() * 4, cudaMemcpyDeviceToHost));
    if (scale) VLTK_CUDA(cudaMemcpy(hs.data(), scale, (size_t)cout * 4, cudaMemcpyDeviceToHost));
    for (int o = 0; o < cout; ++o)
      for (int c = 0; c < cin; ++c)
        for (int t = 0; t < kh * kw; ++t) wk[(size_t)o * K + (size_t)t * cin + c] = hw[((size_t)o * cin + c) * kh * kw + t];
    std::vector<__half> pl;
    std::vector<float> sx;
    build_h3(wk.data(), cout, K, cout, hs.data(), pl, sx);
    void *w3 = nullptr, *xh = nullptr, *yh = nullptr, *rh = nullptr;
    float* dsx = nullptr;
    VLTK_CUDA(scratch.get(&w3, pl.size() * 2));
    VLTK_CUDA(scratch.get(&dsx, sx.size() * 4));
    VLTK_CUDA(scratch.get(&xh, (size_t)Min * cin * 4));
    VLTK_CUDA(scratch.get(&yh, (size_t)Mout * cout * 4));
    VLTK_CUDA(cudaMemcpyAsync(w3, pl.data(), pl.size() * 2, cudaMemcpyHostToDevice, st));
    VLTK_CUDA(cudaMemcpyAsync(dsx, sx.data(), sx.size() * 4, cudaMemcpyHostToDevice, st));
    rc = split_f32_h2((const float*)x, nullptr, 0, xh, Min, cin, st);
    if (!rc && residual) {
      VLTK_CUDA(scratch.get(&rh, (size_t)Mout * cout * 4));
      rc = split_f32_h2((const float*)residual, nullptr, 0, rh, Mout, cout, st);
    }
    p.x = xh; p.ldx = 2 * cin; p.in_dtype = DT_H2; p.scale = dsx;
    if (f32out) { p.out_dtype = DT_F32; p.ldy = cout; p.residual = nullptr; }
    else { p.y = yh; p.ldy = 2 * cout; p.out_dtype = DT_H2; p.residual = rh; p.ldr = 2 * cout; }
    TensorMapCache cache;
    if (!rc) rc = conv_tcx_launch(p, w3, cout, &cache, st);
    if (!rc && !f32out) rc = widen_h2(yh, (float*)y, Mout, cout, st);
    cudaStreamSynchronize(st);
    if (!rc) VLTK_LAUNCH_CHECK();
    return rc;
  }
  if (use_tc) {
    VLTK_CHECK(d == DT_BF16 && cin % 64 == 0, "conv2d: tensor-core path needs bf16 and cin %% 64 == 0");
    const int cout_pad = round_up(cout, 64);
    bf16* w_nk = nullptr;
    float *sc = nullptr, *sh = nullptr;
    VLTK_CUDA(scratch.get(&w_nk, (size_t)cout_pad * K * 2));
    VLTK_CUDA(cudaMemsetAsync(w_nk, 0, (size_t)cout_pad * K * 2, st));
    VLTK_CUDA(scratch.get(&sc, (size_t)cout_pad * 4));
    VLTK_CUDA(scratch.get(&sh, (size_t)cout_pad * 4));
    rc = pack_weight_nk(weight, w_nk, cout, cin, kh * kw, st);
    if (!rc) rc = pad_vector(scale, sc, cout, cout_pad, 1.f, st);
    if (!rc) rc = pad_vector(shift, sh, cout, cout_pad, 0.f, st);
    p.scale = sc; p.shift = sh;
    TensorMapCache cache;
    if (!rc) rc = conv_tc_launch(p, w_nk, cout_pad, &cache, st);
    cudaStreamSynchronize(st);
  } else {
    float* w_kn = nullptr;
    VLTK_CUDA(scratch.get(&w_kn, (size_t)K_pad * cout * 4));
    VLTK_CUDA(cudaMemsetAsync(w_kn, 0, (size_t)K_pad * cout * 4, st));
    rc = pack_weight_kn(weight, w_kn, cout, cin, kh * kw, cout, d == DT_BF16, st);
    if (!rc) rc = conv_simt_launch(p, w_kn, cout, st);
    cudaStreamSynchronize(st);
  }
  return rc;
}

int vltk_conv2d_meanpool_nhwc(const void* x, const float* weight, const float* scale, const float* shift,
                              const void* residual, float* pooled, int n, int hh, int ww, int cin, int cout, int kh,
                              int stride, int pad, int dil, int relu, int pool_rows, void* stream) {
  VLTK_CHECK(x && weight && residual && pooled, "conv2d_meanpool: null argument");
  VLTK_CHECK(cin % 64 == 0 && cout % 256 == 0, "conv2d_meanpool: cin %% 64 and cout %% 256 required");
  cudaStream_t st = (cudaStream_t)stream;
  Scratch scratch(st);
  ConvProblem p;
  memset(&p, 0, sizeof(p));
  p.x = x; p.ldx = cin; p.residual = residual; p.ldr = cout; p.ldy = cout;
  p.N = n; p.H = hh; p.W = ww; p.Cin = cin; p.KH = p.KW = kh; p.stride = stride; p.pad = pad; p.dil = dil;
  p.OH = (hh + 2 * pad - (dil * (kh - 1) + 1)) / stride + 1;
  p.OW = (ww + 2 * pad - (dil * (kh - 1) + 1)) / stride + 1;
  p.Cout = cout; p.relu = relu; p.in_dtype = DT_BF16; p.out_dtype = DT_BF16;
  const int64_t M = (int64_t)n * p.OH * p.OW;
  const int K = kh * kh * cin;
  bf16 *w_nk = nullptr, *ydummy = nullptr;
  float *sc = nullptr, *sh = nullptr, *partial = nullptr;
  VLTK_CUDA(scratch.get(&w_nk, (size_t)cout * K * 2));
  VLTK_CUDA(scratch.get(&ydummy, (size_t)128 * cout * 2));        // tensor map target only: never written
  VLTK_CUDA(scratch.get(&sc, (size_t)cout * 4));
  VLTK_CUDA(scratch.get(&sh, (size_t)cout * 4));
  VLTK_CUDA(scratch.get(&partial, conv_tc_pool_partial_bytes(M, cout)));
  p.y = ydummy;
  int rc = pack_weight_nk(weight, w_nk, cout, cin, kh * kh, st);
  if (!rc) rc = pad_vector(scale, sc, cout, cout, 1.f, st);
  if (!rc) rc = pad_vector(shift, sh, cout, cout, 0.f, st);
  p.scale = sc; p.shift = sh;
  TcPool pool; pool.out = pooled; pool.partial = partial; pool.rows = pool_rows;
  TensorMapCache cache;
  if (!rc) rc = conv_tc_launch(p, w_nk, cout, &cache, st, nullptr, &pool);
  cudaStreamSynchronize(st);
  return rc;
}

int vltk_conv2d_meanpool_exact_nhwc(const float* x, const float* weight, const float* scale, const float* shift,
                                    const float* residual, float* pooled, int n, int hh, int ww, int cin, int cout,
                                    int relu, int pool_rows, void* stream) {
  VLTK_CHECK(x && weight && residual && pooled, "conv2d_meanpool_exact: null argument");
  VLTK_CHECK(cin % 64 == 0 && cout % 256 == 0, "conv2d_meanpool_exact: cin %% 64 and cout %% 256 required");
  cudaStream_t st = (cudaStream_t)stream;
  Scratch scratch(st);
  const int64_t M = (int64_t)n * hh * ww;
  std::vector<float> hw((size_t)cout * cin), hs(cout, 1.f);
  VLTK_CUDA(cudaStreamSynchronize(st));
  VLTK_CUDA(cudaMemcpy(hw.data(), weight, hw.size() * 4, cudaMemcpyDeviceToHost));
  if (scale) VLTK_CUDA(cudaMemcpy(hs.data(), scale, (size_t)cout * 4, cudaMemcpyDeviceToHost));
  std::vector<__half> pl;
  std::vector<float> sx;
  build_h3(hw.data(), cout, cin, cout, hs.data(), pl, sx);
  void *w3 = nullptr, *xh = nullptr, *rh = nullptr;
  float *dsx = nullptr, *partial = nullptr;
  VLTK_CUDA(scratch.get(&w3, pl.size() * 2));
  VLTK_CUDA(scratch.get(&dsx, sx.size() * 4));
  VLTK_CUDA(scratch.get(&xh, (size_t)M * cin * 4));
  VLTK_CUDA(scratch.get(&rh, (size_t)M * cout * 4));
  VLTK_CUDA(scratch.get(&partial, conv_tc_pool_partial_bytes(M, cout)));
  VLTK_CUDA(cudaMemcpyAsync(w3, pl.data(), pl.size() * 2, cudaMemcpyHostToDevice, st));
  VLTK_CUDA(cudaMemcpyAsync(dsx, sx.data(), sx.size() * 4, cudaMemcpyHostToDevice, st));
  int rc = split_f32_h2(x, nullptr, 0, xh, M, cin, st);
  if (!rc) rc = split_f32_h2(residual, nullptr, 0, rh, M, cout, st);
  ConvProblem p;
  memset(&p, 0, sizeof(p));
  p.x = xh; p.ldx = 2 * cin; p.residual = rh; p.ldr = 2 * cout; p.ldy = 2 * cout;
  p.N = n; p.H = p.OH = hh; p.W = p.OW = ww; p.Cin = cin; p.KH = p.KW = 1; p.stride = 1; p.dil = 1;
  p.Cout = cout; p.relu = relu; p.in_dtype = DT_H2; p.out_dtype = DT_H2; p.scale = dsx; p.shift = shift;
  TcPool pool; pool.out = pooled; pool.partial = partial; pool.rows = pool_rows;
  TensorMapCache cache;
  if (!rc) rc = conv_tcx_launch(p, w3, cout, &cache, st, nullptr, &pool);
  cudaStreamSynchronize(st);
  if (!rc) VLTK_LAUNCH_CHECK();
  return rc;
}

int vltk_conv2d_dual_nhwc(const void* x, const float* weight, const void* x2, const float* weight2, const float* shift,
                          void* y, int n, int hh, int ww, int cin, int h2, int w2, int cin2, int stride2, int cout,
                          int relu, void* stream) {
  VLTK_CHECK(x && weight && x2 && weight2 && y, "conv2d_dual: null argument");
  VLTK_CHECK(cin % 64 == 0 && cin2 % 64 == 0 && cout % 64 == 0, "conv2d_dual: cin, cin2, cout must be multiples of 64");
  cudaStream_t st = (cudaStream_t)stream;
  Scratch scratch(st);
  ConvProblem p;
  memset(&p, 0, sizeof(p));
  p.x = x; p.ldx = cin; p.y = y; p.ldy = cout; p.N = n; p.H = p.OH = hh; p.W = p.OW = ww; p.Cin = cin;
  p.KH = p.KW = 1; p.stride = 1; p.dil = 1; p.Cout = cout; p.relu = relu; p.in_dtype = DT_BF16; p.out_dtype = DT_BF16;
  bf16 *w1 = nullptr, *wb = nullptr;
  float* sh = nullptr;
  VLTK_CUDA(scratch.get(&w1, (size_t)cout * cin * 2));
  VLTK_CUDA(scratch.get(&wb, (size_t)cout * cin2 * 2));
  VLTK_CUDA(scratch.get(&sh, (size_t)cout * 4));
  int rc = pack_weight_nk(weight, w1, cout, cin, 1, st);
  if (!rc) rc = pack_weight_nk(weight2, wb, cout, cin2, 1, st);
  if (!rc) rc = pad_vector(shift, sh, cout, cout, 0.f, st);
  p.shift = sh;
  TcConcat cc;
  cc.x2 = x2; cc.ldx2 = cin2; cc.H2 = h2; cc.W2 = w2; cc.Cin2 = cin2; cc.stride2 = stride2; cc.w2 = wb;
  TensorMapCache cache;
  if (!rc) rc = conv_tc_launch(p, w1, cout, &cache, st, nullptr, nullptr, &cc);
  cudaStreamSynchronize(st);
  return rc;
}

// One part of the engine on caller-provided fp32 tensors, through the engine's own layers, weights and activation type
// (teacher-forced stage tests: feed the oracle's stage input, compare the stage output).
int vltk_frcnn_run_part(vltk_frcnn_t* h, int part, int block_begin, int block_end, const float* x, int n, int hh, int ww,
                        float* y, int64_t y_cap, int32_t* out_dims, void* stream) {
  VLTK_CHECK(h && h->finalized && x && y && out_dims, "run_part: bad argument");
  VLTK_CHECK(part == 0 || (part >= 2 && part <= 5), "run_part: part must be 0 (stem), 2-4 (res2-res4) or 5 (rpn head)");
  VLTK_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  Scratch scratch(st);
  const DType d = h->act;
  const size_t e = esz(d);
  auto to_act = [&](const float* src, int64_t rows, int C, void** dst) -> int {
    if (d == DT_F32) { *dst = (void*)src; return 0; }
    VLTK_CUDA(scratch.get(dst, (size_t)rows * C * e));
    return d == DT_BF16 ? cast_f32(src, *dst, DT_BF16, rows * C, st) : split_f32_h2(src, nullptr, 0, *dst, rows, C, st);
  };
  auto from_act = [&](const void* src, int64_t rows, int C) -> int {
    VLTK_CHECK(rows * C <= y_cap, "run_part: output has %lld elements, capacity %lld", (long long)(rows * C), (long long)y_cap);
    if (d == DT_F32) { VLTK_CUDA(cudaMemcpyAsync(y, src, (size_t)rows * C * 4, cudaMemcpyDeviceToDevice, st)); return 0; }
    if (d == DT_H2) return widen_h2(src, y, rows, C, st);
    widen_bf16_kernel<<<(unsigned)ceil_div64(rows * C, 256), 256, 0, st>>>((const bf16*)src, y, rows * C);
    VLTK_LAUNCH_CHECK();
    return 0;
  };
  const vltk_frcnn_config& c = h->cfg;
  int rc = 0;
  if (part == 0) {            // x: images NCHW f32 [n,3,hh,ww] -> pooled stem output NHWC
    const Shapes s = make_shapes(c, n, hh, ww);
    void *in4 = nullptr, *stema = nullptr, *stem = nullptr, *pool = nullptr;
    VLTK_CUDA(scratch.get(&in4, (size_t)n * hh * ww * 16));
    VLTK_CUDA(scratch.get(&stema, (size_t)n * s.Hs * s.Ws * 192 * 2));
    VLTK_CUDA(scratch.get(&stem, (size_t)n * s.Hs * s.Ws * c.stem_out_channels * 4));
    VLTK_CUDA(scratch.get(&pool, (size_t)n * s.Hp * s.Wp * c.stem_out_channels * 4));
    rc = run_stem(h, x, n, hh, ww, s, in4, stema, stem, pool, st);
    if (!rc) rc = from_act(pool, (int64_t)n * s.Hp * s.Wp, c.stem_out_channels);
    out_dims[0] = s.Hp; out_dims[1] = s.Wp; out_dims[2] = c.stem_out_channels;
  } else if (part <= 4) {     // x: NHWC f32 [n,hh,ww,cin] -> stage output NHWC
    const std::vector<Block>& full = h->stages[part - 2];
    const int b0 = std::max(block_begin, 0), b1 = block_end < 0 ? (int)full.size() : std::min(block_end, (int)full.size());
    VLTK_CHECK(b0 < b1, "run_part: empty block range [%d, %d) of a %d-block stage", b0, b1, (int)full.size());
    const std::vector<Block> stage(full.begin() + b0, full.begin() + b1);   // (LayerW holds device pointers only)
    const int cin = stage[0].c1.cin, cout = stage[0].c3.cout, mid = stage[0].c1.cout, stride = stage[0].c1.stride;
    const int h1 = (hh - 1) / stride + 1, w1 = (ww - 1) / stride + 1;
    void *xa = nullptr, *pa = nullptr, *pb = nullptr, *t1 = nullptr, *t2 = nullptr, *sb = nullptr;
    if (to_act(x, (int64_t)n * hh * ww, cin, &xa)) return -1;
    const size_t big = (size_t)n * h1 * w1 * cout * e, midb = (size_t)n * h1 * w1 * mid * e;
    VLTK_CUDA(scratch.get(&pa, big)); VLTK_CUDA(scratch.get(&pb, big)); VLTK_CUDA(scratch.get(&sb, big));
    VLTK_CUDA(scratch.get(&t1, midb)); VLTK_CUDA(scratch.get(&t2, midb));
    const void* cur = xa;
    void* pp[2] = {pa, pb};
    int ch = hh, cw = ww, flip = 0;
    for (const Block& blk : stage) {
      int oh = 0, ow = 0;
      if (run_block(h, blk, cur, n, ch, cw, pp[flip], t1, t2, sb, st, &oh, &ow)) return -1;
      cur = pp[flip]; flip ^= 1; ch = oh; cw = ow;
    }
    rc = from_act(cur, (int64_t)n * ch * cw, cout);
    out_dims[0] = ch; out_dims[1] = cw; out_dims[2] = cout;
  } else {                    // x: res4 NHWC f32 [n,hh,ww,C4] -> fp32 head rows [n*hh*ww, ldh]
    const int C4 = h->rpn_conv.cin;
    void *xa = nullptr, *hid = nullptr, *head = nullptr;
    if (to_act(x, (int64_t)n * hh * ww, C4, &xa)) return -1;
    VLTK_CUDA(scratch.get(&hid, (size_t)n * hh * ww * c.rpn_hidden * e));
    VLTK_CUDA(scratch.get(&head, (size_t)n * hh * ww * std::max(h->rpn_head.ldw, h->rpn_head.cout_pad) * 4));
    int ldh = 0;
    rc = run_rpn_head(h, xa, n, hh, ww, hid, head, &ldh, st);
    VLTK_CHECK((int64_t)n * hh * ww * ldh <= y_cap, "run_part: head output exceeds the capacity");
    if (!rc) VLTK_CUDA(cudaMemcpyAsync(y, head, (size_t)n * hh * ww * ldh * 4, cudaMemcpyDeviceToDevice, st));
    out_dims[0] = hh; out_dims[1] = ww; out_dims[2] = ldh;
  }
  cudaStreamSynchronize(st);
  if (!rc) VLTK_LAUNCH_CHECK();
  return rc;
}

int vltk_conv_tc_set_trace(void* dev_buf, int cap_per_role, int cta) { return conv_tc_set_trace(dev_buf, cap_per_role, cta); }

int vltk_conv_tc_set_cta_pairs(int min_pixels, int residual_layers) {
  conv_tc_set_cta_pairs(min_pixels, residual_layers);
  return 0;
}

int vltk_conv_tcx_set_cta_pairs(int min_pixels) {
  conv_tcx_set_cta_pairs(min_pixels);
  return 0;
}

int vltk_linear_tc3(const float* x, const float* weight, const float* bias, float* y, int m, int k, int n, int relu,
                    void* stream) {
  VLTK_CHECK(x && weight && y, "linear_tc3: null argument");
  VLTK_CHECK(k % 64 == 0 && n % 64 == 0, "linear_tc3: k and n must be multiples of 64");
  cudaStream_t st = (cudaStream_t)stream;
  Scratch scratch(st);
  bf16 *xhi = nullptr, *xlo = nullptr, *whi = nullptr, *wlo = nullptr;
  float* sh = nullptr;
  VLTK_CUDA(scratch.get(&xhi, (size_t)m * k * 2)); VLTK_CUDA(scratch.get(&xlo, (size_t)m * k * 2));
  VLTK_CUDA(scratch.get(&whi, (size_t)n * k * 2)); VLTK_CUDA(scratch.get(&wlo, (size_t)n * k * 2));
  VLTK_CUDA(scratch.get(&sh, (size_t)n * 4));
  int rc = split_f32(x, nullptr, 0, xhi, xlo, (int64_t)m * k, st);
  if (!rc) rc = split_f32(weight, nullptr, 0, whi, wlo, (int64_t)n * k, st);   // [n][k] is already the B layout
  if (!rc) rc = pad_vector(bias, sh, n, n, 0.f, st);
  ConvProblem q;
  memset(&q, 0, sizeof(q));
  q.x = xhi; q.ldx = k; q.y = y; q.ldy = n; q.N = m; q.H = q.W = q.OH = q.OW = 1; q.Cin = k; q.Cout = n;
  q.KH = q.KW = 1; q.stride = 1; q.dil = 1; q.shift = sh; q.relu = relu; q.in_dtype = DT_BF16; q.out_dtype = DT_F32;
  TcSplit sp; sp.x_lo = xlo; sp.w_lo = wlo;
  TensorMapCache cache;
  if (!rc) rc = conv_tc_launch(q, whi, n, &cache, st, &sp);
  cudaStreamSynchronize(st);
  return rc;
}

int vltk_rpn_proposals(const float* logits, const float* deltas, const float* cell_host, const int32_t* sizes_hw,
                       int n, int a, int h4, int w4, int stride, int pre_topk, int post_topk, float nms_thresh,
                       float min_size, const float weights[4], float* proposals, float* proposal_logits,
                       int32_t* counts, void* stream) {
  VLTK_CHECK(logits && deltas && cell_host && sizes_hw && proposals && counts, "rpn_proposals: null argument");
  VLTK_CHECK(pre_topk >= 1 && pre_topk <= 8192 && post_topk >= 1 && post_topk <= 8192, "rpn_proposals: topk out of range");
  cudaStream_t st = (cudaStream_t)stream;
  Scratch scratch(st);
  const int HW = h4 * w4, ldh = round_up(5 * a, 4), K = std::min(pre_topk, HW * a);
  float *head = nullptr, *cell = nullptr, *sbox = nullptr, *ssc = nullptr, *psc = nullptr;
  int *sizes = nullptr, *sidx = nullptr, *pidx = nullptr;
  uint8_t* valid = nullptr;
  unsigned long long* mask = nullptr;
  VLTK_CUDA(scratch.get(&head, (size_t)n * HW * ldh * 4));
  VLTK_CUDA(scratch.get(&cell, (size_t)a * 16));
  VLTK_CUDA(scratch.get(&sizes, (size_t)n * 8));
  VLTK_CUDA(scratch.get(&sbox, (size_t)n * K * 16));
  VLTK_CUDA(scratch.get(&ssc, (size_t)n * K * 4));
  VLTK_CUDA(scratch.get(&sidx, (size_t)n * K * 4));
  VLTK_CUDA(scratch.get(&valid, (size_t)n * K));
  VLTK_CUDA(scratch.get(&mask, nms_mask_bytes(n, K)));
  VLTK_CUDA(scratch.get(&pidx, (size_t)n * post_topk * 4));
  int* done = nullptr;
  VLTK_CUDA(scratch.get(&done, (size_t)n * 4));
  if (!proposal_logits) VLTK_CUDA(scratch.get(&psc, (size_t)n * post_topk * 4));
  VLTK_CUDA(cudaMemcpyAsync(cell, cell_host, (size_t)a * 16, cudaMemcpyHostToDevice, st));
  VLTK_CUDA(cudaMemcpyAsync(sizes, sizes_hw, (size_t)n * 8, cudaMemcpyHostToDevice, st));
  int64_t t1 = (int64_t)n * 4 * a * HW, t2 = (int64_t)n * a * HW;
  nchw_to_nhwc_kernel<<<(unsigned)ceil_div64(t1, 256), 256, 0, st>>>(deltas, head, 4 * a, HW, ldh, 0, t1);
  nchw_to_nhwc_kernel<<<(unsigned)ceil_div64(t2, 256), 256, 0, st>>>(logits, head, a, HW, ldh, 4 * a, t2);
  RpnSelectArgs ra;
  memset(&ra, 0, sizeof(ra));
  ra.head = head; ra.ldh = ldh; ra.delta_off = 0; ra.logit_off = 4 * a; ra.N = n; ra.H4 = h4; ra.W4 = w4; ra.A = a;
  ra.stride = stride; ra.cell = cell; ra.sizes_hw = sizes; ra.pre_topk = pre_topk; ra.min_size = min_size;
  ra.wx = weights[0]; ra.wy = weights[1]; ra.ww = weights[2]; ra.wh = weights[3];
  ra.boxes = sbox; ra.scores = ssc; ra.anchor_idx = sidx; ra.valid = valid; ra.K = K;
  int rc = rpn_select(ra, st);
  NmsArgs na;
  memset(&na, 0, sizeof(na));
  na.boxes = sbox; na.scores = ssc; na.valid = valid; na.N = n; na.K = K; na.thresh = nms_thresh; na.max_keep = post_topk;
  na.mask = mask; na.out_boxes = proposals; na.out_scores = proposal_logits ? proposal_logits : psc; na.out_idx = pidx;
  na.out_count = counts; na.done = done;
  if (!rc) rc = nms_sorted(na, st);
  cudaStreamSynchronize(st);
  return rc;
}

int vltk_nms(const float* boxes, const float* scores, int k, float thresh, int max_keep, int32_t* keep,
             int32_t* count, void* stream) {
  VLTK_CHECK(boxes && scores && keep && count, "nms: null argument");
  VLTK_CHECK(k >= 0 && k <= 8192 && max_keep >= 1 && max_keep <= 8192, "nms: k must be <= 8192");
  cudaStream_t st = (cudaStream_t)stream;
  Scratch scratch(st);
  float *sbox = nullptr, *ssc = nullptr, *obox = nullptr;
  int *order = nullptr, *oidx = nullptr;
  unsigned long long* mask = nullptr;
  VLTK_CUDA(scratch.get(&sbox, (size_t)std::max(k, 1) * 16));
  VLTK_CUDA(scratch.get(&ssc, (size_t)std::max(k, 1) * 4));
  VLTK_CUDA(scratch.get(&order, (size_t)std::max(k, 1) * 4));
  VLTK_CUDA(scratch.get(&obox, (size_t)max_keep * 16));
  VLTK_CUDA(scratch.get(&oidx, (size_t)max_keep * 4));
  VLTK_CUDA(scratch.get(&mask, nms_mask_bytes(1, std::max(k, 1))));
  int* done = nullptr;
  VLTK_CUDA(scratch.get(&done, 4));
  int rc = sort_boxes_desc(boxes, scores, k, sbox, ssc, order, st);
  NmsArgs na;
  memset(&na, 0, sizeof(na));
  na.boxes = sbox; na.scores = ssc; na.valid = nullptr; na.N = 1; na.K = k; na.thresh = thresh; na.max_keep = max_keep;
  na.mask = mask; na.out_boxes = obox; na.out_scores = nullptr; na.out_idx = oidx; na.out_count = count; na.done = done;
  if (!rc) rc = nms_sorted(na, st);
  if (!rc) rc = remap_indices(oidx, order, max_keep, keep, st);
  cudaStreamSynchronize(st);
  return rc;
}

int vltk_roi_pool_nchw(const float* feat, int n, int c, int hh, int ww, const float* rois5, int r, int pp, float scale,
                       float* out, void* stream) {
  VLTK_CHECK(feat && rois5 && out, "roi_pool: null argument");
  VLTK_CHECK(c % 4 == 0, "roi_pool: C must be a multiple of 4");
  cudaStream_t st = (cudaStream_t)stream;
  Scratch scratch(st);
  // the stage entry accepts arbitrary batch indices: pool each ROI as its own "image slot"
  float *nhwc = nullptr, *boxes = nullptr, *tmp = nullptr;
  int *bidx = nullptr, *ones = nullptr;
  VLTK_CUDA(scratch.get(&nhwc, (size_t)n * hh * ww * c * 4));
  VLTK_CUDA(scratch.get(&boxes, (size_t)std::max(r, 1) * 16));
  VLTK_CUDA(scratch.get(&bidx, (size_t)std::max(r, 1) * 4));
  VLTK_CUDA(scratch.get(&ones, (size_t)std::max(n, 1) * 4));
  VLTK_CUDA(scratch.get(&tmp, (size_t)std::max(r, 1) * pp * pp * c * 4));
  int64_t tot = (int64_t)n * c * hh * ww;
  nchw_to_nhwc_kernel<<<(unsigned)ceil_div64(tot, 256), 256, 0, st>>>(feat, nhwc, c, hh * ww, c, 0, tot);
  int rc = 0;
  if (r > 0) {
    rois5_split_kernel<<<ceil_div(r, 128), 128, 0, st>>>(rois5, r, boxes, bidx);
    rc = roi_pool_indexed(nhwc, DT_F32, hh, ww, c, boxes, bidx, r, pp, scale, tmp, st);
    int64_t to = (int64_t)r * c * pp * pp;
    nhwc_to_nchw_kernel<<<(unsigned)ceil_div64(to, 256), 256, 0, st>>>(tmp, out, c, pp * pp, to);
  }
  cudaStreamSynchronize(st);
  VLTK_LAUNCH_CHECK();
  return rc;
}

// one CTA = 16 KB of one output row: 256 threads x 4 independent 16-byte loads (streaming: every byte is used once)
__global__ void __launch_bounds__(256)
feature_gather_kernel(const float4* __restrict__ table, int64_t ld4, const int32_t* __restrict__ idx, int64_t n_rows,
                      int cols4, float4* __restrict__ out) {
  const int r = blockIdx.y;
  const int64_t src = idx[r];
  if (src < 0 || src >= n_rows) return;                 // out-of-range index: row left untouched (host validates)
  const float4* __restrict__ s = table + src * ld4;
  float4* __restrict__ d = out + (int64_t)r * cols4;
  const int c0 = blockIdx.x * 1024 + threadIdx.x;
  float4 v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) if (c0 + k * 256 < cols4) v[k] = __ldcs(s + c0 + k * 256);
#pragma unroll
  for (int k = 0; k < 4; ++k) if (c0 + k * 256 < cols4) __stcs(d + c0 + k * 256, v[k]);
}

int vltk_gather_rows_f32(const float* table, int64_t n_rows, int64_t ld, const int32_t* idx, int rows, int cols,
                         float* out, void* stream) {
  VLTK_CHECK(table && idx && out, "gather_rows: null argument");
  VLTK_CHECK(cols % 4 == 0 && ld % 4 == 0 && cols <= ld, "gather_rows: cols and ld must be multiples of 4 floats, cols <= ld");
  VLTK_CHECK(((uintptr_t)table % 16 == 0) && ((uintptr_t)out % 16 == 0), "gather_rows: pointers must be 16-byte aligned");
  if (rows <= 0 || cols == 0) return 0;
  VLTK_CHECK(rows <= 65535, "gather_rows: at most 65535 rows per call");
  dim3 grid((unsigned)ceil_div(cols / 4, 1024), (unsigned)rows);
  feature_gather_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float4*)table, ld / 4, idx, n_rows, cols / 4, (float4*)out);
  VLTK_LAUNCH_CHECK();
  return 0;
}

int vltk_roi_outputs(const float* obj_logits, const float* attr_logits, const float* box_deltas, const float* feats,
                     const float* proposals, const int32_t* counts, const int32_t* sizes_hw, const float* scales_yx,
                     int n, int r, int num_classes, int num_attrs, int d, const float weights[4],
                     const vltk_frcnn_knobs* knobs, const vltk_frcnn_out* out, void* stream) {
  VLTK_CHECK(obj_logits && attr_logits && box_deltas && feats && proposals && counts && sizes_hw && knobs && out,
             "roi_outputs: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  Scratch scratch(st);
  int* sizes = nullptr;
  float* scales = nullptr;
  VLTK_CUDA(scratch.get(&sizes, (size_t)n * 8));
  VLTK_CUDA(scratch.get(&scales, (size_t)n * 8));
  VLTK_CUDA(cudaMemcpyAsync(sizes, sizes_hw, (size_t)n * 8, cudaMemcpyHostToDevice, st));
  if (scales_yx) VLTK_CUDA(cudaMemcpyAsync(scales, scales_yx, (size_t)n * 8, cudaMemcpyHostToDevice, st));
  TailArgs ta;
  memset(&ta, 0, sizeof(ta));
  ta.N = n; ta.R = r; ta.cls_logits = obj_logits; ta.ldc = num_classes + 1; ta.bbox_deltas = box_deltas; ta.ldb = num_classes * 4;
  ta.attr_logits = attr_logits; ta.lda = num_attrs + 1; ta.feats = feats; ta.D = d; ta.proposals = proposals;
  ta.count = counts; ta.sizes_hw = sizes; ta.scales_yx = scales_yx ? scales : nullptr;
  ta.num_classes = num_classes; ta.num_attrs = num_attrs;
  ta.wx = weights[0]; ta.wy = weights[1]; ta.ww = weights[2]; ta.wh = weights[3];
  ta.nms_thresh = knobs->nms_thresh; ta.n_thresh = knobs->n_nms_thresh; ta.min_det = knobs->min_detections;
  ta.max_det = knobs->max_detections; ta.pad_value = knobs->pad_value;
  ta.boxes = out->boxes; ta.norm_boxes = out->normalized_boxes; ta.obj_ids = (long long*)out->obj_ids;
  ta.obj_probs = out->obj_probs; ta.attr_ids = (long long*)out->attr_ids; ta.attr_probs = out->attr_probs;
  ta.roi_features = out->roi_features; ta.preds_per_image = out->preds_per_image; ta.keep_idx = out->keep_idx;
  VLTK_CUDA(scratch.get(&ta.stats, (size_t)n * r * 32));
  int rc = roi_tail(ta, st);
  cudaStreamSynchronize(st);
  return rc;
}

}  // extern "C"
