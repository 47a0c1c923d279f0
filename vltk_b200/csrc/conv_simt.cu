// fp32-accumulate implicit-GEMM convolution on the CUDA cores (register-tiled 128xBNx16).
//
// Role: (1) the arithmetic of the exact ("fp32") engine mode, where end-to-end index parity
// with the reference is asserted; (2) the on-device cross-check for the tcgen05 kernel
// (same bf16 operands, fp32 accumulate); (3) layers whose shapes do not suit the tensor
// pipe yet (3-channel stem, predictor heads).  Reference layers: frcnn.py:794-822, 963-979.
#include "conv.cuh"

namespace vltk {

namespace {

constexpr int BM = 128;
constexpr int BK = 16;
constexpr int NT = 256;

template <typename TI, typename TO, int BN>
__global__ void __launch_bounds__(NT)
conv_simt_kernel(ConvProblem p, const float* __restrict__ w, int ldw) {
  constexpr int CN = BN / 64;  // column groups of 4 per thread (2 for BN=128, 1 for BN=64)
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const int tid = threadIdx.x;
  const int64_t M = (int64_t)p.N * p.OH * p.OW;
  const int K = p.KH * p.KW * p.Cin;
  const int nk = (K + BK - 1) / BK;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const TI* __restrict__ x = reinterpret_cast<const TI*>(p.x);

  // ---- A gather bookkeeping: this thread loads k-quad `aq` of rows `ar` and `ar+64`
  const int aq = tid & 3, ar = tid >> 2;
  int ih0[2], iw0[2];
  const TI* xb[2];
  bool rv[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int64_t m = m0 + ar + 64 * i;
    rv[i] = m < M;
    int64_t mm = rv[i] ? m : 0;
    int ow = (int)(mm % p.OW);
    int64_t t = mm / p.OW;
    int oh = (int)(t % p.OH);
    int n = (int)(t / p.OH);
    ih0[i] = oh * p.stride - p.pad;
    iw0[i] = ow * p.stride - p.pad;
    xb[i] = x + (int64_t)n * p.H * p.W * p.ldx;
  }

  float4 ra[2], rb[CN];
  auto load_tiles = [&](int kt) {
    const int kq = kt * BK + 4 * aq;
    int tap = kq / p.Cin;
    int c = kq - tap * p.Cin;
    int kh = tap / p.KW;
    int kw = tap - kh * p.KW;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int ih = ih0[i] + kh * p.dil, iw = iw0[i] + kw * p.dil;
      bool ok = rv[i] && kq < K && ih >= 0 && ih < p.H && iw >= 0 && iw < p.W;
      ra[i] = ok ? load4(xb[i] + ((int64_t)ih * p.W + iw) * p.ldx + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < CN; ++i) {
      int idx = tid + i * NT;
      int br = idx / (BN / 4), bc = (idx % (BN / 4)) * 4;
      int n = n0 + bc;
      rb[i] = (n < ldw) ? *reinterpret_cast<const float4*>(w + (int64_t)(kt * BK + br) * ldw + n)
                        : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int r = ar + 64 * i;
      As[buf][4 * aq + 0][r] = ra[i].x;
      As[buf][4 * aq + 1][r] = ra[i].y;
      As[buf][4 * aq + 2][r] = ra[i].z;
      As[buf][4 * aq + 3][r] = ra[i].w;
    }
#pragma unroll
    for (int i = 0; i < CN; ++i) {
      int idx = tid + i * NT;
      int br = idx / (BN / 4), bc = (idx % (BN / 4)) * 4;
      *reinterpret_cast<float4*>(&Bs[buf][br][bc]) = rb[i];
    }
  };

  const int tx = tid & 15, ty = tid >> 4;
  float acc[8][4 * CN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4 * CN; ++j) acc[i][j] = 0.f;

  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles(kt + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[8], b[4 * CN];
      *reinterpret_cast<float4*>(&a[0]) = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      *reinterpret_cast<float4*>(&a[4]) = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
#pragma unroll
      for (int g = 0; g < CN; ++g)
        *reinterpret_cast<float4*>(&b[4 * g]) = *reinterpret_cast<const float4*>(&Bs[buf][k][g * 64 + tx * 4]);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4 * CN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) store_tiles(buf ^ 1);
    __syncthreads();
  }

  // ---- epilogue: scale/shift (+residual) (+ReLU), vector stores along channels
  TO* __restrict__ y = reinterpret_cast<TO*>(p.y);
  const TO* __restrict__ res = reinterpret_cast<const TO*>(p.residual);
#pragma unroll
  for (int g = 0; g < CN; ++g) {
    const int n = n0 + g * 64 + tx * 4;
    if (n >= ldw) continue;
    float4 sc = p.scale ? *reinterpret_cast<const float4*>(p.scale + n) : make_float4(1.f, 1.f, 1.f, 1.f);
    float4 sh = p.shift ? *reinterpret_cast<const float4*>(p.shift + n) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
      if (m >= M) continue;
      float4 v;
      v.x = fmaf(acc[i][4 * g + 0], sc.x, sh.x);
      v.y = fmaf(acc[i][4 * g + 1], sc.y, sh.y);
      v.z = fmaf(acc[i][4 * g + 2], sc.z, sh.z);
      v.w = fmaf(acc[i][4 * g + 3], sc.w, sh.w);
      if (res) {
        float4 r = load4(res + m * p.ldr + n);
        v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
      }
      if (p.relu) {
        v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
      }
      store4(y + m * p.ldy + n, v);
    }
  }
}

template <typename TI, typename TO>
int launch_typed(const ConvProblem& p, const float* w, int ldw, cudaStream_t st) {
  const int64_t M = (int64_t)p.N * p.OH * p.OW;
  if (M == 0) return 0;
  dim3 block(NT);
  if (ldw > 64) {
    dim3 grid((unsigned)ceil_div64(M, BM), (unsigned)ceil_div(ldw, 128));
    conv_simt_kernel<TI, TO, 128><<<grid, block, 0, st>>>(p, w, ldw);
  } else {
    dim3 grid((unsigned)ceil_div64(M, BM), 1);
    conv_simt_kernel<TI, TO, 64><<<grid, block, 0, st>>>(p, w, ldw);
  }
  VLTK_LAUNCH_CHECK();
  return 0;
}

}  // namespace

int conv_simt_launch(const ConvProblem& p, const float* w, int ldw, cudaStream_t st) {
  VLTK_CHECK(p.Cin % 4 == 0, "conv_simt: Cin=%d must be a multiple of 4", p.Cin);
  VLTK_CHECK(ldw % 4 == 0 && ldw >= p.Cout, "conv_simt: bad ldw=%d for Cout=%d", ldw, p.Cout);
  VLTK_CHECK(p.ldy >= ldw && (p.ldy % 4) == 0, "conv_simt: ldy=%d must be >= ldw=%d", p.ldy, ldw);
  VLTK_CHECK(p.ldx % 4 == 0, "conv_simt: ldx=%d must be a multiple of 4", p.ldx);
  if (p.in_dtype == DT_F32 && p.out_dtype == DT_F32) return launch_typed<float, float>(p, w, ldw, st);
  if (p.in_dtype == DT_BF16 && p.out_dtype == DT_BF16) return launch_typed<bf16, bf16>(p, w, ldw, st);
  if (p.in_dtype == DT_BF16 && p.out_dtype == DT_F32) return launch_typed<bf16, float>(p, w, ldw, st);
  if (p.in_dtype == DT_F32 && p.out_dtype == DT_BF16) return launch_typed<float, bf16>(p, w, ldw, st);
  set_error("conv_simt: unsupported dtype combination");
  return -2;
}

}  // namespace vltk
