// Implicit-GEMM convolution on the 5th-generation tensor cores (sm_100a):
//
//   D[128 pixels, BN couts] (fp32, TMEM) += A[128, 64] (bf16, smem) * W[BN, 64]^T (bf16, smem)
//
//   * A tiles are fetched by TMA in IM2COL mode straight from the NHWC activation tensor: one
//     cp.async.bulk.tensor.4d.im2col per (filter tap, 64-channel block) gathers 128 consecutive
//     output pixels, applies stride/dilation and zero-fills the padding halo in hardware.
//   * W tiles are fetched by tiled TMA from the [cout][tap*cin] weight matrix.
//   * both land in 128B-swizzled K-major smem, a STAGES-deep mbarrier ring feeds one elected
//     thread that issues tcgen05.mma (UMMA 128 x BN x 16, kind::f16, fp32 accumulate in TMEM),
//     tcgen05.commit recycles the stages, and four epilogue warps drain TMEM with tcgen05.ld,
//     apply frozen-BN scale/shift (+ residual) (+ ReLU) and store bf16 NHWC rows.
//
// Reference layers replaced: every Conv2d+BN(+ReLU) of res2-res5 and the RPN 3x3
// (frcnn.py:794-822, 963-979, 1345-1355, 1569).
#include "conv_tc.cuh"

namespace vltk {

namespace {

constexpr int BM = 128;       // UMMA M (one TMEM lane per output pixel)
constexpr int BK = 64;        // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int TC_THREADS = 256;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a protocol bug must surface as a launch failure, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin) {
    if (spin > (1u << 26)) {
      printf("conv_tc: mbarrier timeout (block %d,%d thread %d bar %u parity %u)\n", blockIdx.x, blockIdx.y,
             threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c, int w,
                                                   int h, int n, uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}

template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "n"(NCOLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major operand in 128B-swizzled rows (8-row atoms of 1024 B):
//   [0,14) start>>4 | [16,30) LBO>>4 (unused for swizzled K-major) | [32,46) SBO>>4 = 1024>>4
//   [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor (kind::f16): D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1, both K-major,
// N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

struct TcParams {
  bf16* y; const bf16* residual; const float* scale; const float* shift;
  int64_t M;
  int ldy, ldr, Cout, relu;
  int OH, OW, stride, pad, dil, KW, taps, cblocks;  // cblocks = Cin / 64
};

template <int BN, int STAGES>
struct Smem {
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int TILE_BYTES = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = TILE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, TcParams p) {
  using S = Smem<BN, STAGES>;
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024 B alignment
  const uint32_t sA = base, sB = base + STAGES * A_STAGE_BYTES;
  const uint32_t bars = base + S::TILE_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  const uint32_t tmem_full_bar = bars + 8u * (2 * STAGES);
  const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 1);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_dyn + (tmem_slot - smem_u32(smem_dyn)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int64_t m0 = (int64_t)blockIdx.y * BM;
  const int num_kb = p.taps * p.cblocks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot_ptr;

  if (warp == 0 && lane == 0) {
    // ================= TMA producer =================
    const int ow0 = (int)(m0 % p.OW);
    const int64_t t = m0 / p.OW;
    const int oh0 = (int)(t % p.OH);
    const int img0 = (int)(t / p.OH);
    const int bw = ow0 * p.stride - p.pad, bh = oh0 * p.stride - p.pad;
    int stage = 0; uint32_t phase = 0;
    for (int tap = 0; tap < p.taps; ++tap) {
      const int kh = tap / p.KW, kw = tap - kh * p.KW;
      for (int cb = 0; cb < p.cblocks; ++cb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        mbar_expect_tx(full_bar(stage), S::STAGE_BYTES);
        tma_load_im2col_4d(sA + stage * A_STAGE_BYTES, &tmA, full_bar(stage), cb * BK, bw, bh, img0,
                           (uint16_t)(kw * p.dil), (uint16_t)(kh * p.dil));
        tma_load_2d(sB + stage * S::B_STAGE_BYTES, &tmB, full_bar(stage), (tap * p.cblocks + cb) * BK, n0);
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ================= MMA issuer (single thread) =================
    constexpr uint32_t idesc = make_idesc(BM, BN);
    int stage = 0; uint32_t phase = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(full_bar(stage), phase);
      tc_fence_after();
      const uint32_t a = sA + stage * A_STAGE_BYTES, b = sB + stage * S::B_STAGE_BYTES;
#pragma unroll
      for (int k = 0; k < BK / UMMA_K; ++k) {
        umma_bf16(tmem_acc, make_smem_desc(a + k * UMMA_K * 2), make_smem_desc(b + k * UMMA_K * 2), idesc,
                  (kb | k) ? 1u : 0u);
      }
      umma_commit(empty_bar(stage));  // frees the smem stage once these MMAs retire
      if (++stage == STAGES) { stage = 0; phase ^= 1u; }
    }
    umma_commit(tmem_full_bar);       // accumulator complete
  } else if (warp >= 4) {
    // ================= epilogue: TMEM -> regs -> scale/shift/residual/ReLU -> bf16 NHWC =================
    const int e = warp - 4;  // TMEM lane quarter == warp id % 4
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int64_t m = m0 + e * 32 + lane;
    const bool row_ok = m < p.M;
    bf16* yrow = p.y + m * p.ldy;
    const bf16* rrow = p.residual ? p.residual + m * p.ldr : nullptr;
#pragma unroll 1
    for (int cc = 0; cc < BN / 32; ++cc) {
      uint32_t v[32];
      tmem_ld32(tmem_acc + ((uint32_t)(e * 32) << 16) + (uint32_t)(cc * 32), v);
      tmem_ld_wait();
      const int n = n0 + cc * 32;
      if (row_ok && n < p.Cout) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {  // 8 channels (16 B of bf16) per store
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int c = n + q * 8 + j;
            float x = __uint_as_float(v[q * 8 + j]);
            f[j] = fmaf(x, p.scale ? __ldg(p.scale + c) : 1.f, p.shift ? __ldg(p.shift + c) : 0.f);
          }
          if (rrow) {
            uint4 r = *reinterpret_cast<const uint4*>(rrow + n + q * 8);
            const __nv_bfloat162* rb = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float2 rf = __bfloat1622float2(rb[j]);
              f[2 * j] += rf.x; f[2 * j + 1] += rf.y;
            }
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
          }
          uint4 o;
          __nv_bfloat162* ob = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
          for (int j = 0; j < 4; ++j) ob[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
          *reinterpret_cast<uint4*>(yrow + n + q * 8) = o;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<BN>(tmem_acc);
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode_tiled = nullptr;
EncodeIm2colFn g_encode_im2col = nullptr;

int load_driver_entry_points() {
  if (g_encode_tiled && g_encode_im2col) return 0;
  cudaDriverEntryPointQueryResult q;
  void* fn = nullptr;
  VLTK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  VLTK_CHECK(fn && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available from the driver");
  g_encode_tiled = (EncodeTiledFn)fn;
  fn = nullptr;
  VLTK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &q));
  VLTK_CHECK(fn && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeIm2col not available from the driver");
  g_encode_im2col = (EncodeIm2colFn)fn;
  return 0;
}

int make_a_map(const ConvProblem& p, CUtensorMap* out) {
  cuuint64_t dims[4] = {(cuuint64_t)p.Cin, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.N};
  cuuint64_t strides[3] = {(cuuint64_t)p.ldx * 2, (cuuint64_t)p.W * p.ldx * 2, (cuuint64_t)p.H * p.W * p.ldx * 2};
  int lower[2] = {-p.pad, -p.pad};
  int upper[2] = {p.pad - (p.KW - 1) * p.dil, p.pad - (p.KH - 1) * p.dil};
  cuuint32_t estr[4] = {1, (cuuint32_t)p.stride, (cuuint32_t)p.stride, 1};
  CUresult r = g_encode_im2col(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(p.x), dims, strides, lower,
                               upper, BK, BM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VLTK_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeIm2col failed (%d) for x[%d,%d,%d,%d] k%d s%d p%d d%d", (int)r, p.N, p.H,
             p.W, p.Cin, p.KH, p.stride, p.pad, p.dil);
  // Same fix-up NVIDIA's own CUTLASS applies to im2col descriptors of tensors < 128 KiB on
  // drivers <= 13.1 (cute/atom/copy_traits_sm90_im2col.hpp): clear bit 21 of descriptor word 1.
  int drv = 0;
  if (cudaDriverGetVersion(&drv) == cudaSuccess && drv <= 13010 &&
      (uint64_t)p.N * p.H * p.W * p.ldx * 2 < 131072)
    reinterpret_cast<uint64_t*>(out)[1] &= ~(1ull << 21);
  return 0;
}

int make_b_map(const bf16* w, int K, int cout_pad, int bn, CUtensorMap* out) {
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)cout_pad};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {BK, (cuuint32_t)bn};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode_tiled(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(w), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VLTK_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) for w[%d,%d]", (int)r, cout_pad, K);
  return 0;
}

template <int BN, int STAGES>
int launch(const CUtensorMap& a, const CUtensorMap& b, const TcParams& tp, int cout_pad, cudaStream_t st) {
  using S = Smem<BN, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    VLTK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    attr_set = true;
  }
  dim3 grid(cout_pad / BN, (unsigned)ceil_div64(tp.M, BM));
  conv_tc_kernel<BN, STAGES><<<grid, TC_THREADS, S::TOTAL, st>>>(a, b, tp);
  VLTK_LAUNCH_CHECK();
  return 0;
}

}  // namespace

int conv_tc_launch(const ConvProblem& p, const bf16* w_nk, int cout_pad, TensorMapCache* cache, cudaStream_t st) {
  VLTK_CHECK(p.in_dtype == DT_BF16 && p.out_dtype == DT_BF16, "conv_tc: bf16 activations only");
  VLTK_CHECK(p.Cin % BK == 0, "conv_tc: Cin=%d must be a multiple of %d", p.Cin, BK);
  VLTK_CHECK(cout_pad % 64 == 0 && p.Cout <= cout_pad, "conv_tc: bad cout_pad");
  VLTK_CHECK(p.ldy % 8 == 0 && p.ldx % 8 == 0 && (!p.residual || p.ldr % 8 == 0), "conv_tc: rows must be 16-byte aligned");
  VLTK_CHECK(p.Cout % 32 == 0, "conv_tc: Cout=%d must be a multiple of 32", p.Cout);
  if (load_driver_entry_points()) return -1;
  const int64_t M = (int64_t)p.N * p.OH * p.OW;
  if (M == 0) return 0;
  const int K = p.KH * p.KW * p.Cin;
  const int bn = (cout_pad % 256 == 0) ? 256 : (cout_pad % 128 == 0 ? 128 : 64);

  CUtensorMap ta, tb;
  TensorMapCache::Key ka(p.x, p.N, p.H, p.W, p.Cin, p.ldx, p.KH, p.stride, p.pad, p.dil, 0);
  TensorMapCache::Key kb(w_nk, K, cout_pad, bn, 0, 0, 0, 0, 0, 0, 1);
  auto ia = cache->maps.find(ka);
  if (ia == cache->maps.end()) {
    if (make_a_map(p, &ta)) return -1;
    cache->maps[ka] = ta;
  } else ta = ia->second;
  auto ib = cache->maps.find(kb);
  if (ib == cache->maps.end()) {
    if (make_b_map(w_nk, K, cout_pad, bn, &tb)) return -1;
    cache->maps[kb] = tb;
  } else tb = ib->second;

  TcParams tp;
  tp.y = (bf16*)p.y; tp.residual = (const bf16*)p.residual; tp.scale = p.scale; tp.shift = p.shift;
  tp.M = M; tp.ldy = p.ldy; tp.ldr = p.ldr; tp.Cout = p.Cout; tp.relu = p.relu;
  tp.OH = p.OH; tp.OW = p.OW; tp.stride = p.stride; tp.pad = p.pad; tp.dil = p.dil; tp.KW = p.KW;
  tp.taps = p.KH * p.KW; tp.cblocks = p.Cin / BK;
  if (bn == 256) return launch<256, 4>(ta, tb, tp, cout_pad, st);
  if (bn == 128) return launch<128, 4>(ta, tb, tp, cout_pad, st);
  return launch<64, 4>(ta, tb, tp, cout_pad, st);
}

}  // namespace vltk
