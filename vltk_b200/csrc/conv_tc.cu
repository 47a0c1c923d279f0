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
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>

#include "conv_tc.cuh"
#include "tc_ptx.cuh"

namespace vltk {

namespace {


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

// =========================================================================================
// v2: persistent, fully warp-specialised, every global access through TMA.
//
//   warp 0   TMA producer     A (im2col) + W tiles -> STAGES-deep smem ring
//   warp 1   MMA issuer       tcgen05.mma into one of TWO TMEM accumulators (2 x BN columns), so
//                             tile i+1's MMAs overlap tile i's epilogue
//   warp 2   TMEM allocator
//   warp 3   residual producer  [128 x 64] bf16 slabs of the shortcut tensor -> RS-deep smem ring
//   warps 4-7 epilogue        tcgen05.ld -> scale/shift (+residual from smem) (+ReLU) -> bf16 ->
//                             128B-swizzled smem staging -> TMA store (double buffered)
//
// One CTA per SM loops over tiles t = blockIdx.x, +gridDim.x, ... with the cout tile fastest, so
// the CTAs running concurrently share A tiles in L2.  Rows past M are zero-filled on load and
// clipped on store by the TMA unit: no tail code.
constexpr int SLAB = 64;                       // epilogue column slab: 64 bf16 = one 128 B swizzle row
constexpr int SLAB_BYTES = BM * SLAB * 2;      // 16 KB
// residual ring depth: 3 slabs (48 KB) next to BN=256 operand stages; 5 slabs (80 KB) with the smaller BN=128 stages.
template <int BN, bool HAS_RES> struct ResRing { static constexpr int DEPTH = (HAS_RES && BN == 128) ? 5 : 3; };


struct TcParams2 {
  const float* scale; const float* shift;
  FastDiv fd_ntiles, fd_ow, fd_oh;             // divisors n_tiles, OW, OH
  int64_t M;
  int Cout, relu;
  int OH, OW, stride, pad, dil, KH, KW, taps, cblocks;
  int n_tiles, num_tiles;                      // cout tiles, total tiles
  // split-precision GEMM: the K loop runs `npass` times; pass i reads A from map pass_a[i] (0: tmA,
  // 1: tmA2) and W from map pass_b[i] (0: tmB, 1: tmB2), all accumulating into the same TMEM tile.
  // npass = 1 is the plain bf16 product; {hi*hi, lo*hi, hi*lo} gives an fp32-faithful product.
  // Each pass has its own K extent and sampling stride (K-concatenated dual GEMM, TcConcat).
  int npass, pass_a[3], pass_b[3], pass_cblocks[3], pass_stride[3];
  // POOL epilogue (res5 tail, frcnn.py:1401): rows are grouped in ROIs of `pool_rows` consecutive pixels
  // (128 < pool_rows <= 256).  Row tiles are ROI-ALIGNED — tile mt covers rows [0,128) (mt even) or [128,pool_rows)
  // (mt odd) of ROI mt/2; rows past the ROI are computed and masked — so a tile never mixes ROIs and an ROI's sums
  // do not depend on where it sits in the batch.  Instead of storing the tile, each tile writes the fp32 column
  // sums of its valid rows to pool_partial[mt * Cout + c]; pool_finish() adds an ROI's two partials and divides.
  float* pool_partial;
  int pool_rows;
  // pipeline trace (diagnosis builds only, -DVLTK_TC_TRACE): per-role (tag, clock64) records of CTA `trace_cta`
  long long* trace;
  int trace_cap, trace_cta;
};

// Pipeline trace of one CTA (tools/tc_trace.py): role r appends (tag, clock64) pairs to trace[r * cap ...].  Compiled out
// of the shipped library.  tag = event << 40 | tile iteration << 16 | index (k-block / slab).
#ifdef VLTK_TC_TRACE
#define TC_TRACE_DECL(role) int tr_n = 0; const bool tr_on = p.trace && (int)blockIdx.x == p.trace_cta; const int tr_role = (role);
#define TC_TRACE(ev, it, idx)                                                                                      \
  do {                                                                                                             \
    if (tr_on && tr_n < p.trace_cap) {                                                                             \
      long long* tq = p.trace + ((size_t)tr_role * p.trace_cap + tr_n) * 2;                                        \
      tq[0] = ((long long)(ev) << 40) | ((long long)(it) << 16) | (long long)(idx);                                \
      tq[1] = clock64();                                                                                           \
      ++tr_n;                                                                                                      \
    }                                                                                                              \
  } while (0)
#define TC_TRACE_AT(ev, clk)                                                                                        \
  do {                                                                                                             \
    if (tr_on && tr_n < p.trace_cap) {                                                                             \
      long long* tq = p.trace + ((size_t)tr_role * p.trace_cap + tr_n) * 2;                                        \
      tq[0] = ((long long)(ev) << 40);                                                                             \
      tq[1] = (clk);                                                                                               \
      ++tr_n;                                                                                                      \
    }                                                                                                              \
  } while (0)
#define TC_TRACE_ENTRY() const long long tr_entry = clock64();
#else
#define TC_TRACE_DECL(role)
#define TC_TRACE(ev, it, idx) do {} while (0)
#define TC_TRACE_AT(ev, clk) do {} while (0)
#define TC_TRACE_ENTRY()
#endif

template <int BN, int STAGES, bool HAS_RES, int EG = 1>
struct Smem2 {
  static constexpr int GSC = (BN / EG > 64) ? BN / EG : 64;   // scale (and shift) floats staged per epilogue group
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int OFF_B = STAGES * A_STAGE_BYTES;
  static constexpr int OFF_OUT = STAGES * STAGE_BYTES;
  static constexpr int OFF_RES = OFF_OUT + 2 * SLAB_BYTES;
  static constexpr int RS = ResRing<BN, HAS_RES>::DEPTH;
  static constexpr int OFF_SCALE = OFF_RES + (HAS_RES ? RS * SLAB_BYTES : 0);
  static constexpr int OFF_BARS = OFF_SCALE + 2 * EG * GSC * 4;
  static constexpr int NUM_BARS = 2 * STAGES + 4 + 2 * RS;
  static constexpr int TOTAL = OFF_BARS + NUM_BARS * 8 + 16;   // + tmem slot (4 B) + seen[2] (8 B)
  static_assert(TOTAL <= 232448, "exceeds the 227 KB shared memory of one sm_100 CTA");
};

template <int BN, int STAGES, bool HAS_RES, bool OUT_F32, bool POOL = false, int EG = 1>
__global__ void __launch_bounds__(128 + 128 * EG, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmB2,
                const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmR, TcParams2 p) {
  using S = Smem2<BN, STAGES, HAS_RES, EG>;
  TC_TRACE_ENTRY()
  static_assert(!(HAS_RES && OUT_F32), "fp32 output has no residual path");
  static_assert(!POOL || !OUT_F32, "the pooled epilogue reduces the bf16-path tile");
  static_assert(EG == 1 || EG == 2, "one or two epilogue warpgroups");
  constexpr int RS = S::RS;
  constexpr int SLABC = OUT_F32 ? 32 : SLAB;   // columns per 128 B staging row (fp32: 32, bf16: 64)
  constexpr int NSLAB = BN / SLABC;
  // EG epilogue warpgroups (warps 4-7, 8-11) share the slabs round-robin over the CTA's global slab sequence
  // c = it * NSLAB + s: group g owns the slabs with c % EG == g, has its own staging buffers (2 / EG), its own
  // named barrier and its own scale/shift cache, so the two groups never synchronise with each other and two
  // warps per SM sub-partition hide each other's tcgen05.ld / shared-memory / barrier latencies.
  constexpr int NBUF = 2 / EG;                 // 128 x 128 B staging buffers per group
  constexpr int TE_COUNT = 4 * (NSLAB < EG ? NSLAB : EG);   // warps that drain one accumulator
  // no static smem in this kernel: the dynamic window starts at the CTA's (1024 B aligned) base
  extern __shared__ __align__(1024) unsigned char smem_dyn2[];
  const uint32_t base = smem_u32(smem_dyn2);
  if (base & 1023u) {
    if (threadIdx.x == 0) printf("conv_tc2: dynamic smem base %u is not 1024 B aligned\n", base);
    __trap();
  }
  unsigned char* gbase = smem_dyn2;
  const uint32_t sA = base, sB = base + S::OFF_B, sOut = base + S::OFF_OUT, sRes = base + S::OFF_RES;
  float* s_scale_all = reinterpret_cast<float*>(gbase + S::OFF_SCALE);
  const uint32_t bars = base + S::OFF_BARS;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * STAGES + 2 + a); };
  auto rfull_bar = [&](int s) { return bars + 8u * (2 * STAGES + 4 + s); };
  auto rempty_bar = [&](int s) { return bars + 8u * (2 * STAGES + 4 + RS + s); };
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gbase + S::OFF_BARS + S::NUM_BARS * 8);
  const uint32_t tmem_slot = bars + S::NUM_BARS * 8;
  // seen[g] = last residual slab whose arrival epilogue group g has OBSERVED (see the epilogue's ring guard)
  volatile int* seen = reinterpret_cast<volatile int*>(gbase + S::OFF_BARS + S::NUM_BARS * 8 + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int num_kb = 0;
  for (int ps = 0; ps < p.npass; ++ps) num_kb += p.taps * p.pass_cblocks[ps];

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmY);
    if (p.npass > 1) { tma_prefetch_desc(&tmA2); tma_prefetch_desc(&tmB2); }
    if (HAS_RES) tma_prefetch_desc(&tmR);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), TE_COUNT); }
    for (int s = 0; s < RS; ++s) { mbar_init(rfull_bar(s), 1); mbar_init(rempty_bar(s), 4); }
    seen[0] = -1; seen[1] = -1;
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<2 * BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // Programmatic dependent launch: everything above (smem carve-up, mbarrier init, TMEM allocation, descriptor
  // prefetch) touched no activation, so it may overlap the previous kernel's tail.  Let OUR successor become
  // resident as early as SM resources allow, then wait for the predecessor grid to complete and flush before the
  // first global access of any role.  Each kernel waits for its predecessor's FULL completion, so completion is
  // transitive along the stream and the engine's ping-pong buffers stay hazard-free.  (No-ops when the kernel
  // was launched without the attribute.)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0 && lane == 0) {
    // ================= TMA producer =================
    int stage = 0; uint32_t phase = 0;
    TC_TRACE_DECL(0)
    int tr_kb = 0; (void)tr_kb;
    const int KH = p.KH;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
      uint32_t mt, nt, q, ow0, img0, oh0;
      p.fd_ntiles.divmod((uint32_t)t, mt, nt);
      const int n0 = (int)nt * BN;
      const uint32_t m0 = POOL ? (mt >> 1) * (uint32_t)p.pool_rows + (mt & 1) * BM : mt * BM;   // M + 256 < 2^31 (host check)
      p.fd_ow.divmod(m0, q, ow0);
      p.fd_oh.divmod(q, img0, oh0);
      for (int ps = 0; ps < p.npass; ++ps) {
        const CUtensorMap* ma = p.pass_a[ps] ? &tmA2 : &tmA;
        const CUtensorMap* mb = p.pass_b[ps] ? &tmB2 : &tmB;
        const int cblocks = p.pass_cblocks[ps];
        const int bw = (int)ow0 * p.pass_stride[ps] - p.pad, bh = (int)oh0 * p.pass_stride[ps] - p.pad;
        int kcol = 0;                              // (tap * cblocks + cb) * BK
        for (int kh = 0; kh < KH; ++kh) {
          for (int kw = 0; kw < p.KW; ++kw) {
            for (int cb = 0; cb < cblocks; ++cb, kcol += BK) {
              mbar_wait(empty_bar(stage), phase ^ 1u);
              TC_TRACE(1, 0, tr_kb++);
              mbar_expect_tx(full_bar(stage), S::STAGE_BYTES);
              tma_load_im2col_4d(sA + stage * A_STAGE_BYTES, ma, full_bar(stage), cb * BK, bw, bh, (int)img0,
                                 (uint16_t)(kw * p.dil), (uint16_t)(kh * p.dil));
              tma_load_2d(sB + stage * S::B_STAGE_BYTES, mb, full_bar(stage), kcol, n0);
              if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc = make_idesc(BM, BN);
    int stage = 0; uint32_t phase = 0;
    int it = 0;
    TC_TRACE_DECL(1)
    TC_TRACE_AT(4, tr_entry);                  // kernel entry
    TC_TRACE(5, 0, 0);                         // set-up done, predecessor grid complete (griddepcontrol.wait returned)
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t use = (uint32_t)(it >> 1) & 1u;
      TC_TRACE(0, it, 0);
      mbar_wait(tempty_bar(acc), use ^ 1u);   // epilogue has drained this accumulator
      TC_TRACE(1, it, 0);
      tc_fence_after();
      const uint32_t d = tmem_base + (uint32_t)(acc * BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full_bar(stage), phase);
        TC_TRACE(2, it, kb);
        tc_fence_after();
        const uint32_t a = sA + stage * A_STAGE_BYTES, b = sB + stage * S::B_STAGE_BYTES;
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k)
          umma_bf16(d, make_smem_desc(a + k * UMMA_K * 2), make_smem_desc(b + k * UMMA_K * 2), idesc,
                    (kb | k) ? 1u : 0u);
        umma_commit(empty_bar(stage));
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      umma_commit(tfull_bar(acc));
      TC_TRACE(3, it, 0);
    }
  } else if (HAS_RES && warp == 3 && lane == 0) {
    // ================= residual producer =================
    int slot = 0; uint32_t phase = 0;
    TC_TRACE_DECL(2)
    int tr_c = 0; (void)tr_c;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
      uint32_t umt, unt;
      p.fd_ntiles.divmod((uint32_t)t, umt, unt);
      const int n0 = (int)unt * BN, mt = (int)umt;
      const int m0 = POOL ? (mt >> 1) * p.pool_rows + (mt & 1) * BM : mt * BM;
      for (int s = 0; s < NSLAB; ++s) {
        mbar_wait(rempty_bar(slot), phase ^ 1u);
        TC_TRACE(1, 0, tr_c++);
        mbar_expect_tx(rfull_bar(slot), SLAB_BYTES);
        tma_load_2d(sRes + slot * SLAB_BYTES, &tmR, rfull_bar(slot), n0 + s * SLAB, m0);
        if (++slot == RS) { slot = 0; phase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue =================
    const int g = (warp - 4) >> 2;            // epilogue warpgroup
    const int e = warp & 3;                   // TMEM lane quarter == warp id % 4
    const int row = e * 32 + lane;            // row of the tile owned by this thread
    const int et = threadIdx.x - 128 - g * 128;   // 0..127 within the group
    const bool issuer = et == 0;
    const uint32_t swz = (uint32_t)(row & 7);
    const uint32_t sOutG = sOut + (uint32_t)(g * NBUF) * SLAB_BYTES;
    float* s_scale = s_scale_all + g * 2 * S::GSC;
    float* s_shift = s_scale + S::GSC;
    auto group_barrier = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory"); };
    int obuf = 0, it = 0;
#ifdef VLTK_TC_TRACE
    int tr_n = 0; const bool tr_on = p.trace && (int)blockIdx.x == p.trace_cta && issuer; const int tr_role = 3 + g;
#endif
    // entry i of my scale/shift cache, for the it_-th tile of this CTA (first cout n0_), holds column ss_col(...); -1: unused
    constexpr bool SS_PREFETCH = S::GSC <= 128;     // one entry per thread
    auto ss_col = [&](int n0_, int it_, int i) -> int {
      const int sf = (g - (it_ * NSLAB) % EG + EG) % EG;
      const int j = i / SLABC, s = sf + j * EG;
      return s < NSLAB ? n0_ + s * SLABC + (i - j * SLABC) : -1;
    };
    float pf_sc = 1.f, pf_sh = 0.f;
    if (SS_PREFETCH && et < S::GSC && (int)blockIdx.x < p.num_tiles) {
      const int c = ss_col((int)p.fd_ntiles.mod(blockIdx.x) * BN, 0, et);
      if (c >= 0) { pf_sc = p.scale ? __ldg(p.scale + c) : 1.f; pf_sh = p.shift ? __ldg(p.shift + c) : 0.f; }
    }
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      uint32_t umt, unt;
      p.fd_ntiles.divmod((uint32_t)t, umt, unt);
      const int n0 = (int)unt * BN, mt = (int)umt;
      const int m0 = POOL ? (mt >> 1) * p.pool_rows + (mt & 1) * BM : mt * BM;
      (void)m0;
      const int acc = it & 1;
      const uint32_t use = (uint32_t)(it >> 1) & 1u;
      const int c0 = it * NSLAB;              // global index of this tile's first slab
      const int s_first = (g - c0 % EG + EG) % EG;
      // my slabs' BN scale/shift -> my smem cache (this group's previous readers are past their last barrier).  The values
      // were fetched from global memory one tile ago (ss_col / pf_sc / pf_sh above), so their latency, which the pipeline
      // trace showed as a ~800-cycle bubble per tile in both epilogue groups, is off the tile's critical path.
      if constexpr (SS_PREFETCH) {
        if (et < S::GSC) {
          if (ss_col(n0, it, et) >= 0) { s_scale[et] = pf_sc; s_shift[et] = pf_sh; }
          const int tn = t + (int)gridDim.x;
          const int cn = tn < p.num_tiles ? ss_col((int)p.fd_ntiles.mod((uint32_t)tn) * BN, it + 1, et) : -1;
          if (cn >= 0) { pf_sc = p.scale ? __ldg(p.scale + cn) : 1.f; pf_sh = p.shift ? __ldg(p.shift + cn) : 0.f; }
        }
      }
      if (s_first >= NSLAB) continue;         // (NSLAB < EG) this tile belongs to the other group
      if constexpr (!SS_PREFETCH) {
        for (int i = et; i < S::GSC; i += 128) {
          const int j = i / SLABC, s = s_first + j * EG;
          if (s < NSLAB) {
            const int col = n0 + s * SLABC + (i - j * SLABC);
            s_scale[i] = p.scale ? p.scale[col] : 1.f;
            s_shift[i] = p.shift ? p.shift[col] : 0.f;
          }
        }
      }
      TC_TRACE(0, it, 0);
      mbar_wait(tfull_bar(acc), use);
      TC_TRACE(1, it, 0);
      tc_fence_after();
      const uint32_t tacc = tmem_base + ((uint32_t)(e * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
      for (int s = s_first, j = 0; s < NSLAB; s += EG, ++j) {
        const int c = c0 + s;
        const int slot = c % RS;
        const uint32_t rphase = (uint32_t)(c / RS) & 1u;
        (void)slot; (void)rphase;
        uint32_t v[SLABC];
        if constexpr (OUT_F32) {
          tmem_ld32(tacc + (uint32_t)(s * SLABC), v);
          tmem_ld_wait();
        } else {
          uint32_t lo[32], hi[32];
          tmem_ld32(tacc + (uint32_t)(s * SLABC), lo);
          tmem_ld32(tacc + (uint32_t)(s * SLABC + 32), hi);
          tmem_ld_wait();
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) { v[jj] = lo[jj]; v[32 + jj] = hi[jj]; }
        }
        TC_TRACE(2, it, s);
        if (s + EG >= NSLAB) {                // my last slab of this accumulator: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(acc));
        }
        if (HAS_RES) {
          if constexpr (EG == 2) {
            // Parity waits are only sound when the slot's PREVIOUS phase is known complete.  With one consumer that
            // is implied (it consumed that phase itself); with two groups sharing one ring the previous occupant of
            // this slot (slab c - RS, RS odd) belongs to the OTHER group, and its bytes may still be in flight when
            // we get here (TMA boxes land out of order under HBM load) — a parity wait would then return at once
            // on the stale phase.  So first wait until the other group has seen slab c - RS land.  This never adds a
            // stall: slab c is only issued after the other group RELEASED slab c - RS.
            static_assert(RS % 2 == 1, "slab c - RS must belong to the other epilogue group");
            if (c >= RS) {
              for (uint32_t spin = 0; seen[g ^ 1] < c - RS; ++spin)
                if (spin > (1u << 26)) { printf("conv_tc2: residual ring guard timeout\n"); __trap(); }
            }
          }
          mbar_wait(rfull_bar(slot), rphase);
          if (EG == 2 && issuer) seen[g] = c;
        }
        TC_TRACE(3, it, s);
        if (!POOL && issuer) bulk_wait_read<NBUF - 1>();   // the store that last read sOutG[obuf] has drained it
        TC_TRACE(4, it, s);
        group_barrier();                      // sOutG[obuf] reusable (POOL: last slab's column readers done); scale/shift visible
        TC_TRACE(5, it, s);
        const uint32_t orow = sOutG + obuf * SLAB_BYTES + (uint32_t)row * 128u;
        const uint32_t rrow = sRes + slot * SLAB_BYTES + (uint32_t)row * 128u;
        float pv[POOL ? 64 : 1];              // POOL: this row's 64 fp32 epilogue values of the slab (then its column sums)
        (void)pv;
        if constexpr (OUT_F32) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {       // 4 fp32 channels = one 16 B chunk
            const float4 sc = *reinterpret_cast<const float4*>(s_scale + j * SLABC + q * 4);
            const float4 sh = *reinterpret_cast<const float4*>(s_shift + j * SLABC + q * 4);
            float f0 = fmaf(__uint_as_float(v[q * 4 + 0]), sc.x, sh.x), f1 = fmaf(__uint_as_float(v[q * 4 + 1]), sc.y, sh.y);
            float f2 = fmaf(__uint_as_float(v[q * 4 + 2]), sc.z, sh.z), f3 = fmaf(__uint_as_float(v[q * 4 + 3]), sc.w, sh.w);
            if (p.relu) { f0 = fmaxf(f0, 0.f); f1 = fmaxf(f1, 0.f); f2 = fmaxf(f2, 0.f); f3 = fmaxf(f3, 0.f); }
            sts128(orow + (((uint32_t)q ^ swz) << 4),
                   make_uint4(__float_as_uint(f0), __float_as_uint(f1), __float_as_uint(f2), __float_as_uint(f3)));
          }
        } else {
          // software-pipelined over the eight 16 B chunks: the residual chunks are all loaded first and chunk q+1's
          // scale / shift before chunk q's arithmetic and store (two warps per SM sub-partition cannot hide an LDS round
          // trip per chunk on their own: the trace showed 180-250 cycles per chunk)
          float4 sc0[2], sc1[2], sh0[2], sh1[2];
          uint4 rr[HAS_RES ? 8 : 1];
          if (HAS_RES) {
#pragma unroll
            for (int q = 0; q < 8; ++q) rr[q] = lds128(rrow + (((uint32_t)q ^ swz) << 4));
          }
          auto load_chunk = [&](int q, int b) {
            sc0[b] = *reinterpret_cast<const float4*>(s_scale + j * SLABC + q * 8);
            sc1[b] = *reinterpret_cast<const float4*>(s_scale + j * SLABC + q * 8 + 4);
            sh0[b] = *reinterpret_cast<const float4*>(s_shift + j * SLABC + q * 8);
            sh1[b] = *reinterpret_cast<const float4*>(s_shift + j * SLABC + q * 8 + 4);
          };
          load_chunk(0, 0);
#pragma unroll
          for (int q = 0; q < 8; ++q) {       // 8 bf16 channels = one 16 B chunk, stored at chunk q ^ (row % 8)
            const int b = q & 1;
            if (q + 1 < 8) load_chunk(q + 1, b ^ 1);
            const uint32_t coff = ((uint32_t)q ^ swz) << 4;
            // packed fp32x2 arithmetic (FFMA2 / FADD2, IEEE round-to-nearest per lane: same bits as the scalar form,
            // half the issue slots — the slab loop is issue-bound, summary §29)
            float2 g2[4];
            g2[0] = __ffma2_rn(make_float2(__uint_as_float(v[q * 8 + 0]), __uint_as_float(v[q * 8 + 1])), make_float2(sc0[b].x, sc0[b].y), make_float2(sh0[b].x, sh0[b].y));
            g2[1] = __ffma2_rn(make_float2(__uint_as_float(v[q * 8 + 2]), __uint_as_float(v[q * 8 + 3])), make_float2(sc0[b].z, sc0[b].w), make_float2(sh0[b].z, sh0[b].w));
            g2[2] = __ffma2_rn(make_float2(__uint_as_float(v[q * 8 + 4]), __uint_as_float(v[q * 8 + 5])), make_float2(sc1[b].x, sc1[b].y), make_float2(sh1[b].x, sh1[b].y));
            g2[3] = __ffma2_rn(make_float2(__uint_as_float(v[q * 8 + 6]), __uint_as_float(v[q * 8 + 7])), make_float2(sc1[b].z, sc1[b].w), make_float2(sh1[b].z, sh1[b].w));
            if (HAS_RES) {
              const uint32_t rw[4] = {rr[q].x, rr[q].y, rr[q].z, rr[q].w};
#pragma unroll
              for (int j = 0; j < 4; ++j)     // bf16 pair -> fp32 pair: low half << 16, high half masked
                g2[j] = __fadd2_rn(g2[j], make_float2(__uint_as_float(rw[j] << 16), __uint_as_float(rw[j] & 0xFFFF0000u)));
            }
            float f[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) { f[2 * j] = g2[j].x; f[2 * j + 1] = g2[j].y; }
            if (p.relu) {
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
            }
            if constexpr (POOL) {
#pragma unroll
              for (int j = 0; j < 8; ++j) pv[q * 8 + j] = f[j];     // keep the fp32 row in registers
            } else {
              uint4 o;
              __nv_bfloat162* ob = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
              for (int j = 0; j < 4; ++j) ob[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
              sts128(orow + coff, o);
            }
          }
        }
        TC_TRACE(6, it, s);
        if (HAS_RES) {                        // this warp is done with the residual slab
          __syncwarp();
          if (lane == 0) mbar_arrive(rempty_bar(slot));
        }
        if constexpr (POOL) {
          // column sums of this slab's valid rows without staging the tile: each warp reduces its 32 rows with a
          // 62-shuffle butterfly, the group's 4 warps are combined through (double-buffered) staging memory in a
          // fixed order, 64 threads write the tile's partial sums
          const int valid = (mt & 1) ? p.pool_rows - BM : BM;             // rows of this tile that belong to its ROI
          float* comb = reinterpret_cast<float*>(gbase + S::OFF_OUT + (size_t)(g * NBUF) * SLAB_BYTES) + (j & 1) * 256;
          float2 ts = make_float2(0.f, 0.f);
          if (e * 32 < valid) {                                            // warp-uniform
            const bool in = row < valid;
#pragma unroll
            for (int q = 0; q < 64; ++q) pv[q] = in ? pv[q] : 0.f;
            warp_colsum64(pv, lane);
            ts = make_float2(pv[0], pv[1]);
          }
          *reinterpret_cast<float2*>(comb + e * 64 + 2 * lane) = ts;       // lane L owns columns 2L, 2L+1
          group_barrier();
          if (et < 64) {
            const float* c4 = comb + et;
            const float tot = ((c4[0] + c4[64]) + c4[128]) + c4[192];      // fixed warp order
            p.pool_partial[(int64_t)mt * p.Cout + n0 + s * SLABC + et] = tot;
          }
        } else {
          fence_proxy_async_smem();           // generic-proxy smem writes -> visible to the TMA unit
          group_barrier();
          if (issuer) {
            tma_store_2d(&tmY, sOutG + obuf * SLAB_BYTES, n0 + s * SLABC, m0);
            bulk_commit();
          }
          TC_TRACE(7, it, s);
          if (NBUF > 1) obuf ^= 1;
        }
      }
    }
    if (!POOL && issuer) bulk_wait_all();     // all output bytes are in global memory
    TC_TRACE(8, it, 0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<2 * BN>(tmem_base);
}


// =========================================================================================
// v3: CTA pairs (tcgen05 cta_group::2).  Two CTAs of a cluster compute a 256 x 256 output tile with ONE
// tcgen05.mma stream issued by the leader CTA (rank 0): each CTA stages its own 128 pixel rows of A and only HALF of
// the W tile (128 of the 256 couts), the tensor cores of both SMs read both halves.  A stage is 16 + 16 KB instead
// of 16 + 32 KB, so the same 192 KB hold SIX stages instead of four — the mainloop is bound by how many operand
// bytes it keeps in flight (profiles/r01_summary.md §26) — and the W operand crosses L2 -> SM once per pair.
//   * both CTAs' TMA loads complete on the LEADER's full barrier (cta_group::2 loads, peer bit cleared);
//   * tcgen05.commit multicasts to both CTAs' empty / tmem-full barriers;
//   * each CTA drains its own TMEM half; the peer's epilogue warps release the accumulator on the leader's barrier.
// Scope: bf16 out, BN = 256 (the multi-pass K loop of TcConcat is supported).  HAS_RES adds the shortcut tensor through
// a CTA-local 4-slab ring (4 operand stages instead of 6 make room for it); each epilogue group owns two fixed slots of
// the ring, so — unlike v2's shared odd-depth ring — a slot's previous occupant was consumed by the SAME group and plain
// parity waits are sound.  POOL is v2's ROI-aligned fused 14x14 mean: the pair covers one ROI (rank 0 rows [0,128),
// rank 1 rows [128, pool_rows)), so tile numbering and the partial-sum layout are identical to v2's.
constexpr int B3_STAGE_BYTES = 128 * BK * 2;   // half a 256-cout W tile
template <int STAGES, bool HAS_RES, int RING = 4>
struct Smem3 {
  static constexpr int RS = HAS_RES ? RING : 0;     // residual ring slots (half of them private to each epilogue group)
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B3_STAGE_BYTES;   // 32 KB
  static constexpr int OFF_B = STAGES * A_STAGE_BYTES;
  static constexpr int OFF_OUT = STAGES * STAGE_BYTES;
  static constexpr int OFF_RES = OFF_OUT + 2 * SLAB_BYTES;
  static constexpr int OFF_SCALE = OFF_RES + RS * SLAB_BYTES;
  static constexpr int OFF_BARS = OFF_SCALE + 2 * 2 * 128 * 4;
  static constexpr int NUM_BARS = 2 * STAGES + 4 + 2 * RS;
  static constexpr int TOTAL = OFF_BARS + NUM_BARS * 8 + 16;
  static_assert(TOTAL <= 232448, "exceeds the 227 KB shared memory of one sm_100 CTA");
};

template <int STAGES, bool HAS_RES, bool POOL, int RING = 4>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
conv_tc3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmB2,
                const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmR, TcParams2 p) {
  using S = Smem3<STAGES, HAS_RES, RING>;
  static_assert(!POOL || HAS_RES, "the pooled epilogue is the res5 conv3 tail (has a shortcut)");
  static_assert(RING % 2 == 0, "each epilogue group owns half of the ring");
  constexpr int BN = 256, NSLAB = BN / SLAB, EG = 2, RS = S::RS, GRS = RING / 2;
  extern __shared__ __align__(1024) unsigned char smem_dyn3[];
  const uint32_t base = smem_u32(smem_dyn3);
  if (base & 1023u) { if (threadIdx.x == 0) printf("conv_tc3: dynamic smem base %u is not 1024 B aligned\n", base); __trap(); }
  unsigned char* gbase = smem_dyn3;
  const uint32_t sA = base, sB = base + S::OFF_B, sOut = base + S::OFF_OUT, sRes = base + S::OFF_RES;
  float* s_scale_all = reinterpret_cast<float*>(gbase + S::OFF_SCALE);
  const uint32_t bars = base + S::OFF_BARS;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * STAGES + 2 + a); };
  auto rfull_bar = [&](int s) { return bars + 8u * (2 * STAGES + 4 + s); };
  auto rempty_bar = [&](int s) { return bars + 8u * (2 * STAGES + 4 + RS + s); };
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gbase + S::OFF_BARS + S::NUM_BARS * 8);
  const uint32_t tmem_slot = bars + S::NUM_BARS * 8;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  int num_kb = 0;
  for (int ps = 0; ps < p.npass; ++ps) num_kb += p.taps * p.pass_cblocks[ps];
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int num_pair_tiles = p.num_tiles;              // (row-tile pairs) x (cout tiles)
  // first output row of this CTA's half of pair tile t (POOL: one ROI per pair, second half starts at row 128 of the ROI)
  auto tile_m0 = [&](uint32_t pt) -> uint32_t {       // pt = t / n_tiles; M + 256 < 2^31 (host check)
    return POOL ? pt * (uint32_t)p.pool_rows + rank * BM : (pt * 2 + rank) * BM;
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmY);
    if (p.npass > 1) { tma_prefetch_desc(&tmA2); tma_prefetch_desc(&tmB2); }
    if (HAS_RES) tma_prefetch_desc(&tmR);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 2 * 4 * EG); }
    for (int s = 0; s < RS; ++s) { mbar_init(rfull_bar(s), 1); mbar_init(rempty_bar(s), 4); }
    fence_barrier_init();
  }
  cluster_sync_all();                                  // barriers of both CTAs are initialised before anyone signals them
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  cluster_sync_all();

  // Programmatic dependent launch, as in v2: nothing above touched an activation.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0 && lane == 0) {
    // ================= TMA producer (both CTAs: own A rows, own half of W) =================
    int stage = 0; uint32_t phase = 0;
    const int KH = p.KH;
    TC_TRACE_DECL(0)
    int tr_kb = 0; (void)tr_kb;
    for (int t = pair; t < num_pair_tiles; t += npairs) {
      uint32_t pt, nt, q, ow0, img0, oh0;
      p.fd_ntiles.divmod((uint32_t)t, pt, nt);
      const int n0 = (int)nt * BN;
      const uint32_t m0 = tile_m0(pt);
      p.fd_ow.divmod(m0, q, ow0);
      p.fd_oh.divmod(q, img0, oh0);
      for (int ps = 0; ps < p.npass; ++ps) {
        const CUtensorMap* ma = p.pass_a[ps] ? &tmA2 : &tmA;
        const CUtensorMap* mb = p.pass_b[ps] ? &tmB2 : &tmB;
        const int cblocks = p.pass_cblocks[ps];
        const int bw = (int)ow0 * p.pass_stride[ps] - p.pad, bh = (int)oh0 * p.pass_stride[ps] - p.pad;
        int kcol = 0;                                  // (tap * cblocks + cb) * BK
        for (int kh = 0; kh < KH; ++kh) {
          for (int kw = 0; kw < p.KW; ++kw) {
            for (int cb = 0; cb < cblocks; ++cb, kcol += BK) {
              mbar_wait(empty_bar(stage), phase ^ 1u);
              TC_TRACE(1, 0, tr_kb++);
              if (leader) mbar_expect_tx(full_bar(stage), 2 * S::STAGE_BYTES);   // both CTAs' bytes land on this barrier
              tma2_load_im2col_4d(sA + stage * A_STAGE_BYTES, ma, full_bar(stage), cb * BK, bw, bh, (int)img0,
                                  (uint16_t)(kw * p.dil), (uint16_t)(kh * p.dil));
              tma2_load_2d(sB + stage * B3_STAGE_BYTES, mb, full_bar(stage), kcol, n0 + (int)rank * 128);
              if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1 && lane == 0 && leader) {
    // ================= MMA issuer (leader CTA only) =================
    constexpr uint32_t idesc = make_idesc(256, BN);
    int stage = 0; uint32_t phase = 0;
    int it = 0;
    TC_TRACE_DECL(1)
    TC_TRACE(5, 0, 0);
    for (int t = pair; t < num_pair_tiles; t += npairs, ++it) {
      const int acc = it & 1;
      const uint32_t use = (uint32_t)(it >> 1) & 1u;
      TC_TRACE(0, it, 0);
      mbar_wait(tempty_bar(acc), use ^ 1u);            // both CTAs' epilogues have drained this accumulator
      TC_TRACE(1, it, 0);
      tc_fence_after();
      const uint32_t d = tmem_base + (uint32_t)(acc * BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full_bar(stage), phase);
        TC_TRACE(2, it, kb);
        tc_fence_after();
        const uint32_t a = sA + stage * A_STAGE_BYTES, b = sB + stage * B3_STAGE_BYTES;
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k)
          umma2_bf16(d, make_smem_desc(a + k * UMMA_K * 2), make_smem_desc(b + k * UMMA_K * 2), idesc, (kb | k) ? 1u : 0u);
        umma2_commit_mc(empty_bar(stage));
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      umma2_commit_mc(tfull_bar(acc));
      TC_TRACE(3, it, 0);
    }
  } else if (HAS_RES && warp == 3 && lane == 0) {
    // ================= residual producer (CTA-local ring; group g = s % 2 owns slots [g * GRS, (g + 1) * GRS)) =================
    int it = 0;
    for (int t = pair; t < num_pair_tiles; t += npairs, ++it) {
      uint32_t pt, nt;
      p.fd_ntiles.divmod((uint32_t)t, pt, nt);
      const int n0 = (int)nt * BN;
      const int m0 = (int)tile_m0(pt);
      for (int s = 0; s < NSLAB; ++s) {
        const int cg = it * 2 + (s >> 1);              // this slab's index in its group's sequence
        const int slot = (s & 1) * GRS + cg % GRS;
        const uint32_t rphase = (uint32_t)(cg / GRS) & 1u;
        mbar_wait(rempty_bar(slot), rphase ^ 1u);
        mbar_expect_tx(rfull_bar(slot), SLAB_BYTES);
        tma_load_2d(sRes + slot * SLAB_BYTES, &tmR, rfull_bar(slot), n0 + s * SLAB, m0);
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue (each CTA: its own 128 rows) =================
    const int g = (warp - 4) >> 2, e = warp & 3;
    const int row = e * 32 + lane;
    const int et = threadIdx.x - 128 - g * 128;
    const bool issuer = et == 0;
    const uint32_t swz = (uint32_t)(row & 7);
    const uint32_t sOutG = sOut + (uint32_t)g * SLAB_BYTES;
    float* s_scale = s_scale_all + g * 2 * 128;
    float* s_shift = s_scale + 128;
    auto group_barrier = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory"); };
#ifdef VLTK_TC_TRACE
    int tr_n = 0; const bool tr_on = p.trace && (int)blockIdx.x == p.trace_cta && issuer; const int tr_role = 3 + g;
#endif
    // entry `et` of my scale/shift cache holds column ss_col(t) of tile t; fetched one tile ahead (as in v2)
    auto ss_col = [&](int t_) { return (int)p.fd_ntiles.mod((uint32_t)t_) * BN + (g + (et >> 6) * EG) * SLAB + (et & 63); };
    float pf_sc = 1.f, pf_sh = 0.f;
    if (pair < num_pair_tiles) {
      const int c = ss_col(pair);
      pf_sc = p.scale ? __ldg(p.scale + c) : 1.f; pf_sh = p.shift ? __ldg(p.shift + c) : 0.f;
    }
    int it = 0;
    for (int t = pair; t < num_pair_tiles; t += npairs, ++it) {
      uint32_t pt, nt;
      p.fd_ntiles.divmod((uint32_t)t, pt, nt);
      const int n0 = (int)nt * BN;
      const int m0 = (int)tile_m0(pt);
      (void)m0;
      const int acc = it & 1;
      const uint32_t use = (uint32_t)(it >> 1) & 1u;
      {                                                // my slabs (s = g, g + 2): 2 x 64 columns, one entry per thread
        s_scale[et] = pf_sc; s_shift[et] = pf_sh;
        const int tn = t + npairs;
        if (tn < num_pair_tiles) {
          const int cn = ss_col(tn);
          pf_sc = p.scale ? __ldg(p.scale + cn) : 1.f; pf_sh = p.shift ? __ldg(p.shift + cn) : 0.f;
        }
      }
      TC_TRACE(0, it, 0);
      mbar_wait(tfull_bar(acc), use);
      TC_TRACE(1, it, 0);
      tc_fence_after();
      const uint32_t tacc = tmem_base + ((uint32_t)(e * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
      for (int s = g, j = 0; s < NSLAB; s += EG, ++j) {
        const int cg = it * 2 + j;
        const int slot = g * GRS + cg % GRS;
        const uint32_t rphase = (uint32_t)(cg / GRS) & 1u;
        (void)slot; (void)rphase;
        uint32_t v[64];
        {
          uint32_t lo[32], hi[32];
          tmem_ld32(tacc + (uint32_t)(s * SLAB), lo);
          tmem_ld32(tacc + (uint32_t)(s * SLAB + 32), hi);
          tmem_ld_wait();
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) { v[jj] = lo[jj]; v[32 + jj] = hi[jj]; }
        }
        if (s + EG >= NSLAB) {                         // my last slab: release the accumulator on the LEADER's barrier
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if (leader) mbar_arrive(tempty_bar(acc)); else mbar_arrive_leader(tempty_bar(acc)); }
        }
        TC_TRACE(2, it, s);
        if (HAS_RES) mbar_wait(rfull_bar(slot), rphase);
        TC_TRACE(3, it, s);
        if (!POOL && issuer) bulk_wait_read<0>();      // the store that last read sOutG has drained it
        TC_TRACE(4, it, s);
        group_barrier();                               // sOutG reusable (POOL: last slab's column readers done); scale/shift visible
        TC_TRACE(5, it, s);
        const uint32_t orow = sOutG + (uint32_t)row * 128u;
        const uint32_t rrow = sRes + (uint32_t)slot * SLAB_BYTES + (uint32_t)row * 128u;
        (void)rrow;
        float pv[POOL ? 64 : 1];
        (void)pv;
        float4 sc0[2], sc1[2], sh0[2], sh1[2];         // software-pipelined over the 16 B chunks, as in v2
        uint4 rr[HAS_RES ? 8 : 1];
        if (HAS_RES) {
#pragma unroll
          for (int q = 0; q < 8; ++q) rr[q] = lds128(rrow + (((uint32_t)q ^ swz) << 4));
        }
        auto load_chunk = [&](int q, int b) {
          sc0[b] = *reinterpret_cast<const float4*>(s_scale + j * SLAB + q * 8);
          sc1[b] = *reinterpret_cast<const float4*>(s_scale + j * SLAB + q * 8 + 4);
          sh0[b] = *reinterpret_cast<const float4*>(s_shift + j * SLAB + q * 8);
          sh1[b] = *reinterpret_cast<const float4*>(s_shift + j * SLAB + q * 8 + 4);
        };
        load_chunk(0, 0);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int b = q & 1;
          if (q + 1 < 8) load_chunk(q + 1, b ^ 1);
          const uint32_t coff = ((uint32_t)q ^ swz) << 4;
          float2 g2[4];                                // packed fp32x2 arithmetic, as in v2
          g2[0] = __ffma2_rn(make_float2(__uint_as_float(v[q * 8 + 0]), __uint_as_float(v[q * 8 + 1])), make_float2(sc0[b].x, sc0[b].y), make_float2(sh0[b].x, sh0[b].y));
          g2[1] = __ffma2_rn(make_float2(__uint_as_float(v[q * 8 + 2]), __uint_as_float(v[q * 8 + 3])), make_float2(sc0[b].z, sc0[b].w), make_float2(sh0[b].z, sh0[b].w));
          g2[2] = __ffma2_rn(make_float2(__uint_as_float(v[q * 8 + 4]), __uint_as_float(v[q * 8 + 5])), make_float2(sc1[b].x, sc1[b].y), make_float2(sh1[b].x, sh1[b].y));
          g2[3] = __ffma2_rn(make_float2(__uint_as_float(v[q * 8 + 6]), __uint_as_float(v[q * 8 + 7])), make_float2(sc1[b].z, sc1[b].w), make_float2(sh1[b].z, sh1[b].w));
          if (HAS_RES) {
            const uint32_t rw[4] = {rr[q].x, rr[q].y, rr[q].z, rr[q].w};
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
              g2[jj] = __fadd2_rn(g2[jj], make_float2(__uint_as_float(rw[jj] << 16), __uint_as_float(rw[jj] & 0xFFFF0000u)));
          }
          float f[8];
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) { f[2 * jj] = g2[jj].x; f[2 * jj + 1] = g2[jj].y; }
          if (p.relu) {
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) f[jj] = fmaxf(f[jj], 0.f);
          }
          if constexpr (POOL) {
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) pv[q * 8 + jj] = f[jj];
          } else {
            uint4 o;
            __nv_bfloat162* ob = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) ob[jj] = __floats2bfloat162_rn(f[2 * jj], f[2 * jj + 1]);
            sts128(orow + coff, o);
          }
        }
        if (HAS_RES) {                                 // this warp is done with the residual slab
          __syncwarp();
          if (lane == 0) mbar_arrive(rempty_bar(slot));
        }
        if constexpr (POOL) {
          // same reduction tree as v2's POOL epilogue (bit-identical partial sums)
          const int mt = (int)pt * 2 + (int)rank;
          const int valid = rank ? p.pool_rows - BM : BM;
          float* comb = reinterpret_cast<float*>(gbase + S::OFF_OUT + (size_t)g * SLAB_BYTES) + (j & 1) * 256;
          float2 ts = make_float2(0.f, 0.f);
          if (e * 32 < valid) {                        // warp-uniform
            const bool in = row < valid;
#pragma unroll
            for (int q = 0; q < 64; ++q) pv[q] = in ? pv[q] : 0.f;
            warp_colsum64(pv, lane);
            ts = make_float2(pv[0], pv[1]);
          }
          *reinterpret_cast<float2*>(comb + e * 64 + 2 * lane) = ts;
          group_barrier();
          if (et < 64) {
            const float* c4 = comb + et;
            const float tot = ((c4[0] + c4[64]) + c4[128]) + c4[192];
            p.pool_partial[(int64_t)mt * p.Cout + n0 + s * SLAB + et] = tot;
          }
        } else {
          TC_TRACE(6, it, s);
          fence_proxy_async_smem();
          group_barrier();
          if (issuer) {
            tma_store_2d(&tmY, sOutG, n0 + s * SLAB, m0);
            bulk_commit();
          }
          TC_TRACE(7, it, s);
        }
      }
    }
    if (!POOL && issuer) bulk_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                  // the peer may still be reading / the leader still issuing into its TMEM
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
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

// [rows, cols] bf16 row-major (row stride ld elements), box = one 128 x 64 epilogue slab
int make_rowmajor_map(const void* ptr, int64_t rows, int cols, int ld, bool f32, CUtensorMap* out) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * (f32 ? 4 : 2)};
  cuuint32_t box[2] = {(cuuint32_t)(f32 ? 32 : SLAB), (cuuint32_t)BM};   // 128 B rows either way
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode_tiled(out, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                              const_cast<void*>(ptr), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VLTK_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) for [%lld,%d] ld %d", (int)r, (long long)rows, cols, ld);
  return 0;
}

int num_sms() {            // of the CURRENT device (a process may drive several)
  static int cache[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return 148;
  if (!cache[dev]) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cache[dev] = n > 0 ? n : 148;
  }
  return cache[dev];
}

// feats[roi][c] = (partial of the ROI's first tile + partial of its second tile) / rows
__global__ void pool_finish_kernel(const float* __restrict__ partial, float* __restrict__ out, int rois, int rows, int C) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int c4n = C / 4;
  if (i >= (int64_t)rois * c4n) return;
  const int c = (int)(i % c4n) * 4, r = (int)(i / c4n);
  const float4 a = *reinterpret_cast<const float4*>(partial + ((int64_t)r * 2) * C + c);
  const float4 b = *reinterpret_cast<const float4*>(partial + ((int64_t)r * 2 + 1) * C + c);
  const float d = (float)rows;
  *reinterpret_cast<float4*>(out + (int64_t)r * C + c) = make_float4((a.x + b.x) / d, (a.y + b.y) / d, (a.z + b.z) / d, (a.w + b.w) / d);
}

struct Maps { CUtensorMap a, a2, b, b2, y, r; };

template <int BN, int STAGES, bool HAS_RES, bool OUT_F32, bool POOL = false, int EG = 1>
int launch2e(const Maps& m, TcParams2 tp, int cout_pad, cudaStream_t st) {
  using S = Smem2<BN, STAGES, HAS_RES, EG>;
  static DeviceOnce once;
  if (once.first()) {
    VLTK_CUDA(cudaFuncSetAttribute(conv_tc2_kernel<BN, STAGES, HAS_RES, OUT_F32, POOL, EG>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
  }
  tp.n_tiles = cout_pad / BN;
  tp.fd_ntiles.init((uint32_t)tp.n_tiles);
  const int64_t tiles = (POOL ? 2 * (tp.M / tp.pool_rows) : ceil_div64(tp.M, BM)) * tp.n_tiles;
  VLTK_CHECK(tiles < (1ll << 31), "conv_tc: too many tiles");
  tp.num_tiles = (int)tiles;
  const int grid = (int)std::min<int64_t>(tiles, num_sms());  // persistent: one CTA per SM
  static const bool use_pdl = [] { const char* e = getenv("VLTK_PDL"); return !(e && e[0] == '0'); }();
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128 + 128 * EG); cfg.dynamicSmemBytes = S::TOTAL; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = use_pdl ? 1 : 0;
  VLTK_CUDA(cudaLaunchKernelEx(&cfg, conv_tc2_kernel<BN, STAGES, HAS_RES, OUT_F32, POOL, EG>, m.a, m.a2, m.b, m.b2, m.y, m.r, tp));
  VLTK_LAUNCH_CHECK();
  return 0;
}

template <int STAGES, bool HAS_RES, bool POOL, int RING = 4>
int launch3(const Maps& m, TcParams2 tp, int cout_pad, cudaStream_t st) {
  using S = Smem3<STAGES, HAS_RES, RING>;
  static DeviceOnce once;
  if (once.first()) {
    VLTK_CUDA(cudaFuncSetAttribute(conv_tc3_kernel<STAGES, HAS_RES, POOL, RING>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
  }
  tp.n_tiles = cout_pad / 256;
  tp.fd_ntiles.init((uint32_t)tp.n_tiles);
  const int64_t pair_tiles = (POOL ? tp.M / tp.pool_rows : ceil_div64(ceil_div64(tp.M, BM), 2)) * tp.n_tiles;
  VLTK_CHECK(pair_tiles < (1ll << 31), "conv_tc: too many tiles");
  tp.num_tiles = (int)pair_tiles;
  const int pairs = (int)std::min<int64_t>(pair_tiles, num_sms() / 2);
  static const bool use_pdl = [] { const char* e = getenv("VLTK_PDL"); return !(e && e[0] == '0'); }();
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(384); cfg.dynamicSmemBytes = S::TOTAL; cfg.stream = st;   // cluster dims are compiled in
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = use_pdl ? 1 : 0;
  VLTK_CUDA(cudaLaunchKernelEx(&cfg, conv_tc3_kernel<STAGES, HAS_RES, POOL, RING>, m.a, m.a2, m.b, m.b2, m.y, m.r, tp));
  VLTK_LAUNCH_CHECK();
  return 0;
}

// Two epilogue warpgroups unless VLTK_EPI_GROUPS=1 (A/B switch).
template <int BN, int STAGES, bool HAS_RES, bool OUT_F32, bool POOL = false>
int launch2(const Maps& m, TcParams2 tp, int cout_pad, cudaStream_t st) {
  static const bool one = [] { const char* e = getenv("VLTK_EPI_GROUPS"); return e && e[0] == '1'; }();
  if (one) return launch2e<BN, STAGES, HAS_RES, OUT_F32, POOL, 1>(m, tp, cout_pad, st);
  return launch2e<BN, STAGES, HAS_RES, OUT_F32, POOL, 2>(m, tp, cout_pad, st);
}

template <int BN, int STAGES>
int launch(const CUtensorMap& a, const CUtensorMap& b, const TcParams& tp, int cout_pad, cudaStream_t st) {
  using S = Smem<BN, STAGES>;
  static DeviceOnce once;
  if (once.first()) {
    VLTK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
  }
  dim3 grid(cout_pad / BN, (unsigned)ceil_div64(tp.M, BM));
  conv_tc_kernel<BN, STAGES><<<grid, TC_THREADS, S::TOTAL, st>>>(a, b, tp);
  VLTK_LAUNCH_CHECK();
  return 0;
}

}  // namespace

// ---- descriptor encoders shared with conv_tcx.cu (fp16 split-plane kernels)
int tc_num_sms() { return num_sms(); }

int tc_pool_finish(const float* partial, float* out, int rois, int rows, int C, cudaStream_t st) {
  const int64_t tot = (int64_t)rois * (C / 4);
  if (tot == 0) return 0;
  pool_finish_kernel<<<(unsigned)ceil_div64(tot, 256), 256, 0, st>>>(partial, out, rois, rows, C);
  VLTK_LAUNCH_CHECK();
  return 0;
}

int tc_encode_tiled(CUtensorMap* out, CUtensorMapDataType dt, const void* ptr, uint64_t cols, uint64_t rows,
                    uint64_t row_stride_bytes, uint32_t box_cols, uint32_t box_rows, bool promote256) {
  if (load_driver_entry_points()) return -1;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)row_stride_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode_tiled(out, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, promote256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VLTK_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) for [%llu,%llu]", (int)r, (unsigned long long)rows, (unsigned long long)cols);
  return 0;
}

int tc_encode_im2col(CUtensorMap* out, CUtensorMapDataType dt, const void* ptr, int N, int H, int W, int C, int ld_elems,
                     int esz, int KH, int KW, int stride, int pad, int dil) {
  if (load_driver_entry_points()) return -1;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)ld_elems * esz, (cuuint64_t)W * ld_elems * esz, (cuuint64_t)H * W * ld_elems * esz};
  int lower[2] = {-pad, -pad};
  int upper[2] = {pad - (KW - 1) * dil, pad - (KH - 1) * dil};
  cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = g_encode_im2col(out, dt, 4, const_cast<void*>(ptr), dims, strides, lower, upper, BK, BM, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VLTK_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeIm2col failed (%d) for x[%d,%d,%d,%d] k%d s%d p%d d%d", (int)r, N, H, W, C, KH, stride, pad, dil);
  int drv = 0;   // same small-tensor descriptor fix-up as make_a_map
  if (cudaDriverGetVersion(&drv) == cudaSuccess && drv <= 13010 && (uint64_t)N * H * W * ld_elems * esz < 131072)
    reinterpret_cast<uint64_t*>(out)[1] &= ~(1ull << 21);
  return 0;
}

// cta_group::2 dispatch knobs: environment defaults, overridable at run time (vltk_conv_tc_set_cta_pairs)
std::atomic<int> g_cta2_min_m{[] { const char* e = getenv("VLTK_CTA2"); return e ? atoi(e) : 32768; }()};
std::atomic<int> g_cta2_res{[] { const char* e = getenv("VLTK_CTA2_RES"); return e ? atoi(e) : 0; }()};

std::atomic<long long*> g_trace_buf{nullptr};
std::atomic<int> g_trace_cap{0}, g_trace_cta{0};

int conv_tc_set_trace(void* dev_buf, int cap_per_role, int cta) {
#ifdef VLTK_TC_TRACE
  g_trace_buf.store((long long*)dev_buf); g_trace_cap.store(cap_per_role); g_trace_cta.store(cta);
  return 0;
#else
  (void)dev_buf; (void)cap_per_role; (void)cta;
  VLTK_CHECK(false, "conv_tc: this library was built without -DVLTK_TC_TRACE (VLTK_TRACE=1 csrc/build.sh)");
  return -1;
#endif
}

void conv_tc_set_cta_pairs(int min_pixels, int residual_layers) {
  if (min_pixels >= 0) g_cta2_min_m.store(min_pixels);
  if (residual_layers >= 0) g_cta2_res.store(residual_layers);
}

size_t conv_tc_pool_partial_bytes(int64_t M, int cout) { return ((size_t)(M / (BM + 1)) + 1) * 2 * cout * sizeof(float); }   // 2 tiles per ROI, rows > BM

int conv_tc_launch(const ConvProblem& p, const bf16* w_nk, int cout_pad, TensorMapCache* cache, cudaStream_t st,
                   const TcSplit* split, const TcPool* pool, const TcConcat* concat) {
  const bool out_f32 = p.out_dtype == DT_F32;
  const bool is_split = split && split->x_lo && split->w_lo;
  const bool is_wsplit = split && !split->x_lo && split->w_lo;      // exact bf16 activations, weights = hi + lo
  const bool is_concat = concat && concat->x2 && concat->w2;
  VLTK_CHECK(!(is_concat && (is_split || is_wsplit || (pool && pool->out))), "conv_tc: concat excludes split / pooled epilogue");
  VLTK_CHECK(!is_concat || (p.KH == 1 && p.KW == 1 && p.pad == 0 && concat->Cin2 % BK == 0 && concat->ldx2 % 8 == 0 &&
                            (concat->H2 - 1) / concat->stride2 + 1 == p.OH && (concat->W2 - 1) / concat->stride2 + 1 == p.OW),
             "conv_tc: concat needs a 1x1 primary conv and a second operand on the same output grid");
  VLTK_CHECK(p.in_dtype == DT_BF16, "conv_tc: bf16 operands only");
  VLTK_CHECK(!(out_f32 && p.residual), "conv_tc: fp32 output has no residual path");
  VLTK_CHECK(p.Cin % BK == 0, "conv_tc: Cin=%d must be a multiple of %d", p.Cin, BK);
  VLTK_CHECK(cout_pad % 64 == 0 && p.Cout <= cout_pad, "conv_tc: bad cout_pad");
  VLTK_CHECK(p.ldy % (out_f32 ? 4 : 8) == 0 && p.ldx % 8 == 0 && (!p.residual || p.ldr % 8 == 0), "conv_tc: rows must be 16-byte aligned");
  VLTK_CHECK(p.Cout % 32 == 0, "conv_tc: Cout=%d must be a multiple of 32", p.Cout);
  if (load_driver_entry_points()) return -1;
  const int64_t M = (int64_t)p.N * p.OH * p.OW;
  if (M == 0) return 0;
  const int K = p.KH * p.KW * p.Cin;
  // Cout tile.  Measured on B200 (profiles/r01_summary.md §5): layers with K >= 512 are MMA-bound and want
  // the widest tile (N=128 MMAs read A 4 KB + B 4 KB per 64 cycles = the whole 128 B/clk smem port, so
  // they do NOT run at half the cost of N=256); layers with K <= 256 are epilogue-bound (1-4 k-blocks
  // per tile) and run 7-34 % faster with BN=128, whose finer tiles keep both TMEM accumulators busy.
  int bn = (cout_pad % 256 == 0) ? 256 : (cout_pad % 128 == 0 ? 128 : 64);
  static const int smallk = [] { const char* e = getenv("VLTK_SMALLK"); return e ? atoi(e) : 256; }();   // tuning knob
  if (K + (is_concat ? concat->Cin2 : 0) <= smallk && cout_pad % 128 == 0) bn = 128;
  // (BN=128 + a 5-slab residual ring for the K=512 residual layers of res5 was measured: 1.07 -> 1.21 ms, slower — the
  //  doubled A traffic and the N=128 MMA rate cost more than the deeper ring returns; profiles/r01_summary.md §25)
  if (cache->maps.size() > 8192) cache->maps.clear();   // keys hold buffer addresses: bound growth across reallocations
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

  static const bool use_v1 = [] { const char* e = getenv("VLTK_TC_V1"); return e && e[0] == '1'; }();
  VLTK_CHECK(!(use_v1 && is_concat), "conv_tc: the v1 kernel has no concat path");
  if (!use_v1 || out_f32 || is_split || is_wsplit) {
    VLTK_CHECK(p.Cout % 64 == 0 && cout_pad == p.Cout, "conv_tc: Cout=%d must be a multiple of 64", p.Cout);
    Maps m;
    m.a = ta; m.b = tb; m.a2 = ta; m.b2 = tb;
    auto cached = [&](const TensorMapCache::Key& k, CUtensorMap* dst, auto make) -> int {
      auto it = cache->maps.find(k);
      if (it == cache->maps.end()) {
        if (make(dst)) return -1;
        cache->maps[k] = *dst;
      } else *dst = it->second;
      return 0;
    };
    if (cached(TensorMapCache::Key(p.y, (int)M, p.Cout, p.ldy, out_f32 ? 1 : 0, 0, 0, 0, 0, 0, 2), &m.y,
               [&](CUtensorMap* d) { return make_rowmajor_map(p.y, M, p.Cout, p.ldy, out_f32, d); })) return -1;
    m.r = m.y;
    if (p.residual &&
        cached(TensorMapCache::Key(p.residual, (int)M, p.Cout, p.ldr, 0, 0, 0, 0, 0, 0, 3), &m.r,
               [&](CUtensorMap* d) { return make_rowmajor_map(p.residual, M, p.Cout, p.ldr, false, d); })) return -1;
    if (is_split) {
      ConvProblem plo = p;
      plo.x = split->x_lo;
      if (cached(TensorMapCache::Key(plo.x, p.N, p.H, p.W, p.Cin, p.ldx, p.KH, p.stride, p.pad, p.dil, 0), &m.a2,
                 [&](CUtensorMap* d) { return make_a_map(plo, d); })) return -1;
      if (cached(TensorMapCache::Key(split->w_lo, K, cout_pad, bn, 0, 0, 0, 0, 0, 0, 1), &m.b2,
                 [&](CUtensorMap* d) { return make_b_map(split->w_lo, K, cout_pad, bn, d); })) return -1;
    }
    if (is_wsplit &&
        cached(TensorMapCache::Key(split->w_lo, K, cout_pad, bn, 0, 0, 0, 0, 0, 0, 1), &m.b2,
               [&](CUtensorMap* d) { return make_b_map(split->w_lo, K, cout_pad, bn, d); })) return -1;
    if (is_concat) {
      ConvProblem p2 = p;
      p2.x = concat->x2; p2.ldx = concat->ldx2; p2.H = concat->H2; p2.W = concat->W2; p2.Cin = concat->Cin2;
      p2.stride = concat->stride2;
      if (cached(TensorMapCache::Key(p2.x, p2.N, p2.H, p2.W, p2.Cin, p2.ldx, p2.KH, p2.stride, p2.pad, p2.dil, 0), &m.a2,
                 [&](CUtensorMap* d) { return make_a_map(p2, d); })) return -1;
      if (cached(TensorMapCache::Key(concat->w2, concat->Cin2, cout_pad, bn, 0, 0, 0, 0, 0, 0, 1), &m.b2,
                 [&](CUtensorMap* d) { return make_b_map(concat->w2, concat->Cin2, cout_pad, bn, d); })) return -1;
    }
    TcParams2 t2;
    t2.scale = p.scale; t2.shift = p.shift; t2.M = M; t2.Cout = p.Cout; t2.relu = p.relu;
    t2.OH = p.OH; t2.OW = p.OW; t2.stride = p.stride; t2.pad = p.pad; t2.dil = p.dil; t2.KH = p.KH; t2.KW = p.KW;
    t2.taps = p.KH * p.KW; t2.cblocks = p.Cin / BK; t2.n_tiles = 0; t2.num_tiles = 0;
    t2.pool_partial = nullptr; t2.pool_rows = 1;
    VLTK_CHECK(M < (1ll << 31) - 512, "conv_tc: M=%lld output pixels exceed the 32-bit tile arithmetic", (long long)M);
    t2.fd_ow.init((uint32_t)p.OW); t2.fd_oh.init((uint32_t)p.OH);      // fd_ntiles: set by the launcher with n_tiles
    t2.trace = g_trace_buf.load(std::memory_order_relaxed); t2.trace_cap = g_trace_cap.load(std::memory_order_relaxed);
    t2.trace_cta = g_trace_cta.load(std::memory_order_relaxed);
    // (an L2 prefetch of the residual tensor was tried and measured slower: the residual layers are DRAM-bandwidth-
    // bound, not latency-bound — profiles/r01_summary.md §19)
    t2.npass = is_split ? 3 : 1;                       // hi*hi, lo*hi, hi*lo
    t2.pass_a[0] = 0; t2.pass_a[1] = 1; t2.pass_a[2] = 0;
    t2.pass_b[0] = 0; t2.pass_b[1] = 0; t2.pass_b[2] = 1;
    for (int i = 0; i < 3; ++i) { t2.pass_cblocks[i] = t2.cblocks; t2.pass_stride[i] = p.stride; }
    if (is_wsplit) { t2.npass = 2; t2.pass_a[1] = 0; t2.pass_b[1] = 1; }   // x*w_hi + x*w_lo
    if (is_concat) {                                   // pass 1 = the second operand pair
      t2.npass = 2; t2.pass_a[1] = 1; t2.pass_b[1] = 1;
      t2.pass_cblocks[1] = concat->Cin2 / BK; t2.pass_stride[1] = concat->stride2;
    }
    // CTA pairs (cta_group::2, conv_tc3_kernel) for the BN = 256 bf16-out layers with at least VLTK_CTA2 output pixels
    // (default 32768; 0 = never): +5 % on the whole step (profiles/r01_summary.md §27).  The residual / pooled layers
    // stay on v2 unless VLTK_CTA2_RES=1: measured 3 % slower on pairs (K = 512 leaves 8 k-blocks per tile, and the
    // accumulator hand-off then waits for the slower of TWO residual streams; §27).
    const int cta2 = g_cta2_min_m.load(std::memory_order_relaxed);
    const int cta2_res = g_cta2_res.load(std::memory_order_relaxed);   // 0: residual layers on v2; 1: pairs, 4 stages + 4-slab ring; 2: 3 + 6
    const bool pooled = pool && pool->out;
    const bool use3 = cta2 > 0 && bn == 256 && !out_f32 && !is_split && !is_wsplit && M >= cta2 &&
                      (cta2_res || !(p.residual || pooled));
    Maps m3 = m;
    if (use3) {                                        // each CTA of the pair loads half of the W tile
      if (cached(TensorMapCache::Key(w_nk, K, cout_pad, 128, 0, 0, 0, 0, 0, 0, 1), &m3.b,
                 [&](CUtensorMap* d) { return make_b_map(w_nk, K, cout_pad, 128, d); })) return -1;
      m3.b2 = m3.b;
      if (is_concat &&
          cached(TensorMapCache::Key(concat->w2, concat->Cin2, cout_pad, 128, 0, 0, 0, 0, 0, 0, 1), &m3.b2,
                 [&](CUtensorMap* d) { return make_b_map(concat->w2, concat->Cin2, cout_pad, 128, d); })) return -1;
    }
    if (out_f32) {
      if (bn == 256) return launch2<256, 4, false, true>(m, t2, cout_pad, st);
      if (bn == 128) return launch2<128, 4, false, true>(m, t2, cout_pad, st);
      return launch2<64, 4, false, true>(m, t2, cout_pad, st);
    }
    if (pooled) {
      VLTK_CHECK(p.residual && bn == 256 && !out_f32 && !is_split, "conv_tc: the pooled epilogue is built for the res5 conv3 shape (residual, Cout %% 256 == 0, K > 256)");
      VLTK_CHECK(pool->rows > BM && pool->rows <= 2 * BM && M % pool->rows == 0 && p.Cout % 4 == 0,
                 "conv_tc: pool_rows=%d must be in (%d, %d] and divide M", pool->rows, BM, 2 * BM);
      t2.pool_partial = pool->partial; t2.pool_rows = pool->rows;
      if ((use3 && cta2_res == 1) ? launch3<4, true, true>(m3, t2, cout_pad, st) : launch2<256, 3, true, false, true>(m, t2, cout_pad, st)) return -1;
      const int rois = (int)(M / pool->rows);
      const int64_t tot = (int64_t)rois * (p.Cout / 4);
      pool_finish_kernel<<<(unsigned)ceil_div64(tot, 256), 256, 0, st>>>(pool->partial, pool->out, rois, pool->rows, p.Cout);
      VLTK_LAUNCH_CHECK();
      return 0;
    }
    if (p.residual) {
      if (use3) return cta2_res == 2 ? launch3<3, true, false, 6>(m3, t2, cout_pad, st) : launch3<4, true, false>(m3, t2, cout_pad, st);
      if (bn == 256) return launch2<256, 3, true, false>(m, t2, cout_pad, st);
      if (bn == 128) return launch2<128, 3, true, false>(m, t2, cout_pad, st);
      return launch2<64, 4, true, false>(m, t2, cout_pad, st);
    }
    if (use3) return launch3<6, false, false>(m3, t2, cout_pad, st);
    static const bool probe3 = [] { const char* e = getenv("VLTK_PROBE_STAGES3"); return e && e[0] == '1'; }();
    if (bn == 256 && probe3) return launch2<256, 3, false, false>(m, t2, cout_pad, st);   // diagnosis knob: 3 instead of 4 stages
    if (bn == 256) return launch2<256, 4, false, false>(m, t2, cout_pad, st);
    if (bn == 128) return launch2<128, 4, false, false>(m, t2, cout_pad, st);
    return launch2<64, 4, false, false>(m, t2, cout_pad, st);
  }
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
