// fp32-FAITHFUL implicit-GEMM convolution on the tensor cores ("exact_tc" mode):
//
//   every fp32 activation x is stored as two fp16 planes   x = hi + lo' * 2^-11   (hi = fp16(x), lo' = fp16((x - hi) * 2^11))
//   every fp32 weight row (scaled by a per-row power of two 2^s, folded back into the epilogue scale) as three planes
//       WA = fp16(2^s w) ,  WB = WA * 2^-11 ,  WC = fp16(2^s w - WA)
//   and the product is   x w 2^s  =  lo' * WB  +  hi * WC  +  hi * WA     (the dropped lo * w_lo term is 2^-24 relative)
//
// i.e. THREE tcgen05.mma kind::f16 passes (fp16 x fp16 products are exact in the multiplier) instead of one.  The tensor
// core's fp32 accumulator TRUNCATES towards zero on every MMA (measured: tools/acc_probe.py, profiles/r02_acc_probe.json —
// ~0.85 ulp of the running sum per instruction, 240 ulp over a K = 4608 reduction), so a plain three-pass K loop is 50x
// noisier than an fp32 FMA chain.  Hence TWO-LEVEL ACCUMULATION: the K loop is cut into chunks of `chunk_kb` k-blocks
// (default 4 = 256 input channels); within a chunk the two small passes run first and the hi*hi pass last, into one of
// two alternating TMEM accumulators that starts from zero; the epilogue warps drain every finished chunk with
// tcgen05.ld and add it to a REGISTER-resident fp32 running sum with IEEE round-to-nearest while the tensor core
// fills the other accumulator.  Truncation then acts on chunk-sized partial sums only (<= 16 MMAs of the big pass).
//
// Pipeline roles as in conv_tc2_kernel (TMA producer / MMA issuer / TMEM allocator / epilogue warpgroups); the
// epilogue writes the two output planes through two 128B-swizzled staging buffers per group and TMA stores.  A
// residual tensor (two planes) reaches it by two routes: the first 64-column unit of a group is TMA-prefetched INTO
// those staging buffers a whole tile ahead and consumed in place, later units are fetched into registers with 32-byte
// loads behind the first unit's arithmetic (the staging buffers are busy until the first unit's store has read them).
//
// Variants (template parameters):  PAIR = CTA pairs, tcgen05 cta_group::2 (all res5 layers: five 32 KB stages, half a W
// tile per CTA);  K-concatenation (run time, kb1 < num_kb) = a second 1x1 operand accumulated into the same tile, i.e. a
// projection block's conv3 + shortcut as ONE GEMM;  POOL = the res5 tail's 14x14 mean reduced in the epilogue instead of
// storing the tile;  OUT_F32 = fp32 rows out (RPN head, predictor linears).
//
// Reference layers: every Conv2d+BN(+ReLU)(+residual) of res2-res5, the RPN convs and the predictor linears
// (frcnn.py:794-822, 963-979, 1345-1355, 1561-1572, 1726-1740), computed in fp32 by the reference.
#include <cuda_fp16.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>

#include "conv_tc.cuh"
#include "tc_ptx.cuh"

namespace vltk {

namespace {

constexpr int XSLAB = 64;                      // columns per epilogue unit (one 128 B fp16 staging row)
constexpr int XBUF_BYTES = BM * 128;           // one staging buffer: 128 rows x 128 B
constexpr float LO_SCALE = 2048.f, LO_INV = 1.f / 2048.f;

// kind::f16 instruction descriptor with fp16 A and B (format 0), fp32 accumulate, both K-major
__host__ __device__ constexpr uint32_t make_idesc_f16(int m, int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

struct TcxParams {
  const float* scale; const float* shift;      // per output channel; scale already carries the weight row's 2^-s
  FastDiv fd_ntiles, fd_ow, fd_oh;
  int64_t M;
  int Cout, relu;
  int stride, pad, dil, KW, num_kb, lg_cblocks;
  int cin;                                     // channel offset of the lo' plane in x (= Cin)
  int K;                                       // element offset between the weight planes WA | WB | WC
  int n_tiles, num_tiles, chunk_kb;
  const __half* res; int ldr;                  // residual tensor (two planes per row) and its row stride in elements
  float* pool_partial; int pool_rows;          // POOL: per (ROI half, cout) column sums; rows per ROI
  int kb1, cin2, stride2;                      // K-concatenated second operand: k-blocks [kb1, num_kb) come from tmA2
  int out_plane, res_plane;                    // column offset of the lo' plane in y / residual
};

template <int BN, int STAGES, int EG, bool PAIR = false>
struct SmemX {
  static constexpr int NS_OWN = BN / XSLAB / EG;                 // 64-column units per epilogue group
  static constexpr int B_STAGE_BYTES = (PAIR ? BN / 2 : BN) * BK * 2;   // a CTA of a pair stages half of the W tile
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int OFF_B = STAGES * A_STAGE_BYTES;
  static constexpr int OFF_OUT = STAGES * STAGE_BYTES;
  static constexpr int OFF_SCALE = OFF_OUT + EG * 2 * XBUF_BYTES;
  static constexpr int OFF_BARS = OFF_SCALE + EG * NS_OWN * XSLAB * 2 * 4;
  static constexpr int NUM_BARS = 2 * STAGES + 4 + EG;
  static constexpr int TOTAL = OFF_BARS + NUM_BARS * 8 + 16;
  static_assert(NS_OWN >= 1 && NS_OWN * EG * XSLAB == BN, "BN must split evenly over the epilogue groups");
  static_assert(TOTAL <= 232448, "exceeds the 227 KB shared memory of one sm_100 CTA");
};

// (x0, x1) -> packed fp16 pairs hi = fp16(x), lo' = fp16((x - hi) * 2^11); saturating to the largest finite fp16
// (|x| > 65504 cannot be represented: documented limit of the mode).  7 instructions per pair: the tile finish is what
// bounds the K <= 512 layers, so it works on packed fp32x2 / f16x2 throughout.
__device__ __forceinline__ void split_pair(float2 x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x.y), "f"(x.x));
  const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  const float2 d = __fmul2_rn(make_float2(x.x - h.x, x.y - h.y), make_float2(LO_SCALE, LO_SCALE));   // both exact
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(d.y), "f"(d.x));
}

// PAIR: CTA pairs (tcgen05 cta_group::2, launched as 2-CTA clusters), as conv_tc3_kernel does for the bf16 mode: the pair
// computes a 256-pixel x 256-cout tile with ONE MMA stream issued by the leader (rank 0); each CTA stages its own 128
// im2col rows and HALF of the W tile (a stage is 32 KB instead of 48 KB -> five stages instead of three, and W crosses
// L2 -> SM once per pair); all TMA bytes of a stage complete on the leader's full barrier, tcgen05.commit multicasts to
// both CTAs' empty / accumulator-full barriers, each CTA promotes and finishes its own 128 TMEM lanes and releases the
// chunk accumulator on the leader's barrier.  Same k-block / pass / chunk order, same arithmetic: bit-identical outputs.
// POOL (pairs only): the res5 tail's 14x14 mean (frcnn.py:1401) fused as in conv_tc3_kernel — a pair tile is ONE ROI
// (rank 0 rows [0,128), rank 1 rows [128, pool_rows), the rows past the ROI computed and masked), the finish reduces the
// fp32 tile over its rows (fixed-order warp butterfly + fixed-order combine of the group's four warps) and writes one
// partial sum per (ROI half, cout) instead of storing the tile; an ROI's mean does not depend on its place in the batch.
template <int BN, int STAGES, int EG, bool HAS_RES, bool OUT_F32, bool BIGREG, bool PAIR = false, bool POOL = false>
__global__ void __launch_bounds__(128 + 128 * EG, 1)
conv_tcx_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmY,
                const __grid_constant__ CUtensorMap tmR, TcxParams p) {
  using S = SmemX<BN, STAGES, EG, PAIR>;
  static_assert(!(HAS_RES && OUT_F32), "fp32 output has no residual path");
  static_assert(!PAIR || BN == 256, "CTA pairs share one 256-cout W tile");
  static_assert(!POOL || (PAIR && HAS_RES && !OUT_F32), "the pooled finish is the res5 conv3 tail on CTA pairs");
  constexpr int NS_OWN = S::NS_OWN;
  extern __shared__ __align__(1024) unsigned char smem_dynx[];
  const uint32_t base = smem_u32(smem_dynx);
  if (base & 1023u) {
    if (threadIdx.x == 0) printf("conv_tcx: dynamic smem base %u is not 1024 B aligned\n", base);
    __trap();
  }
  unsigned char* gbase = smem_dynx;
  const uint32_t sA = base, sB = base + S::OFF_B, sOut = base + S::OFF_OUT;
  float* s_scale_all = reinterpret_cast<float*>(gbase + S::OFF_SCALE);
  const uint32_t bars = base + S::OFF_BARS;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * STAGES + 2 + a); };
  auto res_bar = [&](int g) { return bars + 8u * (2 * STAGES + 4 + g); };
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gbase + S::OFF_BARS + S::NUM_BARS * 8);
  const uint32_t tmem_slot = bars + S::NUM_BARS * 8;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunks = (p.num_kb + p.chunk_kb - 1) / p.chunk_kb;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  // work unit w (a tile, or a pair tile) -> first output row of THIS CTA and first cout
  const int unit0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, unit_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  auto unit_m0 = [&](uint32_t mt) -> uint32_t {
    return POOL ? mt * (uint32_t)p.pool_rows + rank * BM : (PAIR ? (mt * 2 + rank) * BM : mt * BM);
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB);
    if (!POOL) tma_prefetch_desc(&tmY);
    if (p.kb1 < p.num_kb) tma_prefetch_desc(&tmA2);
    if (HAS_RES) tma_prefetch_desc(&tmR);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), (PAIR ? 2 : 1) * 4 * EG); }
    for (int g = 0; g < EG; ++g) mbar_init(res_bar(g), 1);
    fence_barrier_init();
  }
  if constexpr (PAIR) {
    cluster_sync_all();                            // barriers of both CTAs are initialised before anyone signals them
    if (warp == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(2 * BN) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  } else {
    if (warp == 2) tmem_alloc<2 * BN>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  if constexpr (PAIR) cluster_sync_all();

  // Programmatic dependent launch (see conv_tc2_kernel): nothing above touched an activation.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  // The epilogue threads keep a [NS_OWN x 64] fp32 running sum in registers: give them the register file of the
  // (single-thread) producer / issuer warps.
  // pass ps of a chunk: 0 = lo' x WB, 1 = hi x WC, 2 = hi x WA   (small terms first, see the header)
  if (warp < 4) {
  if constexpr (BIGREG) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  if (warp == 0 && lane == 0) {
    // ================= TMA producer =================
    int stage = 0; uint32_t phase = 0;
    const int cb_mask = (1 << p.lg_cblocks) - 1;
    for (int t = unit0; t < p.num_tiles; t += unit_step) {
      uint32_t mt, nt, q, ow0, img0, oh0;
      p.fd_ntiles.divmod((uint32_t)t, mt, nt);
      const int n0 = (int)nt * BN;
      const uint32_t m0 = unit_m0(mt);
      p.fd_ow.divmod(m0, q, ow0);
      p.fd_oh.divmod(q, img0, oh0);
      const int bw = (int)ow0 * p.stride - p.pad, bh = (int)oh0 * p.stride - p.pad;
      const int bw2 = (int)ow0 * p.stride2, bh2 = (int)oh0 * p.stride2;
      for (int c0 = 0; c0 < p.num_kb; c0 += p.chunk_kb) {
        const int c1 = min(c0 + p.chunk_kb, p.num_kb);
        for (int ps = 0; ps < 3; ++ps) {
          const int a_off = ps == 0 ? p.cin : 0;
          const int b_off = ps == 0 ? p.K : (ps == 1 ? 2 * p.K : 0);
          for (int kb = c0; kb < c1; ++kb) {
            // A box of this k-block: primary tensor (tap, channel block) or, past kb1, the concatenated 1x1 operand
            const bool second = kb >= p.kb1;
            const int tap = kb >> p.lg_cblocks, cb = kb & cb_mask;
            const int kh = p.KW == 1 ? tap : (tap * 11) >> 5;      // tap / 3 for tap < 9
            const int kw = tap - kh * p.KW;
            const CUtensorMap* ma = second ? &tmA2 : &tmA;
            const int ac = second ? (ps == 0 ? p.cin2 : 0) + (kb - p.kb1) * BK : a_off + cb * BK;
            const int aw = second ? bw2 : bw, ah = second ? bh2 : bh;
            const uint16_t ow_off = second ? (uint16_t)0 : (uint16_t)(kw * p.dil), oh_off = second ? (uint16_t)0 : (uint16_t)(kh * p.dil);
            mbar_wait(empty_bar(stage), phase ^ 1u);
            if constexpr (PAIR) {
              if (leader) mbar_expect_tx(full_bar(stage), 2 * S::STAGE_BYTES);   // both CTAs' bytes land on the leader's barrier
              tma2_load_im2col_4d(sA + stage * A_STAGE_BYTES, ma, full_bar(stage), ac, aw, ah, (int)img0, ow_off, oh_off);
              tma2_load_2d(sB + stage * S::B_STAGE_BYTES, &tmB, full_bar(stage), b_off + kb * BK, n0 + (int)rank * (BN / 2));
            } else {
              mbar_expect_tx(full_bar(stage), S::STAGE_BYTES);
              tma_load_im2col_4d(sA + stage * A_STAGE_BYTES, ma, full_bar(stage), ac, aw, ah, (int)img0, ow_off, oh_off);
              tma_load_2d(sB + stage * S::B_STAGE_BYTES, &tmB, full_bar(stage), b_off + kb * BK, n0);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1 && lane == 0 && leader) {
    // ================= MMA issuer (PAIR: the leader CTA issues for both SMs) =================
    constexpr uint32_t idesc = make_idesc_f16(PAIR ? 2 * BM : BM, BN);
    int stage = 0; uint32_t phase = 0;
    int seq = 0;                                   // chunk sequence number of this CTA (across tiles)
    for (int t = unit0; t < p.num_tiles; t += unit_step) {
      for (int c0 = 0; c0 < p.num_kb; c0 += p.chunk_kb, ++seq) {
        const int nkb = min(p.chunk_kb, p.num_kb - c0) * 3;
        const int acc = seq & 1;
        const uint32_t use = (uint32_t)(seq >> 1) & 1u;
        mbar_wait(tempty_bar(acc), use ^ 1u);      // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a = sA + stage * A_STAGE_BYTES, b = sB + stage * S::B_STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            if constexpr (PAIR) umma2_bf16(d, make_smem_desc(a + k * UMMA_K * 2), make_smem_desc(b + k * UMMA_K * 2), idesc, (kb | k) ? 1u : 0u);
            else umma_bf16(d, make_smem_desc(a + k * UMMA_K * 2), make_smem_desc(b + k * UMMA_K * 2), idesc, (kb | k) ? 1u : 0u);
          }
          if constexpr (PAIR) umma2_commit_mc(empty_bar(stage)); else umma_commit(empty_bar(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        if constexpr (PAIR) umma2_commit_mc(tfull_bar(acc)); else umma_commit(tfull_bar(acc));
      }
    }
  }
  } else {
    // ================= epilogue: chunk promotion + tile finish =================
    if constexpr (BIGREG) asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int g = (warp - 4) >> 2;                 // epilogue warpgroup
    const int e = warp & 3;                        // TMEM lane quarter == warp id % 4
    const int row = e * 32 + lane;
    const int et = threadIdx.x - 128 - g * 128;
    const bool issuer = et == 0;
    const uint32_t swz = (uint32_t)(row & 7);
    const uint32_t buf0 = sOut + (uint32_t)(g * 2) * XBUF_BYTES + (uint32_t)row * 128u;   // this thread's row in the group's buffers
    const uint32_t buf1 = buf0 + XBUF_BYTES;
    const uint32_t gbuf0 = sOut + (uint32_t)(g * 2) * XBUF_BYTES, gbuf1 = gbuf0 + XBUF_BYTES;
    float* s_scale = s_scale_all + g * 2 * NS_OWN * XSLAB;
    float* s_shift = s_scale + NS_OWN * XSLAB;
    auto group_barrier = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory"); };
    uint32_t rphase = 0;
    int seq = 0;
    float run[NS_OWN][XSLAB];
    for (int t = unit0; t < p.num_tiles; t += unit_step) {
      uint32_t umt, unt;
      p.fd_ntiles.divmod((uint32_t)t, umt, unt);
      const int n0 = (int)unt * BN, m0 = (int)unit_m0(umt);
      // scale / shift of my columns (read in the tile finish, at least one chunk of MMAs from now)
      for (int i = et; i < NS_OWN * XSLAB; i += 128) {
        const int col = n0 + (g + (i >> 6) * EG) * XSLAB + (i & 63);
        s_scale[i] = p.scale ? __ldg(p.scale + col) : 1.f;
        s_shift[i] = p.shift ? __ldg(p.shift + col) : 0.f;
      }
      for (int ci = 0; ci < nchunks; ++ci, ++seq) {
        const int acc = seq & 1;
        const uint32_t use = (uint32_t)(seq >> 1) & 1u;
        if (HAS_RES && ci == 0 && issuer) {
          // residual planes of my first unit -> my staging buffers (free until the tile finish), a whole tile of MMAs ahead
          if constexpr (POOL) fence_proxy_async_smem();   // the last tile's column-sum scratch lived in these buffers
          else bulk_wait_read<0>();                // the stores that last read the buffers have drained them
          mbar_expect_tx(res_bar(g), 2 * XBUF_BYTES);
          tma_load_2d(gbuf0, &tmR, res_bar(g), n0 + g * XSLAB, m0);
          tma_load_2d(gbuf1, &tmR, res_bar(g), p.res_plane + n0 + g * XSLAB, m0);
        }
        mbar_wait(tfull_bar(acc), use);
        tc_fence_after();
        const uint32_t tacc = tmem_base + ((uint32_t)(e * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll
        for (int j = 0; j < NS_OWN; ++j) {
          uint32_t v0[32], v1[32];
          tmem_ld32(tacc + (uint32_t)((g + j * EG) * XSLAB), v0);
          tmem_ld32(tacc + (uint32_t)((g + j * EG) * XSLAB + 32), v1);
          tmem_ld_wait();
          if (ci == 0) {
#pragma unroll
            for (int i = 0; i < 32; ++i) { run[j][i] = __uint_as_float(v0[i]); run[j][32 + i] = __uint_as_float(v1[i]); }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              run[j][i] = __fadd_rn(run[j][i], __uint_as_float(v0[i]));
              run[j][32 + i] = __fadd_rn(run[j][32 + i], __uint_as_float(v1[i]));
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {                               // hand the accumulator back to the MMA warp (PAIR: the leader's)
          if (PAIR && !leader) mbar_arrive_leader(tempty_bar(acc)); else mbar_arrive(tempty_bar(acc));
        }
      }
      // ---- tile finish: scale/shift (+ residual) (+ ReLU) -> two fp16 planes (or fp32) -> TMA store
      // Residual of the units after the first: the staging buffers are busy with the first unit until its store has
      // read them, so a TMA load into them would start a full memory latency before it is needed (ncu: 21 % of the
      // epilogue's time).  Each thread owns one output row, i.e. 128 contiguous bytes per plane: fetch them straight into
      // registers with 32-byte (sector-sized) loads now, behind the first unit's arithmetic.
      uint32_t rr[HAS_RES && NS_OWN > 1 ? (NS_OWN - 1) * 64 : 1];
      (void)rr;
      if constexpr (HAS_RES && NS_OWN > 1) {
        const bool in = (int64_t)m0 + row < p.M;
#pragma unroll
        for (int j = 1; j < NS_OWN; ++j) {
          const __half* rp = p.res + (int64_t)(m0 + row) * p.ldr + n0 + (g + j * EG) * XSLAB;
#pragma unroll
          for (int pl = 0; pl < 2; ++pl)
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              uint32_t* d = rr + (j - 1) * 64 + pl * 32 + v * 8;
              if (in) ldg256_nc(rp + pl * p.res_plane + v * 16, d);
              else {
#pragma unroll
                for (int i = 0; i < 8; ++i) d[i] = 0u;
              }
            }
        }
      }
#pragma unroll
      for (int j = 0; j < NS_OWN; ++j) {
        const int col0 = n0 + (g + j * EG) * XSLAB;
        if (HAS_RES && j == 0) {
          mbar_wait(res_bar(g), rphase);
          rphase ^= 1u;
        } else if (!POOL && issuer) {
          bulk_wait_read<0>();
        }
        group_barrier();                           // staging reusable; scale/shift visible
        float (&pv)[XSLAB] = run[j];               // POOL: the finished values replace the sums they came from
        if constexpr (OUT_F32) {
#pragma unroll
          for (int hb = 0; hb < 2; ++hb) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {          // 4 fp32 channels = one 16 B chunk
              const int c = j * XSLAB + hb * 32 + q * 4;
              const float4 sc = *reinterpret_cast<const float4*>(s_scale + c);
              const float4 sh = *reinterpret_cast<const float4*>(s_shift + c);
              float f0 = fmaf(run[j][hb * 32 + q * 4 + 0], sc.x, sh.x), f1 = fmaf(run[j][hb * 32 + q * 4 + 1], sc.y, sh.y);
              float f2 = fmaf(run[j][hb * 32 + q * 4 + 2], sc.z, sh.z), f3 = fmaf(run[j][hb * 32 + q * 4 + 3], sc.w, sh.w);
              if (p.relu) { f0 = fmaxf(f0, 0.f); f1 = fmaxf(f1, 0.f); f2 = fmaxf(f2, 0.f); f3 = fmaxf(f3, 0.f); }
              sts128((hb ? buf1 : buf0) + (((uint32_t)q ^ swz) << 4),
                     make_uint4(__float_as_uint(f0), __float_as_uint(f1), __float_as_uint(f2), __float_as_uint(f3)));
            }
          }
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) {            // 8 channels = one 16 B fp16 chunk per plane
            const uint32_t coff = ((uint32_t)q ^ swz) << 4;
            const int c = j * XSLAB + q * 8;
            const float4 sc0 = *reinterpret_cast<const float4*>(s_scale + c), sc1 = *reinterpret_cast<const float4*>(s_scale + c + 4);
            const float4 sh0 = *reinterpret_cast<const float4*>(s_shift + c), sh1 = *reinterpret_cast<const float4*>(s_shift + c + 4);
            float2 y[4];                           // packed fp32x2 arithmetic (FFMA2 / FADD2), same roundings as scalar code
            y[0] = __ffma2_rn(make_float2(run[j][q * 8 + 0], run[j][q * 8 + 1]), make_float2(sc0.x, sc0.y), make_float2(sh0.x, sh0.y));
            y[1] = __ffma2_rn(make_float2(run[j][q * 8 + 2], run[j][q * 8 + 3]), make_float2(sc0.z, sc0.w), make_float2(sh0.z, sh0.w));
            y[2] = __ffma2_rn(make_float2(run[j][q * 8 + 4], run[j][q * 8 + 5]), make_float2(sc1.x, sc1.y), make_float2(sh1.x, sh1.y));
            y[3] = __ffma2_rn(make_float2(run[j][q * 8 + 6], run[j][q * 8 + 7]), make_float2(sc1.z, sc1.w), make_float2(sh1.z, sh1.w));
            if (HAS_RES) {
              uint32_t hw[4], lw[4];
              if (j == 0) {
                const uint4 rh = lds128(buf0 + coff), rl = lds128(buf1 + coff);
                hw[0] = rh.x; hw[1] = rh.y; hw[2] = rh.z; hw[3] = rh.w; lw[0] = rl.x; lw[1] = rl.y; lw[2] = rl.z; lw[3] = rl.w;
              } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) { hw[i] = rr[(j > 0 ? j - 1 : 0) * 64 + q * 4 + i]; lw[i] = rr[(j > 0 ? j - 1 : 0) * 64 + 32 + q * 4 + i]; }
              }
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 fh = __half22float2(*reinterpret_cast<const __half2*>(&hw[i]));
                const float2 fl = __half22float2(*reinterpret_cast<const __half2*>(&lw[i]));
                y[i] = __fadd2_rn(y[i], __ffma2_rn(fl, make_float2(LO_INV, LO_INV), fh));
              }
            }
            if (p.relu) {
#pragma unroll
              for (int i = 0; i < 4; ++i) { y[i].x = fmaxf(y[i].x, 0.f); y[i].y = fmaxf(y[i].y, 0.f); }
            }
            if constexpr (POOL) {
#pragma unroll
              for (int i = 0; i < 4; ++i) { pv[q * 8 + 2 * i] = y[i].x; pv[q * 8 + 2 * i + 1] = y[i].y; }
            } else {
              uint32_t hi[4], lo[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) split_pair(y[i], hi[i], lo[i]);
              sts128(buf0 + coff, make_uint4(hi[0], hi[1], hi[2], hi[3]));
              sts128(buf1 + coff, make_uint4(lo[0], lo[1], lo[2], lo[3]));
            }
          }
        }
        if constexpr (POOL) {
          // column sums over this CTA's rows of the ROI: rows past the ROI contribute nothing
          const int valid = rank ? p.pool_rows - BM : BM;
          float* comb = reinterpret_cast<float*>(gbase + S::OFF_OUT + (size_t)(g * 2) * XBUF_BYTES) + (j & 1) * 256;
          float2 ts = make_float2(0.f, 0.f);
          if (j == 0) group_barrier();             // every thread has read its residual out of the staging buffers
          if (e * 32 < valid) {                    // warp-uniform
            const bool in = row < valid;
#pragma unroll
            for (int i = 0; i < XSLAB; ++i) pv[i] = in ? pv[i] : 0.f;
            warp_colsum64(pv, lane);
            ts = make_float2(pv[0], pv[1]);
          }
          *reinterpret_cast<float2*>(comb + e * 64 + 2 * lane) = ts;
          group_barrier();
          if (et < 64) {
            const float* c4 = comb + et;
            const float tot = ((c4[0] + c4[64]) + c4[128]) + c4[192];
            p.pool_partial[((int64_t)umt * 2 + rank) * p.Cout + col0 + et] = tot;
          }
          if (j == NS_OWN - 1) group_barrier();    // scratch readers done before the next tile's residual lands on it
        } else {
          fence_proxy_async_smem();                // generic-proxy smem writes -> visible to the TMA unit
          group_barrier();
          if (issuer) {
            tma_store_2d(&tmY, gbuf0, col0, m0);
            tma_store_2d(&tmY, gbuf1, col0 + (OUT_F32 ? 32 : p.out_plane), m0);
            bulk_commit();
          }
        }
      }
    }
    if (!POOL && issuer) bulk_wait_all();          // all output bytes are in global memory
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) {
    cluster_sync_all();                            // the peer may still be reading / the leader still issuing into its TMEM
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN) : "memory");
  } else {
    if (warp == 2) tmem_dealloc<2 * BN>(tmem_base);
  }
}

template <int BN, int STAGES, int EG, bool HAS_RES, bool OUT_F32, bool BIGREG, bool PAIR = false, bool POOL = false>
int launchx(const CUtensorMap& ma, const CUtensorMap& ma2, const CUtensorMap& mb, const CUtensorMap& my, const CUtensorMap& mr,
            TcxParams tp, int cout_pad, cudaStream_t st) {
  using S = SmemX<BN, STAGES, EG, PAIR>;
  auto kern = conv_tcx_kernel<BN, STAGES, EG, HAS_RES, OUT_F32, BIGREG, PAIR, POOL>;
  static DeviceOnce once;
  if (once.first()) VLTK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
  tp.n_tiles = cout_pad / BN;
  tp.fd_ntiles.init((uint32_t)tp.n_tiles);
  const int64_t row_tiles = ceil_div64(tp.M, BM);
  const int64_t tiles = (POOL ? tp.M / tp.pool_rows : (PAIR ? ceil_div64(row_tiles, 2) : row_tiles)) * tp.n_tiles;   // PAIR: 256-row pair tiles; POOL: one ROI each
  VLTK_CHECK(tiles < (1ll << 31), "conv_tcx: too many tiles");
  tp.num_tiles = (int)tiles;
  // persistent: one CTA (or one CTA pair) per SM (pair of SMs)
  const int grid = PAIR ? 2 * (int)std::min<int64_t>(tiles, tc_num_sms() / 2) : (int)std::min<int64_t>(tiles, tc_num_sms());
  static const bool use_pdl = [] { const char* e = getenv("VLTK_PDL"); return !(e && e[0] == '0'); }();
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128 + 128 * EG); cfg.dynamicSmemBytes = S::TOTAL; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (PAIR) { attr[na].id = cudaLaunchAttributeClusterDimension; attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1; ++na; }
  if (use_pdl) { attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[na].val.programmaticStreamSerializationAllowed = 1; ++na; }
  cfg.attrs = attr; cfg.numAttrs = na;
  VLTK_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, ma2, mb, my, mr, tp));
  VLTK_LAUNCH_CHECK();
  return 0;
}

std::atomic<int> g_tcx_cta2_min_m{[] { const char* e = getenv("VLTK_TCX_CTA2"); return e ? atoi(e) : 32768; }()};

}  // namespace

void conv_tcx_set_cta_pairs(int min_pixels) { if (min_pixels >= 0) g_tcx_cta2_min_m.store(min_pixels); }

int conv_tcx_launch(const ConvProblem& p, const void* w3, int cout_pad, TensorMapCache* cache, cudaStream_t st,
                    const TcConcat* cc, const TcPool* pool) {
  const bool pooled = pool && pool->out;
  const bool out_f32 = p.out_dtype == DT_F32;
  VLTK_CHECK(p.in_dtype == DT_H2 && (out_f32 || p.out_dtype == DT_H2), "conv_tcx: split-fp16 activations in, split-fp16 or fp32 out");
  VLTK_CHECK(p.Cin % BK == 0 && ((p.Cin / BK) & (p.Cin / BK - 1)) == 0, "conv_tcx: Cin=%d must be 64 * 2^k", p.Cin);
  VLTK_CHECK(p.ldx == 2 * p.Cin, "conv_tcx: x must hold the two planes of a pixel back to back (ldx = 2 Cin)");
  VLTK_CHECK((p.KH == 1 && p.KW == 1) || (p.KH == 3 && p.KW == 3), "conv_tcx: 1x1 and 3x3 kernels only");
  VLTK_CHECK(!(out_f32 && p.residual), "conv_tcx: fp32 output has no residual path");
  VLTK_CHECK(cout_pad % 64 == 0 && p.Cout == cout_pad, "conv_tcx: Cout=%d must equal cout_pad=%d (a multiple of 64)", p.Cout, cout_pad);
  VLTK_CHECK(out_f32 ? (p.ldy % 4 == 0 && p.ldy >= p.Cout) : p.ldy == 2 * p.Cout, "conv_tcx: bad output row stride");
  VLTK_CHECK(!p.residual || p.ldr == 2 * p.Cout, "conv_tcx: the residual must be a split-fp16 tensor of the output shape");
  const int64_t M = (int64_t)p.N * p.OH * p.OW;
  if (M == 0) return 0;
  VLTK_CHECK(M < (1ll << 31) - 512, "conv_tcx: M=%lld output pixels exceed the 32-bit tile arithmetic", (long long)M);
  if (cc) {
    VLTK_CHECK(p.KH == 1 && p.KW == 1 && p.pad == 0, "conv_tcx: K-concatenation needs a 1x1 primary convolution");
    VLTK_CHECK(cc->x2 && cc->Cin2 % BK == 0 && cc->ldx2 == 2 * cc->Cin2 && cc->stride2 >= 1, "conv_tcx: bad concatenated operand");
    VLTK_CHECK((cc->H2 - 1) / cc->stride2 + 1 == p.OH && (cc->W2 - 1) / cc->stride2 + 1 == p.OW, "conv_tcx: concatenated operand does not map onto the output grid");
  }
  const int K = p.KH * p.KW * p.Cin + (cc ? cc->Cin2 : 0);     // row length of one weight plane
  static const int smallk = [] { const char* e = getenv("VLTK_TCX_SMALLK"); return e ? atoi(e) : 256; }();
  static const int chunk = [] { const char* e = getenv("VLTK_TCX_CHUNK"); return e ? std::max(1, atoi(e)) : 4; }();
  int bn = (cout_pad % 256 == 0) ? 256 : (cout_pad % 128 == 0 ? 128 : 64);
  if (K <= smallk && cout_pad % 128 == 0) bn = 128;
  if (out_f32) { VLTK_CHECK(cout_pad % 128 == 0, "conv_tcx: fp32 output needs cout_pad %% 128 == 0"); bn = 128; }
  static const int res_bn = [] { const char* e = getenv("VLTK_TCX_RES_BN"); return e ? atoi(e) : 256; }();   // tuning knob
  if (p.residual && bn == 256 && res_bn == 128) bn = 128;
  if (p.residual && bn == 64) { VLTK_CHECK(false, "conv_tcx: residual layers need Cout %% 128 == 0"); }
  // CTA pairs for the wide layers with enough rows to fill 74 pairs (all of res5); VLTK_TCX_CTA2=0 switches them off
  const bool pair = bn == 256 && !out_f32 && (pooled || (g_tcx_cta2_min_m.load() > 0 && M >= g_tcx_cta2_min_m.load()));
  if (pooled) {
    VLTK_CHECK(bn == 256 && !out_f32 && p.residual && !cc, "conv_tcx: the pooled finish needs a residual layer with Cout %% 256 == 0");
    VLTK_CHECK(pool->partial && pool->rows > BM && pool->rows <= 2 * BM && M % pool->rows == 0, "conv_tcx: pooled rows=%d must be in (128, 256] and divide M", pool->rows);
  }
  if (cache->maps.size() > 8192) cache->maps.clear();
  auto cached = [&](const TensorMapCache::Key& k, CUtensorMap* dst, auto make) -> int {
    auto it = cache->maps.find(k);
    if (it == cache->maps.end()) {
      if (make(dst)) return -1;
      cache->maps[k] = *dst;
    } else *dst = it->second;
    return 0;
  };
  CUtensorMap ma, ma2, mb, my, mr;
  if (cached(TensorMapCache::Key(p.x, p.N, p.H, p.W, p.Cin, p.ldx, p.KH, p.stride, p.pad, p.dil, 10), &ma, [&](CUtensorMap* d) {
        return tc_encode_im2col(d, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, p.x, p.N, p.H, p.W, 2 * p.Cin, p.ldx, 2, p.KH, p.KW, p.stride, p.pad, p.dil);
      })) return -1;
  ma2 = ma;
  if (cc &&
      cached(TensorMapCache::Key(cc->x2, p.N, cc->H2, cc->W2, cc->Cin2, cc->ldx2, 1, cc->stride2, 0, 1, 10), &ma2, [&](CUtensorMap* d) {
        return tc_encode_im2col(d, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, cc->x2, p.N, cc->H2, cc->W2, 2 * cc->Cin2, cc->ldx2, 2, 1, 1, cc->stride2, 0, 1);
      })) return -1;
  const int b_rows = pair ? bn / 2 : bn;           // a CTA of a pair loads half of the W tile
  if (cached(TensorMapCache::Key(w3, K, cout_pad, b_rows, 0, 0, 0, 0, 0, 0, 11), &mb, [&](CUtensorMap* d) {
        return tc_encode_tiled(d, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, w3, (uint64_t)3 * K, (uint64_t)cout_pad, (uint64_t)3 * K * 2, BK, (uint32_t)b_rows, true);
      })) return -1;
  if (pooled) my = ma;                             // no output tensor: the tile is reduced, not stored
  else if (cached(TensorMapCache::Key(p.y, (int)M, p.Cout, p.ldy, out_f32 ? 1 : 0, 0, 0, 0, 0, 0, 12), &my, [&](CUtensorMap* d) {
        return out_f32 ? tc_encode_tiled(d, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, p.y, (uint64_t)p.Cout, (uint64_t)M, (uint64_t)p.ldy * 4, 32, BM, false)
                       : tc_encode_tiled(d, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, p.y, (uint64_t)2 * p.Cout, (uint64_t)M, (uint64_t)p.ldy * 2, XSLAB, BM, false);
      })) return -1;
  mr = ma;
  if (p.residual &&
      cached(TensorMapCache::Key(p.residual, (int)M, p.Cout, p.ldr, 0, 0, 0, 0, 0, 0, 13), &mr, [&](CUtensorMap* d) {
        return tc_encode_tiled(d, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, p.residual, (uint64_t)2 * p.Cout, (uint64_t)M, (uint64_t)p.ldr * 2, XSLAB, BM, false);
      })) return -1;
  TcxParams t;
  memset(&t, 0, sizeof(t));
  t.scale = p.scale; t.shift = p.shift; t.M = M; t.Cout = p.Cout; t.relu = p.relu;
  t.stride = p.stride; t.pad = p.pad; t.dil = p.dil; t.KW = p.KW;
  const int cblocks = p.Cin / BK;
  t.lg_cblocks = 0;
  while ((1 << t.lg_cblocks) < cblocks) ++t.lg_cblocks;
  t.kb1 = p.KH * p.KW * cblocks;
  t.num_kb = t.kb1 + (cc ? cc->Cin2 / BK : 0); t.cin = p.Cin; t.K = K; t.chunk_kb = chunk;
  t.cin2 = cc ? cc->Cin2 : 0; t.stride2 = cc ? cc->stride2 : 1;
  t.out_plane = p.Cout; t.res_plane = p.Cout;
  t.res = (const __half*)p.residual; t.ldr = p.ldr;
  t.fd_ow.init((uint32_t)p.OW); t.fd_oh.init((uint32_t)p.OH);
  if (out_f32) return launchx<128, 5, 2, false, true, false>(ma, ma2, mb, my, mr, t, cout_pad, st);
  t.pool_partial = pooled ? pool->partial : nullptr; t.pool_rows = pooled ? pool->rows : 0;
  if (pooled) {
    if (launchx<256, 5, 2, true, false, true, true, true>(ma, ma2, mb, my, mr, t, cout_pad, st)) return -1;
    return tc_pool_finish(pool->partial, pool->out, (int)(M / pool->rows), pool->rows, p.Cout, st);
  }
  if (pair) {
    VLTK_CHECK(M + 256 < (1ll << 31), "conv_tcx: M too large for pair tiles");
    static const int probe_stages = [] { const char* e = getenv("VLTK_TCX_PROBE_STAGES"); return e ? atoi(e) : 5; }();   // diagnosis knob
    if (!p.residual && probe_stages == 4) return launchx<256, 4, 2, false, false, true, true>(ma, ma2, mb, my, mr, t, cout_pad, st);
    if (!p.residual && probe_stages == 3) return launchx<256, 3, 2, false, false, true, true>(ma, ma2, mb, my, mr, t, cout_pad, st);
    return p.residual ? launchx<256, 5, 2, true, false, true, true>(ma, ma2, mb, my, mr, t, cout_pad, st)
                      : launchx<256, 5, 2, false, false, true, true>(ma, ma2, mb, my, mr, t, cout_pad, st);
  }
  if (p.residual) {
    if (bn == 256) return launchx<256, 3, 2, true, false, true>(ma, ma2, mb, my, mr, t, cout_pad, st);
    return launchx<128, 5, 2, true, false, false>(ma, ma2, mb, my, mr, t, cout_pad, st);
  }
  if (bn == 256) return launchx<256, 3, 2, false, false, true>(ma, ma2, mb, my, mr, t, cout_pad, st);
  if (bn == 128) return launchx<128, 5, 2, false, false, false>(ma, ma2, mb, my, mr, t, cout_pad, st);
  return launchx<64, 6, 1, false, false, false>(ma, ma2, mb, my, mr, t, cout_pad, st);
}

}  // namespace vltk
