// Device half of the JPEG front end: dequantisation + inverse DCT + chroma upsampling + colour
// conversion of the coefficients produced by jpeg_host.cpp, writing the BGR u8 [h, w, 3] image that
// `cv2.imread` returns at the reference's call site (vltk/compat.py:573-579) — bit for bit: the integer
// algorithms below are libjpeg(-turbo)'s defaults (JDCT_ISLOW, fancy upsampling, 16-bit fixed-point YCbCr->RGB),
// restated from their published description (IJG libjpeg jidctint.c / jdsample.c / jdcolor.c; the SIMD code of
// libjpeg-turbo is specified to produce identical output).
#include "../../include/vltk_frcnn.h"
#include "common.cuh"
#include "jpeg_dev.h"

namespace vltk {
namespace {

struct JpegGeom {
  int width, height, ncomp;
  int blocks_w[3], blocks_h[3], comp_w[3], comp_h[3];
  long long coef_offset[3], plane_offset[3];
  int h2, v2;            // luma is sampled 2x horizontally / vertically relative to chroma
  int fancy;             // triangle-filter upsampling (downsampled_width > 2), else sample replication
  int ycc;               // 1: YCbCr -> RGB, 0: components are already R, G, B
  unsigned short qt[3][64];
};

// jidctint.c constants: FIX(x) = round(x * 2^13)
constexpr int CONST_BITS = 13, PASS1_BITS = 2;
constexpr int F_0_298631336 = 2446, F_0_390180644 = 3196, F_0_541196100 = 4433, F_0_765366865 = 6270,
              F_0_899976223 = 7373, F_1_175875602 = 9633, F_1_501321110 = 12299, F_1_847759065 = 15137,
              F_1_961570560 = 16069, F_2_053119869 = 16819, F_2_562915447 = 20995, F_3_072711026 = 25172;

__device__ __forceinline__ int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

// One 1-D pass of the Loeffler-Ligtenberg-Moschytz butterfly on 8 values; `shift` is the final descale.
__device__ __forceinline__ void idct8(const int (&in)[8], int (&out)[8], int shift) {
  int z2 = in[2], z3 = in[6];
  int z1 = (z2 + z3) * F_0_541196100;
  int tmp2 = z1 + z3 * (-F_1_847759065);
  int tmp3 = z1 + z2 * F_0_765366865;
  z2 = in[0]; z3 = in[4];
  int tmp0 = (z2 + z3) << CONST_BITS;
  int tmp1 = (z2 - z3) << CONST_BITS;
  const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
  tmp0 = in[7]; tmp1 = in[5]; tmp2 = in[3]; tmp3 = in[1];
  z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
  int z4 = tmp1 + tmp3;
  const int z5 = (z3 + z4) * F_1_175875602;
  tmp0 *= F_0_298631336; tmp1 *= F_2_053119869; tmp2 *= F_3_072711026; tmp3 *= F_1_501321110;
  z1 *= -F_0_899976223; z2 *= -F_2_562915447; z3 *= -F_1_961570560; z4 *= -F_0_390180644;
  z3 += z5; z4 += z5;
  tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
  out[0] = descale(tmp10 + tmp3, shift); out[7] = descale(tmp10 - tmp3, shift);
  out[1] = descale(tmp11 + tmp2, shift); out[6] = descale(tmp11 - tmp2, shift);
  out[2] = descale(tmp12 + tmp1, shift); out[5] = descale(tmp12 - tmp1, shift);
  out[3] = descale(tmp13 + tmp0, shift); out[4] = descale(tmp13 - tmp0, shift);
}

// One thread per 8x8 block: dequantise, columns pass (kept at PASS1_BITS extra precision), rows pass, +128, clamp.
__global__ void __launch_bounds__(128)
jpeg_idct_kernel(const short* __restrict__ coef, unsigned char* __restrict__ planes, JpegGeom g, int nb0, int nb1, int nb2) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  int c = 0;
  if (b >= nb0) { b -= nb0; c = 1; if (b >= nb1) { b -= nb1; c = 2; if (b >= nb2) return; } }
  const short* __restrict__ src = coef + g.coef_offset[c] + (long long)b * 64;
  int ws[8][8];
  {
    short raw[64];
    const int4* s4 = reinterpret_cast<const int4*>(src);
#pragma unroll
    for (int i = 0; i < 8; ++i) reinterpret_cast<int4*>(raw)[i] = s4[i];
#pragma unroll
    for (int col = 0; col < 8; ++col) {
      int in[8], out[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) in[r] = (int)raw[r * 8 + col] * (int)g.qt[c][r * 8 + col];
      idct8(in, out, CONST_BITS - PASS1_BITS);
#pragma unroll
      for (int r = 0; r < 8; ++r) ws[r][col] = out[r];
    }
  }
  const int bw = g.blocks_w[c];
  const int by = b / bw, bx = b - by * bw;
  unsigned char* dst = planes + g.plane_offset[c] + ((long long)by * 8) * (bw * 8) + bx * 8;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    int out[8];
    idct8(ws[r], out, CONST_BITS + PASS1_BITS + 3);
    unsigned int lo = 0, hi = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      lo |= (unsigned int)min(max(out[k] + 128, 0), 255) << (8 * k);
      hi |= (unsigned int)min(max(out[4 + k] + 128, 0), 255) << (8 * k);
    }
    *reinterpret_cast<uint2*>(dst + (long long)r * (bw * 8)) = make_uint2(lo, hi);
  }
}

// chroma sample at full resolution (jdsample.c): h2v2 / h2v1 triangle filters with replicated edges, or replication
__device__ __forceinline__ int chroma_at(const unsigned char* __restrict__ p, int stride, int cw, int ch, int x, int y,
                                         const JpegGeom& g) {
  if (!g.h2) return p[(long long)y * stride + x];
  const int cx = x >> 1;
  if (!g.fancy) return p[(long long)(g.v2 ? (y >> 1) : y) * stride + cx];
  const int nb = min(max((x & 1) ? cx + 1 : cx - 1, 0), cw - 1);
  if (g.v2) {
    const int cy = y >> 1;
    const int far = min(max((y & 1) ? cy + 1 : cy - 1, 0), ch - 1);
    const unsigned char* r0 = p + (long long)cy * stride;
    const unsigned char* r1 = p + (long long)far * stride;
    const int cs = r0[cx] * 3 + r1[cx], ns = r0[nb] * 3 + r1[nb];
    return (cs * 3 + ns + ((x & 1) ? 7 : 8)) >> 4;
  }
  const unsigned char* r0 = p + (long long)y * stride;
  return (r0[cx] * 3 + r0[nb] + ((x & 1) ? 2 : 1)) >> 2;
}

__global__ void __launch_bounds__(256)
jpeg_color_kernel(const unsigned char* __restrict__ planes, unsigned char* __restrict__ bgr, JpegGeom g) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= g.width) return;
  const unsigned char* py = planes + g.plane_offset[0];
  const int Y = py[(long long)y * (g.blocks_w[0] * 8) + x];
  int R = Y, G = Y, B = Y;
  if (g.ncomp == 3) {
    const int c1 = chroma_at(planes + g.plane_offset[1], g.blocks_w[1] * 8, g.comp_w[1], g.comp_h[1], x, y, g);
    const int c2 = chroma_at(planes + g.plane_offset[2], g.blocks_w[2] * 8, g.comp_w[2], g.comp_h[2], x, y, g);
    if (g.ycc) {                          // jdcolor.c build_ycc_rgb_table: SCALEBITS 16, ONE_HALF 32768
      const int cb = c1 - 128, cr = c2 - 128;
      R = Y + ((91881 * cr + 32768) >> 16);
      B = Y + ((116130 * cb + 32768) >> 16);
      G = Y + ((-22554 * cb - 46802 * cr + 32768) >> 16);
      R = min(max(R, 0), 255); G = min(max(G, 0), 255); B = min(max(B, 0), 255);
    } else { R = Y; G = c1; B = c2; }
  }
  unsigned char* o = bgr + ((long long)y * g.width + x) * 3;
  o[0] = (unsigned char)B; o[1] = (unsigned char)G; o[2] = (unsigned char)R;
}


// =========================================================================================
// GPU Huffman entropy decoder.  A JPEG scan is ONE serial bit stream, but Huffman decoders self-synchronise:
// a decoder started at an arbitrary bit falls into step with the true symbol sequence after a few blocks
// (Klein & Wiseman 2003; Weissenberger & Schmidt 2018 for JPEG on GPUs).  One CTA per image:
//   A. the stream is cut into subsequences of S bits; thread t decodes subsequence t from a GUESSED state
//      (bit t*S, start of a block, first block of an MCU) and records the state in which it crosses into t+1;
//   B. Jacobi iterations: every thread restarts from its predecessor's recorded end state and re-decodes its own
//      subsequence; when no recorded state changes any more, every start state is the true one (subsequence 0
//      starts from the true state, so truth propagates at least one subsequence per iteration and, thanks to
//      self-synchronisation, usually everywhere at once);
//   C. a prefix sum over the blocks completed per subsequence gives every thread its first output block, and a
//      final pass writes the coefficients (DC still as differences);
// then jpeg_dc_scan_kernel turns the DC differences into values with a per-component prefix sum in scan order.
// The state machine (including what it does on invalid codes) is identical in all passes, which is what makes
// the fixed point meaningful.  Output == the host decoder's coefficients, bit for bit (tests/test_jpeg.py).
struct HState { unsigned int p; unsigned short kb; };   // bit position; k | (block-in-MCU << 8)

__device__ __constant__ unsigned char c_zigzag[64] = {
    0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
    41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
    30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

__device__ __forceinline__ unsigned int bits_at(const unsigned int* __restrict__ w, unsigned int p) {
  const unsigned int i = p >> 5, sh = p & 31;
  return __funnelshift_l(w[i + 1], w[i], sh);
}

__device__ __forceinline__ void huff_lookup(const DevHuff& T, unsigned int win, int& len, int& sym) {
  const unsigned int e = T.look[win >> 23];
  if (e) { len = e >> 8; sym = e & 255; return; }
  int l = 10;
  int code = (int)(win >> 22);
  while (l <= 16 && code > T.maxcode[l]) { ++l; code = (int)(win >> (32 - l)); }
  if (l > 16) { len = 16; sym = 0; return; }            // invalid code: a fixed, deterministic transition
  len = l;
  sym = T.vals[(code + T.valoffset[l]) & 255];
}

__device__ __forceinline__ int extend_bits(unsigned int win, int len, int s) {
  const int v = (int)((win << len) >> (32 - s));
  return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v;
}

// Decodes symbols that START before end_p.  WRITE: coefficients go to their blocks, starting at scan-order block g.
// DCPRED (restart-interval streams): the thread owns whole intervals, so it integrates the DC differences itself
// (prediction starts at 0) and stops after `max_blocks` blocks.
template <bool WRITE, bool DCPRED = false>
__device__ __forceinline__ HState decode_range(const unsigned int* __restrict__ words, HState st, unsigned int end_p,
                                               const DevHuff* __restrict__ T, const DevImage& D, int& nblk, int g,
                                               short* __restrict__ coef, int max_blocks = 0x7fffffff) {
  unsigned int p = st.p;
  int k = st.kb & 255, b = st.kb >> 8;
  nblk = 0;
  int pred[3] = {0, 0, 0};
  (void)pred;
  short* out = nullptr;
  auto locate = [&](int gg) -> short* {
    const int mcu = gg / D.B, j = gg - mcu * D.B;
    const int c = D.comp_of_block[j];
    const int my = mcu / D.mcus_x, mx = mcu - my * D.mcus_x;
    const long long blk = (long long)(my * D.vs[c] + D.by_of_block[j]) * D.blocks_w[c] + (mx * D.hs[c] + D.bx_of_block[j]);
    return coef + D.coef_off + D.comp_coef_off[c] + blk * 64;
  };
  if (WRITE) { if (g >= D.total_blocks) return st; out = locate(g); }
  while (p < end_p) {
    const unsigned int win = bits_at(words, p);
    const int c = D.comp_of_block[b];
    if (k == 0) {
      int len, sym;
      huff_lookup(T[2 * c], win, len, sym);
      const int s = sym & 15;
      if (DCPRED) {
        if (s) pred[c] += extend_bits(win, len, s);
        out[0] = (short)pred[c];
      } else if (WRITE && s) out[0] = (short)extend_bits(win, len, s);
      p += len + s;
      k = 1;
    } else {
      const DevHuff& A = T[2 * c + 1];
      const int fa = A.fast_ac[win >> 23];
      if (fa) {
        k += (fa >> 4) & 15;
        if (WRITE && k < 64) out[c_zigzag[k]] = (short)(fa >> 8);
        ++k;
        p += fa & 15;
      } else {
        int len, sym;
        huff_lookup(A, win, len, sym);
        const int r = sym >> 4, s = sym & 15;
        if (s == 0) {
          k = (r == 15) ? k + 16 : 64;                   // ZRL | EOB
          p += len;
        } else {
          k += r;
          if (WRITE && k < 64) out[c_zigzag[k]] = (short)extend_bits(win, len, s);
          ++k;
          p += len + s;
        }
      }
    }
    if (k >= 64) {                                       // block complete (or overrun by garbage: same rule)
      k = 0;
      b = (b + 1 == D.B) ? 0 : b + 1;
      ++nblk;
      if (WRITE) {
        ++g;
        if (g >= D.total_blocks || nblk >= max_blocks) break;   // the rest of the stream / interval is padding
        out = locate(g);
      }
    }
  }
  HState e;
  e.p = p; e.kb = (unsigned short)(k | (b << 8));
  return e;
}

constexpr int HUFF_THREADS = 1024;

__global__ void __launch_bounds__(HUFF_THREADS, 1)
jpeg_huffman_kernel(const unsigned char* __restrict__ blob, short* __restrict__ coef, int* __restrict__ iters_out) {
  extern __shared__ __align__(16) unsigned char hsm[];
  DevHuff* T = reinterpret_cast<DevHuff*>(hsm);
  unsigned int* Ep = reinterpret_cast<unsigned int*>(hsm + 6 * sizeof(DevHuff));          // [2][MAX]
  unsigned short* Ekb = reinterpret_cast<unsigned short*>(Ep + 2 * JPEG_MAX_SUBSEQ);     // [2][MAX]
  int* nb = reinterpret_cast<int*>(Ekb + 2 * JPEG_MAX_SUBSEQ);                           // [MAX] blocks per subsequence
  __shared__ DevImage D;
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  const int tid = threadIdx.x;
  const DevImage* imgs = reinterpret_cast<const DevImage*>(blob);
  if (tid < (int)(sizeof(DevImage) / 4)) reinterpret_cast<int*>(&D)[tid] = reinterpret_cast<const int*>(&imgs[blockIdx.x])[tid];
  __syncthreads();
  if (D.restart_interval > 0) {                          // exact start states: one pass, no synchronisation
    {
      const uint2* src = reinterpret_cast<const uint2*>(blob + D.tables_off);
      uint2* dst = reinterpret_cast<uint2*>(T);
      for (int i = tid; i < (int)(6 * sizeof(DevHuff) / 8); i += HUFF_THREADS) dst[i] = src[i];
    }
    __syncthreads();
    const unsigned int* words = reinterpret_cast<const unsigned int*>(blob + D.words_off);
    const unsigned int* starts = reinterpret_cast<const unsigned int*>(blob + D.starts_off);
    const int per = D.restart_interval * D.B;
    for (int t = tid; t < D.n_intervals; t += HUFF_THREADS) {
      HState st; st.p = starts[t]; st.kb = 0;
      int n;
      decode_range<true, true>(words, st, (unsigned int)D.total_bits, T, D, n, t * per, coef, per);
    }
    if (tid == 0 && iters_out) iters_out[blockIdx.x] = 0;
    return;
  }
  if (D.nsub == 0) return;                               // empty scan
  {
    const uint2* src = reinterpret_cast<const uint2*>(blob + D.tables_off);
    uint2* dst = reinterpret_cast<uint2*>(T);
    for (int i = tid; i < (int)(6 * sizeof(DevHuff) / 8); i += HUFF_THREADS) dst[i] = src[i];
  }
  __syncthreads();
  const unsigned int* words = reinterpret_cast<const unsigned int*>(blob + D.words_off);
  const unsigned int total = (unsigned int)D.total_bits, S = (unsigned int)D.S;
  const int nsub = D.nsub;
  auto end_of = [&](int t) { const unsigned long long e = (unsigned long long)(t + 1) * S; return e < total ? (unsigned int)e : total; };

  // ---- A: guessed starts
  for (int t = tid; t < nsub; t += HUFF_THREADS) {
    HState st; st.p = (unsigned int)t * S; st.kb = 0;
    int n;
    const HState e = decode_range<false>(words, st, end_of(t), T, D, n, 0, nullptr);
    Ep[t] = e.p; Ekb[t] = e.kb; nb[t] = n;
  }
  __syncthreads();
  // ---- B: iterate to the fixed point (double-buffered states: read `cur`, write `cur ^ 1`)
  int cur = 0, iters = 0;
  for (;; ++iters) {
    int changed = 0;
    for (int t = tid; t < nsub; t += HUFF_THREADS) {
      HState e;
      int n = nb[t];
      if (t == 0) { e.p = Ep[cur * JPEG_MAX_SUBSEQ]; e.kb = Ekb[cur * JPEG_MAX_SUBSEQ]; }
      else {
        HState st; st.p = Ep[cur * JPEG_MAX_SUBSEQ + t - 1]; st.kb = Ekb[cur * JPEG_MAX_SUBSEQ + t - 1];
        e = decode_range<false>(words, st, end_of(t), T, D, n, 0, nullptr);
        if (e.p != Ep[cur * JPEG_MAX_SUBSEQ + t] || e.kb != Ekb[cur * JPEG_MAX_SUBSEQ + t] || n != nb[t]) changed = 1;
      }
      Ep[(cur ^ 1) * JPEG_MAX_SUBSEQ + t] = e.p; Ekb[(cur ^ 1) * JPEG_MAX_SUBSEQ + t] = e.kb; nb[t] = n;
    }
    cur ^= 1;
    if (!__syncthreads_or(changed) || iters > nsub) break;
  }
  if (tid == 0 && iters_out) iters_out[blockIdx.x] = iters;
  // ---- exclusive prefix sum of nb[] (chunks of 1024 with a running carry)
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < nsub; base += HUFF_THREADS) {
    const int t = base + tid;
    const int v = t < nsub ? nb[t] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if ((tid & 31) >= o) x += y; }
    if ((tid & 31) == 31) s_warp[tid >> 5] = x;
    __syncthreads();
    if (tid < 32) {
      int w = s_warp[tid];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, w, o); if (tid >= o) w += y; }
      s_warp[tid] = w;
    }
    __syncthreads();
    const int incl = x + ((tid >> 5) ? s_warp[(tid >> 5) - 1] : 0) + s_carry;
    if (t < nsub) nb[t] = incl - v;                      // exclusive
    __syncthreads();
    if (tid == HUFF_THREADS - 1) s_carry = incl;
    __syncthreads();
  }
  // ---- C: write
  for (int t = tid; t < nsub; t += HUFF_THREADS) {
    HState st;
    if (t == 0) { st.p = 0; st.kb = 0; }
    else { st.p = Ep[cur * JPEG_MAX_SUBSEQ + t - 1]; st.kb = Ekb[cur * JPEG_MAX_SUBSEQ + t - 1]; }
    int n;
    decode_range<true>(words, st, end_of(t), T, D, n, nb[t], coef);
  }
}

// DC differences -> DC values: inclusive prefix sum over each component's blocks in scan order (one CTA each).
__global__ void __launch_bounds__(1024)
jpeg_dc_scan_kernel(const unsigned char* __restrict__ blob, short* __restrict__ coef) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  const DevImage& D = reinterpret_cast<const DevImage*>(blob)[blockIdx.x];
  const int c = blockIdx.y, tid = threadIdx.x;
  if (D.nsub == 0 || c >= D.ncomp || D.restart_interval > 0) return;   // interval streams carry final DC values already
  const int per_mcu = D.hs[c] * D.vs[c];
  const int n = (D.total_blocks / D.B) * per_mcu;
  short* base = coef + D.coef_off + D.comp_coef_off[c];
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int b0 = 0; b0 < n; b0 += 1024) {
    const int i = b0 + tid;
    long long pos = 0;
    int v = 0;
    if (i < n) {
      const int mcu = i / per_mcu, j = i - mcu * per_mcu;
      const int by = j / D.hs[c], bx = j - by * D.hs[c];
      const int my = mcu / D.mcus_x, mx = mcu - my * D.mcus_x;
      pos = ((long long)(my * D.vs[c] + by) * D.blocks_w[c] + (mx * D.hs[c] + bx)) * 64;
      v = base[pos];
    }
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if ((tid & 31) >= o) x += y; }
    if ((tid & 31) == 31) s_warp[tid >> 5] = x;
    __syncthreads();
    if (tid < 32) {
      int w = s_warp[tid];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, w, o); if (tid >= o) w += y; }
      s_warp[tid] = w;
    }
    __syncthreads();
    const int incl = x + ((tid >> 5) ? s_warp[(tid >> 5) - 1] : 0) + s_carry;
    if (i < n) base[pos] = (short)incl;
    __syncthreads();
    if (tid == 1023) s_carry = incl;
    __syncthreads();
  }
}

}  // namespace
}  // namespace vltk

using namespace vltk;

extern "C" {

int vltk_jpeg_reconstruct(const int16_t* coef, const vltk_jpeg_info* I, uint8_t* planes, uint8_t* bgr, void* stream) {
  VLTK_CHECK(coef && I && planes && bgr, "jpeg_reconstruct: null argument");
  VLTK_CHECK((I->ncomp == 1 || I->ncomp == 3) && I->width >= 1 && I->height >= 1, "jpeg_reconstruct: bad geometry");
  VLTK_CHECK(((uintptr_t)coef % 16 == 0) && ((uintptr_t)planes % 8 == 0), "jpeg_reconstruct: coef must be 16-byte and planes 8-byte aligned");
  JpegGeom g;
  memset(&g, 0, sizeof(g));
  g.width = I->width; g.height = I->height; g.ncomp = I->ncomp;
  for (int c = 0; c < I->ncomp; ++c) {
    g.blocks_w[c] = I->blocks_w[c]; g.blocks_h[c] = I->blocks_h[c]; g.comp_w[c] = I->comp_w[c]; g.comp_h[c] = I->comp_h[c];
    g.coef_offset[c] = I->coef_offset[c]; g.plane_offset[c] = I->plane_offset[c];
    for (int i = 0; i < 64; ++i) g.qt[c][i] = I->qt[c][i];
  }
  g.h2 = I->ncomp == 3 && I->hs[0] == 2; g.v2 = I->ncomp == 3 && I->vs[0] == 2;
  g.fancy = g.h2 && I->comp_w[1] > 2;
  g.ycc = !(I->color_transform == 0 || (I->comp_id[0] == 'R' && I->comp_id[1] == 'G' && I->comp_id[2] == 'B'));
  int nb[3] = {0, 0, 0};
  for (int c = 0; c < I->ncomp; ++c) nb[c] = I->blocks_w[c] * I->blocks_h[c];
  const int total = nb[0] + nb[1] + nb[2];
  cudaStream_t st = (cudaStream_t)stream;
  jpeg_idct_kernel<<<ceil_div(total, 128), 128, 0, st>>>(coef, planes, g, nb[0], nb[1], nb[2]);
  VLTK_LAUNCH_CHECK();
  dim3 grid(ceil_div(I->width, 256), I->height);
  jpeg_color_kernel<<<grid, 256, 0, st>>>(planes, bgr, g);
  VLTK_LAUNCH_CHECK();
  return 0;
}

int vltk_jpeg_gpu_entropy_decode(int n, const uint8_t* blob, int16_t* coef, int64_t coef_total, int32_t* iterations,
                                 void* stream) {
  VLTK_CHECK(blob && coef && n >= 0, "jpeg_gpu_entropy_decode: null argument");
  VLTK_CHECK(((uintptr_t)blob % 8 == 0) && ((uintptr_t)coef % 16 == 0), "jpeg_gpu_entropy_decode: blob must be 8-byte and coef 16-byte aligned");
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = 6 * sizeof(DevHuff) + (size_t)JPEG_MAX_SUBSEQ * (2 * 4 + 2 * 2 + 4);
  static DeviceOnce once;
  if (once.first()) VLTK_CUDA(cudaFuncSetAttribute(jpeg_huffman_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  VLTK_CUDA(cudaMemsetAsync(coef, 0, (size_t)coef_total * sizeof(int16_t), st));
  jpeg_huffman_kernel<<<n, HUFF_THREADS, smem, st>>>(blob, coef, iterations);
  VLTK_LAUNCH_CHECK();
  jpeg_dc_scan_kernel<<<dim3(n, 3), 1024, 0, st>>>(blob, coef);
  VLTK_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
