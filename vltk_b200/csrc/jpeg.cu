// Device half of the JPEG front end: dequantisation + inverse DCT + chroma upsampling + colour
// conversion of the coefficients produced by jpeg_host.cpp, writing the BGR u8 [h, w, 3] image that
// `cv2.imread` returns at the reference's call site (vltk/compat.py:573-579) — bit for bit: the integer
// algorithms below are libjpeg(-turbo)'s defaults (JDCT_ISLOW, fancy upsampling, 16-bit fixed-point YCbCr->RGB),
// restated from their published description (IJG libjpeg jidctint.c / jdsample.c / jdcolor.c; the SIMD code of
// libjpeg-turbo is specified to produce identical output).
#include "../../include/vltk_frcnn.h"
#include "common.cuh"

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

}  // extern "C"
