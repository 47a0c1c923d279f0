// Host half of the JPEG front end: marker parsing + Huffman entropy decoding of baseline /
// extended-sequential 8-bit JPEG (ITU-T T.81 Annex B, F.2) into quantised DCT coefficients.  Everything after
// the entropy decoder — dequantisation, IDCT, chroma upsampling, colour conversion — runs on the GPU (jpeg.cu).
//
// Replaces the decode inside `cv2.imread` at the reference's image-loading call site (vltk/compat.py:573-579
// `img_tensorize`, reached from legacy/processing.py:119-129); the pixel arithmetic being matched is
// libjpeg-turbo's (bundled in the cv2 / PIL wheels, not part of the reference tree).
//
// Coefficient layout handed to the GPU: per component, blocks in raster order over the component's padded block
// grid (blocks_w = mcus_x * h, blocks_h = mcus_y * v), 64 int16 per block in NATURAL (row-major) order.
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/vltk_frcnn.h"
#include "jpeg_dev.h"

namespace vltk {
void set_error(const char* fmt, ...);
}
using vltk::set_error;

namespace {

const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct HuffTable {
  bool present = false;
  uint8_t bits[17] = {0};
  uint8_t vals[256] = {0};
  // canonical decode tables (T.81 F.2.2.3) + a 9-bit lookahead
  int32_t maxcode[18];
  int32_t valoffset[17];
  uint16_t look[512];  // (len << 8) | symbol, 0 = longer than 9 bits
  // AC tables: code + magnitude bits resolved by ONE lookup when both fit in 9 bits:
  // (value << 8) | (run << 4) | (code length + magnitude bits), 0 = take the general path
  int16_t fast_ac[512];

  bool build() {
    int nsym = 0;
    for (int l = 1; l <= 16; ++l) nsym += bits[l];
    if (nsym > 256) return false;
    uint32_t code = 0;
    int p = 0;
    memset(look, 0, sizeof(look));
    for (int l = 1; l <= 16; ++l) {
      valoffset[l] = p - (int32_t)code;
      for (int i = 0; i < bits[l]; ++i, ++p, ++code) {
        if (code >= (1u << l)) return false;
        if (l <= 9) {
          const uint32_t first = code << (9 - l);
          for (uint32_t k = 0; k < (1u << (9 - l)); ++k) look[first + k] = (uint16_t)((l << 8) | vals[p]);
        }
      }
      maxcode[l] = bits[l] ? (int32_t)code - 1 : -1;
      code <<= 1;
    }
    maxcode[17] = 0x7fffffff;
    for (int i = 0; i < 512; ++i) {
      fast_ac[i] = 0;
      const uint16_t e = look[i];
      if (!e) continue;
      const int len = e >> 8, run = (e & 255) >> 4, mag = e & 15;
      if (mag && len + mag <= 9) {
        int k = ((i << len) & 511) >> (9 - mag);
        if (k < (1 << (mag - 1))) k += (int)((~0u) << mag) + 1;
        if (k >= -128 && k <= 127) fast_ac[i] = (int16_t)(k * 256 + run * 16 + len + mag);
      }
    }
    return true;
  }
};

struct Parsed {
  vltk_jpeg_info info;
  HuffTable dc[4], ac[4];
  int td[3], ta[3];
  size_t scan_offset = 0;   // first entropy-coded byte
};

inline int rd16(const uint8_t* p) { return (p[0] << 8) | p[1]; }

// EXIF orientation (TIFF tag 0x0112) of an APP1 segment, 0 if absent
int exif_orientation(const uint8_t* p, int len) {
  if (len < 14 || memcmp(p, "Exif\0\0", 6) != 0) return 0;
  const uint8_t* t = p + 6;
  const int n = len - 6;
  const bool le = t[0] == 'I' && t[1] == 'I';
  if (!le && !(t[0] == 'M' && t[1] == 'M')) return 0;
  auto u16 = [&](int o) { return le ? (t[o] | (t[o + 1] << 8)) : ((t[o] << 8) | t[o + 1]); };
  auto u32 = [&](int o) {
    return le ? (uint32_t)(t[o] | (t[o + 1] << 8) | (t[o + 2] << 16) | ((uint32_t)t[o + 3] << 24))
              : (uint32_t)(((uint32_t)t[o] << 24) | (t[o + 1] << 16) | (t[o + 2] << 8) | t[o + 3]);
  };
  if (n < 8) return 0;
  const uint32_t ifd = u32(4);
  if (ifd > (uint32_t)n - 2) return 0;                  // (no `ifd + 2`: it wraps for offsets >= 0xFFFFFFFE)
  const int cnt = u16((int)ifd);
  for (int i = 0; i < cnt; ++i) {
    const uint64_t e = (uint64_t)ifd + 2 + 12ull * i;
    if (e + 12 > (uint64_t)n) return 0;
    if (u16((int)e) == 0x0112) return u16((int)e + 8);
  }
  return 0;
}

int parse_dht(const uint8_t* s, int n, Parsed* P) {
  int o = 0;
  while (o + 17 <= n) {
    const int tc = s[o] >> 4, th = s[o] & 15;
    if (tc > 1 || th > 3) { set_error("jpeg: bad DHT id"); return -2; }
    HuffTable& H = tc ? P->ac[th] : P->dc[th];
    int cnt = 0;
    H.bits[0] = 0;
    for (int i = 1; i <= 16; ++i) { H.bits[i] = s[o + i]; cnt += s[o + i]; }
    o += 17;
    if (cnt > 256 || o + cnt > n) { set_error("jpeg: bad DHT length"); return -2; }
    memcpy(H.vals, s + o, cnt);
    o += cnt;
    if (!H.build()) { set_error("jpeg: invalid Huffman table"); return -2; }
    H.present = true;
  }
  return 0;
}

int parse(const uint8_t* d, size_t len, Parsed* P) {
  vltk_jpeg_info& I = P->info;
  memset(&I, 0, sizeof(I));
  I.color_transform = -1;
  if (len < 4 || d[0] != 0xFF || d[1] != 0xD8) { set_error("jpeg: missing SOI marker"); return -2; }
  uint16_t qt_zz[4][64];
  bool qt_ok[4] = {false, false, false, false};
  int tq[3] = {0, 0, 0};
  bool have_sof = false;
  size_t pos = 2;
  while (pos + 4 <= len) {
    if (d[pos] != 0xFF) { set_error("jpeg: expected a marker at byte %zu", pos); return -2; }
    while (pos < len && d[pos] == 0xFF) ++pos;          // fill bytes
    if (pos >= len) break;
    const int m = d[pos++];
    if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) continue;
    if (m == 0xD9) break;
    if (pos + 2 > len) break;
    const int L = rd16(d + pos);
    if (L < 2 || pos + L > len) { set_error("jpeg: truncated segment 0xFF%02X", m); return -2; }
    const uint8_t* s = d + pos + 2;
    const int n = L - 2;
    if (m == 0xDB) {                                     // DQT
      int o = 0;
      while (o < n) {
        const int pq = s[o] >> 4, t = s[o] & 15;
        ++o;
        if (t > 3 || o + (pq ? 128 : 64) > n) { set_error("jpeg: bad DQT"); return -2; }
        for (int i = 0; i < 64; ++i) { qt_zz[t][i] = pq ? (uint16_t)rd16(s + o + 2 * i) : s[o + i]; }
        o += pq ? 128 : 64;
        qt_ok[t] = true;
      }
    } else if (m == 0xC4) {                              // DHT
      const int rc = parse_dht(s, n, P);
      if (rc) return rc;
    } else if (m == 0xC0 || m == 0xC1 || m == 0xC2) {    // SOF0 / SOF1 (sequential) / SOF2 (progressive), Huffman
      I.progressive = m == 0xC2;
      if (n < 6) { set_error("jpeg: bad SOF"); return -2; }
      if (s[0] != 8) { set_error("jpeg: %d-bit samples are not supported (8 only)", s[0]); return -3; }
      I.height = rd16(s + 1); I.width = rd16(s + 3); I.ncomp = s[5];
      if (I.width < 1 || I.height < 1) { set_error("jpeg: empty image"); return -2; }
      if (I.ncomp != 1 && I.ncomp != 3) { set_error("jpeg: %d components are not supported (1 or 3)", I.ncomp); return -3; }
      if (n < 6 + 3 * I.ncomp) { set_error("jpeg: bad SOF"); return -2; }
      for (int c = 0; c < I.ncomp; ++c) {
        I.comp_id[c] = s[6 + 3 * c];
        I.hs[c] = s[7 + 3 * c] >> 4; I.vs[c] = s[7 + 3 * c] & 15;
        tq[c] = s[8 + 3 * c];
        if (tq[c] > 3) { set_error("jpeg: bad quantisation table id"); return -2; }
      }
      have_sof = true;
    } else if (m >= 0xC3 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) {
      set_error("jpeg: SOF%d (lossless / hierarchical / arithmetic coding) is not supported", m - 0xC0);
      return -3;
    } else if (m == 0xDD) {
      if (n >= 2) I.restart_interval = rd16(s);
    } else if (m == 0xE1) {
      const int o = exif_orientation(s, n);
      if (o) I.orientation = o;
    } else if (m == 0xEE) {                              // Adobe: colour transform flag
      if (n >= 12 && memcmp(s, "Adobe", 5) == 0) I.color_transform = s[11];
    } else if (m == 0xDA) {                              // SOS
      if (!have_sof) { set_error("jpeg: SOS before SOF"); return -2; }
      if (I.progressive) { P->scan_offset = pos - 2; break; }   // the scans are walked by decode_progressive()
      if (n < 1) { set_error("jpeg: empty SOS segment"); return -2; }
      const int ns = s[0];
      if (ns != I.ncomp || n < 1 + 2 * ns + 3) {
        set_error("jpeg: only one interleaved scan covering all components is supported (scan has %d of %d)", ns, I.ncomp);
        return -3;
      }
      for (int k = 0; k < ns; ++k) {
        int c = -1;
        for (int j = 0; j < I.ncomp; ++j) if (I.comp_id[j] == s[1 + 2 * k]) c = j;
        if (c != k) { set_error("jpeg: scan component order differs from the frame"); return -3; }
        P->td[c] = s[2 + 2 * k] >> 4; P->ta[c] = s[2 + 2 * k] & 15;
        if (P->td[c] > 3 || P->ta[c] > 3 || !P->dc[P->td[c]].present || !P->ac[P->ta[c]].present) {
          set_error("jpeg: scan references a missing Huffman table"); return -2;
        }
      }
      if (s[1 + 2 * ns] != 0 || s[2 + 2 * ns] != 63 || s[3 + 2 * ns] != 0) { set_error("jpeg: not a sequential scan"); return -3; }
      P->scan_offset = pos + L;
      break;
    }
    pos += L;
  }
  if (!have_sof || !P->scan_offset) { set_error("jpeg: no frame / scan found"); return -2; }
  // geometry
  int hmax = 1, vmax = 1;
  for (int c = 0; c < I.ncomp; ++c) { hmax = I.hs[c] > hmax ? I.hs[c] : hmax; vmax = I.vs[c] > vmax ? I.vs[c] : vmax; }
  if (I.ncomp == 1) { I.hs[0] = I.vs[0] = 1; hmax = vmax = 1; }       // a single-component scan is non-interleaved (A.2.2)
  if (I.ncomp == 3) {
    const bool ok = I.hs[1] == 1 && I.vs[1] == 1 && I.hs[2] == 1 && I.vs[2] == 1 &&
                    ((I.hs[0] == 1 && I.vs[0] == 1) || (I.hs[0] == 2 && I.vs[0] == 1) || (I.hs[0] == 2 && I.vs[0] == 2));
    if (!ok) {
      set_error("jpeg: sampling %dx%d,%dx%d,%dx%d is not supported (4:4:4, 4:2:2, 4:2:0 only)", I.hs[0], I.vs[0], I.hs[1],
                I.vs[1], I.hs[2], I.vs[2]);
      return -3;
    }
  }
  I.hmax = hmax; I.vmax = vmax;
  I.mcus_x = (I.width + 8 * hmax - 1) / (8 * hmax);
  I.mcus_y = (I.height + 8 * vmax - 1) / (8 * vmax);
  int64_t off = 0, poff = 0;
  for (int c = 0; c < I.ncomp; ++c) {
    if (!qt_ok[tq[c]]) { set_error("jpeg: missing quantisation table %d", tq[c]); return -2; }
    for (int i = 0; i < 64; ++i) I.qt[c][kZigzag[i]] = qt_zz[tq[c]][i];
    I.blocks_w[c] = I.mcus_x * I.hs[c]; I.blocks_h[c] = I.mcus_y * I.vs[c];
    I.comp_w[c] = (I.width * I.hs[c] + hmax - 1) / hmax;               // downsampled_width / height (jdmaster.c)
    I.comp_h[c] = (I.height * I.vs[c] + vmax - 1) / vmax;
    I.coef_offset[c] = off;
    off += (int64_t)I.blocks_w[c] * I.blocks_h[c] * 64;
    I.plane_offset[c] = poff;
    poff += (int64_t)I.blocks_w[c] * I.blocks_h[c] * 64;
  }
  I.coef_count = off;
  I.plane_bytes = poff;
  return 0;
}

struct BitReader {
  const uint8_t* p;
  const uint8_t* end;
  uint64_t acc = 0;
  int nbits = 0;
  bool hit_marker = false;
  BitReader(const uint8_t* b, const uint8_t* e) : p(b), end(e) {}
  inline void fill() {
    if (nbits > 56) return;
    if (!hit_marker && end - p >= 8) {                   // 8 bytes without an 0xFF: take every whole byte that fits
      uint64_t v;
      memcpy(&v, p, 8);
      v = __builtin_bswap64(v);
      const uint64_t t = ~v;                             // a zero byte of t <=> an 0xFF byte of v
      if (!((t - 0x0101010101010101ULL) & ~t & 0x8080808080808080ULL)) {
        const int adv = (64 - nbits) >> 3, rem = 64 - nbits - adv * 8;
        acc |= (v >> nbits) & (~0ULL << rem);
        p += adv;
        nbits += adv * 8;
        return;
      }
    }
    while (nbits <= 56) {
      uint32_t b = 0;
      if (!hit_marker && p < end) {
        b = *p;
        if (b == 0xFF) {
          if (p + 1 < end && p[1] == 0) p += 2;          // stuffed zero
          else { hit_marker = true; b = 0; }             // a marker: feed zeros (T.81 F.2.2.5 / libjpeg behaviour)
        } else ++p;
      } else hit_marker = true;
      acc |= (uint64_t)b << (56 - nbits);
      nbits += 8;
    }
  }
  inline uint32_t peek(int n) { return (uint32_t)(acc >> (64 - n)); }
  inline void skip(int n) { acc <<= n; nbits -= n; }
  inline int receive_extend(int s) {                     // F.2.2.1 RECEIVE + EXTEND
    if (!s) return 0;
    if (nbits < s) fill();
    const int v = (int)peek(s);
    skip(s);
    return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v;
  }
  inline int decode(const HuffTable& H) {
    if (nbits < 16) fill();
    const uint16_t e = H.look[peek(9)];
    if (e) { skip(e >> 8); return e & 255; }
    int32_t code = (int32_t)peek(10);
    int l = 10;
    while (l <= 16 && code > H.maxcode[l]) { ++l; code = (int32_t)peek(l); }
    if (l > 16) return -1;
    skip(l);
    return H.vals[(code + H.valoffset[l]) & 255];
  }
  // byte-align and consume an RSTn marker
  bool restart() {
    acc = 0; nbits = 0;
    hit_marker = false;
    while (p + 1 < end) {
      if (p[0] == 0xFF && p[1] >= 0xD0 && p[1] <= 0xD7) { p += 2; return true; }
      if (p[0] == 0xFF && p[1] == 0xFF) { ++p; continue; }
      if (p[0] == 0xFF && p[1] != 0) return false;       // some other marker
      ++p;                                               // garbage / padding before the marker
    }
    return false;
  }
};

// ---- progressive JPEG (SOF2, T.81 Annex G): several scans refine one coefficient array.  Entropy decoding is
// serial per scan and runs here on the host; the GPU stages that follow are the same as for baseline files.
inline int get_bits(BitReader& br, int n) {
  if (!n) return 0;
  if (br.nbits < n) br.fill();
  const int v = (int)br.peek(n);
  br.skip(n);
  return v;
}

int decode_progressive(const uint8_t* d, size_t len, Parsed& P, int16_t* coef) {
  const vltk_jpeg_info& I = P.info;
  memset(coef, 0, (size_t)I.coef_count * sizeof(int16_t));
  size_t pos = P.scan_offset;                            // at the 0xFF of the first SOS
  int restart_interval = I.restart_interval;
  while (pos + 4 <= len) {
    if (d[pos] != 0xFF) { ++pos; continue; }
    while (pos < len && d[pos] == 0xFF) ++pos;
    if (pos >= len) break;
    const int m = d[pos++];
    if (m == 0xD9) break;
    if (m == 0 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) continue;
    if (pos + 2 > len) break;
    const int L = rd16(d + pos);
    if (L < 2 || pos + L > len) { set_error("jpeg: truncated segment 0xFF%02X", m); return -2; }
    const uint8_t* s = d + pos + 2;
    const int n = L - 2;
    if (m == 0xC4) { const int rc = parse_dht(s, n, &P); if (rc) return rc; pos += L; continue; }
    if (m == 0xDD) { if (n >= 2) restart_interval = rd16(s); pos += L; continue; }
    if (m != 0xDA) { pos += L; continue; }
    // ---- one scan
    if (n < 1) { set_error("jpeg: empty SOS segment"); return -2; }
    const int ns = s[0];
    if (ns < 1 || ns > I.ncomp || n < 1 + 2 * ns + 3) { set_error("jpeg: bad SOS"); return -2; }
    int comp[3], td[3], ta[3];
    for (int k = 0; k < ns; ++k) {
      int c = -1;
      for (int j = 0; j < I.ncomp; ++j) if (I.comp_id[j] == s[1 + 2 * k]) c = j;
      if (c < 0) { set_error("jpeg: scan names an unknown component"); return -2; }
      comp[k] = c; td[k] = s[2 + 2 * k] >> 4; ta[k] = s[2 + 2 * k] & 15;
      if (td[k] > 3 || ta[k] > 3) { set_error("jpeg: bad table id in SOS"); return -2; }
    }
    const int Ss = s[1 + 2 * ns], Se = s[2 + 2 * ns], Ah = s[3 + 2 * ns] >> 4, Al = s[3 + 2 * ns] & 15;
    if (Ss > Se || Se > 63 || (Ss == 0 && Se != 0) || (Ss > 0 && ns != 1) || Al > 13) { set_error("jpeg: invalid progressive scan parameters"); return -2; }
    for (int k = 0; k < ns; ++k) {
      if (Ss == 0 && Ah == 0 && !P.dc[td[k]].present) { set_error("jpeg: scan references a missing DC table"); return -2; }
      if (Ss > 0 && !P.ac[ta[k]].present) { set_error("jpeg: scan references a missing AC table"); return -2; }
    }
    BitReader br(d + pos + L, d + len);
    int pred[3] = {0, 0, 0};
    int eobrun = 0;
    int until_restart = restart_interval;
    // block iteration: interleaved scans walk MCUs; a single-component scan walks that component's own block grid
    const bool inter = ns > 1;
    const int c0 = comp[0];
    const int bw = inter ? I.mcus_x : (I.comp_w[c0] + 7) / 8, bh = inter ? I.mcus_y : (I.comp_h[c0] + 7) / 8;
    for (int uy = 0; uy < bh; ++uy)
      for (int ux = 0; ux < bw; ++ux) {
        if (restart_interval && until_restart == 0) {
          if (!br.restart()) { set_error("jpeg: missing restart marker in a progressive scan"); return -2; }
          pred[0] = pred[1] = pred[2] = 0; eobrun = 0;
          until_restart = restart_interval;
        }
        for (int k = 0; k < ns; ++k) {
          const int c = comp[k];
          const int nby = inter ? I.vs[c] : 1, nbx = inter ? I.hs[c] : 1;
          for (int by = 0; by < nby; ++by)
            for (int bx = 0; bx < nbx; ++bx) {
              const int64_t blk = inter ? (int64_t)(uy * I.vs[c] + by) * I.blocks_w[c] + (ux * I.hs[c] + bx)
                                        : (int64_t)uy * I.blocks_w[c] + ux;
              int16_t* out = coef + I.coef_offset[c] + blk * 64;
              if (Ss == 0) {
                if (Ah == 0) {                           // DC first (G.1.2.1)
                  const int t = br.decode(P.dc[td[k]]);
                  if (t < 0 || t > 15) { set_error("jpeg: corrupt DC code in a progressive scan"); return -2; }
                  pred[c] += br.receive_extend(t);
                  out[0] = (int16_t)(pred[c] * (1 << Al));
                } else if (get_bits(br, 1)) out[0] |= (int16_t)(1 << Al);   // DC refinement
              } else if (Ah == 0) {                      // AC first (G.1.2.2)
                if (eobrun > 0) { --eobrun; continue; }
                const HuffTable& A = P.ac[ta[k]];
                for (int kk = Ss; kk <= Se;) {
                  const int rs = br.decode(A);
                  if (rs < 0) { set_error("jpeg: corrupt AC code in a progressive scan"); return -2; }
                  const int r = rs >> 4, sz = rs & 15;
                  if (sz == 0) {
                    if (r < 15) { eobrun = (1 << r) - 1; if (r) eobrun += get_bits(br, r); break; }
                    kk += 16;
                  } else {
                    kk += r;
                    if (kk > 63) { set_error("jpeg: AC run past the block in a progressive scan"); return -2; }
                    out[kZigzag[kk]] = (int16_t)(br.receive_extend(sz) * (1 << Al));
                    ++kk;
                  }
                }
              } else {                                   // AC refinement (G.1.2.3)
                const HuffTable& A = P.ac[ta[k]];
                const int p1 = 1 << Al, m1 = -(1 << Al);
                int kk = Ss;
                if (eobrun == 0) {
                  for (; kk <= Se; ++kk) {
                    const int rs = br.decode(A);
                    if (rs < 0) { set_error("jpeg: corrupt AC code in a progressive scan"); return -2; }
                    int r = rs >> 4, sz = rs & 15, val = 0;
                    if (sz) val = get_bits(br, 1) ? p1 : m1;
                    else if (r != 15) { eobrun = 1 << r; if (r) eobrun += get_bits(br, r); break; }
                    do {                                 // skip r zero-history coefficients, refining the others
                      int16_t* cf = out + kZigzag[kk];
                      if (*cf) {
                        if (get_bits(br, 1) && (*cf & p1) == 0) *cf = (int16_t)(*cf + (*cf >= 0 ? p1 : m1));
                      } else if (--r < 0) break;
                      ++kk;
                    } while (kk <= Se);
                    if (val && kk <= 63) out[kZigzag[kk]] = (int16_t)val;
                  }
                }
                if (eobrun > 0) {
                  for (; kk <= Se; ++kk) {
                    int16_t* cf = out + kZigzag[kk];
                    if (*cf && get_bits(br, 1) && (*cf & p1) == 0) *cf = (int16_t)(*cf + (*cf >= 0 ? p1 : m1));
                  }
                  --eobrun;
                }
              }
            }
        }
        --until_restart;
      }
    // continue after this scan's entropy-coded data: the next marker that is not RSTn / stuffing
    size_t q = (size_t)(br.p - d);
    if (q > pos + L + 8) q -= 8;                         // the reader may have run up to 8 bytes ahead
    else q = pos + L;
    while (q + 1 < len && !(d[q] == 0xFF && d[q + 1] != 0 && d[q + 1] != 0xFF && !(d[q + 1] >= 0xD0 && d[q + 1] <= 0xD7))) ++q;
    pos = q;
  }
  return 0;
}

int decode_scan(const uint8_t* d, size_t len, const Parsed& P, int16_t* coef) {
  const vltk_jpeg_info& I = P.info;
  memset(coef, 0, (size_t)I.coef_count * sizeof(int16_t));
  BitReader br(d + P.scan_offset, d + len);
  int pred[3] = {0, 0, 0};
  int until_restart = I.restart_interval;
  for (int my = 0; my < I.mcus_y; ++my) {
    for (int mx = 0; mx < I.mcus_x; ++mx) {
      if (I.restart_interval && until_restart == 0) {
        if (!br.restart()) { set_error("jpeg: missing restart marker at MCU (%d,%d)", mx, my); return -2; }
        pred[0] = pred[1] = pred[2] = 0;
        until_restart = I.restart_interval;
      }
      for (int c = 0; c < I.ncomp; ++c) {
        const HuffTable& DC = P.dc[P.td[c]];
        const HuffTable& AC = P.ac[P.ta[c]];
        for (int by = 0; by < I.vs[c]; ++by)
          for (int bx = 0; bx < I.hs[c]; ++bx) {
            const int64_t blk = (int64_t)(my * I.vs[c] + by) * I.blocks_w[c] + (mx * I.hs[c] + bx);
            int16_t* out = coef + I.coef_offset[c] + blk * 64;
            int s = br.decode(DC);
            if (s < 0 || s > 11) { set_error("jpeg: corrupt DC code at MCU (%d,%d)", mx, my); return -2; }
            pred[c] += br.receive_extend(s);
            out[0] = (int16_t)pred[c];
            for (int k = 1; k < 64;) {
              if (br.nbits < 32) br.fill();
              const int fa = AC.fast_ac[br.peek(9)];
              if (fa) {                                  // run, code and value in one lookup
                k += (fa >> 4) & 15;
                if (k > 63) { set_error("jpeg: AC run past the block at MCU (%d,%d)", mx, my); return -2; }
                out[kZigzag[k++]] = (int16_t)(fa >> 8);
                br.skip(fa & 15);
                continue;
              }
              const int rs = br.decode(AC);
              if (rs < 0) { set_error("jpeg: corrupt AC code at MCU (%d,%d)", mx, my); return -2; }
              const int r = rs >> 4;
              s = rs & 15;
              if (s == 0) {
                if (r != 15) break;                      // EOB
                k += 16;
                continue;
              }
              k += r;
              if (k > 63) { set_error("jpeg: AC run past the block at MCU (%d,%d)", mx, my); return -2; }
              out[kZigzag[k]] = (int16_t)br.receive_extend(s);
              ++k;
            }
          }
      }
      --until_restart;
    }
  }
  return 0;
}

}  // namespace

extern "C" {

int vltk_jpeg_parse(const uint8_t* data, size_t len, vltk_jpeg_info* info) {
  if (!data || !info) { set_error("jpeg_parse: null argument"); return -2; }
  Parsed P;
  const int rc = parse(data, len, &P);
  *info = P.info;
  return rc;
}

int vltk_jpeg_decode_coefficients(const uint8_t* data, size_t len, int16_t* coef, int64_t coef_capacity) {
  if (!data || !coef) { set_error("jpeg_decode_coefficients: null argument"); return -2; }
  Parsed P;
  int rc = parse(data, len, &P);
  if (rc) return rc;
  if (P.info.coef_count > coef_capacity) { set_error("jpeg_decode_coefficients: buffer holds %lld of %lld coefficients", (long long)coef_capacity, (long long)P.info.coef_count); return -2; }
  return P.info.progressive ? decode_progressive(data, len, P, coef) : decode_scan(data, len, P, coef);
}

int vltk_jpeg_decode_coefficients_batch(int n, const uint8_t* const* datas, const size_t* lens, int16_t* const* coefs,
                                        const int64_t* capacities, int n_threads, int* status) {
  if (n < 0 || !datas || !lens || !coefs || !capacities || !status) { set_error("jpeg batch: null argument"); return -2; }
  std::atomic<int> next(0);
  std::vector<std::string> errs(n);
  auto work = [&]() {
    for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) {
      status[i] = vltk_jpeg_decode_coefficients(datas[i], lens[i], coefs[i], capacities[i]);
      if (status[i]) errs[i] = vltk_frcnn_last_error();   // thread-local message of this worker
    }
  };
  const int nt = n_threads < 1 ? 1 : (n_threads > n ? (n > 0 ? n : 1) : n_threads);
  if (nt == 1) work();
  else {
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) th.emplace_back(work);
    for (auto& t : th) t.join();
  }
  for (int i = 0; i < n; ++i)
    if (status[i]) { set_error("image %d: %s", i, errs[i].c_str()); return status[i]; }
  return 0;
}

}  // extern "C"

// ---- preparation for the GPU entropy decoder: tables + destuffed big-endian words, no Huffman work on the host
namespace {

void fill_dev_huff(const HuffTable& H, vltk::DevHuff* D) {
  memset(D, 0, sizeof(*D));
  memcpy(D->look, H.look, sizeof(D->look));
  memcpy(D->fast_ac, H.fast_ac, sizeof(D->fast_ac));
  memcpy(D->maxcode, H.maxcode, sizeof(D->maxcode));
  memcpy(D->valoffset, H.valoffset, sizeof(D->valoffset));
  memcpy(D->vals, H.vals, sizeof(D->vals));
}

// Copies the entropy-coded segment without its stuffed zero bytes, stops at the first marker that is not RSTn, packs
// the bytes into big-endian 32-bit words (+ 3 zero guard words).  RSTn markers are dropped and the byte offset of the
// interval that follows each of them is appended to `starts` (the first interval starts at 0).  Returns the payload bytes.
size_t destuff(const uint8_t* s, const uint8_t* end, uint32_t* words, std::vector<uint32_t>* starts) {
  size_t nb = 0;
  uint8_t* out = reinterpret_cast<uint8_t*>(words);
  if (starts) starts->push_back(0);
  while (s < end) {
    const uint8_t* f = (const uint8_t*)memchr(s, 0xFF, end - s);
    const size_t run = (f ? f : end) - s;
    memcpy(out + nb, s, run);
    nb += run;
    if (!f) break;
    if (f + 1 < end && f[1] == 0) { out[nb++] = 0xFF; s = f + 2; continue; }
    if (f + 1 < end && f[1] == 0xFF) { s = f + 1; continue; }                    // fill byte
    if (starts && f + 1 < end && f[1] >= 0xD0 && f[1] <= 0xD7) { starts->push_back((uint32_t)nb); s = f + 2; continue; }
    break;                                               // a marker (EOI): end of the scan
  }
  const size_t nw = (nb + 3) / 4 + 3;
  memset(out + nb, 0, nw * 4 - nb);
  for (size_t i = 0; i < nw; ++i) words[i] = __builtin_bswap32(words[i]);
  return nb;
}

}  // namespace

extern "C" {

size_t vltk_jpeg_gpu_blob_bound(int n, const size_t* lens) {
  size_t tot = (size_t)n * sizeof(vltk::DevImage);
  // tables + destuffed words + (worst case: one restart interval every 2 bytes of scan data) interval starts
  for (int i = 0; i < n; ++i) tot += 6 * sizeof(vltk::DevHuff) + ((lens[i] + 3) / 4 + 3) * 4 + (lens[i] / 2 + 2) * 4 + 32;
  return tot + 64;
}

int vltk_jpeg_gpu_prepare_batch(int n, const uint8_t* const* datas, const size_t* lens, vltk_jpeg_info* infos,
                                uint8_t* blob, size_t cap, size_t* used, int64_t* coef_offsets, int64_t* coef_total,
                                int* on_gpu) {
  if (n < 0 || !datas || !lens || !infos || !blob || !used || !coef_offsets || !coef_total || !on_gpu) {
    set_error("jpeg_gpu_prepare_batch: null argument"); return -2;
  }
  if (vltk_jpeg_gpu_blob_bound(n, lens) > cap) { set_error("jpeg_gpu_prepare_batch: blob capacity %zu is too small", cap); return -2; }
  vltk::DevImage* imgs = reinterpret_cast<vltk::DevImage*>(blob);
  size_t off = ((size_t)n * sizeof(vltk::DevImage) + 7) & ~(size_t)7;
  int64_t coff = 0;
  for (int i = 0; i < n; ++i) {
    Parsed P;
    const int rc = parse(datas[i], lens[i], &P);
    infos[i] = P.info;
    if (rc) { const std::string m = vltk_frcnn_last_error(); set_error("image %d: %s", i, m.c_str()); return rc; }
    const vltk_jpeg_info& I = P.info;
    vltk::DevImage& D = imgs[i];
    memset(&D, 0, sizeof(D));
    coef_offsets[i] = coff;
    D.coef_off = coff;
    coff += (I.coef_count + 7) / 8 * 8;
    on_gpu[i] = I.progressive ? 0 : 1;                   // progressive scans are entropy-decoded on the host
    if (!on_gpu[i]) continue;
    D.tables_off = (int64_t)off;
    vltk::DevHuff* T = reinterpret_cast<vltk::DevHuff*>(blob + off);
    for (int c = 0; c < 3; ++c) {
      const int cc = c < I.ncomp ? c : 0;
      fill_dev_huff(P.dc[P.td[cc]], T + 2 * c);
      fill_dev_huff(P.ac[P.ta[cc]], T + 2 * c + 1);
    }
    off += 6 * sizeof(vltk::DevHuff);
    D.words_off = (int64_t)off;
    std::vector<uint32_t> starts;
    const size_t nb = destuff(datas[i] + P.scan_offset, datas[i] + lens[i], reinterpret_cast<uint32_t*>(blob + off),
                              I.restart_interval ? &starts : nullptr);
    off += ((nb + 3) / 4 + 3) * 4;
    off = (off + 7) & ~(size_t)7;
    D.total_bits = (int64_t)nb * 8;
    if (I.restart_interval) {
      const int64_t mcus = (int64_t)I.mcus_x * I.mcus_y;
      const int64_t need = (mcus + I.restart_interval - 1) / I.restart_interval;
      if ((int64_t)starts.size() < need) { set_error("image %d: jpeg: %zu restart intervals found, %lld expected", i, starts.size(), (long long)need); return -2; }
      D.restart_interval = I.restart_interval;
      D.n_intervals = (int32_t)need;
      D.starts_off = (int64_t)off;
      uint32_t* so = reinterpret_cast<uint32_t*>(blob + off);
      for (int64_t k = 0; k < need; ++k) so[k] = starts[k] * 8u;
      off += (size_t)need * 4;
      off = (off + 7) & ~(size_t)7;
    }
    int64_t S = (D.total_bits + 1023) / 1024;            // ~1 subsequence per thread of the 1024-thread CTA
    S = (S + 31) / 32 * 32;
    if (S < vltk::JPEG_MIN_SUBSEQ_BITS) S = vltk::JPEG_MIN_SUBSEQ_BITS;
    D.S = (int32_t)S;
    D.nsub = (int32_t)((D.total_bits + S - 1) / S);
    if (D.nsub > vltk::JPEG_MAX_SUBSEQ) { set_error("image %d: internal: too many subsequences", i); return -2; }
    D.ncomp = I.ncomp; D.mcus_x = I.mcus_x;
    int B = 0;
    for (int c = 0; c < I.ncomp; ++c) {
      D.hs[c] = I.hs[c]; D.vs[c] = I.vs[c]; D.blocks_w[c] = I.blocks_w[c]; D.comp_coef_off[c] = I.coef_offset[c];
      for (int by = 0; by < I.vs[c]; ++by)
        for (int bx = 0; bx < I.hs[c]; ++bx) { D.comp_of_block[B] = c; D.bx_of_block[B] = bx; D.by_of_block[B] = by; ++B; }
    }
    D.B = B;
    D.total_blocks = I.mcus_x * I.mcus_y * B;
  }
  *used = off;
  *coef_total = coff;
  return 0;
}

}  // extern "C"
