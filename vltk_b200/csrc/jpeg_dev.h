// Layout of the upload blob shared by the host preparation (jpeg_host.cpp) and the GPU entropy decoder (jpeg.cu):
//   [DevImage x n] [per image: DevHuff x 6 (DC, AC per component) | destuffed entropy-coded words, big-endian u32]
#pragma once
#include <stdint.h>

namespace vltk {

struct DevHuff {
  uint16_t look[512];     // 9-bit lookahead: (len << 8) | symbol, 0 = longer code
  int16_t fast_ac[512];   // (value << 8) | (run << 4) | (len + magnitude bits), 0 = general path
  int32_t maxcode[18];
  int32_t valoffset[17];
  uint8_t vals[256];
  uint8_t pad[4];
};
static_assert(sizeof(DevHuff) % 8 == 0, "DevHuff must keep 8-byte alignment");

constexpr int JPEG_MAX_SUBSEQ = 4096;   // subsequences per image (shared-memory state arrays)
constexpr int JPEG_MIN_SUBSEQ_BITS = 1024;

struct DevImage {
  int64_t tables_off, words_off;   // bytes from the blob start (8-byte aligned)
  int64_t total_bits;              // destuffed entropy-coded bits
  int64_t coef_off;                // int16 elements into the batch coefficient buffer
  int64_t comp_coef_off[3];        // relative to coef_off
  int32_t S, nsub;                 // bits per subsequence, subsequences
  int32_t ncomp, B, total_blocks, mcus_x;
  int32_t hs[3], vs[3], blocks_w[3];
  int32_t comp_of_block[12], bx_of_block[12], by_of_block[12];
  // restart-interval streams: every interval starts byte-aligned in a known state, so no synchronisation is needed —
  // thread t decodes interval t from starts[t] (bit offsets in the destuffed stream, n_intervals entries at starts_off)
  int32_t restart_interval, n_intervals;
  int64_t starts_off;
};

}  // namespace vltk
