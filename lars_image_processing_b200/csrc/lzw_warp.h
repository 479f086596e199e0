// TIFF LZW (TIFF 6.0 section 13) decoded by one warp per strip -- the device form of lzw_chunk() in
// tiff_host.h, written once for both compilers: nvcc builds the warp version, tests/hostcheck builds the same
// source for the host with the 32 lanes run one after the other, so the control flow and the table logic
// are checked against the host decoder on the CPU (TEST INFRASTRUCTURE: the product only runs the device build).
//
// All lanes of the warp run the code stream in lockstep with identical state (bit buffer, table, positions);
// what is shared out is the copy of each string: lane l moves bytes l, l + 32, ... .  As in the host decoder
// the table holds (position, length) of every string inside the output already written, so emitting a string
// is a copy from earlier output, never a walk down a prefix chain.  The table lives in shared memory (one
// 32-bit word per code: position in bits 0..19, length in bits 20..31, hence strips of at most 1 MB); every
// lane writes every entry itself (same value, same address: one shared-memory transaction), so table reads
// never depend on another lane and only the output copy needs a warp barrier.  (Table slots are reused after a Clear
// code; a lane could only overwrite a slot another lane's pending write still targets if it ran a whole epoch of
// literal codes ahead between two barriers, which no encoder's stream allows -- a hostile stream could at worst turn its
// own pixels into different garbage, positions and lengths stay inside the output.  Variant 2 closes even that with a
// barrier at every Clear.)
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDA_ARCH__)
#define LARS_LZW_FN __device__ __forceinline__
#define LARS_LZW_FOR_LANES(lane) const int lane = (int)(threadIdx.x & 31u);
#define LARS_LZW_SYNC() __syncwarp()
#define LARS_LZW_LOAD(p) __ldg(p)
#else
#define LARS_LZW_FN static inline
#define LARS_LZW_FOR_LANES(lane) for (int lane = 0; lane < 32; ++lane)
#define LARS_LZW_SYNC() ((void)0)
#define LARS_LZW_LOAD(p) (*(p))
#endif

#define LARS_LZW_MAX_CHUNK (1u << 20)   /* decoded bytes per strip the (position, length) word can address */

// Decodes one LZW stream of n_in bytes into out[0, cap); returns the number of bytes produced, 0 for a
// corrupt stream.  cap <= LARS_LZW_MAX_CHUNK.  `table`: 4096 words private to the warp.
LARS_LZW_FN uint32_t lars_lzw_decode_warp(const uint8_t* in, uint32_t n_in, uint8_t* out, uint32_t cap,
                                          uint32_t* table) {
  uint64_t acc = 0;
  int have = 0, nbits = 9, next_code = 258, old = -1;
  uint32_t ip = 0, op = 0, old_pos = 0, old_len = 0;
  while (op < cap) {
    if (have < nbits) {                               // refill: several codes' worth of bits at a time
      while (have <= 56 && ip < n_in) { acc = (acc << 8) | (uint64_t)LARS_LZW_LOAD(in + ip); ++ip; have += 8; }
      if (have < nbits) break;                        // ran out of input: treat as EndOfInformation
    }
    const int code = (int)((acc >> (have - nbits)) & ((1u << nbits) - 1u));
    have -= nbits;
    if (code == 256) { nbits = 9; next_code = 258; old = -1; continue; }
    if (code == 257) break;
    const uint32_t at = op;
    uint32_t len;
    if (code < 256) {                                 // a literal
      LARS_LZW_FOR_LANES(lane) { if (lane == 0) out[op] = (uint8_t)code; }
      op += 1;
      len = 1;
    } else if (old < 0) {
      return 0;                                       // the first code after a Clear must be a literal
    } else if (code < next_code) {                    // a string made earlier: it ends before `op`
      const uint32_t e = table[code];
      const uint32_t pos = e & 0xFFFFFu;
      len = e >> 20;
      const uint32_t keep = len < cap - op ? len : cap - op;
      LARS_LZW_SYNC();                                // the bytes other lanes wrote are visible
      LARS_LZW_FOR_LANES(lane) { for (uint32_t i = (uint32_t)lane; i < keep; i += 32u) out[op + i] = out[pos + i]; }
      op += keep;
    } else if (code == next_code && next_code < 4096) {   // the string being defined: old + first(old)
      len = old_len + 1;
      const uint32_t keep = len < cap - op ? len : cap - op;
      LARS_LZW_SYNC();
      LARS_LZW_FOR_LANES(lane) {
        for (uint32_t i = (uint32_t)lane; i < keep; i += 32u) out[op + i] = out[old_pos + (i < old_len ? i : 0u)];
      }
      op += keep;
    } else {
      return 0;                                       // a code the table cannot hold yet: corrupt stream
    }
    if (old >= 0) {
      if (next_code < 4096) {
        table[next_code] = old_pos | ((old_len + 1u) << 20);
        ++next_code;
      }
      if (next_code >= (1 << nbits) - 1 && nbits < 12) ++nbits;
    }
    old = code;
    old_pos = at;
    old_len = len;
  }
  return op;
}

// A second variant (compressed stream and output staged through shared-memory rings, 6 warps per SM) was measured on
// B200 at 118-133 ms against 68-88 ms for the decoder above on 16 x 12 MP frames (profiles/r02_device_decode.log):
// occupancy, not the scattered global accesses, is what hides the per-code latency.  It was removed.

LARS_LZW_FN uint32_t lars_lzw_bswap32(uint32_t w) {
  return (w >> 24) | ((w >> 8) & 0xFF00u) | ((w << 8) & 0xFF0000u) | (w << 24);
}
