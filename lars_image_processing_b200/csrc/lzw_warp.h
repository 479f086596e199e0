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

// ---- variant 2 (opt-in, LARS_LZW_VARIANT=2) -----------------------------------------------------------------
// Measured history (B200, 16 x 12 MP noise frames = 9,600 strips): variant 1 above 88 ms; a first variant 2 that
// only mirrored recent output in a shared-memory ring (string copies on chip, 7 warps per SM instead of 12)
// 118 ms -- per warp just 1.27x faster, so the L2 round trip of the copy was not what a code costs.  What is
// left are the scattered small accesses on both sides with almost no L1 behind them (shared memory takes most
// of the SM's array): byte loads of the compressed stream and 1-3 byte stores of the output.  This variant
// (written after that measurement, NOT yet timed on hardware) removes both:
//   * input: the warp fetches the compressed stream 128 bytes at a time -- every lane one aligned 32-bit word,
//     the next segment already in flight in a register while the current one is consumed -- into a 1 KB
//     shared ring; the bit buffer refills from there four bytes at a time.  `in - (in & 3)` up to the next
//     multiple of 4 after the stream must be readable (file starts are 8-byte aligned and padded).
//   * output: strings are assembled in a 16 KB shared ring that is also the source of every copy no further
//     back than LARS_LZW_RING - 4096 bytes; completed 128-byte stretches leave for global memory as aligned
//     32-bit stores, 32 lanes side by side.
#define LARS_LZW_RING 16384u
#define LARS_LZW_INBUF_WORDS 256u   /* 1 KB ring of the compressed stream */

LARS_LZW_FN uint32_t lars_lzw_bswap32(uint32_t w) {
  return (w >> 24) | ((w >> 8) & 0xFF00u) | ((w << 8) & 0xFF0000u) | (w << 24);
}

LARS_LZW_FN uint32_t lars_lzw_decode_warp_v2(const uint8_t* in, uint32_t n_in, uint8_t* out, uint32_t cap,
                                             uint32_t* table, uint8_t* ring, uint32_t* inbuf) {
  // ---- input side: stream byte k lives at byte (skew + k) of the aligned word array `words`
  const uint32_t skew = (uint32_t)((uintptr_t)in & 3u);
  const uint32_t* words = reinterpret_cast<const uint32_t*>(in - skew);
  const uint32_t end_q = skew + n_in;                      // one past the last stream byte, buffer coordinates
  const uint32_t n_words = (end_q + 3u) >> 2;              // words that may be read
  uint32_t loaded = 0;                                     // words already in inbuf
  uint32_t fetched = 0;                                    // words already requested into `pending`
#if defined(__CUDA_ARCH__)
  uint32_t pending = 0;                                    // this lane's word of the segment in flight
#else
  uint32_t pending[32];
#endif
  // request the words [fetched, fetched + 32) into the lanes' registers
#define LARS_LZW_FETCH()                                                                              \
  do {                                                                                                \
    LARS_LZW_FOR_LANES(lane) {                                                                        \
      const uint32_t w_ = fetched + (uint32_t)lane;                                                   \
      LARS_LZW_PENDING(lane) = w_ < n_words ? LARS_LZW_LOAD(words + w_) : 0u;                         \
    }                                                                                                 \
    fetched += 32u;                                                                                   \
  } while (0)
  // park the segment in flight in the shared ring and request the next one
#define LARS_LZW_COMMIT()                                                                             \
  do {                                                                                                \
    LARS_LZW_FOR_LANES(lane) { inbuf[(loaded + (uint32_t)lane) & (LARS_LZW_INBUF_WORDS - 1u)] = LARS_LZW_PENDING(lane); } \
    loaded += 32u;                                                                                    \
    LARS_LZW_SYNC();                                                                                  \
    LARS_LZW_FETCH();                                                                                 \
  } while (0)
#if defined(__CUDA_ARCH__)
#define LARS_LZW_PENDING(lane) pending
#else
#define LARS_LZW_PENDING(lane) pending[lane]
#endif
  LARS_LZW_FETCH();
  LARS_LZW_COMMIT();

  uint64_t acc = 0;
  int have = 0, nbits = 9, next_code = 258, old = -1;
  uint32_t q = skew;                                       // next stream byte to enter the bit buffer
  uint32_t op = 0, old_pos = 0, old_len = 0;
  // ---- output side: bytes [flushed, op) exist only in the ring
  uint32_t flushed = 0;
  const uint32_t out_skew = (uint32_t)((uintptr_t)out & 3u);
  while (op < cap) {
    if (have < nbits) {
      // keep at least one whole segment ahead of the reader (the ring holds four)
      while (loaded < n_words && (loaded << 2) < q + 256u) LARS_LZW_COMMIT();
      while (have <= 56 && (q & 3u) != 0u && q < end_q) {  // unaligned head of the stream: single bytes
        acc = (acc << 8) | (uint64_t)((inbuf[(q >> 2) & (LARS_LZW_INBUF_WORDS - 1u)] >> (8u * (q & 3u))) & 0xFFu);
        ++q;
        have += 8;
      }
      if (have <= 32 && (q & 3u) == 0u && q + 4u <= end_q) {   // four stream bytes at once, first byte on top
        acc = (acc << 32) | (uint64_t)lars_lzw_bswap32(inbuf[(q >> 2) & (LARS_LZW_INBUF_WORDS - 1u)]);
        q += 4u;
        have += 32;
      }
      while (have <= 56 && q < end_q && q + 4u > end_q) {  // tail shorter than a word
        acc = (acc << 8) | (uint64_t)((inbuf[(q >> 2) & (LARS_LZW_INBUF_WORDS - 1u)] >> (8u * (q & 3u))) & 0xFFu);
        ++q;
        have += 8;
      }
      if (have < nbits) break;                             // ran out of input: treat as EndOfInformation
    }
    const int code = (int)((acc >> (have - nbits)) & ((1u << nbits) - 1u));
    have -= nbits;
    if (code == 256) { LARS_LZW_SYNC(); nbits = 9; next_code = 258; old = -1; continue; }   // no lane drifts across an epoch
    if (code == 257) break;
    const uint32_t at = op;
    uint32_t len;
    if (code < 256) {
      LARS_LZW_FOR_LANES(lane) { if (lane == 0) ring[op & (LARS_LZW_RING - 1u)] = (uint8_t)code; }
      op += 1;
      len = 1;
    } else if (old < 0) {
      return 0;
    } else if (code < next_code || (code == next_code && next_code < 4096)) {
      uint32_t pos, wrap;                                  // string = bytes [pos, pos + wrap) followed by byte pos again
      if (code < next_code) {
        const uint32_t e = table[code];
        pos = e & 0xFFFFFu;
        len = e >> 20;
        wrap = len;
      } else {                                             // the string being defined: old + first(old)
        pos = old_pos;
        len = old_len + 1;
        wrap = old_len;
      }
      const uint32_t keep = len < cap - op ? len : cap - op;
      const bool near = op - pos <= LARS_LZW_RING - 4096u;  // else older than the ring keeps: flushed long ago
      LARS_LZW_SYNC();                                     // ring bytes / global bytes of other lanes are visible
      LARS_LZW_FOR_LANES(lane) {
        for (uint32_t i = (uint32_t)lane; i < keep; i += 32u) {
          const uint32_t from = pos + (i < wrap ? i : 0u);
          ring[(op + i) & (LARS_LZW_RING - 1u)] = near ? ring[from & (LARS_LZW_RING - 1u)] : out[from];
        }
      }
      LARS_LZW_SYNC();                                     // nobody still reads the ring when a lane that runs ahead writes on
      op += keep;
    } else {
      return 0;
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
    // ---- flush completed stretches: first up to a 4-byte boundary of the destination, then 128 bytes at a time
    if (op - flushed >= 256u) {
      LARS_LZW_SYNC();
      const uint32_t head = (4u - ((out_skew + flushed) & 3u)) & 3u;
      LARS_LZW_FOR_LANES(lane) { if ((uint32_t)lane < head) out[flushed + lane] = ring[(flushed + lane) & (LARS_LZW_RING - 1u)]; }
      flushed += head;
      while (op - flushed >= 128u) {
        LARS_LZW_FOR_LANES(lane) {
          const uint32_t b = flushed + 4u * (uint32_t)lane;
          const uint32_t w = (uint32_t)ring[b & (LARS_LZW_RING - 1u)] | ((uint32_t)ring[(b + 1u) & (LARS_LZW_RING - 1u)] << 8) |
                             ((uint32_t)ring[(b + 2u) & (LARS_LZW_RING - 1u)] << 16) |
                             ((uint32_t)ring[(b + 3u) & (LARS_LZW_RING - 1u)] << 24);
          *reinterpret_cast<uint32_t*>(out + b) = w;
        }
        flushed += 128u;
      }
    }
  }
  LARS_LZW_SYNC();
  LARS_LZW_FOR_LANES(lane) {                               // what is left in the ring, byte by byte
    for (uint32_t b = flushed + (uint32_t)lane; b < op; b += 32u) out[b] = ring[b & (LARS_LZW_RING - 1u)];
  }
#undef LARS_LZW_FETCH
#undef LARS_LZW_COMMIT
#undef LARS_LZW_PENDING
  return op;
}
