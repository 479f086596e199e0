// Deflate (RFC 1951, inside the zlib wrapper of RFC 1950) decoded by one warp per stream -- the device form of
// what zlib's inflate does for Deflate-compressed TIFF strips / tiles.  Written once for
// both compilers like lzw_warp.h: nvcc builds the warp version, tests/hostcheck builds the same source for the host
// with the 32 lanes run one after the other, so every table and every control path is checked against zlib on the
// CPU (TEST INFRASTRUCTURE: the product only runs the device build).
// STATUS: pinned against zlib on the CPU and against the host reader on B200 (round 2: 16 x 12 MP frames = 9,600
// strips in 147 ms on noise, 37 ms on blocky content, against 222 / 64 ms for 16 host threads;
// profiles/r02_device_decode.log).  PNG image data -- one stream per image -- is NOT decoded this way: the device
// path measured 4-11x slower than the host reader and was removed.
//
// All lanes run the bit stream in lockstep with identical state.  Per warp, in shared memory:
//   * a 1 KB ring of the compressed stream, filled 128 bytes at a time (one aligned word per lane, the next
//     segment already in flight in a register);
//   * two 1,024-entry tables that resolve every Huffman code of up to 10 bits with one lookup (literal / length and
//     distance alphabets), filled lane-parallel, plus the canonical (count, sorted symbol) form for longer codes;
//   * a 32 KB ring of the most recent output -- the whole Deflate window -- which is the source of every match
//     (lane l moves bytes l, l + 32, ...; overlapping matches index their period, so no lane waits for another) and
//     from which completed 128-byte stretches leave for global memory as aligned 32-bit stores.
// The Adler-32 trailer IS verified: the checksum is accumulated as the output leaves the ring (lane-parallel partial
// sums per 128-byte stretch), a stream that decodes to the wrong bytes is reported as corrupt like on the host.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "lzw_warp.h"   // the lane / barrier / load macros

#define LARS_INF_RING 32768u
#define LARS_INF_INBUF_WORDS 256u
#define LARS_INF_FAST_BITS 10
#define LARS_INF_FAST_SIZE (1u << LARS_INF_FAST_BITS)

// shared-memory working set of one warp (39,776 bytes)
struct LarsInflateSmem {
  uint8_t ring[LARS_INF_RING];
  uint32_t inbuf[LARS_INF_INBUF_WORDS];
  uint16_t fast_ll[LARS_INF_FAST_SIZE];   // (code length << 9) | symbol, 0 = longer than the fast bits
  uint16_t fast_d[LARS_INF_FAST_SIZE];
  uint16_t sym_ll[288];                   // symbols in canonical order (by length, then value)
  uint16_t sym_d[32];
  uint16_t cnt_ll[16], cnt_d[16];         // codes per length
  uint16_t work[16];                      // next code / offsets while a table is built; [0] = verdict
  uint16_t codes[288];                    // canonical code of every symbol of the alphabet being built
  uint8_t lens[352];                      // code lengths of the block being set up (32 + 286 + 30, rounded)
  uint32_t adler_part[64];                // per-lane partial sums of the checksum of one 128-byte stretch
};

struct LarsInflateBits {
  uint64_t acc;
  int have;
  uint32_t next_word;      // next word of the stream to enter the bit buffer
  uint64_t used_bits;      // bits handed out so far (to notice reads past the end)
};

LARS_LZW_FN uint32_t lars_inf_bitrev(uint32_t v, int n) {
  uint32_t r = 0;
  for (int i = 0; i < n; ++i) { r = (r << 1) | (v & 1u); v >>= 1; }
  return r;
}

// Builds the decode tables of one alphabet from lens[0, n): returns false for an over-subscribed set of lengths.
// The counting and the code assignment are read-modify-write sequences on shared memory, so ONE lane runs them (the
// lanes of a warp are not guaranteed to move in lockstep); the others wait at the barrier and then share the fill.
LARS_LZW_FN bool lars_inf_build(const uint8_t* lens, int n, uint16_t* fast, uint16_t* sym, uint16_t* cnt, uint16_t* work,
                                uint16_t* codes) {
  LARS_LZW_SYNC();                                  // nobody still decodes with the tables being replaced
  LARS_LZW_FOR_LANES(lane) {
    if (lane == 0) {
      for (int l = 0; l < 16; ++l) cnt[l] = 0;
      for (int s = 0; s < n; ++s) cnt[lens[s]] = (uint16_t)(cnt[lens[s]] + 1);
      int left = 1;
      bool ok = true;
      for (int l = 1; l < 16; ++l) {
        left = (left << 1) - (int)cnt[l];
        if (left < 0) { ok = false; break; }        // more codes of this length than the prefix tree has room for
      }
      // canonical order of the symbols (for codes longer than the fast bits)
      work[1] = 0;
      for (int l = 1; l < 15; ++l) work[l + 1] = (uint16_t)(work[l] + cnt[l]);
      for (int s = 0; s < n; ++s)
        if (lens[s]) { sym[work[lens[s]]] = (uint16_t)s; work[lens[s]] = (uint16_t)(work[lens[s]] + 1); }
      // first code of every length, then the code of every symbol
      uint32_t code = 0;
      for (int l = 1; l < 16; ++l) { code = (code + (l > 1 ? cnt[l - 1] : 0u)) << 1; work[l] = (uint16_t)code; }
      for (int s = 0; s < n; ++s)
        if (lens[s]) { codes[s] = work[lens[s]]; work[lens[s]] = (uint16_t)(work[lens[s]] + 1); }
      work[0] = ok ? 1 : 0;
    }
  }
  LARS_LZW_SYNC();
  if (work[0] == 0) return false;
  { LARS_LZW_FOR_LANES(lane) { for (uint32_t i = (uint32_t)lane; i < LARS_INF_FAST_SIZE; i += 32u) fast[i] = 0; } }
  LARS_LZW_SYNC();
  // the fast table: every slot whose low `len` bits are the bit-reversed code of a symbol
  for (int s = 0; s < n; ++s) {
    const int l = lens[s];
    if (l == 0 || l > LARS_INF_FAST_BITS) continue;
    const uint32_t r = lars_inf_bitrev(codes[s], l), step = 1u << l;
    const uint16_t e = (uint16_t)((l << 9) | s);
    LARS_LZW_FOR_LANES(lane) { for (uint32_t i = r + step * (uint32_t)lane; i < LARS_INF_FAST_SIZE; i += step * 32u) fast[i] = e; }
  }
  LARS_LZW_SYNC();
  return true;
}

// Decodes one zlib stream of n_in bytes into out[0, cap); returns the number of bytes produced, 0 for a corrupt
// stream.  `in - (in & 3)` up to the next multiple of 4 after the stream must be readable.
LARS_LZW_FN uint32_t lars_inflate_warp(const uint8_t* in, uint32_t n_in, uint8_t* out, uint32_t cap, LarsInflateSmem* sm) {
  static const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
  static const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
  static const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
  static const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
  static const uint8_t kClOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

  // ---- input side (as in lars_lzw_decode_warp_v2): aligned words through a shared ring, one segment in flight
  const uint32_t skew = (uint32_t)((uintptr_t)in & 3u);
  const uint32_t* words = reinterpret_cast<const uint32_t*>(in - skew);
  const uint32_t n_words = (skew + n_in + 3u) >> 2;
  uint32_t loaded = 0, fetched = 0;
#if defined(__CUDA_ARCH__)
  uint32_t pending = 0;
#define LARS_INF_PENDING(lane) pending
#else
  uint32_t pending[32];
#define LARS_INF_PENDING(lane) pending[lane]
#endif
#define LARS_INF_FETCH()                                                                              \
  do {                                                                                                \
    LARS_LZW_FOR_LANES(lane) {                                                                        \
      const uint32_t w_ = fetched + (uint32_t)lane;                                                   \
      LARS_INF_PENDING(lane) = w_ < n_words ? LARS_LZW_LOAD(words + w_) : 0u;                         \
    }                                                                                                 \
    fetched += 32u;                                                                                   \
  } while (0)
#define LARS_INF_COMMIT()                                                                             \
  do {                                                                                                \
    LARS_LZW_FOR_LANES(lane) { sm->inbuf[(loaded + (uint32_t)lane) & (LARS_INF_INBUF_WORDS - 1u)] = LARS_INF_PENDING(lane); } \
    loaded += 32u;                                                                                    \
    LARS_LZW_SYNC();                                                                                  \
    LARS_INF_FETCH();                                                                                 \
  } while (0)
  LARS_INF_FETCH();
  LARS_INF_COMMIT();
  LarsInflateBits b;
  b.acc = 0; b.have = 0; b.next_word = 0; b.used_bits = 0;
  // makes at least `n` (<= 32) bits available; words past the stream read as zero and are caught by used_bits
#define LARS_INF_NEED(n)                                                                              \
  do {                                                                                                \
    if (b.have < (int)(n)) {                                                                          \
      while (loaded < n_words && loaded < b.next_word + 64u) LARS_INF_COMMIT();                       \
      const uint32_t w_ = b.next_word < n_words ? sm->inbuf[b.next_word & (LARS_INF_INBUF_WORDS - 1u)] : 0u; \
      b.acc |= (uint64_t)w_ << b.have;                                                                \
      b.have += 32;                                                                                   \
      b.next_word += 1u;                                                                              \
    }                                                                                                 \
  } while (0)
#define LARS_INF_DROP(n) do { b.acc >>= (n); b.have -= (int)(n); b.used_bits += (uint64_t)(n); } while (0)
  // completed stretches of the output ring leave for global memory: up to a 4-byte boundary of the destination
  // byte by byte, then 128 bytes per step as aligned words
#define LARS_INF_FLUSH()                                                                              \
  do {                                                                                                \
    LARS_LZW_SYNC();                                                                                  \
    const uint32_t head_ = (4u - ((out_skew + flushed) & 3u)) & 3u;                                   \
    for (uint32_t i_ = 0; i_ < head_; ++i_) {        /* every lane the same few bytes: checksum */   \
      ad1 += sm->ring[(flushed + i_) & (LARS_INF_RING - 1u)];                                         \
      ad2 += ad1;                                                                                     \
    }                                                                                                 \
    LARS_LZW_FOR_LANES(lane) { if ((uint32_t)lane < head_) out[flushed + lane] = sm->ring[(flushed + lane) & (LARS_INF_RING - 1u)]; } \
    flushed += head_;                                                                                 \
    while (op - flushed >= 128u) {                                                                    \
      LARS_LZW_FOR_LANES(lane) {                                                                      \
        const uint32_t a_ = flushed + 4u * (uint32_t)lane;                                            \
        const uint32_t b0_ = sm->ring[a_ & (LARS_INF_RING - 1u)], b1_ = sm->ring[(a_ + 1u) & (LARS_INF_RING - 1u)],  \
                       b2_ = sm->ring[(a_ + 2u) & (LARS_INF_RING - 1u)], b3_ = sm->ring[(a_ + 3u) & (LARS_INF_RING - 1u)]; \
        *reinterpret_cast<uint32_t*>(out + a_) = b0_ | (b1_ << 8) | (b2_ << 16) | (b3_ << 24);        \
        /* Adler-32 of the stretch: s1 += sum b, s2 += 128 s1 + sum (128 - offset) b */               \
        const uint32_t o_ = 4u * (uint32_t)lane;                                                      \
        sm->adler_part[lane] = b0_ + b1_ + b2_ + b3_;                                                 \
        sm->adler_part[32 + lane] = (128u - o_) * b0_ + (127u - o_) * b1_ + (126u - o_) * b2_ + (125u - o_) * b3_; \
      }                                                                                               \
      LARS_LZW_SYNC();                                                                                \
      {                                                                                               \
        uint32_t sa_ = 0, sb_ = 0;                                                                    \
        for (int l_ = 0; l_ < 32; ++l_) { sa_ += sm->adler_part[l_]; sb_ += sm->adler_part[32 + l_]; } \
        ad2 = (ad2 + 128u * ad1 + sb_) % 65521u;                                                      \
        ad1 = (ad1 + sa_) % 65521u;                                                                   \
      }                                                                                               \
      LARS_LZW_SYNC();                               /* the partial sums are free for the next stretch */ \
      flushed += 128u;                                                                                \
    }                                                                                                 \
    ad1 %= 65521u; ad2 %= 65521u;                                                                     \
  } while (0)
  const uint64_t total_bits = 8ull * n_in;
  LARS_INF_NEED(32);
  LARS_INF_DROP(8u * skew);                       // the bytes in front of the stream inside its first word
  b.used_bits = 0;

  // ---- zlib header
  LARS_INF_NEED(16);
  {
    const uint32_t cmf = (uint32_t)(b.acc & 0xFFu), flg = (uint32_t)((b.acc >> 8) & 0xFFu);
    LARS_INF_DROP(16);
    if ((cmf & 0x0Fu) != 8u || (cmf >> 4) > 7u || ((cmf << 8) | flg) % 31u != 0u || (flg & 0x20u)) return 0;
  }

  uint32_t op = 0, flushed = 0;
  uint32_t ad1 = 1, ad2 = 0;                       // Adler-32 of the bytes that have left the ring
  const uint32_t out_skew = (uint32_t)((uintptr_t)out & 3u);
  bool last = false, fixed_ready = false;
  bool cut = false;                                // the output filled up before the stream ended
  while (!last && !cut) {
    LARS_LZW_SYNC();                               // no lane still decodes with the tables the next block replaces
    LARS_INF_NEED(3);
    last = (b.acc & 1u) != 0;
    const uint32_t type = (uint32_t)((b.acc >> 1) & 3u);
    LARS_INF_DROP(3);
    if (type == 3u) return 0;
    if (type == 0u) {                              // stored: LEN, ~LEN, then LEN bytes from the next byte boundary
      LARS_INF_DROP((uint32_t)b.have & 7u);
      LARS_INF_NEED(32);
      const uint32_t len = (uint32_t)(b.acc & 0xFFFFu), nlen = (uint32_t)((b.acc >> 16) & 0xFFFFu);
      LARS_INF_DROP(32);
      if ((len ^ nlen) != 0xFFFFu) return 0;
      for (uint32_t i = 0; i < len; ++i) {
        if (op >= cap) { cut = true; break; }
        LARS_INF_NEED(8);
        const uint8_t v = (uint8_t)(b.acc & 0xFFu);
        LARS_INF_DROP(8);
        LARS_LZW_FOR_LANES(lane) { if (lane == 0) sm->ring[op & (LARS_INF_RING - 1u)] = v; }
        ++op;
        if (op - flushed >= 512u) LARS_INF_FLUSH();
      }
      if (b.used_bits > total_bits) return 0;
      continue;
    }
    if (type == 1u) {                              // fixed codes (RFC 1951 3.2.6)
      if (!fixed_ready) {
        for (int s = 0; s < 288; ++s) sm->lens[s] = (uint8_t)(s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8);
        if (!lars_inf_build(sm->lens, 288, sm->fast_ll, sm->sym_ll, sm->cnt_ll, sm->work, sm->codes)) return 0;
        for (int s = 0; s < 30; ++s) sm->lens[s] = 5;
        if (!lars_inf_build(sm->lens, 30, sm->fast_d, sm->sym_d, sm->cnt_d, sm->work, sm->codes)) return 0;
        fixed_ready = true;
      }
    } else {                                       // dynamic codes (RFC 1951 3.2.7)
      fixed_ready = false;
      LARS_INF_NEED(14);
      const int hlit = (int)(b.acc & 31u) + 257, hdist = (int)((b.acc >> 5) & 31u) + 1, hclen = (int)((b.acc >> 10) & 15u) + 4;
      LARS_INF_DROP(14);
      if (hlit > 286 || hdist > 30) return 0;
      for (int i = 0; i < 19; ++i) sm->lens[i] = 0;
      for (int i = 0; i < hclen; ++i) {
        LARS_INF_NEED(3);
        sm->lens[kClOrder[i]] = (uint8_t)(b.acc & 7u);
        LARS_INF_DROP(3);
      }
      // the code-length alphabet borrows the distance tables
      if (!lars_inf_build(sm->lens, 19, sm->fast_d, sm->sym_d, sm->cnt_d, sm->work, sm->codes)) return 0;
      int idx = 0;
      uint8_t prev = 0;
      // lengths are collected behind the 19 code-length lengths and moved down afterwards
      while (idx < hlit + hdist) {
        LARS_INF_NEED(7 + 7);
        const uint16_t e = sm->fast_d[b.acc & (LARS_INF_FAST_SIZE - 1u)];   // code-length codes are at most 7 bits
        if (e == 0) return 0;
        LARS_INF_DROP((uint32_t)(e >> 9));
        const int s = e & 511;
        if (s < 16) {
          prev = (uint8_t)s;
          sm->lens[32 + idx++] = prev;
        } else {
          int rep;
          uint8_t v = 0;
          if (s == 16) { if (idx == 0) return 0; v = prev; rep = 3 + (int)(b.acc & 3u); LARS_INF_DROP(2); }
          else if (s == 17) { rep = 3 + (int)(b.acc & 7u); LARS_INF_DROP(3); }
          else { rep = 11 + (int)(b.acc & 127u); LARS_INF_DROP(7); }
          if (idx + rep > hlit + hdist) return 0;
          while (rep--) sm->lens[32 + idx++] = v;
          prev = v;
        }
      }
      if (b.used_bits > total_bits) return 0;
      if (sm->lens[32 + 256] == 0) return 0;       // no end-of-block code
      // the collected lengths are used where they are: literal / length ones first, the distance ones behind them
      if (!lars_inf_build(sm->lens + 32 + hlit, hdist, sm->fast_d, sm->sym_d, sm->cnt_d, sm->work, sm->codes)) return 0;
      if (!lars_inf_build(sm->lens + 32, hlit, sm->fast_ll, sm->sym_ll, sm->cnt_ll, sm->work, sm->codes)) return 0;
    }

    // ---- the symbols of the block
    for (;;) {
      LARS_INF_NEED(15 + 5);
      int sym;
      {
        const uint16_t e = sm->fast_ll[b.acc & (LARS_INF_FAST_SIZE - 1u)];
        if (e) {
          LARS_INF_DROP((uint32_t)(e >> 9));
          sym = e & 511;
        } else {                                   // longer than the fast bits: canonical walk, one bit at a time
          int code = 0, first = 0, index = 0, l = 1;
          sym = -1;
          for (; l < 16; ++l) {
            code |= (int)((b.acc >> (l - 1)) & 1u);
            const int count = sm->cnt_ll[l];
            if (code - count < first) { sym = sm->sym_ll[index + (code - first)]; break; }
            index += count; first += count; first <<= 1; code <<= 1;
          }
          if (sym < 0) return 0;
          LARS_INF_DROP((uint32_t)l);
        }
      }
      if (sym == 256) break;                       // end of block (also when the output is exactly full)
      if (op >= cap) { cut = true; break; }        // more data than the caller has room for: stop here
      if (sym < 256) {
        LARS_LZW_FOR_LANES(lane) { if (lane == 0) sm->ring[op & (LARS_INF_RING - 1u)] = (uint8_t)sym; }
        ++op;
      } else {
        sym -= 257;
        if (sym >= 29) return 0;
        LARS_INF_NEED(5 + 15);
        uint32_t len = kLenBase[sym] + (uint32_t)(b.acc & ((1u << kLenExtra[sym]) - 1u));
        LARS_INF_DROP(kLenExtra[sym]);
        LARS_INF_NEED(15 + 13);
        int ds;
        {
          const uint16_t e = sm->fast_d[b.acc & (LARS_INF_FAST_SIZE - 1u)];
          if (e) {
            LARS_INF_DROP((uint32_t)(e >> 9));
            ds = e & 511;
          } else {
            int code = 0, first = 0, index = 0, l = 1;
            ds = -1;
            for (; l < 16; ++l) {
              code |= (int)((b.acc >> (l - 1)) & 1u);
              const int count = sm->cnt_d[l];
              if (code - count < first) { ds = sm->sym_d[index + (code - first)]; break; }
              index += count; first += count; first <<= 1; code <<= 1;
            }
            if (ds < 0) return 0;
            LARS_INF_DROP((uint32_t)l);
          }
        }
        if (ds >= 30) return 0;
        const uint32_t dist = kDistBase[ds] + (uint32_t)(b.acc & ((1u << kDistExtra[ds]) - 1u));
        LARS_INF_DROP(kDistExtra[ds]);
        if (dist > op) return 0;                   // reaches in front of the output
        const uint32_t keep = len < cap - op ? len : cap - op;
        const uint32_t from = op - dist;
        const bool near = dist <= LARS_INF_RING - 512u;   // else the source was flushed long ago: read it back
        LARS_LZW_SYNC();
        LARS_LZW_FOR_LANES(lane) {
          for (uint32_t i = (uint32_t)lane; i < keep; i += 32u) {
            const uint32_t s = from + (i < dist ? i : i % dist);   // an overlapping match repeats its period
            sm->ring[(op + i) & (LARS_INF_RING - 1u)] = near ? sm->ring[s & (LARS_INF_RING - 1u)] : out[s];
          }
        }
        LARS_LZW_SYNC();                           // nobody still reads the ring when a lane that runs ahead writes on
        op += keep;
        if (keep < len) { cut = true; break; }     // the match did not fit: the output is full
      }
      if (b.used_bits > total_bits) return 0;      // the stream ended inside this block
      if (op - flushed >= 512u) LARS_INF_FLUSH();
    }
  }
  if (b.used_bits > total_bits) return 0;
  LARS_LZW_SYNC();
  LARS_LZW_FOR_LANES(lane) {                       // what is left in the ring, byte by byte
    for (uint32_t a = flushed + (uint32_t)lane; a < op; a += 32u) out[a] = sm->ring[a & (LARS_INF_RING - 1u)];
  }
  for (uint32_t a = flushed; a < op; ++a) {        // ... and its part of the checksum (fewer than 900 bytes)
    ad1 += sm->ring[a & (LARS_INF_RING - 1u)];
    if (ad1 >= 65521u) ad1 -= 65521u;
    ad2 += ad1;
    if (ad2 >= 65521u) ad2 -= 65521u;
  }
  if (!cut) {                                      // the whole stream was decoded: its Adler-32 trailer must agree
    LARS_INF_DROP((uint32_t)b.have & 7u);
    LARS_INF_NEED(32);
    const uint32_t t = (uint32_t)b.acc;
    LARS_INF_DROP(32);
    if (b.used_bits > total_bits) return 0;
    if (lars_lzw_bswap32(t) != ((ad2 << 16) | ad1)) return 0;
  }
#undef LARS_INF_FETCH
#undef LARS_INF_COMMIT
#undef LARS_INF_PENDING
#undef LARS_INF_NEED
#undef LARS_INF_DROP
#undef LARS_INF_FLUSH
  return op;
}
