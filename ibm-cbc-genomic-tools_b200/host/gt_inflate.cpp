// gt_inflate.cpp -- see gt_inflate.h.  DEFLATE is RFC 1951, the gzip container RFC 1952.
#include "gt_inflate.h"
#include <errno.h>
#include <string.h>
#include <unistd.h>
#include <zlib.h>                                     // crc32() only

namespace gt {
namespace {

// ---- decode table entries -------------------------------------------------------------------------------------------------------
// bits 0-7: bits of the code that this lookup consumes; bits 8-12: extra bits that follow the code (or, for a subtable
// pointer, the index bits of the subtable); bits 13-15: what it is; bits 16-31: literal / base length / base distance / first
// entry of the subtable
enum : uint32_t { T_LITERAL = 0, T_BASE = 1, T_END = 2, T_SUB = 3, T_INVALID = 7 };
constexpr uint32_t kInvalid = (uint32_t)T_INVALID << 13;
constexpr int kLitlenBits = 11, kDistBits = 8, kPrecodeBits = 7;
inline uint32_t Entry(uint32_t type, uint32_t bits, uint32_t extra, uint32_t value) { return bits | (extra << 8) | (type << 13) | (value << 16); }
inline uint32_t EType(uint32_t e) { return (e >> 13) & 7u; }
inline uint32_t EBits(uint32_t e) { return e & 0xFFu; }
inline uint32_t EExtra(uint32_t e) { return (e >> 8) & 31u; }
inline uint32_t EValue(uint32_t e) { return e >> 16; }

const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
const uint8_t kPrecodeOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

inline uint32_t ReverseBits(uint32_t code, int len) {
  uint32_t r = 0;
  for (int i = 0; i < len; i++) { r = (r << 1) | (code & 1u); code >>= 1; }
  return r;
}

// what symbol `sym` of the alphabet stands for, as a table entry without its bit count
enum Alphabet { A_LITLEN, A_DIST, A_PRECODE };
inline uint32_t SymbolEntry(Alphabet a, int sym) {
  if (a == A_PRECODE) return Entry(T_LITERAL, 0, 0, (uint32_t)sym);
  if (a == A_DIST) return sym < 30 ? Entry(T_BASE, 0, kDistExtra[sym], kDistBase[sym]) : kInvalid;
  if (sym < 256) return Entry(T_LITERAL, 0, 0, (uint32_t)sym);
  if (sym == 256) return Entry(T_END, 0, 0, 0);
  return sym < 286 ? Entry(T_BASE, 0, kLenExtra[sym - 257], kLenBase[sym - 257]) : kInvalid;
}

// Canonical Huffman code of the lengths lens[0..n) (0: symbol unused) as a lookup table of 2^primary entries followed by
// subtables for the codes longer than that.  False: the lengths are over-subscribed.  An incomplete code is accepted (a lone
// distance code is one, RFC 1951 3.2.7); the patterns it leaves unassigned stay T_INVALID and are a fault only if the data uses one.
bool BuildTable(const uint8_t *lens, int n, int primary, Alphabet alphabet, std::vector<uint32_t> *table) {
  int count[16] = {0};
  for (int s = 0; s < n; s++) count[lens[s]]++;
  count[0] = 0;
  long left = 1;
  for (int l = 1; l <= 15; l++) { left = (left << 1) - count[l]; if (left < 0) return false; }
  uint32_t next_code[16];
  uint32_t code = 0;
  for (int l = 1; l <= 15; l++) { code = (code + (uint32_t)count[l - 1]) << 1; next_code[l] = code; }
  const uint32_t pmask = (1u << primary) - 1u;
  table->assign((size_t)1 << primary, kInvalid);
  // index bits of the subtable behind every primary slot that needs one
  std::vector<uint8_t> sub_bits((size_t)1 << primary, 0);
  std::vector<uint32_t> rev((size_t)n, 0);
  {
    uint32_t nc[16];
    memcpy(nc, next_code, sizeof nc);
    for (int s = 0; s < n; s++) {
      const int l = lens[s];
      if (l == 0) continue;
      rev[(size_t)s] = ReverseBits(nc[l]++, l);
      if (l > primary) { uint8_t &b = sub_bits[rev[(size_t)s] & pmask]; if (l - primary > b) b = (uint8_t)(l - primary); }
    }
  }
  for (uint32_t p = 0; p <= pmask; p++)
    if (sub_bits[p]) {
      (*table)[p] = Entry(T_SUB, (uint32_t)primary, sub_bits[p], 0) | ((uint32_t)table->size() << 16);
      table->resize(table->size() + ((size_t)1 << sub_bits[p]), kInvalid);
      if (table->size() > 0xFFFF) return false;
    }
  for (int s = 0; s < n; s++) {
    const int l = lens[s];
    if (l == 0) continue;
    const uint32_t what = SymbolEntry(alphabet, s);
    const uint32_t r = rev[(size_t)s];
    if (l <= primary) {
      for (uint32_t i = r; i <= pmask; i += 1u << l) (*table)[i] = what | (uint32_t)l;
    } else {
      const uint32_t p = r & pmask;
      const uint32_t start = EValue((*table)[p]), bits = sub_bits[p];
      for (uint32_t i = r >> primary; i < (1u << bits); i += 1u << (l - primary)) (*table)[start + i] = what | (uint32_t)(l - primary);
    }
  }
  return true;
}

inline uint64_t Load64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }

}  // namespace

GzipStream::GzipStream(int fd) : fd_(fd) {
  in_.resize(kInCap + kInPad);
  out_.resize(kHistory + kChunk + 258 + kOutSlack);
}

GzipStream::GzipStream() : GzipStream(-1) { in_eof_ = true; state_ = S_END; }

void GzipStream::Reset(const uint8_t *data, size_t n) {
  mem_ = data; mem_left_ = n;
  in_pos_ = in_end_ = 0;
  in_eof_ = false;
  bitbuf_ = 0; bitcnt_ = 0;
  out_lo_ = out_hi_ = valid_lo_ = summed_ = kHistory;
  state_ = S_HEADER;
  last_block_ = failed_ = any_member_ = false;
  stored_left_ = 0;
}

// Moves what is left of the input to the front of the buffer and reads more behind it.
void GzipStream::FillInput() {
  if (in_eof_) return;
  // (the whole bytes waiting in the bit buffer go back to the input first: positions are counted from the buffer's start, and
  // what has been moved out of the buffer could not be handed back later)
  in_pos_ -= (size_t)(bitcnt_ >> 3);
  bitcnt_ &= 7;
  bitbuf_ &= (1ull << bitcnt_) - 1;
  if (in_pos_ > 0) {
    memmove(in_.data(), in_.data() + in_pos_, in_end_ - in_pos_);
    in_end_ -= in_pos_;
    in_pos_ = 0;
  }
  if (fd_ < 0) {                                                        // input from memory
    const size_t n = mem_left_ < (size_t)kInCap - in_end_ ? mem_left_ : (size_t)kInCap - in_end_;
    if (n) memcpy(in_.data() + in_end_, mem_, n);
    mem_ += n; mem_left_ -= n; in_end_ += n;
    if (mem_left_ == 0) { in_eof_ = true; memset(in_.data() + in_end_, 0, kInPad); }
    return;
  }
  while (in_end_ < (size_t)kInCap) {
    const ssize_t got = read(fd_, in_.data() + in_end_, (size_t)kInCap - in_end_);
    if (got < 0 && errno == EINTR) continue;
    if (got <= 0) { in_eof_ = true; break; }
    in_end_ += (size_t)got;
    if (in_end_ >= (size_t)kInCap / 2) break;
  }
  if (in_eof_) memset(in_.data() + in_end_, 0, kInPad);
}

size_t GzipStream::InputLeft() const {
  const size_t consumed = in_pos_ - (size_t)(bitcnt_ >> 3);
  return consumed <= in_end_ ? in_end_ - consumed : 0;
}

// drops the bits up to the next byte boundary of the stream and hands the whole bytes waiting in the bit buffer back to the input
void GzipStream::AlignToByte() {
  const int drop = bitcnt_ & 7;
  bitbuf_ >>= drop;
  bitcnt_ -= drop;
  in_pos_ -= (size_t)(bitcnt_ >> 3);
  bitbuf_ = 0;
  bitcnt_ = 0;
}

// RFC 1952 2.3: magic, CM = 8, FLG, MTIME(4), XFL, OS, then the optional fields FLG announces.  Called at a byte boundary with
// an empty bit buffer.  False: the header is not (all) there.
bool GzipStream::ParseGzipHeader() {
  const uint8_t *p = in_.data() + in_pos_;
  const size_t n = in_end_ - in_pos_;
  if (n < 10 || p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || (p[3] & 0xE0)) return false;
  const unsigned flg = p[3];
  size_t at = 10;
  if (flg & 4) {                                                        // FEXTRA
    if (n < at + 2) return false;
    const size_t xlen = (size_t)p[at] | (size_t)p[at + 1] << 8;
    at += 2 + xlen;
    if (n < at) return false;
  }
  for (unsigned f : {8u, 16u})                                          // FNAME, FCOMMENT: zero-terminated
    if (flg & f) {
      const void *z = at < n ? memchr(p + at, 0, n - at) : nullptr;
      if (z == nullptr) return false;
      at = (size_t)((const uint8_t *)z - p) + 1;
    }
  if (flg & 2) { at += 2; if (n < at) return false; }                   // FHCRC
  in_pos_ += at;
  return true;
}

bool GzipStream::BuildTables(const uint8_t *lens, int n_litlen, int n_dist) {
  return BuildTable(lens, n_litlen, kLitlenBits, A_LITLEN, &litlen_) && BuildTable(lens + n_litlen, n_dist, kDistBits, A_DIST, &dist_);
}

// BFINAL, BTYPE and what the type brings with it (RFC 1951 3.2.3-3.2.7).  The caller has seen to it that the whole header is
// in the buffer (or that the input has ended, in which case the zero padding is read and the overrun is noticed afterwards).
bool GzipStream::ParseBlockHeader() {
  const uint8_t *in = in_.data();
  auto refill = [&] {
    bitbuf_ |= Load64(in + in_pos_) << bitcnt_;
    in_pos_ += (size_t)((63 - bitcnt_) >> 3);
    bitcnt_ |= 56;
  };
  auto take = [&](int nbits) { const uint32_t v = (uint32_t)(bitbuf_ & ((1ull << nbits) - 1)); bitbuf_ >>= nbits; bitcnt_ -= nbits; return v; };
  refill();
  last_block_ = take(1) != 0;
  const uint32_t type = take(2);
  if (type == 0) {
    const int drop = bitcnt_ & 7;
    bitbuf_ >>= drop; bitcnt_ -= drop;
    if (bitcnt_ < 32) refill();
    const uint32_t len = take(16), nlen = take(16);
    if ((len ^ nlen) != 0xFFFFu) return false;
    AlignToByte();
    stored_left_ = len;
    state_ = S_STORED;
    return true;
  }
  if (type == 3) return false;
  uint8_t lens[320];
  int n_litlen = 288, n_dist = 32;
  if (type == 1) {
    for (int s = 0; s < 288; s++) lens[s] = s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8;
    for (int s = 0; s < 32; s++) lens[288 + s] = 5;
  } else {
    n_litlen = (int)take(5) + 257;
    n_dist = (int)take(5) + 1;
    const int n_pre = (int)take(4) + 4;
    if (n_litlen > 286 || n_dist > 30) return false;
    uint8_t pre_lens[19] = {0};
    for (int i = 0; i < n_pre; i++) {
      if (bitcnt_ < 3) refill();
      pre_lens[kPrecodeOrder[i]] = (uint8_t)take(3);
    }
    std::vector<uint32_t> pre;
    if (!BuildTable(pre_lens, 19, kPrecodeBits, A_PRECODE, &pre)) return false;
    int have = 0;
    while (have < n_litlen + n_dist) {
      if (bitcnt_ < 7 + 7) refill();
      const uint32_t e = pre[bitbuf_ & ((1u << kPrecodeBits) - 1)];
      if (EType(e) != T_LITERAL) return false;
      take((int)EBits(e));
      const uint32_t sym = EValue(e);
      if (sym < 16) { lens[have++] = (uint8_t)sym; continue; }
      int rep;
      uint8_t what = 0;
      if (sym == 16) { if (have == 0) return false; what = lens[have - 1]; rep = 3 + (int)take(2); }
      else if (sym == 17) rep = 3 + (int)take(3);
      else rep = 11 + (int)take(7);
      if (have + rep > n_litlen + n_dist) return false;
      while (rep--) lens[have++] = what;
    }
    if (lens[256] == 0) return false;                                   // no end-of-block code
  }
  if (!BuildTables(lens, n_litlen, n_dist)) return false;
  state_ = S_HUFFMAN;
  return true;
}

// The symbols of a Huffman block, until the block ends (0), the output chunk is full (1), the input buffer runs low (2; at the
// end of the input: the data is used up, i.e. the stream is truncated) or the data is faulty (-1).  Stops between symbols only.
int GzipStream::DecodeHuffman() {
  uint8_t *const out = out_.data();
  const uint8_t *const in = in_.data();
  const uint32_t *const lt = litlen_.data(), *const dt = dist_.data();
  size_t op = out_hi_, ip = in_pos_;
  uint64_t bb = bitbuf_;
  int bc = bitcnt_;
  const size_t out_limit = (size_t)kHistory + kChunk;
  const size_t window_lo = valid_lo_;
  // (not at the end of the input: stop while a whole symbol's worth of bytes and a refill's reach are still in the buffer; at
  // the end: run into the zero padding and notice afterwards that a symbol has used bits that are not there)
  const size_t in_limit = in_eof_ ? in_end_ + 8 : (in_end_ > 24 ? in_end_ - 24 : 0);
  int rc;
  // ---- the fast loop: while a symbol's worth of input (three refills' reach) is in the buffer, nothing can run off its end
  {
    const size_t in_fast = in_end_ > 32 ? in_end_ - 32 : 0;
    constexpr uint32_t lmask = (1u << kLitlenBits) - 1, dmask = (1u << kDistBits) - 1;
    rc = 3;                                                             // (3: the careful loop below takes over)
    while (op < out_limit && ip <= in_fast) {
      bb |= Load64(in + ip) << bc;
      ip += (size_t)((63 - bc) >> 3);
      bc |= 56;
      uint32_t e = lt[bb & lmask];
      // up to three literals per refill (15 bits each at most)
      if (EType(e) == T_LITERAL) {
        bb >>= EBits(e); bc -= (int)EBits(e);
        out[op++] = (uint8_t)EValue(e);
        e = lt[bb & lmask];
        if (EType(e) == T_LITERAL) {
          bb >>= EBits(e); bc -= (int)EBits(e);
          out[op++] = (uint8_t)EValue(e);
          e = lt[bb & lmask];
          if (EType(e) == T_LITERAL) {
            bb >>= EBits(e); bc -= (int)EBits(e);
            out[op++] = (uint8_t)EValue(e);
            continue;
          }
        }
        if (bc < 48) {
          bb |= Load64(in + ip) << bc;
          ip += (size_t)((63 - bc) >> 3);
          bc |= 56;
        }
      }
      if (EType(e) == T_SUB) {
        bb >>= kLitlenBits; bc -= kLitlenBits;
        e = lt[EValue(e) + (uint32_t)(bb & ((1u << EExtra(e)) - 1))];
      }
      bb >>= EBits(e); bc -= (int)EBits(e);
      const uint32_t type = EType(e);
      if (type == T_LITERAL) { out[op++] = (uint8_t)EValue(e); continue; }
      if (type == T_END) { rc = 0; break; }
      if (type != T_BASE) { rc = -1; break; }
      const uint32_t ex = EExtra(e);
      const uint32_t len = EValue(e) + (uint32_t)(bb & ((1u << ex) - 1));
      bb >>= ex; bc -= (int)ex;
      uint32_t d = dt[bb & dmask];
      if (EType(d) == T_SUB) {
        bb >>= kDistBits; bc -= kDistBits;
        d = dt[EValue(d) + (uint32_t)(bb & ((1u << EExtra(d)) - 1))];
      }
      if (EType(d) != T_BASE) { rc = -1; break; }
      bb >>= EBits(d); bc -= (int)EBits(d);
      const uint32_t dx = EExtra(d);
      const size_t dist = EValue(d) + (size_t)(bb & ((1u << dx) - 1));
      bb >>= dx; bc -= (int)dx;
      if (dist > op - window_lo) { rc = -1; break; }
      uint8_t *dst = out + op;
      const uint8_t *src = dst - dist;
      op += len;
      if (dist >= 8) {
        uint8_t *const end = dst + len;
        do { memcpy(dst, src, 8); dst += 8; src += 8; } while (dst < end);
      } else if (dist == 1) {
        memset(dst, *src, len);
      } else {
        for (uint32_t i = 0; i < len; i++) dst[i] = src[i];
      }
    }
    if (rc != 3) { out_hi_ = op; in_pos_ = ip; bitbuf_ = bb; bitcnt_ = bc; return rc; }
    if (op >= out_limit) { out_hi_ = op; in_pos_ = ip; bitbuf_ = bb; bitcnt_ = bc; return 1; }
    if (!in_eof_) { out_hi_ = op; in_pos_ = ip; bitbuf_ = bb; bitcnt_ = bc; return 2; }
  }
  // ---- the careful loop: the last bytes of the input
  for (;;) {
    if (op >= out_limit) { rc = 1; break; }
    if (ip > in_limit) { rc = 2; break; }
    if (bc < 48) {
      bb |= Load64(in + ip) << bc;
      ip += (size_t)((63 - bc) >> 3);
      bc |= 56;
    }
    // (the symbol is decoded on copies, so that one that turns out to lie beyond the end of the input leaves no trace)
    uint64_t b = bb;
    int c = bc;
    uint32_t e = lt[b & ((1u << kLitlenBits) - 1)];
    if (EType(e) == T_SUB) {
      b >>= kLitlenBits; c -= kLitlenBits;
      e = lt[EValue(e) + (uint32_t)(b & ((1u << EExtra(e)) - 1))];
    }
    b >>= EBits(e); c -= (int)EBits(e);
    const uint32_t type = EType(e);
    if (type == T_LITERAL) {
      if (in_eof_ && (long)8 * ((long)in_end_ - (long)ip) + c < 0) { rc = 2; break; }
      out[op++] = (uint8_t)EValue(e);
      bb = b; bc = c;
      continue;
    }
    if (type == T_END) {
      if (in_eof_ && (long)8 * ((long)in_end_ - (long)ip) + c < 0) { rc = 2; break; }
      bb = b; bc = c;
      rc = 0;
      break;
    }
    if (type != T_BASE) { rc = -1; break; }
    const uint32_t ex = EExtra(e);
    const uint32_t len = EValue(e) + (uint32_t)(b & ((1u << ex) - 1));
    b >>= ex; c -= (int)ex;
    uint32_t d = dt[b & ((1u << kDistBits) - 1)];
    if (EType(d) == T_SUB) {
      b >>= kDistBits; c -= kDistBits;
      d = dt[EValue(d) + (uint32_t)(b & ((1u << EExtra(d)) - 1))];
    }
    if (EType(d) != T_BASE) { rc = -1; break; }
    b >>= EBits(d); c -= (int)EBits(d);
    const uint32_t dx = EExtra(d);
    const size_t dist = EValue(d) + (size_t)(b & ((1u << dx) - 1));
    b >>= dx; c -= (int)dx;
    if (in_eof_ && (long)8 * ((long)in_end_ - (long)ip) + c < 0) { rc = 2; break; }
    if (dist > op - window_lo) { rc = -1; break; }
    bb = b; bc = c;
    uint8_t *dst = out + op;
    const uint8_t *src = dst - dist;
    op += len;
    if (dist >= 8) {
      uint8_t *const end = dst + len;
      do { memcpy(dst, src, 8); dst += 8; src += 8; } while (dst < end);
    } else {
      for (uint32_t i = 0; i < len; i++) dst[i] = src[i];
    }
  }
  out_hi_ = op; in_pos_ = ip; bitbuf_ = bb; bitcnt_ = bc;
  return rc;
}

void GzipStream::SumUp() {
  crc_ = (uint32_t)crc32(crc_, out_.data() + summed_, (unsigned)(out_hi_ - summed_));
  isize_ += (uint32_t)(out_hi_ - summed_);
  summed_ = out_hi_;
}

// Inflates until the output chunk is full or the stream cannot go on.  False: nothing new and nothing more to come.
bool GzipStream::Produce() {
  // the window: the last 32 KB (of this member) move in front of the chunk
  {
    const size_t h = out_hi_ - valid_lo_ < (size_t)kHistory ? out_hi_ - valid_lo_ : (size_t)kHistory;
    if (out_hi_ != (size_t)kHistory) memmove(out_.data() + kHistory - h, out_.data() + out_hi_ - h, h);
    valid_lo_ = (size_t)kHistory - h;
    out_lo_ = out_hi_ = summed_ = kHistory;
  }
  const size_t out_limit = (size_t)kHistory + kChunk;
  auto fault = [&] { failed_ = true; state_ = S_END; };
  while (state_ != S_END && out_hi_ < out_limit) {
    switch (state_) {
      case S_HEADER: {
        if (!in_eof_ && in_end_ - in_pos_ < (size_t)kInCap / 4) FillInput();
        if (in_end_ == in_pos_) { state_ = S_END; break; }              // the stream ends between members
        if (!ParseGzipHeader()) {
          // behind a member: bytes that are no gzip header are ignored, as zlib ignores them; a header cut short ends the stream
          if (!any_member_) failed_ = true;
          state_ = S_END;
          break;
        }
        crc_ = (uint32_t)crc32(0L, Z_NULL, 0);
        isize_ = 0;
        valid_lo_ = summed_ = out_hi_;
        state_ = S_BLOCK_HEADER;
        break;
      }
      case S_BLOCK_HEADER: {
        if (!in_eof_ && InputLeft() < 1024) FillInput();
        if (!ParseBlockHeader()) { fault(); break; }
        if (in_pos_ - (size_t)(bitcnt_ >> 3) > in_end_) state_ = S_END;  // the header ran past the end of the input: truncated
        break;
      }
      case S_STORED: {
        if (stored_left_ == 0) { state_ = last_block_ ? S_TRAILER : S_BLOCK_HEADER; break; }
        if (in_pos_ == in_end_) {
          if (in_eof_) state_ = S_END; else FillInput();
          break;
        }
        size_t n = stored_left_;
        if (n > in_end_ - in_pos_) n = in_end_ - in_pos_;
        if (n > out_limit - out_hi_) n = out_limit - out_hi_;
        memcpy(out_.data() + out_hi_, in_.data() + in_pos_, n);
        out_hi_ += n; in_pos_ += n; stored_left_ -= n;
        break;
      }
      case S_HUFFMAN: {
        const int rc = DecodeHuffman();
        if (rc == 0) state_ = last_block_ ? S_TRAILER : S_BLOCK_HEADER;
        else if (rc == 2) { if (in_eof_) state_ = S_END; else FillInput(); }
        else if (rc == -1) fault();
        break;
      }
      case S_TRAILER: {
        AlignToByte();
        if (in_end_ - in_pos_ < 8 && !in_eof_) FillInput();
        if (in_end_ - in_pos_ < 8) { state_ = S_END; break; }            // no trailer: the file ends here
        SumUp();
        const uint8_t *t = in_.data() + in_pos_;
        const uint32_t want_crc = (uint32_t)t[0] | (uint32_t)t[1] << 8 | (uint32_t)t[2] << 16 | (uint32_t)t[3] << 24;
        const uint32_t want_len = (uint32_t)t[4] | (uint32_t)t[5] << 8 | (uint32_t)t[6] << 16 | (uint32_t)t[7] << 24;
        in_pos_ += 8;
        any_member_ = true;
        state_ = S_HEADER;
        if (want_crc != crc_ || want_len != isize_) fault();
        break;
      }
      case S_END: break;
    }
  }
  if (state_ == S_BLOCK_HEADER || state_ == S_STORED || state_ == S_HUFFMAN || state_ == S_TRAILER) SumUp();   // a member that goes on
  return out_hi_ > out_lo_;
}

long GzipStream::Read(void *dst, size_t want) {
  size_t done = 0;
  while (done < want) {
    if (out_lo_ == out_hi_) {
      if (state_ == S_END) break;
      if (!Produce()) { if (state_ == S_END) break; continue; }
    }
    size_t n = out_hi_ - out_lo_;
    if (n > want - done) n = want - done;
    memcpy((char *)dst + done, out_.data() + out_lo_, n);
    out_lo_ += n;
    done += n;
  }
  return (long)done;
}

}  // namespace gt
