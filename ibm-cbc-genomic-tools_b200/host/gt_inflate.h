// gt_inflate.h -- a gzip stream reader for the host side's line reader.
//
// Plain gzip (one member, or several one after the other) cannot be cut into pieces the way BGZF can, so the pace of the
// decoder is the pace of the reader.  zlib's inflate delivers 170-200 MB/s of text on the boxes this was measured on; the
// decoder here is built the way the fast ones are (a 64-bit bit buffer refilled a word at a time, Huffman tables that resolve
// a code and its extra-bit count in one lookup, copies by words) and delivers what gzread delivers: the data of every member
// in order; from a file that ends inside a member, everything that inflates before the end; nothing after a fault.  Every
// member's CRC-32 and length are checked against its trailer.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <vector>

namespace gt {

class GzipStream {
 public:
  explicit GzipStream(int fd);                       // reads from the descriptor's current position; does not close it
  GzipStream();                                      // no input yet: Reset() names it
  void Reset(const uint8_t *data, size_t n);         // a new stream: the n bytes at data (they must stay where they are while it is read)
  // up to `want` inflated bytes into dst; fewer only at the end of the stream (0: nothing is left).  After a fault the data in
  // front of it has been handed out and the stream is at its end.
  long Read(void *dst, size_t want);
  bool failed() const { return failed_; }            // the stream ended on a fault (damaged data, a CRC or length that does not match)

 private:
  enum { kHistory = 32768, kChunk = 1 << 20, kOutSlack = 320, kInCap = 2 << 20, kInPad = 64 };
  enum State { S_HEADER, S_BLOCK_HEADER, S_STORED, S_HUFFMAN, S_TRAILER, S_END };
  bool Produce();                                    // more output into the chunk; false: the stream is at its end
  void FillInput();
  bool ParseGzipHeader();                            // false: no (complete) header here
  bool ParseBlockHeader();
  bool BuildTables(const uint8_t *lens, int n_litlen, int n_dist);
  int DecodeHuffman();                               // 0: block finished, 1: output chunk full, 2: input needed, -1: fault
  size_t InputLeft() const;                          // whole bytes not consumed yet (those waiting in the bit buffer included)
  void AlignToByte();
  int fd_;
  const uint8_t *mem_ = nullptr;                     // input from memory instead of fd_ (Reset)
  size_t mem_left_ = 0;
  std::vector<uint8_t> in_;                          // [0, in_end_) valid, kInPad zero bytes behind it once the input has ended
  size_t in_pos_ = 0, in_end_ = 0;
  bool in_eof_ = false;
  uint64_t bitbuf_ = 0;
  int bitcnt_ = 0;
  std::vector<uint8_t> out_;                         // kHistory of history, the chunk, slack for copies that run over
  size_t out_lo_ = kHistory, out_hi_ = kHistory;     // [out_lo_, out_hi_) inflated and not handed out yet
  size_t valid_lo_ = kHistory;                       // where the current member's data begins in out_ (at most 32 KB of it in front of the chunk): no match reaches further back
  size_t summed_ = kHistory;                         // [summed_, out_hi_) has not gone into the member's CRC and length yet
  void SumUp();
  State state_ = S_HEADER;
  bool last_block_ = false, failed_ = false, any_member_ = false;
  size_t stored_left_ = 0;
  uint32_t crc_ = 0, isize_ = 0;
  std::vector<uint32_t> litlen_, dist_;              // decode tables (primary + subtables)
};

}  // namespace gt
