// gt_host.cpp -- see gt_host.h
#include "gt_host.h"
#include <errno.h>
#include <fcntl.h>
#include <limits.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>
#include <algorithm>
#include <atomic>
#include <functional>
#include <iostream>

namespace gt {

void die(const std::string &msg) {
  fprintf(stderr, "\n");
  fprintf(stderr, "Error: %s\n", msg.c_str());
  exit(1);
}

void die_line(long line, const std::string &msg) {
  fprintf(stderr, "\n");
  fprintf(stderr, "Error: Line %ld: %s\n", line, msg.c_str());
  exit(1);
}

// ---------------------------------------------------------------------------------------------
// CmdLine
// ---------------------------------------------------------------------------------------------
void CmdLine::AddOperation(const std::string &op, const std::string &usage, const std::string &description, const std::string &details) {
  ops_[op] = Operation{usage, description, details};
}

void CmdLine::AddOption(const char *opt, bool *ptr, bool def, const char *d) { Option o{opt, d, 'b', ptr}; o.def_b = def; options_.push_back(o); }
void CmdLine::AddOption(const char *opt, char *ptr, char def, const char *d) { Option o{opt, d, 'c', ptr}; o.def_c = def; options_.push_back(o); }
void CmdLine::AddOption(const char *opt, long *ptr, long def, const char *d) { Option o{opt, d, 'l', ptr}; o.def_l = def; options_.push_back(o); }
void CmdLine::AddOption(const char *opt, unsigned long *ptr, unsigned long def, const char *d) { Option o{opt, d, 'u', ptr}; o.def_u = def; options_.push_back(o); }
void CmdLine::AddOption(const char *opt, double *ptr, double def, const char *d) { Option o{opt, d, 'd', ptr}; o.def_d = def; options_.push_back(o); }
void CmdLine::AddOption(const char *opt, const char **ptr, const char *def, const char *d) { Option o{opt, d, 's', ptr}; o.def_s = def; options_.push_back(o); }

int CmdLine::Read(char **argv, int argc) {
  // defaults first (CmdLine::Init, core.cpp:2438)
  for (auto &o : options_) {
    switch (o.type) {
      case 'b': *(bool *)o.ptr = o.def_b; break;
      case 'c': *(char *)o.ptr = o.def_c; break;
      case 'l': *(long *)o.ptr = o.def_l; break;
      case 'u': *(unsigned long *)o.ptr = o.def_u; break;
      case 'd': *(double *)o.ptr = o.def_d; break;
      case 's': o.cur_s = o.def_s; *(const char **)o.ptr = o.cur_s.c_str(); break;
    }
  }
  int i = 1;
  while (i < argc) {
    if (argv[i][0] != '-') return i;                                   // first non-option ends the options (core.cpp:2428)
    Option *found = nullptr;
    for (auto &o : options_) if (o.opt == argv[i]) found = &o;        // a later registration of the same name wins, like the map
    if (!found) { fprintf(stderr, "Error: unknown option '%s'!\n", argv[i]); exit(1); }
    if (found->type == 'b') { *(bool *)found->ptr = !*(bool *)found->ptr; i++; continue; }   // flags TOGGLE (core.cpp:2212)
    if (i + 1 >= argc) { fprintf(stderr, "Error: could not set option '%s'!\n", found->opt.c_str()); exit(1); }
    const char *v = argv[i + 1];
    switch (found->type) {
      case 'c': *(char *)found->ptr = v[0]; break;
      case 'l': *(long *)found->ptr = atol(v); break;
      case 'u': *(unsigned long *)found->ptr = (unsigned long)atol(v); break;
      case 'd': *(double *)found->ptr = atof(v); break;
      case 's': found->cur_s = v; *(const char **)found->ptr = found->cur_s.c_str(); break;
    }
    i += 2;
  }
  return argc;
}

void CmdLine::Print() {
  for (auto &o : options_) {
    printf("  %-25s %-80s ", o.opt.c_str(), o.description.c_str());
    switch (o.type) {
      case 'b': printf("[%s]", *(bool *)o.ptr ? "true" : "false"); break;
      case 'c': printf("[%c]", *(char *)o.ptr); break;
      case 'l': printf("[%ld]", *(long *)o.ptr); break;
      case 'u': printf("[%lu]", *(unsigned long *)o.ptr); break;
      case 'd': printf("[%.6e]", *(double *)o.ptr); break;
      case 's': printf("[%s]", *(const char **)o.ptr); break;
    }
    printf("\n");
  }
}

void CmdLine::OperationSummary(const std::string &usage, const std::string &description) {
  std::cout << '\n' << "USAGE: \n" << "  " << program_ << " " << usage << '\n' << '\n';
  if (version_ != "") std::cout << "VERSION: \n" << "  " << version_ << '\n' << '\n';
  std::cout << "DESCRIPTION: \n" << "  " << description << '\n' << '\n' << "OPERATION: \n";
  std::cout.flush();
  for (auto &kv : ops_) printf("  %-15s %s\n", kv.first.c_str(), kv.second.description.c_str());
  fflush(stdout);
  std::cout << '\n';
}

void CmdLine::OperationUsage() {
  auto it = ops_.find(current_);
  if (it == ops_.end()) { fprintf(stderr, "Error: [CmdLine::AddOperation] operation not found!\n"); exit(1); }
  std::cout << '\n' << "USAGE: \n" << "  " << program_ << " " << it->first << " " << it->second.usage << '\n' << '\n';
  std::cout << "DESCRIPTION: \n" << "  " << it->second.description << '\n' << '\n';
  if (it->second.details != "") std::cout << "DETAILS: \n" << "  " << it->second.details << '\n' << '\n';
  std::cout << "OPTIONS: \n";
  std::cout.flush();
  Print();
  fflush(stdout);
  std::cout << '\n';
}

// ---------------------------------------------------------------------------------------------
// LineReader
// ---------------------------------------------------------------------------------------------
static const size_t kBlockBytes = 32u << 20;

// [offset, offset + want) of a regular file into dst; returns the bytes read (short only at the end of the file or on an error)
static size_t PreadAll(int fd, char *dst, size_t want, off_t offset) {
  size_t have = 0;
  while (have < want) {
    const ssize_t got = pread(fd, dst + have, want - have, offset + (off_t)have);
    if (got < 0 && errno == EINTR) continue;
    if (got <= 0) break;
    have += (size_t)got;
  }
  return have;
}

// BGZF -- what bgzip and samtools write (SAM specification, section 4.1): a series of gzip members of at most 64 KB of data each,
// every one carrying its own compressed size in a 'BC' extra field.  zlib's gzread walks them one after the other on one thread
// (170 MB/s of text); here the members of a stretch of the file are found by their size fields and inflated side by side, each
// by this build's decoder (gt_inflate.h).  What
// the reader hands out is what gzread would: the data of every complete member in file order and, from a member the file ends
// in, whatever inflates -- then the end of the input.  A complete member that does not inflate to what its trailer promises ends
// the input in front of it.
struct LineReader::Bgzf {
  int fd;
  off_t offset = 0;
  int threads;
  bool file_end = false, stream_end = false;
  bool damaged = false;                                                 // the stream ended at a complete member that did not inflate to what its trailer says
  std::vector<unsigned char> comp;                                      // compressed bytes not consumed yet
  std::vector<char> out;                                                // the inflated stretch
  size_t out_pos = 0;
  struct Member { size_t begin, size, cdata, at; uint32_t isize; bool ok; };   // the member in comp, its deflate data, its place in out
  std::vector<GzipStream *> decoder;                                    // one per inflating thread
  ~Bgzf() { for (GzipStream *d : decoder) delete d; }
  static constexpr size_t kStretch = 16u << 20;
  explicit Bgzf(int f) : fd(f) {
    const char *env = getenv("GT_INFLATE_THREADS");
    threads = env ? atoi(env) : (int)std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
    if (threads < 1) threads = 1;
  }
  static uint32_t U16(const unsigned char *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8; }
  static uint32_t U32(const unsigned char *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
  // the size of the member whose header starts at p (n bytes available), 0 if the header is not all there, -1 if it is no BGZF header
  static long MemberSize(const unsigned char *p, size_t n, size_t *cdata_at) {
    if (n < 12) return 0;
    if (p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || !(p[3] & 4)) return -1;
    const size_t xlen = U16(p + 10);
    if (n < 12 + xlen) return 0;
    for (size_t x = 12; x + 4 <= 12 + xlen;) {
      const size_t slen = U16(p + x + 2);
      if (p[x] == 'B' && p[x + 1] == 'C' && slen == 2 && x + 6 <= 12 + xlen) {
        const long bsize = (long)U16(p + x + 4) + 1;
        if ((size_t)bsize < 12 + xlen + 8) return -1;
        *cdata_at = 12 + xlen;
        return bsize;
      }
      x += 4 + slen;
    }
    return -1;
  }
  static bool IsBgzf(int fd) {                                          // does the file begin with a BGZF member?
    unsigned char h[512];
    const ssize_t got = pread(fd, h, sizeof h, 0);
    size_t at;
    return got >= 18 && MemberSize(h, (size_t)got, &at) > 0;
  }
  bool Fill() {                                                         // the next stretch; false at the end of the stream
    out.clear(); out_pos = 0;
    while (!stream_end && out.empty()) {
      while (!file_end && comp.size() < kStretch) {
        const size_t have = comp.size();
        comp.resize(have + kStretch);
        const size_t got = PreadAll(fd, (char *)comp.data() + have, kStretch, offset);
        offset += (off_t)got;
        comp.resize(have + got);
        if (got < kStretch) file_end = true;
      }
      // the complete members of the buffer
      std::vector<Member> mem;
      size_t at = 0, total = 0;
      bool broken = false;                                              // what is left at `at` is no complete member and never will be
      bool not_bgzf = false;
      for (;;) {
        size_t cdata = 0;
        const long bsize = MemberSize(comp.data() + at, comp.size() - at, &cdata);
        if (bsize < 0) { broken = not_bgzf = true; break; }             // (bytes that are no BGZF member: the stream ends here, with a warning)
        if (bsize == 0 || (size_t)bsize > comp.size() - at) { broken = file_end && comp.size() > at; break; }
        const unsigned char *m = comp.data() + at;
        if (U32(m + bsize - 4) > (1u << 16)) { broken = true; break; }  // (no BGZF member holds more than 64 KB)
        mem.push_back({at, (size_t)bsize, at + cdata, total, U32(m + bsize - 4), false});
        total += mem.back().isize;
        at += (size_t)bsize;
      }
      out.resize(total);
      std::atomic<size_t> next{0};
      // (every member is a gzip stream of its own: header, data, CRC-32 and length, all of which the decoder checks)
      auto work = [&](int t) {
        GzipStream &gz = *decoder[(size_t)t];
        for (size_t i; (i = next.fetch_add(1)) < mem.size();) {
          Member &b = mem[i];
          gz.Reset(comp.data() + b.begin, b.size);
          b.ok = gz.Read(out.data() + b.at, (size_t)b.isize) == (long)b.isize && !gz.failed();   // (a member is far smaller than the decoder's chunk: its trailer has been checked by now)
        }
      };
      const int n_threads = (int)std::min<size_t>((size_t)threads, std::max<size_t>(1, mem.size() / 8));
      while ((int)decoder.size() < n_threads) decoder.push_back(new GzipStream());
      std::vector<std::thread> th;
      for (int t = 1; t < n_threads; t++) th.emplace_back(work, t);
      work(0);
      for (auto &t : th) t.join();
      // a member that did not inflate as its header promised ends the stream behind whatever it did yield (as gzread would)
      size_t good = 0;
      while (good < mem.size() && mem[good].ok) good++;
      if (good < mem.size() || broken) {
        stream_end = true;
        damaged = good < mem.size() || not_bgzf;
        // (a complete member that is damaged yields nothing: its bytes up to the fault would be noise; a member the file ends in
        // yields what inflates, which is what zlib hands out before it reports the unexpected end of the file)
        size_t keep = good < mem.size() ? mem[good].at : total;
        const unsigned char *src = nullptr;
        size_t n = 0;
        if (good == mem.size()) {                                       // the fragment behind the last complete member
          size_t cdata = 0;
          if (MemberSize(comp.data() + at, comp.size() - at, &cdata) != -1 && comp.size() - at > cdata && cdata > 0) { src = comp.data() + at + cdata; n = comp.size() - at - cdata; }
        }
        out.resize(keep);
        if (src != nullptr && n > 0) {
          z_stream zs;
          memset(&zs, 0, sizeof zs);
          if (inflateInit2(&zs, -15) == Z_OK) {
            out.resize(keep + 65536);
            zs.next_in = const_cast<unsigned char *>(src); zs.avail_in = (unsigned)std::min<size_t>(n, 1u << 20);
            zs.next_out = (unsigned char *)out.data() + keep; zs.avail_out = 65536;
            inflate(&zs, Z_SYNC_FLUSH);
            out.resize(keep + (65536 - zs.avail_out));
            inflateEnd(&zs);
          }
        }
      } else {
        comp.erase(comp.begin(), comp.begin() + (long)at);
        if (file_end && comp.empty()) stream_end = true;
      }
    }
    return !out.empty();
  }
  long Read(void *dst, size_t want) {
    size_t done = 0;
    while (done < want) {
      if (out_pos == out.size() && !Fill()) break;
      const size_t n = std::min(want - done, out.size() - out_pos);
      memcpy((char *)dst + done, out.data() + out_pos, n);
      out_pos += n; done += n;
    }
    return (long)done;
  }
};

// BAM (core.cpp:371-430, FileBufferBAM): the reference hands the header text out line by line and then every alignment as the
// SAM line samtools' bam_format1_core writes for it (samtools/bam.c:256-340), and parses those lines as SAM.  Same here: the
// producer thread inflates the BGZF stream (a series of gzip members: zlib's gzread walks them), decodes the records and fills
// the line blocks with that text.  A truncated file ends the input where the last complete record ends, as samread() < 0 does.
struct LineReader::BamDecoder {
  std::function<long(void *, size_t)> read_inflated;
  std::vector<std::string> ref;                                         // reference sequence names
  std::string out;                                                      // text not handed out yet
  size_t out_pos = 0;
  bool eof = false;
  explicit BamDecoder(std::function<long(void *, size_t)> r) : read_inflated(std::move(r)) {}
  bool ReadExact(void *dst, size_t n) {
    size_t have = 0;
    while (have < n) {
      const long got = read_inflated((char *)dst + have, n - have);
      if (got <= 0) return false;
      have += (size_t)got;
    }
    return true;
  }
  static int32_t I32(const unsigned char *p) { return (int32_t)((uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24); }
  static uint32_t U16(const unsigned char *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8; }
  // after the magic: l_text, text, n_ref, (l_name, name, l_ref) x n_ref
  void ReadHeader() {
    unsigned char w[4];
    if (!ReadExact(w, 4)) { eof = true; return; }
    const int32_t l_text = I32(w);
    std::string text((size_t)std::max(l_text, 0), '\0');
    if (l_text > 0 && !ReadExact(&text[0], (size_t)l_text)) { eof = true; return; }
    text.resize(strlen(text.c_str()));                                  // header->text is a C string to the reference
    // the reference walks the text with GetNextToken(.., '\n') while anything is left (core.cpp:421-425, :613-625): blanks in
    // front of a line are skipped, the last line needs no newline
    for (const char *p = text.c_str(); *p;) {
      while (*p == ' ') p++;
      const size_t k = strcspn(p, "\n");
      out.append(p, k);
      out.push_back('\n');
      p += k;
      if (*p) p++;
    }
    if (!ReadExact(w, 4)) { eof = true; return; }
    const int32_t n_ref = I32(w);
    for (int32_t i = 0; i < n_ref; i++) {
      if (!ReadExact(w, 4)) { eof = true; return; }
      const int32_t l_name = I32(w);
      std::string name((size_t)std::max(l_name, 0), '\0');
      if (l_name > 0 && !ReadExact(&name[0], (size_t)l_name)) { eof = true; return; }
      name.resize(strlen(name.c_str()));
      if (!ReadExact(w, 4)) { eof = true; return; }                       // l_ref
      ref.push_back(name);
    }
  }
  // one alignment (the `block` bytes behind its length field) -> one SAM line appended to *out (bam_format1_core with decimal
  // flags); false: the record is malformed (the stream ends in front of it, as samread() < 0 ends the reference's)
  static void PutInt(std::string *out, long long v) {
    char b[24];
    char *e = b + sizeof b, *p = e;
    unsigned long long u = v < 0 ? 0ull - (unsigned long long)v : (unsigned long long)v;
    do { *--p = (char)('0' + u % 10); u /= 10; } while (u);
    if (v < 0) *--p = '-';
    out->append(p, (size_t)(e - p));
  }
  static bool FormatRecord(const unsigned char *r, int32_t block, const std::vector<std::string> &ref, std::string *outp) {
    std::string &out = *outp;
    const int32_t tid = I32(r), pos = I32(r + 4);
    const uint32_t l_qname = r[8], mapq = r[9], n_cigar = U16(r + 12), flag = U16(r + 14);
    const int32_t l_seq = I32(r + 16), mtid = I32(r + 20), mpos = I32(r + 24), isize = I32(r + 28);
    const size_t fixed = 32 + (size_t)l_qname + 4 * (size_t)n_cigar + ((size_t)std::max(l_seq, 0) + 1) / 2 + (size_t)std::max(l_seq, 0);
    if (l_seq < 0 || l_qname == 0 || fixed > (size_t)block) return false;
    const unsigned char *qname = r + 32, *cigar = qname + l_qname, *seq = cigar + 4 * n_cigar, *qual = seq + (l_seq + 1) / 2, *aux = qual + l_seq;
    const unsigned char *end = r + block;
    auto name_of = [&](int32_t id) { if (id >= 0 && (size_t)id < ref.size()) out += ref[(size_t)id]; else PutInt(&out, id); };
    out.append((const char *)qname, l_qname - 1); out.push_back('\t');
    PutInt(&out, flag); out.push_back('\t');
    if (tid < 0) out += "*\t"; else { name_of(tid); out.push_back('\t'); }
    PutInt(&out, (long long)pos + 1); out.push_back('\t'); PutInt(&out, mapq); out.push_back('\t');
    if (n_cigar == 0) out.push_back('*');
    else for (uint32_t i = 0; i < n_cigar; i++) {
      const uint32_t c = (uint32_t)I32(cigar + 4 * i);
      PutInt(&out, c >> 4); out.push_back("MIDNSHP=XB??????"[c & 15u]);
    }
    out.push_back('\t');
    if (mtid < 0) out += "*\t"; else if (mtid == tid) out += "=\t"; else { name_of(mtid); out.push_back('\t'); }
    PutInt(&out, (long long)mpos + 1); out.push_back('\t'); PutInt(&out, isize); out.push_back('\t');
    if (l_seq) {
      const size_t at = out.size();
      out.resize(at + (size_t)l_seq);
      char *d = &out[at];
      for (int32_t i = 0; i < l_seq; i++) d[i] = "=ACMGRSVTWYHKDBN"[(seq[i >> 1] >> ((~i & 1) << 2)) & 15];
      out.push_back('\t');
      if (qual[0] == 0xff) out.push_back('*');
      else {
        const size_t q_at = out.size();
        out.resize(q_at + (size_t)l_seq);
        char *q = &out[q_at];
        for (int32_t i = 0; i < l_seq; i++) q[i] = (char)(qual[i] + 33);
      }
    } else out += "*\t*";
    char tmp[64];
    for (const unsigned char *s = aux; s + 3 <= end;) {
      out.push_back('\t'); out.append((const char *)s, 2); out.push_back(':');
      const unsigned char type = s[2];
      s += 3;
      auto fits = [&](size_t n) { return (size_t)(end - s) >= n; };
      if (type == 'A' && fits(1)) { out += "A:"; out.push_back((char)*s); s += 1; }
      else if (type == 'C' && fits(1)) { out += "i:"; PutInt(&out, *s); s += 1; }
      else if (type == 'c' && fits(1)) { out += "i:"; PutInt(&out, (int8_t)*s); s += 1; }
      else if (type == 'S' && fits(2)) { out += "i:"; PutInt(&out, U16(s)); s += 2; }
      else if (type == 's' && fits(2)) { out += "i:"; PutInt(&out, (int16_t)U16(s)); s += 2; }
      else if (type == 'I' && fits(4)) { out += "i:"; PutInt(&out, (uint32_t)I32(s)); s += 4; }
      else if (type == 'i' && fits(4)) { out += "i:"; PutInt(&out, I32(s)); s += 4; }
      else if (type == 'f' && fits(4)) { float f; memcpy(&f, s, 4); out.append(tmp, (size_t)snprintf(tmp, sizeof tmp, "f:%g", f)); s += 4; }
      else if (type == 'd' && fits(8)) { double d; memcpy(&d, s, 8); out.append(tmp, (size_t)snprintf(tmp, sizeof tmp, "d:%lg", d)); s += 8; }
      else if (type == 'Z' || type == 'H') { out.push_back((char)type); out.push_back(':'); while (s < end && *s) out.push_back((char)*s++); if (s < end) s++; }
      else if (type == 'B' && fits(5)) {
        const unsigned char sub = *s++;
        const int32_t n = I32(s);
        s += 4;
        out += "B:"; out.push_back((char)sub);
        for (int32_t i = 0; i < n && s < end; i++) {
          out.push_back(',');
          if (sub == 'c' && fits(1)) { PutInt(&out, (int8_t)*s); s += 1; }
          else if (sub == 'C' && fits(1)) { PutInt(&out, *s); s += 1; }
          else if (sub == 's' && fits(2)) { PutInt(&out, (int16_t)U16(s)); s += 2; }
          else if (sub == 'S' && fits(2)) { PutInt(&out, U16(s)); s += 2; }
          else if (sub == 'i' && fits(4)) { PutInt(&out, I32(s)); s += 4; }
          else if (sub == 'I' && fits(4)) { PutInt(&out, (uint32_t)I32(s)); s += 4; }
          else if (sub == 'f' && fits(4)) { float f; memcpy(&f, s, 4); out.append(tmp, (size_t)snprintf(tmp, sizeof tmp, "%g", f)); s += 4; }
          else { s = end; }
        }
      }
      else break;                                                       // (a type samtools does not know ends the line's tags)
    }
    out.push_back('\n');
    return true;
  }
  // The next stretch of the record stream: the inflated bytes are cut into records on this thread (a walk over the length
  // fields), the records are spelt as SAM lines on several (each thread a run of consecutive records, the runs joined in order).
  // The stream ends in front of the first record that is incomplete or malformed.
  std::vector<unsigned char> raw;                                       // inflated bytes: whole records in front, the beginning of one behind them
  std::vector<std::string> text;                                        // the SAM lines of each formatting thread's run of records
  size_t raw_len = 0;
  static constexpr size_t kStretch = 8u << 20;
  size_t target = kStretch;                                             // bytes to have in hand before cutting: a stretch, or all of a longer record
  void NextStretch() {
    if (raw.size() < target) raw.resize(target);
    bool input_end = false;
    while (raw_len < target) {
      const long got = read_inflated(raw.data() + raw_len, target - raw_len);
      if (got <= 0) { input_end = true; break; }
      raw_len += (size_t)got;
    }
    target = kStretch;
    std::vector<std::pair<size_t, int32_t>> recs;                       // (where the record's body begins, its length)
    size_t at = 0;
    bool bad = false;
    for (;;) {
      if (raw_len - at < 4) break;
      const int32_t block = I32(raw.data() + at);
      if (block < 32) { bad = true; break; }
      if ((size_t)block > raw_len - at - 4) {
        target = std::max(kStretch, (size_t)block + 4);                // (a record longer than a stretch: the next round waits for all of it)
        break;
      }
      recs.emplace_back(at + 4, block);
      at += 4 + (size_t)block;
    }
    const int threads = (int)std::min<size_t>(std::min(8u, std::max(1u, std::thread::hardware_concurrency())), std::max<size_t>(1, recs.size() / 2048));
    if (text.size() < (size_t)threads) text.resize((size_t)threads);   // (kept between stretches: their memory is reused)
    for (auto &t : text) t.clear();
    std::vector<size_t> first_bad((size_t)threads, SIZE_MAX);
    auto work = [&](int t) {
      const size_t lo = recs.size() * (size_t)t / (size_t)threads, hi = recs.size() * (size_t)(t + 1) / (size_t)threads;
      text[(size_t)t].reserve((hi - lo) * 200);
      for (size_t k = lo; k < hi; k++)
        if (!FormatRecord(raw.data() + recs[k].first, recs[k].second, ref, &text[(size_t)t])) { first_bad[(size_t)t] = k; break; }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < threads; t++) th.emplace_back(work, t);
    work(0);
    for (auto &t : th) t.join();
    for (int t = 0; t < threads; t++) {
      out += text[(size_t)t];
      if (first_bad[(size_t)t] != SIZE_MAX) { eof = true; return; }
    }
    if (bad) { eof = true; return; }
    memmove(raw.data(), raw.data() + at, raw_len - at);
    raw_len -= at;
    if (input_end) eof = true;                                          // (what is left is the beginning of a record that never ends)
  }
  long Fill(char *dst, size_t want) {
    while (!eof && out.size() - out_pos < want) {
      if (out_pos > 0) { out.erase(0, out_pos); out_pos = 0; }
      NextStretch();
    }
    const size_t n = std::min(want, out.size() - out_pos);
    memcpy(dst, out.data() + out_pos, n);
    out_pos += n;
    if (out_pos == out.size()) { out.clear(); out_pos = 0; }
    return (long)n;
  }
};

LineReader::LineReader(const char *path) {
  if (path == nullptr) fd_ = 0;
  else {
    fd_ = open(path, O_RDONLY);
    if (fd_ < 0) { fprintf(stderr, "[CreateFileBuffer] Error: cannot open file '%s'!\n", path); exit(1); }
    unsigned char magic[2] = {0, 0};
    const ssize_t got = pread(fd_, magic, 2, 0);                       // gzip sniff by magic (core.cpp:1757-1775)
    if (got == 2 && magic[0] == 0x1f && magic[1] == 0x8b) {
      struct stat fst;
      const char *no_bgzf = getenv("GT_NO_BGZF");                        // (tests: the same file through zlib's gzread)
      const char *use_zlib = getenv("GT_ZLIB");                          // (tests: zlib's gzread instead of either decoder of this build)
      const bool own = !(use_zlib && use_zlib[0] == '1');
      if (own && !(no_bgzf && no_bgzf[0] == '1') && fstat(fd_, &fst) == 0 && S_ISREG(fst.st_mode) && Bgzf::IsBgzf(fd_)) bgzf_ = new Bgzf(fd_);
      else if (own) gzs_ = new GzipStream(fd_);
      gz_ = gzdopen(bgzf_ || gzs_ ? dup(fd_) : fd_, "rb");               // (with a decoder of this build the handle only marks the stream as gzip)
      if (gz_ == nullptr) { fprintf(stderr, "[CreateFileBuffer] Error: cannot open file '%s'!\n", path); exit(1); }
      gzbuffer(gz_, 1 << 20);
      const long got4 = ReadInflated(prefix_, 4);                        // BAM or gzipped text (GetFileType, core.cpp:1764-1772)
      prefix_len_ = got4 > 0 ? (int)got4 : 0;
      if (prefix_len_ == 4 && memcmp(prefix_, "BAM\1", 4) == 0) {
        prefix_len_ = 0;
        bam_ = new BamDecoder([this](void *dst, size_t n) { return ReadInflated(dst, n); });
        bam_->ReadHeader();
      }
    }
  }
  struct stat st;
  if (!gz_ && fstat(fd_, &st) == 0 && S_ISREG(st.st_mode)) {
    const off_t at = lseek(fd_, 0, SEEK_CUR);
    if (at >= 0) { regular_ = true; offset_ = at; }
  }
  producer_ = std::thread(&LineReader::Produce, this);
}

LineReader::~LineReader() {
  {
    std::lock_guard<std::mutex> lk(mu_);
    quit_ = true;
  }
  cv_.notify_all();
  producer_.join();
  delete bam_;
  if (gz_) gzclose(gz_);
  if (bgzf_ || gzs_) { delete bgzf_; delete gzs_; close(fd_); }
  else if (!gz_ && fd_ > 0) close(fd_);
  for (auto &b : block_) free(b.data);
}

long LineReader::ReadInflated(void *dst, size_t want) {
  // (a stream that ends on a fault -- damaged data, a CRC or a length that does not match -- ends the input there, as a failing
  // gzread ends the reference's; it is not passed over in silence)
  auto warn = [this] {
    if (!fault_warned_) fprintf(stderr, "Warning: the gzip data is damaged (or its CRC / length does not match); the input ends in front of the fault!\n");
    fault_warned_ = true;
  };
  if (bgzf_) { const long got = bgzf_->Read(dst, want); if ((size_t)got < want && bgzf_->damaged) warn(); return got; }
  if (gzs_) { const long got = gzs_->Read(dst, want); if ((size_t)got < want && gzs_->failed()) warn(); return got; }
  return (long)gzread(gz_, dst, (unsigned)std::min<size_t>(want, 1u << 30));
}

long LineReader::ReadSome(char *dst, size_t want) {
  if (bam_) return bam_->Fill(dst, want);
  if (prefix_pos_ < prefix_len_) {                                       // the bytes the constructor looked at
    const size_t n = std::min(want, (size_t)(prefix_len_ - prefix_pos_));
    memcpy(dst, prefix_ + prefix_pos_, n);
    prefix_pos_ += (int)n;
    return (long)n;
  }
  if (gz_) return ReadInflated(dst, want);
  if (regular_) {
    // a regular file: one read() copies out of the page cache at a few GB/s, which the parsing threads outrun -- large requests
    // are split over a few threads
    const size_t slice = 8u << 20;
    const int parts = (int)std::min<size_t>(4, want / slice);
    if (parts <= 1) {
      const size_t got = PreadAll(fd_, dst, want, offset_);
      offset_ += (off_t)got;
      return (long)got;
    }
    size_t got[4] = {0, 0, 0, 0};
    const size_t each = want / (size_t)parts;
    std::vector<std::thread> th;
    for (int i = 1; i < parts; i++)
      th.emplace_back([&, i] { got[i] = PreadAll(fd_, dst + each * (size_t)i, i + 1 == parts ? want - each * (size_t)i : each, offset_ + (off_t)(each * (size_t)i)); });
    got[0] = PreadAll(fd_, dst, each, offset_);
    for (auto &t : th) t.join();
    size_t total = 0;                                                  // the contiguous prefix that arrived (a short part ends it: end of file)
    for (int i = 0; i < parts; i++) {
      total += got[i];
      if (got[i] < (i + 1 == parts ? want - each * (size_t)i : each)) break;
    }
    offset_ += (off_t)total;
    return (long)total;
  }
  for (;;) {
    const ssize_t got = read(fd_, dst, want);
    if (got < 0 && errno == EINTR) continue;
    return (long)got;
  }
}

void LineReader::Produce() {
  std::vector<char> carry;                                             // the incomplete line at the end of the previous block
  bool eof = false;
  for (int tail = 0; !eof; tail = (tail + 1) % kBlocks) {
    {
      std::unique_lock<std::mutex> lk(mu_);
      cv_.wait(lk, [&] { return quit_ || ready_ < kBlocks; });
      if (quit_) return;
    }
    Block &b = block_[tail];
    // the first block is small, so that a short file (or the format sniffing of a long one) is not held up by a large read
    const size_t want = tail == 0 && b.data == nullptr ? (1u << 20) : kBlockBytes;
    if (b.cap < carry.size() + want) {
      b.cap = carry.size() + want;
      b.data = (char *)realloc(b.data, b.cap + 16);                     // (16 spare bytes: the parsers read whole words)
      if (b.data == nullptr) { fprintf(stderr, "Error: out of memory!\n"); exit(1); }
    }
    if (!carry.empty()) memcpy(b.data, carry.data(), carry.size());
    size_t have = carry.size();
    b.len = 0;
    for (;;) {
      if (have == b.cap) {                                             // a line longer than the block
        b.cap *= 2;
        b.data = (char *)realloc(b.data, b.cap + 16);
        if (b.data == nullptr) { fprintf(stderr, "Error: out of memory!\n"); exit(1); }
      }
      const long got = ReadSome(b.data + have, b.cap - have);
      if (got <= 0) { eof = true; break; }
      have += (size_t)got;
      if (have < b.cap) continue;                                      // short reads (pipe, gzip): keep filling the block
      const char *nl = (const char *)memrchr(b.data, '\n', have);
      if (nl != nullptr) { b.len = (size_t)(nl - b.data) + 1; break; }
    }
    if (eof) {                                                         // an unterminated last line is dropped (core.cpp:243)
      const char *nl = have ? (const char *)memrchr(b.data, '\n', have) : nullptr;
      b.len = nl ? (size_t)(nl - b.data) + 1 : 0;
    }
    carry.assign(b.data + b.len, b.data + have);
    std::lock_guard<std::mutex> lk(mu_);
    if (b.len > 0) ready_++;
    if (eof) done_ = true;
    cv_.notify_all();
  }
}

bool LineReader::Acquire() {
  std::unique_lock<std::mutex> lk(mu_);
  if (cur_ >= 0) { ready_--; cur_ = -1; cv_.notify_all(); }
  cv_.wait(lk, [&] { return ready_ > 0 || done_; });
  if (ready_ == 0) return false;
  cur_ = head_;
  head_ = (head_ + 1) % kBlocks;
  pos_ = 0;
  return true;
}

char *LineReader::Next() {
  if (cur_ < 0 || pos_ >= block_[cur_].len) {
    if (!Acquire()) return nullptr;
  }
  Block &b = block_[cur_];
  char *line = b.data + pos_;
  char *nl = (char *)memchr(line, '\n', b.len - pos_);                 // a block ends with '\n'
  *nl = 0;
  pos_ = (size_t)(nl - b.data) + 1;
  line_no_++;
  return line;
}

bool LineReader::NextRun(char **begin, char **end) {
  if (cur_ < 0 || pos_ >= block_[cur_].len) {
    if (!Acquire()) return false;
  }
  Block &b = block_[cur_];
  *begin = b.data + pos_;
  *end = b.data + b.len;
  pos_ = b.len;
  return true;
}

// ---------------------------------------------------------------------------------------------
// tokenizer helpers with the reference's semantics (core.cpp:577-625): leading blanks are skipped,
// a token ends at the delimiter; numbers go through atol.
// ---------------------------------------------------------------------------------------------
static int CountTokens(const char *s, char delim) {
  if (s == nullptr) return 0;
  int k = 0, n = 0;
  while (s[k] == ' ') k++;
  for (;;) {
    if (s[k] == 0) return n;
    while (s[k] != 0 && s[k] != delim) k++;
    if (s[k] == delim) k++;
    n++;
    while (s[k] == ' ') k++;
    if (s[k] == 0) return n;
  }
}

static char *NextToken(char **p, char delim) {
  char *b = *p;
  while (*b == ' ') b++;
  int k = 0;
  while (b[k] != 0 && b[k] != delim) k++;
  if (b[k] == 0) *p = b + k;
  else { b[k] = 0; *p = b + k + 1; }
  return b;
}

// atol for the common case (blanks, sign, up to 18 digits); anything longer goes to the library so that overflow behaves alike
static inline long FastAtol(const char *s) {
  const char *p = s;
  while (*p == ' ' || (unsigned)(*p - 9) < 5u) p++;
  const bool neg = *p == '-';
  if (*p == '-' || *p == '+') p++;
  unsigned long v = 0;
  int digits = 0;
  while ((unsigned)(*p - '0') < 10u) { v = v * 10 + (unsigned)(*p - '0'); p++; digits++; }
  if (digits > 18) return atol(s);
  return neg ? -(long)v : (long)v;
}

// '1','+' -> '+'; '-1','-' -> '-'; '.' -> '+'; else 0 (fatal for the caller)
static inline char StrandOf(const char *t) {
  if (t[0] != 0 && t[1] == 0) {
    if (t[0] == '+' || t[0] == '1' || t[0] == '.') return '+';
    if (t[0] == '-') return '-';
    return 0;
  }
  if (t[0] == '-' && t[1] == '1' && t[2] == 0) return '-';
  return 0;
}

char ProcessStrand(const char *t) {
  const char s = StrandOf(t);
  if (s) return s;
  std::cerr << "Error: invalid strand '" << t << "'!\n";
  exit(1);
}

int32_t ChromTable::Get(const char *chrom) {
  std::lock_guard<std::mutex> lk(mu_);
  auto it = id.find(chrom);
  if (it != id.end()) return it->second;
  int32_t v = (int32_t)name.size();
  id.emplace(chrom, v);
  name.push_back(chrom);
  return v;
}

void RegionBatch::Clear() {
  chrom.clear(); start.clear(); stop.clear(); strand.clear(); weight.clear(); label.clear();
  offset.assign(1, 0);
  first_line = 0;
  multi = false;
}

bool RegionWellFormed(const RegionBatch &b, int64_t k) {
  for (int64_t i = b.offset[k] + 1; i < b.offset[k + 1]; i++) {
    if (b.chrom[i] != b.chrom[b.offset[k]] || b.strand[i] != b.strand[b.offset[k]]) return false;
    if (b.start[i] < b.start[i - 1] || b.start[i] <= b.stop[i - 1]) return false;
  }
  return true;
}

static double NowSeconds() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

PhaseTimer::PhaseTimer() : on_(getenv("GT_TIMING") != nullptr), last_(NowSeconds()), t0_(last_) {}
void PhaseTimer::Mark(const char *phase) {
  if (!on_) return;
  const double now = NowSeconds();
  marks_.emplace_back(phase, now - last_);
  last_ = now;
}
PhaseTimer::~PhaseTimer() {
  if (!on_) return;
  fprintf(stderr, "[gt timing]");
  for (auto &m : marks_) fprintf(stderr, " %s=%.3fs", m.first.c_str(), m.second);
  fprintf(stderr, " total=%.3fs\n", NowSeconds() - t0_);
}

bool SortChecker::Accept(const std::string &c, char s, long st) {
  bool ok = true;
  if (have) {
    const int cmp = strcmp(c.c_str(), chrom.c_str());                 // IsBefore, genomic_intervals.cpp:396-401
    if (cmp != 0) ok = !(cmp < 0);
    else if (by_strand && s != strand) ok = !(s < strand);
    else ok = !(st < start);
  }
  have = true; chrom = c; strand = s; start = st;
  return ok;
}

// ---------------------------------------------------------------------------------------------
// RegionReader
// ---------------------------------------------------------------------------------------------
// Chromosome names repeat: a thread remembers the ones it has met and asks the shared table only for new ones.
struct RegionReader::ChromCache {
  ChromTable *table;
  std::string last;
  int32_t last_id = -1;
  std::unordered_map<std::string, int32_t> seen;
  // names of up to 8 bytes (chr1 ... chrUn_x): the bytes themselves are the key of a small open-addressing table
  static const int kSlots = 256;
  uint64_t key[kSlots];
  int32_t val[kSlots];
  int used = 0;
  explicit ChromCache(ChromTable *t) : table(t) { memset(key, 0, sizeof key); }
  int32_t Get(const char *chrom) {
    if (last_id >= 0 && strcmp(chrom, last.c_str()) == 0) return last_id;
    last = chrom;
    auto it = seen.find(last);
    if (it != seen.end()) return last_id = it->second;
    last_id = table->Get(chrom);
    seen.emplace(last, last_id);
    return last_id;
  }
  // name = [p, p + len), 1 <= len <= 8, not NUL-terminated
  int32_t GetShort(const char *p, size_t len) {
    uint64_t k = 0;
    memcpy(&k, p, len);                                                // len < 8 leaves high zero bytes; a name holds no NUL, so keys are unique
    return GetKey(k, p, len);
  }
  int32_t GetKey(uint64_t k, const char *p, size_t len) {
    uint32_t h = (uint32_t)((k * 0x9E3779B97F4A7C15ull) >> 56);
    for (;;) {
      if (key[h] == k) return val[h];
      if (key[h] == 0) break;
      h = (h + 1) & (kSlots - 1);
    }
    char name[9];
    memcpy(name, p, len);
    name[len] = 0;
    const int32_t id = table->Get(name);
    if (used < kSlots / 2) { key[h] = k; val[h] = id; used++; }
    return id;
  }
};

RegionReader::RegionReader(const char *path, ChromTable *chroms, bool keep_labels, long max_label_value)
    : reader_(path), chroms_(chroms), keep_labels_(keep_labels), max_label_value_(max_label_value) {
  const char *env = getenv("GT_PARSE_THREADS");
  threads_ = env ? atoi(env) : (int)std::min(32u, std::max(1u, std::thread::hardware_concurrency()));
  if (threads_ < 1) threads_ = 1;
  // header skipping and format sniffing (genomic_intervals.cpp:3713-3759)
  char *next = reader_.Next();
  auto is_track = [](const char *s) { return strncmp(s, "browser ", 8) == 0 || strncmp(s, "track ", 6) == 0; };
  auto settle = [&] {
    fmt_ = format_ == "BED" ? F_BED : format_ == "REG" ? F_REG : format_ == "GFF" ? F_GFF : format_ == "SAM" ? F_SAM
           : format_ == "SEQ" ? F_SEQ : format_ == "EMPTY" ? F_EMPTY : F_NONE;
  };
  if (next == nullptr) { format_ = "EMPTY"; settle(); return; }
  auto skip = [&] { header_.append(next); header_.push_back('\n'); next = reader_.Next(); };
  if (is_track(next)) { while (next && is_track(next)) skip(); }
  else if (next[0] == '@') { format_ = "SAM"; while (next && next[0] == '@') skip(); }
  else if (next[0] == '#' && next[1] == '#') { format_ = "GFF"; while (next && next[0] == '#' && next[1] == '#') skip(); }
  if (next == nullptr) { format_ = "EMPTY"; settle(); return; }
  pending_ = next;
  if (format_ != "") { settle(); return; }
  if (next[0] == '>') { format_ = "SEQ"; settle(); return; }
  const long n_tokens = CountTokens(next, '\t');
  if (n_tokens == 1) format_ = "BED";                                 // space-separated BED lands here
  if (n_tokens == 2) format_ = "REG";
  if (n_tokens >= 3 && n_tokens <= 6) format_ = "BED";
  else if (n_tokens >= 6) {
    std::string copy(next);
    char *p = &copy[0], *t = nullptr;
    for (int i = 1; i <= 6; i++) t = NextToken(&p, '\t');
    if (strchr(t, '+') != nullptr || strchr(t, '-') != nullptr) format_ = "BED";
    else if (n_tokens >= 11) format_ = "SAM";
    else if (n_tokens >= 8 && n_tokens <= 10) format_ = "GFF";
  }
  settle();
}

static bool LineError(ParseError *err, const char *msg) { err->with_line = true; err->raw = false; err->message = msg; return false; }
static bool StrandError(ParseError *err, const char *token) {
  err->with_line = false; err->raw = true; err->message = std::string("Error: invalid strand '") + token + "'!\n";
  return false;
}

bool RegionReader::Push(RegionBatch *out, ChromCache *cache, const char *chrom, char strand, long start, long stop, ParseError *err) const {
  if (start < INT32_MIN || start > INT32_MAX || stop < INT32_MIN || stop > INT32_MAX)
    return LineError(err, "coordinate does not fit in 32 bits (not supported by the GPU engine)!");
  out->chrom.push_back(cache->Get(chrom));
  out->strand.push_back((int8_t)strand);
  out->start.push_back((int32_t)start);
  out->stop.push_back((int32_t)stop);
  return true;
}

bool RegionReader::ParseLine(char *inp, RegionBatch *out, ChromCache *cache, ParseError *err) const {
  const char *label = "_";
  const size_t first_interval = out->chrom.size();
  if (fmt_ == F_BED) {                                               // GenomicRegionBED::Read, genomic_intervals.cpp:2157-2182
    const char sep = strchr(inp, '\t') == nullptr ? ' ' : '\t';
    const int n_tokens = CountTokens(inp, sep);
    if (n_tokens < 3) return LineError(err, "number of tokens should be at least 3 for BED format!");
    const char *chromosome = NextToken(&inp, sep);
    const long start = FastAtol(NextToken(&inp, sep)) + 1;
    const long stop = FastAtol(NextToken(&inp, sep));
    char strand = '+';
    if (n_tokens != 3) label = NextToken(&inp, sep);
    if (n_tokens >= 5) NextToken(&inp, sep);                          // score
    if (n_tokens >= 6) {
      const char *t = NextToken(&inp, sep);
      if ((strand = StrandOf(t)) == 0) return StrandError(err, t);
    }
    if (n_tokens >= 8) { NextToken(&inp, sep); NextToken(&inp, sep); }  // thickStart, thickEnd
    if (n_tokens >= 9) NextToken(&inp, sep);                          // itemRgb
    if (n_tokens != 12) { if (!Push(out, cache, chromosome, strand, start, stop, err)) return false; }
    else {
      const long n_intervals = atol(NextToken(&inp, sep));
      char *sizes = NextToken(&inp, sep);
      char *starts = NextToken(&inp, sep);
      std::vector<long> bsize((size_t)std::max(n_intervals, 0L)), bstart((size_t)std::max(n_intervals, 0L));
      for (long k = 0; k < n_intervals; k++) bsize[k] = atol(NextToken(&sizes, ','));
      for (long k = 0; k < n_intervals; k++) bstart[k] = atol(NextToken(&starts, ','));
      for (long k = 0; k < n_intervals; k++) {
        const long s = start + bstart[k];
        if (!Push(out, cache, chromosome, strand, s, bsize[k] + s - 1, err)) return false;
      }
    }
  } else if (fmt_ == F_REG) {                                        // GenomicRegion::Read, genomic_intervals.cpp:805-838
    label = NextToken(&inp, '\t');
    if (strchr(inp, ',') == nullptr) {
      const int n_tokens = CountTokens(inp, ' ');
      if (n_tokens < 4 || n_tokens % 4 != 0) return LineError(err, "invalid number of tokens!");
      for (int k = 0; k < n_tokens / 4; k++) {
        const char *chromosome = NextToken(&inp, ' ');
        const char *t = NextToken(&inp, ' ');
        const char strand = StrandOf(t);
        if (strand == 0) return StrandError(err, t);
        const long start = FastAtol(NextToken(&inp, ' '));
        const long stop = FastAtol(NextToken(&inp, ' '));
        if (!Push(out, cache, chromosome, strand, start, stop, err)) return false;
      }
    } else {
      if (CountTokens(inp, ' ') != 4) return LineError(err, "invalid number of tokens in compact format!");
      const char *chromosome = NextToken(&inp, ' ');
      const char *t = NextToken(&inp, ' ');
      const char strand = StrandOf(t);
      if (strand == 0) return StrandError(err, t);
      char *starts = NextToken(&inp, ' ');
      char *stops = NextToken(&inp, ' ');
      const int n_intervals = CountTokens(starts, ',');
      if (CountTokens(stops, ',') != n_intervals) return LineError(err, "number of starts/stops should be equal");
      for (int k = 0; k < n_intervals; k++) {
        const long start = atol(NextToken(&starts, ','));
        const long stop = atol(NextToken(&stops, ','));
        if (!Push(out, cache, chromosome, strand, start, stop, err)) return false;
      }
    }
  } else if (fmt_ == F_GFF) {                                        // GenomicRegionGFF::Read, genomic_intervals.cpp:3501-3517
    const int n_tokens = CountTokens(inp, '\t');
    if (n_tokens < 8 || n_tokens > 10) return LineError(err, "wrong number of tokens for GFF format!");
    const char *seqname = NextToken(&inp, '\t');
    NextToken(&inp, '\t'); NextToken(&inp, '\t');                      // source, feature
    const long start = FastAtol(NextToken(&inp, '\t'));
    const long end = FastAtol(NextToken(&inp, '\t'));
    NextToken(&inp, '\t');                                             // score
    const char strand = NextToken(&inp, '\t')[0];                      // raw character, NOT normalised (:3512)
    NextToken(&inp, '\t');                                             // frame
    if (n_tokens != 8) label = NextToken(&inp, '\t');
    if (!Push(out, cache, seqname, strand, start, end, err)) return false;
  } else if (fmt_ == F_SAM) {                                        // GenomicRegionSAM::Read, genomic_intervals.cpp:2771-2813
    const int n_tokens = CountTokens(inp, '\t');
    if (n_tokens < 11) return LineError(err, "number of tokens should be at least 11 for SAM format!");
    label = NextToken(&inp, '\t');
    const unsigned long flag = (unsigned long)FastAtol(NextToken(&inp, '\t'));
    const char *rname = NextToken(&inp, '\t');
    long start = FastAtol(NextToken(&inp, '\t'));                       // POS: 1-based leftmost mapping position
    NextToken(&inp, '\t');                                              // MAPQ
    const char *cigar = NextToken(&inp, '\t');
    NextToken(&inp, '\t'); NextToken(&inp, '\t'); NextToken(&inp, '\t');  // RNEXT, PNEXT, TLEN
    const char *seq = NextToken(&inp, '\t');
    const char strand = (flag & 0x10ul) ? '-' : '+';                    // CalcStrandFromFlag, :2870-2873
    const long seq_len = (long)strlen(seq);
    const bool star = strcmp(cigar, "*") == 0;                          // "*" stands for <length of SEQ>M (:2791)
    // One pass over the CIGAR (TokenizeCIGAR :2910-2918 + the loop of Read :2801-2809): operations "MD=X" consume the
    // reference, 'N' closes a block and skips; the fragment length is what "MIS=X" add up to (:3021-3026).  '=' is in the
    // reference's reference-consuming set but not in its tokenizer's alphabet "MIDNSHP-X": it is fatal there, and here.
    long reference_len = 0, fragment_len = 0;
    const size_t first = out->chrom.size();
    const char *c = cigar;
    for (bool first_trip = true; star ? first_trip : *c != 0; first_trip = false) {
      long len = 0;
      char type;
      if (star) { len = seq_len; type = 'M'; }
      else {
        const char *d = c;
        while (*c >= '0' && *c <= '9') c++;
        if (c - d > 18) len = atol(std::string(d, c).c_str());
        else for (; d != c; d++) len = len * 10 + (*d - '0');
        type = *c;
        if (type == 0 || strchr("MIDNSHP-X", type) == nullptr) {
          err->with_line = true; err->raw = false;
          err->message = std::string("unknown CIGAR operation type '") + (type ? std::string(1, type) : std::string()) + "'!";
          out->chrom.resize(first); out->start.resize(first); out->stop.resize(first); out->strand.resize(first);
          return false;
        }
        c++;
      }
      if (type == 'M' || type == 'I' || type == 'S' || type == 'X') fragment_len += len;
      if (type != 'N') { if (type == 'M' || type == 'D' || type == 'X') reference_len += len; }
      else {
        if (!Push(out, cache, rname, strand, start, start + reference_len - 1, err)) return false;
        start = start + reference_len + len;
        reference_len = 0;
      }
    }
    if (strcmp(seq, "*") != 0 && seq_len != fragment_len) {
      err->with_line = true; err->raw = false;
      err->message = std::string("length of aligned fragment does not match CIGAR string: \n") + "  LABEL = " + label + "\n" +
                     "  CIGAR = " + (star ? std::to_string(seq_len) + "M" : std::string(cigar)) + "\n" + "  length(SEQ) = " + std::to_string(seq_len) + "\n";
      out->chrom.resize(first); out->start.resize(first); out->stop.resize(first); out->strand.resize(first);
      return false;
    }
    if (reference_len > 0 && !Push(out, cache, rname, strand, start, start + reference_len - 1, err)) return false;
  } else if (fmt_ == F_SEQ) {
    err->with_line = false; err->raw = false;
    err->message = "input format " + format_ + " is not supported by this build (BED, REG, GFF and SAM are)!\n";
    return false;
  } else {
    err->with_line = false; err->raw = false; err->message = "unsupported input format!\n";
    return false;
  }
  if (max_label_value_ > 1) {                                          // GetLabelValue, genomic_intervals.cpp:1081-1085
    const long w = std::min(max_label_value_, atol(label));
    if (w < INT32_MIN || w > INT32_MAX) return LineError(err, "label value does not fit in 32 bits (not supported by the GPU engine)!");
    out->weight.push_back((int32_t)w);
  }
  if (out->chrom.size() - first_interval != 1) out->multi = true;
  out->offset.push_back((int64_t)out->chrom.size());
  if (keep_labels_) out->label.emplace_back(label);
  return true;
}

// eight bytes at p as one word (the blocks of LineReader are allocated with 16 spare bytes, so a word that begins inside a line
// may run past the block's last '\n')
static inline uint64_t Load8(const char *p) { uint64_t w; memcpy(&w, p, 8); return w; }
// high bit of every byte of w that equals c (exact up to and including the first match, which is all the callers look at)
static inline uint64_t EqMask(uint64_t w, unsigned char c) {
  const uint64_t x = w ^ (0x0101010101010101ull * c);
  return (x - 0x0101010101010101ull) & ~x & 0x8080808080808080ull;
}
// The decimal number at *pp: 1 to 18 digits.  Eight digits at a time: the digit bytes are told from the rest in one word, moved
// to the word's top (leading zeros below them) and folded pairwise.  *pp is left behind the last digit; false: no digit, or too many.
static inline bool ParseDigits(const char **pp, unsigned long *value) {
  const char *p = *pp;
  const uint64_t t = Load8(p) ^ 0x3030303030303030ull;                  // digits -> 0..9
  const uint64_t other = ((t + 0x7676767676767676ull) | t) & 0x8080808080808080ull;   // bytes that are not digits
  const int n = other ? (int)(__builtin_ctzll(other) >> 3) : 8;
  if (n == 0) return false;
  uint64_t v = n == 8 ? t : (t & ((1ull << (8 * n)) - 1)) << (8 * (8 - n));
  v = v * 10 + (v >> 8);
  v = (((v & 0x000000FF000000FFull) * 0x000F424000000064ull) + (((v >> 16) & 0x000000FF000000FFull) * 0x0000271000000001ull)) >> 32;
  p += n;
  if (n == 8) {
    int more = 0;
    while ((unsigned)(*p - '0') < 10u) { v = v * 10 + (unsigned)(*p++ - '0'); if (++more > 10) return false; }
  }
  *pp = p;
  *value = (unsigned long)v;
  return true;
}
// the first TAB or '\n' at or behind p
static inline const char *FieldEnd(const char *p) {
  for (;; p += 8) {
    const uint64_t w = Load8(p);
    const uint64_t m = EqMask(w, '\t') | EqMask(w, '\n');
    if (m) return p + (__builtin_ctzll(m) >> 3);
  }
}

// The common case by far -- TAB-separated BED of 3 to 6 clean columns -- in one pass over the line and without writing
// to it.  "Clean": every column non-empty and not starting with a blank, start and stop plain digit strings of at most 18
// digits, strand one of + - . 1 -1.  Anything else is left to ParseLine, so the two can not disagree.  The line ends with '\n'
// (every block does); returns that '\n', or nullptr if the line is not for this path (nothing has been appended then).
char *RegionReader::ParseBedLine(char *line, RegionBatch *out, ChromCache *cache) const {
  const char *p = line;
  if (*p == ' ' || *p == '\t') return nullptr;
  // the chromosome: at most 8 bytes up to the first TAB, no end of line before it; the bytes are the cache's key
  const uint64_t w0 = Load8(p);
  const uint64_t tab = EqMask(w0, '\t'), nl0 = EqMask(w0, '\n');
  size_t chrom_len;
  uint64_t chrom_key;
  if (tab) {
    if (nl0 && (nl0 & (0 - nl0)) < (tab & (0 - tab))) return nullptr;
    chrom_len = (size_t)(__builtin_ctzll(tab) >> 3);
    if (chrom_len == 0) return nullptr;
    chrom_key = w0 & ((1ull << (8 * chrom_len)) - 1);
  } else {
    if (nl0 || p[8] != '\t') return nullptr;
    chrom_len = 8;
    chrom_key = w0;
  }
  p += chrom_len + 1;
  unsigned long start = 0, stop = 0;
  if (!ParseDigits(&p, &start) || *p != '\t') return nullptr;
  ++p;
  if (!ParseDigits(&p, &stop)) return nullptr;
  char strand = '+';
  const char *label = nullptr, *label_end = nullptr;
  if (*p == '\t') {                                                     // column 4: label
    label = ++p;
    if (*p == ' ' || *p == '\t' || *p == '\n') return nullptr;
    p = FieldEnd(p);
    label_end = p;
    if (*p == '\t') {                                                   // column 5: score
      p++;
      if (*p == ' ' || *p == '\t' || *p == '\n') return nullptr;
      p = FieldEnd(p);
      if (*p == '\t') {                                                 // column 6: strand, then the end of the line
        p++;
        if (p[0] != '\n' && p[1] == '\n') {
          if (p[0] == '+' || p[0] == '.' || p[0] == '1') strand = '+';
          else if (p[0] == '-') strand = '-';
          else return nullptr;
          p += 1;
        } else if (p[0] == '-' && p[1] == '1' && p[2] == '\n') { strand = '-'; p += 2; }
        else return nullptr;
      }
    }
  } else if (*p != '\n') return nullptr;
  // p is at the line's '\n' now
  start += 1;
  if (start > (unsigned long)INT32_MAX || stop > (unsigned long)INT32_MAX) return nullptr;
  if (max_label_value_ > 1) {                                          // GetLabelValue, genomic_intervals.cpp:1081-1085
    long w = 0;
    if (label != nullptr) {
      char tmp[32];
      const size_t n = (size_t)(label_end - label);
      if (n > sizeof tmp - 1) return nullptr;
      memcpy(tmp, label, n);
      tmp[n] = 0;
      w = atol(tmp);
    }
    w = std::min(max_label_value_, w);
    if (w < INT32_MIN || w > INT32_MAX) return nullptr;
    out->weight.push_back((int32_t)w);
  }
  out->chrom.push_back(cache->GetKey(chrom_key, line, chrom_len));
  out->strand.push_back((int8_t)strand);
  out->start.push_back((int32_t)start);
  out->stop.push_back((int32_t)stop);
  out->offset.push_back((int64_t)out->chrom.size());
  if (keep_labels_) { if (label) out->label.emplace_back(label, label_end); else out->label.emplace_back("_"); }
  return const_cast<char *>(p);
}

// The same for SAM lines (what BAM records are spelt as, too): the eleven mandatory columns found by their TABs without writing to
// the line, FLAG and POS plain digit strings, a CIGAR of digits and the operations the reference knows (or "*"), at most 16
// blocks, the length of SEQ in agreement with the CIGAR.  Every other line -- and every line when labels are weights -- is left
// to ParseLine, which also words the errors.  Returns the line's '\n', or nullptr (nothing has been appended then).
char *RegionReader::ParseSamLine(char *line, RegionBatch *out, ChromCache *cache) const {
  if (max_label_value_ > 1) return nullptr;
  const char *f[11];
  const char *p = line;
  for (int k = 0; k < 11; k++) {
    f[k] = p;
    if (*p == ' ' || *p == '\t' || *p == '\n') return nullptr;
    p = FieldEnd(p);
    if (k < 10) { if (*p != '\t') return nullptr; p++; }
  }
  const char *nl = *p == '\n' ? p : (const char *)rawmemchr(p, '\n');
  unsigned long flag = 0, pos = 0;
  const char *q = f[1];
  if (!ParseDigits(&q, &flag) || q != f[2] - 1) return nullptr;
  q = f[3];
  if (!ParseDigits(&q, &pos) || q != f[4] - 1 || pos > (unsigned long)INT32_MAX) return nullptr;
  const size_t rname_len = (size_t)(f[3] - 1 - f[2]), seq_len = (size_t)(f[10] - 1 - f[9]);
  if (rname_len > 63) return nullptr;
  const bool seq_star = seq_len == 1 && f[9][0] == '*';
  long b_start[16], b_stop[16];
  int n_blocks = 0;
  long start = (long)pos, reference_len = 0, fragment_len = 0;
  const char *c = f[5], *const c_end = f[6] - 1;
  if (c_end - c == 1 && *c == '*') {                                     // "*" stands for <length of SEQ>M
    reference_len = fragment_len = (long)seq_len;
  } else {
    while (c < c_end) {
      unsigned long len = 0;
      if (!ParseDigits(&c, &len) || c >= c_end || len > (unsigned long)INT32_MAX) return nullptr;
      const char type = *c++;
      if (type == 'M' || type == 'X') { fragment_len += (long)len; reference_len += (long)len; }
      else if (type == 'I' || type == 'S') fragment_len += (long)len;
      else if (type == 'D') reference_len += (long)len;
      else if (type == 'N') {
        if (n_blocks == 16) return nullptr;
        b_start[n_blocks] = start; b_stop[n_blocks] = start + reference_len - 1; n_blocks++;
        start = start + reference_len + (long)len;
        reference_len = 0;
      } else if (type != 'H' && type != 'P' && type != '-') return nullptr;
    }
  }
  if (!seq_star && (long)seq_len != fragment_len) return nullptr;
  if (reference_len > 0) {
    if (n_blocks == 16) return nullptr;
    b_start[n_blocks] = start; b_stop[n_blocks] = start + reference_len - 1; n_blocks++;
  }
  if (n_blocks == 0) return nullptr;
  for (int k = 0; k < n_blocks; k++)
    if (b_start[k] > INT32_MAX || b_stop[k] > INT32_MAX || b_stop[k] < INT32_MIN) return nullptr;
  int32_t chrom;
  if (rname_len <= 8) chrom = cache->GetShort(f[2], rname_len);
  else {
    char name[64];
    memcpy(name, f[2], rname_len);
    name[rname_len] = 0;
    chrom = cache->Get(name);
  }
  const int8_t strand = (flag & 0x10ul) ? '-' : '+';                   // CalcStrandFromFlag, :2870-2873
  for (int k = 0; k < n_blocks; k++) {
    out->chrom.push_back(chrom);
    out->strand.push_back(strand);
    out->start.push_back((int32_t)b_start[k]);
    out->stop.push_back((int32_t)b_stop[k]);
  }
  if (n_blocks != 1) out->multi = true;
  out->offset.push_back((int64_t)out->chrom.size());
  if (keep_labels_) out->label.emplace_back(f[0], f[1] - 1);
  return const_cast<char *>(nl);
}

struct RegionReader::Piece {
  char *lo = nullptr, *hi = nullptr;
  RegionBatch batch;
  long lines = 0;                                                      // lines consumed, the malformed one (if any) not included
  bool bad = false;
  ParseError err;
  int64_t region_at = 0, interval_at = 0;                              // where this piece's regions / intervals go in the joined batch
};

RegionReader::~RegionReader() {
  for (Piece *p : piece_) delete p;
}

// Parses the lines of [begin, end) -- each terminated by '\n' -- and appends their regions to out.  The run is cut into
// pieces at line boundaries, one per thread; each thread fills a batch of its own, then copies it to its place in `out`.
bool RegionReader::ParseRun(char *begin, char *end, RegionBatch *out) {
  const size_t bytes = (size_t)(end - begin);
  static const size_t piece_bytes = getenv("GT_PARSE_PIECE_BYTES") ? (size_t)std::max(1L, atol(getenv("GT_PARSE_PIECE_BYTES"))) : (1u << 18);
  const int pieces = (int)std::max<size_t>(1, std::min<size_t>((size_t)threads_, bytes / piece_bytes));
  while ((int)piece_.size() < pieces) piece_.push_back(new Piece());
  char *at = begin;
  for (int i = 0; i < pieces; i++) {
    Piece &pc = *piece_[i];
    pc.lo = at;
    char *want = i + 1 == pieces ? end : begin + bytes * (size_t)(i + 1) / (size_t)pieces;
    if (want < at) want = at;
    if (want < end && want > begin && want[-1] != '\n') {
      char *nl = (char *)memchr(want, '\n', (size_t)(end - want));
      want = nl ? nl + 1 : end;
    }
    pc.hi = at = want;
    pc.lines = 0; pc.bad = false;
  }
  const bool bed = fmt_ == F_BED, sam = fmt_ == F_SAM;
  auto parse = [&](int i) {
    Piece &pc = *piece_[i];
    RegionBatch *dst = &pc.batch;
    ChromCache cache(chroms_);
    dst->Clear();
    const size_t guess = (size_t)(pc.hi - pc.lo) / 24 + 16;            // avoids most reallocations the first time round
    dst->chrom.reserve(guess); dst->start.reserve(guess); dst->stop.reserve(guess); dst->strand.reserve(guess); dst->offset.reserve(guess + 1);
    for (char *p = pc.lo; p < pc.hi;) {
      char *nl = bed ? ParseBedLine(p, dst, &cache) : sam ? ParseSamLine(p, dst, &cache) : nullptr;
      if (nl == nullptr) {
        nl = (char *)memchr(p, '\n', (size_t)(pc.hi - p));
        *nl = 0;
        if (!ParseLine(p, dst, &cache, &pc.err)) { pc.bad = true; break; }
      }
      pc.lines++;
      p = nl + 1;
    }
    if (pc.bad) {                                                      // drop what the malformed line left behind: whole regions only
      const size_t keep = (size_t)dst->offset.back();
      dst->chrom.resize(keep); dst->start.resize(keep); dst->stop.resize(keep); dst->strand.resize(keep);
      dst->weight.resize(max_label_value_ > 1 ? dst->offset.size() - 1 : 0);
      if (keep_labels_) dst->label.resize(dst->offset.size() - 1);
    }
  };
  auto join = [&](int i) {                                             // this piece's batch to its place in out
    const Piece &pc = *piece_[i];
    const RegionBatch &b = pc.batch;
    const size_t ni = b.chrom.size(), nr = (size_t)b.n_regions();
    memcpy(out->chrom.data() + pc.interval_at, b.chrom.data(), ni * sizeof(int32_t));
    memcpy(out->start.data() + pc.interval_at, b.start.data(), ni * sizeof(int32_t));
    memcpy(out->stop.data() + pc.interval_at, b.stop.data(), ni * sizeof(int32_t));
    memcpy(out->strand.data() + pc.interval_at, b.strand.data(), ni);
    if (!b.weight.empty()) memcpy(out->weight.data() + pc.region_at, b.weight.data(), nr * sizeof(int32_t));
    int64_t *off = out->offset.data() + pc.region_at + 1;
    for (size_t k = 1; k <= nr; k++) off[k - 1] = pc.interval_at + b.offset[k];
  };
  auto run_all = [&](int n, const std::function<void(int)> &fn) {
    if (n == 1) { fn(0); return; }
    std::vector<std::thread> th;
    for (int i = 1; i < n; i++) th.emplace_back(fn, i);
    fn(0);
    for (auto &t : th) t.join();
  };
  run_all(pieces, parse);
  // pieces up to and including the first one that met a malformed line
  int used = pieces;
  for (int i = 0; i < pieces; i++) if (piece_[i]->bad) { used = i + 1; break; }
  int64_t regions = out->n_regions(), intervals = (int64_t)out->chrom.size();
  long lines = 0;
  for (int i = 0; i < used; i++) {
    Piece &pc = *piece_[i];
    pc.region_at = regions; pc.interval_at = intervals;
    regions += pc.batch.n_regions(); intervals += (int64_t)pc.batch.chrom.size();
    lines += pc.lines;
    out->multi = out->multi || pc.batch.multi;
  }
  out->chrom.resize((size_t)intervals); out->start.resize((size_t)intervals); out->stop.resize((size_t)intervals);
  out->strand.resize((size_t)intervals); out->offset.resize((size_t)regions + 1);
  if (max_label_value_ > 1) out->weight.resize((size_t)regions);
  run_all(used, join);
  if (keep_labels_)
    for (int i = 0; i < used; i++)
      for (auto &l : piece_[i]->batch.label) out->label.emplace_back(std::move(l));
  if (used > 0 && piece_[used - 1]->bad) {
    failed_ = true;
    error_ = piece_[used - 1]->err;
    error_.line = reader_.line_no() + lines + 1;
    reader_.Advance(lines + 1);
    return false;
  }
  reader_.Advance(lines);
  return true;
}

int64_t RegionReader::Read(RegionBatch *out, int64_t max_regions) {
  out->Clear();
  if (fmt_ == F_EMPTY || failed_) return 0;
  if (pending_) {
    out->first_line = reader_.line_no();
    ChromCache cache(chroms_);
    char *line = pending_;
    pending_ = nullptr;
    if (!ParseLine(line, out, &cache, &error_)) {
      failed_ = true;
      error_.line = reader_.line_no();
      const size_t keep = (size_t)out->offset.back();
      out->chrom.resize(keep); out->start.resize(keep); out->stop.resize(keep); out->strand.resize(keep);
      return out->n_regions();
    }
  } else {
    out->first_line = reader_.line_no() + 1;
  }
  while (out->n_regions() < max_regions) {
    char *begin, *end;
    if (!reader_.NextRun(&begin, &end)) break;
    if (!ParseRun(begin, end, out)) break;
  }
  return out->n_regions();
}

int64_t RegionReader::ReadKeep(RegionBatch *out, std::vector<std::string> *raw, int64_t max_regions) {
  out->Clear();
  if (fmt_ == F_EMPTY || failed_) return 0;
  ChromCache cache(chroms_);
  bool first = true;
  while (out->n_regions() < max_regions) {
    char *line = pending_;
    if (line) pending_ = nullptr; else line = reader_.Next();
    if (line == nullptr) break;
    if (first) { out->first_line = reader_.line_no(); first = false; }
    raw->emplace_back(line);                                             // before the tokeniser writes its terminators into the line
    if (!ParseLine(line, out, &cache, &error_)) {
      failed_ = true;
      error_.line = reader_.line_no();
      raw->pop_back();
      const size_t keep = (size_t)out->offset.back();
      out->chrom.resize(keep); out->start.resize(keep); out->stop.resize(keep); out->strand.resize(keep);
      out->weight.resize(max_label_value_ > 1 ? out->offset.size() - 1 : 0);
      if (keep_labels_) out->label.resize(out->offset.size() - 1);
      break;
    }
  }
  return out->n_regions();
}

static void AppendLong(std::string *out, long v) { char t[32]; out->append(t, (size_t)snprintf(t, sizeof t, "%ld", v)); }

void PrintModified(const std::string &format, const std::string &raw, const RegionBatch &b, int64_t k, const ChromTable &chroms,
                   const char *label, long start, long stop, std::string *out) {
  const int64_t lo = b.offset[k], hi = b.offset[k + 1];
  for (int64_t i = lo + 1; i < hi; i++)                                  // IsCompatible(false), :1009-1017
    if (b.chrom[i] != b.chrom[lo] || b.strand[i] != b.strand[lo]) die_line(b.line(k), "intervals should have the same chromosome/strand for this operation!");
  const std::string &chrom = chroms.name[b.chrom[lo]];
  const char strand = (char)b.strand[lo];
  std::string copy(raw);
  char *inp = &copy[0];
  if (format == "BED") {                                               // six columns whatever the input had; the score column as read (0 if absent)
    const char sep = strchr(inp, '\t') == nullptr ? ' ' : '\t';
    const int n_tokens = CountTokens(inp, sep);
    long score = 0;
    if (n_tokens >= 5) { for (int i = 0; i < 4; i++) NextToken(&inp, sep); score = atol(NextToken(&inp, sep)); }
    out->append(chrom); out->push_back('\t'); AppendLong(out, start - 1); out->push_back('\t'); AppendLong(out, stop); out->push_back('\t');
    out->append(label); out->push_back('\t'); AppendLong(out, score); out->push_back('\t'); out->push_back(strand); out->push_back('\n');
  } else if (format == "GFF") {
    const int n_tokens = CountTokens(inp, '\t');
    NextToken(&inp, '\t');
    const char *source = NextToken(&inp, '\t'), *feature = NextToken(&inp, '\t');
    NextToken(&inp, '\t'); NextToken(&inp, '\t');
    const char *score = NextToken(&inp, '\t');
    NextToken(&inp, '\t');
    const char frame = NextToken(&inp, '\t')[0];
    out->append(chrom); out->push_back('\t'); out->append(source); out->push_back('\t'); out->append(feature); out->push_back('\t');
    AppendLong(out, start); out->push_back('\t'); AppendLong(out, stop); out->push_back('\t'); out->append(score);
    out->push_back('\t'); out->push_back(strand); out->push_back('\t'); out->push_back(frame);
    if (n_tokens > 8) { out->push_back('\t'); out->append(label); NextToken(&inp, '\t'); }
    if (n_tokens > 9) { out->push_back('\t'); out->append(NextToken(&inp, '\t')); }
    out->push_back('\n');
  } else if (format == "SAM") {
    const int n_tokens = CountTokens(inp, '\t');
    NextToken(&inp, '\t');
    const long flag = atol(NextToken(&inp, '\t'));
    for (int i = 0; i < 4; i++) NextToken(&inp, '\t');
    const char *rnext = NextToken(&inp, '\t');
    for (int i = 0; i < 4; i++) NextToken(&inp, '\t');
    out->append(label); out->push_back('\t'); AppendLong(out, flag); out->push_back('\t'); out->append(chrom); out->push_back('\t');
    AppendLong(out, start); out->append("\t255\t"); AppendLong(out, stop - start + 1); out->append("M\t"); out->append(rnext);
    out->append("\t0\t0\t*\t*");
    if (n_tokens > 11) { out->push_back('\t'); out->append(inp); }
    out->push_back('\n');
  } else {
    out->append(label); out->push_back('\t'); out->append(chrom); out->push_back(' '); out->push_back(strand); out->push_back(' ');
    AppendLong(out, start); out->push_back(' '); AppendLong(out, stop); out->push_back('\n');
  }
}

void PrintRegion(const std::string &format, const std::string &raw, const RegionBatch &b, int64_t k, const ChromTable &chroms, std::string *out) {
  const int64_t lo = b.offset[k], hi = b.offset[k + 1];
  if (hi <= lo) return;                                                // "if (I.size()==0) return;"
  const std::string &chrom = chroms.name[b.chrom[lo]];
  std::string copy(raw);
  char *inp = &copy[0];
  if (format == "BED") {                                               // GenomicRegionBED::Print, :2188-2219
    const char sep = strchr(inp, '\t') == nullptr ? ' ' : '\t';
    const int n_tokens = CountTokens(inp, sep);
    NextToken(&inp, sep); NextToken(&inp, sep); NextToken(&inp, sep);
    out->append(chrom); out->push_back('\t'); AppendLong(out, (long)b.start[lo] - 1); out->push_back('\t'); AppendLong(out, (long)b.stop[hi - 1]);
    if (n_tokens >= 4) {
      out->push_back('\t'); out->append(b.label[k]);
      NextToken(&inp, sep);
      if (n_tokens >= 5) {
        out->push_back('\t'); AppendLong(out, atol(NextToken(&inp, sep)));
        if (n_tokens >= 6) {
          NextToken(&inp, sep);
          out->push_back('\t'); out->push_back((char)b.strand[lo]);
          if (n_tokens >= 8) {
            const long thick_start = atol(NextToken(&inp, sep)), thick_end = atol(NextToken(&inp, sep));
            out->push_back('\t'); AppendLong(out, thick_start); out->push_back('\t'); AppendLong(out, thick_end);
            if (n_tokens >= 9) {
              out->push_back('\t'); out->append(NextToken(&inp, sep));
              if (n_tokens == 12) {
                out->push_back('\t'); AppendLong(out, (long)(hi - lo)); out->push_back('\t');
                for (int64_t i = lo; i < hi; i++) { AppendLong(out, (long)b.stop[i] - (long)b.start[i] + 1); if (i + 1 < hi) out->push_back(','); }
                out->append("\t0");
                for (int64_t i = lo + 1; i < hi; i++) {
                  if (b.start[i] <= b.stop[i - 1]) { fprintf(stderr, "Line %lu: exons cannot overlap!\n", (unsigned long)b.line(k)); exit(1); }
                  out->push_back(','); AppendLong(out, (long)b.start[i] - (long)b.start[lo]);
                }
              }
            }
          }
        }
      }
    }
    out->push_back('\n');
  } else if (format == "GFF") {                                        // GenomicRegionGFF::Print, :3523-3530
    const int n_tokens = CountTokens(inp, '\t');
    NextToken(&inp, '\t');
    const char *source = NextToken(&inp, '\t'), *feature = NextToken(&inp, '\t');
    NextToken(&inp, '\t'); NextToken(&inp, '\t');
    const char *score = NextToken(&inp, '\t');
    NextToken(&inp, '\t');
    const char frame = NextToken(&inp, '\t')[0];
    out->append(chrom); out->push_back('\t'); out->append(source); out->push_back('\t'); out->append(feature); out->push_back('\t');
    AppendLong(out, (long)b.start[lo]); out->push_back('\t'); AppendLong(out, (long)b.stop[lo]); out->push_back('\t'); out->append(score);
    out->push_back('\t'); out->push_back((char)b.strand[lo]); out->push_back('\t'); out->push_back(frame);
    if (n_tokens > 8) { out->push_back('\t'); NextToken(&inp, '\t'); out->append(b.label[k]); }       // LABEL (the caller may have merged another into it)
    if (n_tokens > 9) { out->push_back('\t'); out->append(NextToken(&inp, '\t')); }
    out->push_back('\n');
  } else if (format == "SAM") {                                        // GenomicRegionSAM::Print, :2819-2825
    const int n_tokens = CountTokens(inp, '\t');
    NextToken(&inp, '\t');
    const std::string &label = b.label[k];                              // LABEL (the caller may have merged another into it)
    const long flag = atol(NextToken(&inp, '\t'));
    NextToken(&inp, '\t'); NextToken(&inp, '\t');
    const long mapq = atol(NextToken(&inp, '\t'));
    const char *cigar = NextToken(&inp, '\t'), *rnext = NextToken(&inp, '\t');
    const long pnext = atol(NextToken(&inp, '\t')), tlen = atol(NextToken(&inp, '\t'));
    const char *seq = NextToken(&inp, '\t'), *qual = NextToken(&inp, '\t');
    out->append(label); out->push_back('\t'); AppendLong(out, flag); out->push_back('\t'); out->append(chrom); out->push_back('\t');
    AppendLong(out, (long)b.start[lo]); out->push_back('\t'); AppendLong(out, mapq); out->push_back('\t');
    if (strcmp(cigar, "*") == 0) { AppendLong(out, (long)strlen(seq)); out->push_back('M'); } else out->append(cigar);   // Read replaces "*" (:2791)
    out->push_back('\t'); out->append(rnext); out->push_back('\t'); AppendLong(out, pnext); out->push_back('\t'); AppendLong(out, tlen);
    out->push_back('\t'); out->append(seq); out->push_back('\t'); out->append(qual);
    if (n_tokens > 11) { out->push_back('\t'); out->append(inp); }      // OPTIONAL: the rest of the line
    out->push_back('\n');
  } else {                                                             // REG: GenomicRegion::Print, :843-849
    out->append(b.label[k]); out->push_back('\t');
    for (int64_t i = lo; i < hi; i++) {
      out->append(chroms.name[b.chrom[i]]); out->push_back(' '); out->push_back((char)b.strand[i]); out->push_back(' ');
      AppendLong(out, (long)b.start[i]); out->push_back(' '); AppendLong(out, (long)b.stop[i]);
      if (i + 1 < hi) out->push_back(' ');
    }
    out->push_back('\n');
  }
}

void RegionReader::Fail() const {
  if (error_.raw) { std::cerr << error_.message; exit(1); }
  if (error_.with_line) die_line(error_.line, error_.message);
  die(error_.message);
}

void (*exit_hook)() = nullptr;
#undef exit
void Exit(int code) {
  if (exit_hook) { void (*h)() = exit_hook; exit_hook = nullptr; h(); }
  // GT_FAST_EXIT=1: what the driver has printed is flushed and the process ends at once, leaving the unwinding of the CUDA runtime
  // (modules unloaded, every allocation freed one by one, the context destroyed) to the kernel.  Off by default: measured on the
  // GPU boxes it is lost in the spread of the context's creation (1 to 7 s, profiles/r2ag_cli_exit.json), and profilers that
  // collect at exit need the ordinary way out.
  const char *env = getenv("GT_FAST_EXIT");
  if (env != nullptr && env[0] == '1') {
    fflush(nullptr);
    std::cout.flush();
    std::cerr.flush();
    ::_exit(code);
  }
  ::exit(code);
}
#define exit(code) ::gt::Exit(code)

}  // namespace gt
