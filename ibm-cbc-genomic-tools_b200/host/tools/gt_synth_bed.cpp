// gt_synth_bed -- writes the synthetic hg19 read stream of DESIGN.md section 6 (the one gtb_synth_reads and
// tests/support.py:synth_reads produce) as TAB-separated BED6 text, so that the command-line drivers and the reference
// binaries can be timed end to end on the same input.  Bench / test helper; no GPU needed.
//
//   gt_synth_bed N SEED [READ_LEN] > reads.bed         label = r<i>, score 0
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

static const struct { const char *name; long len; } HG19[] = {
    {"chr1", 249250621}, {"chr2", 243199373}, {"chr3", 198022430}, {"chr4", 191154276}, {"chr5", 180915260}, {"chr6", 171115067},
    {"chr7", 159138663}, {"chr8", 146364022}, {"chr9", 141213431}, {"chr10", 135534747}, {"chr11", 135006516}, {"chr12", 133851895},
    {"chr13", 115169878}, {"chr14", 107349540}, {"chr15", 102531392}, {"chr16", 90354753}, {"chr17", 81195210}, {"chr18", 78077248},
    {"chr19", 59128983}, {"chr20", 63025520}, {"chr21", 48129895}, {"chr22", 51304566}, {"chrX", 155270560}, {"chrY", 59373566},
    {"chrM", 16571}};

static inline uint64_t splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

static inline char *put_num(char *p, uint64_t v) {
  char tmp[24];
  int n = 0;
  do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
  while (n) *p++ = tmp[--n];
  return p;
}

int main(int argc, char **argv) {
  if (argc < 3) { fprintf(stderr, "usage: gt_synth_bed N SEED [READ_LEN]\n"); return 2; }
  const int64_t n = atoll(argv[1]);
  const uint64_t seed = strtoull(argv[2], nullptr, 10);
  const int read_len = argc > 3 ? atoi(argv[3]) : 50;
  // chromosome ids in strcmp order of the names, as everywhere else
  std::vector<std::pair<std::string, long>> chrom;
  for (auto &c : HG19) chrom.emplace_back(c.name, c.len);
  std::sort(chrom.begin(), chrom.end());
  std::vector<uint64_t> cum(chrom.size() + 1, 0);
  for (size_t c = 0; c < chrom.size(); c++) cum[c + 1] = cum[c] + (uint64_t)std::max(0L, chrom[c].second - read_len + 1);
  const uint64_t span = cum.back();
  const int threads = (int)std::min(32u, std::max(1u, std::thread::hardware_concurrency()));
  const int64_t slab = 1 << 18;                                          // lines per thread and round
  std::vector<std::vector<char>> buf((size_t)threads);
  for (auto &b : buf) b.resize((size_t)slab * 64);
  std::vector<size_t> len((size_t)threads);
  for (int64_t base = 0; base < n; base += slab * threads) {
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++)
      th.emplace_back([&, t] {
        const int64_t lo = std::min(n, base + slab * t), hi = std::min(n, lo + slab);
        char *p = buf[t].data();
        for (int64_t i = lo; i < hi; i++) {
          const uint64_t a = splitmix64(seed * 0x9E3779B97F4A7C15ull + (uint64_t)i), b = splitmix64(a);
          const uint64_t pos = a % span;
          const size_t c = (size_t)(std::upper_bound(cum.begin(), cum.end(), pos) - cum.begin()) - 1;
          const uint64_t start0 = pos - cum[c];                          // BED: 0-based, half-open
          memcpy(p, chrom[c].first.data(), chrom[c].first.size()); p += chrom[c].first.size();
          *p++ = '\t'; p = put_num(p, start0);
          *p++ = '\t'; p = put_num(p, start0 + (uint64_t)read_len);
          *p++ = '\t'; *p++ = 'r'; p = put_num(p, (uint64_t)i);
          *p++ = '\t'; *p++ = '0'; *p++ = '\t'; *p++ = (b & 1ull) ? '-' : '+'; *p++ = '\n';
        }
        len[t] = (size_t)(p - buf[t].data());
      });
    for (auto &t : th) t.join();
    for (int t = 0; t < threads; t++) fwrite(buf[t].data(), 1, len[t], stdout);
  }
  return 0;
}
