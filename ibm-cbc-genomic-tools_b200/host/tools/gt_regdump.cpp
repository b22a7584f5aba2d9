// gt_regdump -- prints what the host parser (gt_host.h) makes of a region file, in the reference's REG spelling
// (`genomic_regions reg`, genomic_intervals.cpp:880-890: LABEL <TAB> chrom strand start stop [chrom strand start stop ...]).
// Needs no GPU: the tests use it to hold the multi-threaded parser against the reference's own reader on a CPU-only box, and
// it doubles as a parse-rate probe (-q: parse only, print the number of regions and intervals).
//
//   gt_regdump [-q] [-w MAX_LABEL_VALUE] [-c CHUNK_REGIONS] FILE|-
//   gt_regdump -p FILE|-        every region as GenomicRegion*::Print would write it in the file's own format (gt::PrintRegion: what
//                               `genomic_overlaps subset / overlap` and `genomic_regions gsort` print)
#include <stdlib.h>
#include <string.h>
#include <string>
#include "../gt_host.h"

int main(int argc, char **argv) {
  bool quiet = false, native = false;
  long max_label_value = 1;
  int64_t chunk = INT64_MAX;
  int a = 1;
  for (; a < argc && argv[a][0] == '-' && argv[a][1] != 0; a++) {
    if (!strcmp(argv[a], "-q")) quiet = true;
    else if (!strcmp(argv[a], "-p")) native = true;
    else if (!strcmp(argv[a], "-w") && a + 1 < argc) max_label_value = atol(argv[++a]);
    else if (!strcmp(argv[a], "-c") && a + 1 < argc) chunk = atol(argv[++a]);
    else { fprintf(stderr, "usage: gt_regdump [-q] [-w MAX_LABEL_VALUE] [-c CHUNK_REGIONS] FILE|-\n"); return 2; }
  }
  if (a >= argc) { fprintf(stderr, "usage: gt_regdump [-q] [-w MAX_LABEL_VALUE] [-c CHUNK_REGIONS] FILE|-\n"); return 2; }
  const char *path = strcmp(argv[a], "-") == 0 ? nullptr : argv[a];
  gt::ChromTable chroms;
  gt::RegionReader rr(path, &chroms, !quiet, max_label_value);
  gt::RegionBatch b;
  int64_t regions = 0, intervals = 0;
  std::string text;
  if (native) {
    fwrite(rr.header().data(), 1, rr.header().size(), stdout);
    std::vector<std::string> raw;
    for (;;) {
      raw.clear();
      const int64_t n = rr.ReadKeep(&b, &raw, 1 << 16);
      text.clear();
      for (int64_t k = 0; k < n; k++) gt::PrintRegion(rr.format(), raw[(size_t)k], b, k, chroms, &text);
      fwrite(text.data(), 1, text.size(), stdout);
      if (rr.failed()) { fflush(stdout); rr.Fail(); }
      if (n == 0) break;
    }
    fflush(stdout);
    return 0;
  }
  while (rr.Read(&b, chunk) > 0) {
    regions += b.n_regions();
    intervals += (int64_t)b.chrom.size();
    if (quiet) continue;
    text.clear();
    char num[64];
    for (int64_t k = 0; k < b.n_regions(); k++) {
      if (b.offset[k + 1] == b.offset[k]) continue;                    // a region without intervals prints nothing (GenomicRegion::Print)
      text += b.label[k];
      text += '\t';
      for (int64_t i = b.offset[k]; i < b.offset[k + 1]; i++) {
        if (i > b.offset[k]) text += ' ';
        text += chroms.name[b.chrom[i]];
        const int len = snprintf(num, sizeof num, " %c %d %d", (char)b.strand[i], b.start[i], b.stop[i]);
        text.append(num, (size_t)len);                                 // (a GFF strand column may be empty: the strand byte is then NUL)
      }
      if (max_label_value > 1) { snprintf(num, sizeof num, "\tw=%d", b.weight[k]); text += num; }
      snprintf(num, sizeof num, "\t#%ld\n", b.line(k));
      text += num;
    }
    fwrite(text.data(), 1, text.size(), stdout);
  }
  fflush(stdout);
  if (rr.failed()) rr.Fail();
  if (quiet) printf("format %s regions %ld intervals %ld\n", rr.format().c_str(), (long)regions, (long)intervals);
  return 0;
}
