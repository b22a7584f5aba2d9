// ref_engine.cpp -- engine-only timing harness around the UNMODIFIED reference classes.
// TEST / BASELINE INFRASTRUCTURE ONLY (built into oracle/_ref/ by oracle/Makefile; never part of
// the product).  Links the reference's own object files and times nothing but
//   UnsortedGenomicRegionSetOverlaps + GenomicRegionSetOverlaps::CountIndexOverlaps
// (genomic_intervals.cpp:5593-5764, :5304-5317) with BOTH sets already parsed into memory, so that
// the figure is comparable with the device-timed GPU number (text parsing excluded on both sides,
// SURVEY.md section 8d).
//
//   ref_engine REGIONS.bed SEED FIRST N READ_LEN [coverage]
//
// generates reads [FIRST, FIRST+N) of the counter-based hg19 stream (same as tests/support.py),
// writes them to a temporary BED file, lets the reference parse it (untimed), then times the engine.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unistd.h>
#include "core.h"
#include "genomic_intervals.h"

static const char *NAMES[25] = {"chr1", "chr10", "chr11", "chr12", "chr13", "chr14", "chr15", "chr16", "chr17", "chr18", "chr19", "chr2",
                                "chr20", "chr21", "chr22", "chr3", "chr4", "chr5", "chr6", "chr7", "chr8", "chr9", "chrM", "chrX", "chrY"};
static const long LENS[25] = {249250621, 135534747, 135006516, 133851895, 115169878, 107349540, 102531392, 90354753, 81195210, 78077248,
                              59128983, 243199373, 63025520, 48129895, 51304566, 198022430, 191154276, 180915260, 171115067, 159138663,
                              146364022, 141213431, 16571, 155270560, 59373566};

static unsigned long long splitmix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

int main(int argc, char **argv) {
  if (argc < 6) { fprintf(stderr, "usage: ref_engine REGIONS.bed SEED FIRST N READ_LEN [coverage]\n"); return 2; }
  char *regions_file = argv[1];
  unsigned long long seed = strtoull(argv[2], 0, 10), first = strtoull(argv[3], 0, 10), n = strtoull(argv[4], 0, 10);
  long read_len = atol(argv[5]);
  bool coverage = argc > 6 && strcmp(argv[6], "coverage") == 0;

  unsigned long long cum[26]; cum[0] = 0;
  for (int c = 0; c < 25; c++) cum[c + 1] = cum[c] + (unsigned long long)(LENS[c] - read_len + 1);
  char tmpl[] = "/tmp/gtb_ref_engine_XXXXXX";
  int fd = mkstemp(tmpl);
  FILE *f = fdopen(fd, "w");
  for (unsigned long long i = first; i < first + n; i++) {
    unsigned long long a = splitmix64(seed * 0x9E3779B97F4A7C15ull + i), b = splitmix64(a);
    unsigned long long p = a % cum[25];
    int c = 0; while (cum[c + 1] <= p) c++;
    long start = (long)(p - cum[c]) + 1;
    fprintf(f, "%s\t%ld\t%ld\tr\t0\t%c\n", NAMES[c], start - 1, start + read_len - 1, (b & 1ull) ? '-' : '+');
  }
  fclose(f);

  _MESSAGES_ = false;
  GenomicRegionSet RefRegSet(regions_file, 10000, false, true, true);
  GenomicRegionSet TestRegSet(tmpl, 10000, false, true, true);
  unlink(tmpl);

  auto t0 = std::chrono::steady_clock::now();
  GenomicRegionSetOverlaps *overlaps = new UnsortedGenomicRegionSetOverlaps(&TestRegSet, &RefRegSet, "17,20,23,26");
  unsigned long int *v = coverage ? overlaps->CalcIndexCoverage(false, false, 1) : overlaps->CountIndexOverlaps(false, false, 1);
  auto t1 = std::chrono::steady_clock::now();
  unsigned long long sum = 0;
  for (long k = 0; k < RefRegSet.n_regions; k++) sum += v[k];
  printf("{\"engine_seconds\": %.6f, \"reads\": %llu, \"regions\": %ld, \"checksum\": %llu}\n",
         std::chrono::duration<double>(t1 - t0).count(), n, RefRegSet.n_regions, sum);
  return 0;
}
