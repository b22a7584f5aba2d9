// CPU unit test of the DIRECT engine's shared-memory rank structure (csrc/gtb_direct2_tables.h): the lookup the kernel runs is
// compiled here for the host and compared with lower_bound over the evaluation points -- whenever it says "served here", the slot
// must be the RANK engine's for both ends of the read.  Prints the share of served probes.   g++ -O2 -std=c++17 ...
#include "../../ibm-cbc-genomic-tools_b200/csrc/gtb_direct2_tables.h"
#include <stdio.h>
#include <stdlib.h>
#include <random>

static int check_case(uint64_t seed, int32_t n_chrom, int32_t n_class, int cp, int cm, int64_t chrom_len, int n_points, uint64_t max_records,
                      bool negative, bool exhaustive) {
  std::mt19937_64 rng(seed);
  std::vector<int32_t> goff((size_t)n_chrom * n_class + 1, 0), points;
  for (int g = 0; g < n_chrom * n_class; g++) {
    goff[g] = (int32_t)points.size();
    const int c = g / n_class;
    const int64_t len = std::max<int64_t>(100, chrom_len >> (c % 3));
    int n = (int)(rng() % (uint64_t)(2 * n_points / (n_chrom * n_class) + 1));
    if (g % 7 == 5) n = 0;                                           // an empty group
    std::vector<int32_t> v;
    for (int i = 0; i < n; i++) {
      int64_t x = (int64_t)(rng() % (uint64_t)len);
      if (i % 5 == 0 && i > 0) x = v.back() + (int64_t)(rng() % 40);  // clumps: several points per cell
      if (negative && i % 11 == 0) x = -(int64_t)(rng() % 1000);
      if (g % 7 == 3) x = -(int64_t)(rng() % 50);                     // a group with points <= 0 only
      v.push_back((int32_t)std::min<int64_t>(x, len));
    }
    std::sort(v.begin(), v.end());
    v.erase(std::unique(v.begin(), v.end()), v.end());
    for (int32_t x : v) points.push_back(x);
    if (!v.empty()) points.push_back(INT32_MAX);
  }
  goff[(size_t)n_chrom * n_class] = (int32_t)points.size();
  const int cls_sig[2] = {cp, cm};
  D2Tables t;
  if (!d2_build(n_chrom, n_class, cls_sig, goff, points, max_records, t)) { printf("  (does not fit %llu records)\n", (unsigned long long)max_records); return 0; }
  if (t.recs.size() > max_records) { printf("FAIL: %zu records > %llu\n", t.recs.size(), (unsigned long long)max_records); return 1; }
  const uint32_t nsig = t.p.nsig;
  uint64_t probes = 0, served = 0;
  auto probe = [&](uint32_t c, uint32_t sg, uint32_t s, uint32_t len) -> int {
    uint32_t slot = 0;
    const uint32_t cc = std::min<uint32_t>(c, (uint32_t)n_chrom);
    const bool fast = d2_lookup(t.p, t.gtab.data(), t.recs.data(), cc, sg, s, len, slot);
    probes++;
    if (!fast) return 0;
    served++;
    const int cls = cls_sig[sg];
    if (c >= (uint32_t)n_chrom || cls < 0) { printf("FAIL: served a read of a group that does not exist (c %u sg %u)\n", c, sg); return 1; }
    const int g = (int)c * n_class + cls;
    const int32_t gb = goff[g], ge = goff[g + 1];
    if (ge - gb < 2 || points[ge - 2] < 1) { printf("FAIL: served a read of an unserved group %d\n", g); return 1; }
    const int64_t e = (int64_t)s + len;
    const uint32_t jS = (uint32_t)(std::lower_bound(points.begin() + gb, points.begin() + ge - 1, (int32_t)s) - points.begin());
    const uint32_t jE = (uint32_t)(std::lower_bound(points.begin() + gb, points.begin() + ge - 1, (int32_t)std::min<int64_t>(e, INT32_MAX)) - points.begin());
    if (slot != jS || slot != jE) { printf("FAIL: c %u sg %u s %u len %u: slot %u, lower_bound %u / %u (cell_w %u)\n", c, sg, s, len, slot, jS, jE, t.p.cell_w); return 1; }
    return 0;
  };
  if (exhaustive) {
    for (uint32_t c = 0; c <= (uint32_t)n_chrom; c++)
      for (uint32_t sg = 0; sg < nsig; sg++)
        for (uint32_t s = 1; s < (uint32_t)chrom_len + 3 * t.p.cell_w; s++)
          for (uint32_t len : {0u, 1u, 49u, t.p.cell_w - 1, t.p.cell_w})
            if (probe(c, sg, s, len)) return 1;
  } else {
    for (int i = 0; i < 2000000; i++) {
      const uint32_t c = (uint32_t)(rng() % (uint64_t)(n_chrom + 2));
      const uint32_t sg = (uint32_t)(rng() % nsig);
      uint32_t s = 1 + (uint32_t)(rng() % (uint64_t)(chrom_len + chrom_len / 8));
      if (i % 97 == 0) s = 0x7FFFFFFFu - (uint32_t)(rng() % 100000);       // the far end of the coordinate range
      const uint32_t len = i % 13 == 0 ? (uint32_t)(rng() % 100000) : 49u;
      if ((uint64_t)s + len > 0x7FFFFFFFu) continue;
      if (probe(c, sg, s, len)) return 1;
    }
  }
  printf("  seed %llu chroms %d classes %d (+:%d -:%d) cell %u bp, %u records: %llu probes, %.1f %% served; cells %llu, with points %.1f %%, phantom %.1f %%\n",
         (unsigned long long)seed, n_chrom, n_class, cp, cm, t.p.cell_w, t.p.n_rec, (unsigned long long)probes, 100.0 * served / probes,
         (unsigned long long)t.cells_total, 100.0 * t.cells_point / std::max<uint64_t>(1, t.cells_total),
         100.0 * t.cells_phantom / std::max<uint64_t>(1, t.cells_total));
  return 0;
}

int main() {
  int bad = 0;
  // the cell-number multiply: exact over the whole non-negative int32 range for awkward widths
  for (uint64_t w : {16ull, 17ull, 1000ull, 14931ull, 16384ull, 65535ull, 65536ull, 1000003ull, (1ull << 30)}) {
    uint32_t l = 0; while (((uint64_t)1 << l) < w) l++;
    const uint32_t magic = (uint32_t)((((uint64_t)1 << (31 + l)) + w - 1) / w), shift = l - 1;
    std::mt19937_64 rng(w);
    for (int i = 0; i < 3000000; i++) {
      uint32_t s = (uint32_t)(rng() & 0x7FFFFFFFu);
      if (i < 100000) s = 0x7FFFFFFFu - (uint32_t)i;
      else if (i < 200000) s = (uint32_t)(((uint64_t)(i - 100000) * w) & 0x7FFFFFFFu) - (i & 1);
      s &= 0x7FFFFFFFu;
      if ((d2_umulhi(s, magic) >> shift) != s / w) { printf("FAIL: %u / %llu\n", s, (unsigned long long)w); bad++; break; }
    }
  }
  for (uint32_t k = 0; k < (1u << 17); k++) if (d2_umulhi(k, D2_DIV46) != k / 46) { printf("FAIL: %u / 46\n", k); bad++; break; }
  // small genomes, every start position
  bad += check_case(1, 3, 2, 0, 1, 3000, 60, 40, false, true);
  bad += check_case(2, 3, 2, 1, 0, 3000, 200, 64, true, true);         // '-' got the lower class number
  bad += check_case(3, 2, 1, 0, 0, 5000, 100, 30, true, true);         // -i: one signature
  bad += check_case(4, 4, 3, 2, -1, 2000, 150, 50, false, true);       // no '-' regions at all, a third strand class in between
  // hg19-sized axes, random probes
  bad += check_case(5, 25, 2, 0, 1, 249250621, 120000, 8600, false, false);
  bad += check_case(6, 25, 2, 0, 1, 249250621, 120000, 12000, true, false);
  bad += check_case(7, 1, 1, 0, 0, 197195432, 9570, 8600, false, false);
  bad += check_case(8, 90, 2, 0, 1, 40000000, 180000, 9000, false, false);
  printf(bad ? "FAILED\n" : "ok\n");
  return bad ? 1 : 0;
}
