// gtb_direct2_tables.h -- the shared-memory rank structure of the DIRECT engine's second form (gtb_direct.cu), in plain C++ so
// that the SAME lookup code runs inside the kernel and inside the CPU unit test (tests/cpp/test_direct2_tables.cpp).
//
// What it answers without leaving the SM: "is there an evaluation point anywhere near this read, and if not, which slot is it in?"
//
//   * Every chromosome's axis is cut into coarse cells of `cell_w` bp (any width: the cell of a coordinate is one multiply-high).
//     46 consecutive cells share one 8-byte RECORD: a 46-bit map (bit r set: a read starting in cell r must take the slow way) and
//     the 18-bit absolute slot of the record's first cell, i.e. the number of evaluation points of all groups before it.
//     The records of a chromosome's '+' and '-' groups are interleaved, so the per-chromosome table that says where they start
//     is indexed by the chromosome id alone (26 entries for hg19: one shared-memory wavefront per warp).
//   * For a read [s, e] whose cell has a clear bit and which ends in the same cell,
//         slot = prefix(record) + popc(map bits below the cell)
//     is lower_bound(points, s) == lower_bound(points, e) of the RANK engine -- no search, no global memory.
//   * popc counts CELLS, lower_bound counts POINTS.  A cell holding m > 1 points therefore sets its own bit and borrows the bits of
//     the next m - 1 otherwise clear cells of its record ("phantom" bits: those reads take the slow way although nothing is near
//     them), so that at every clear bit the two counts agree again.  A record starts with an exact prefix, so a debt never
//     crosses a record boundary.  With hg19 x 60 k regions at ~15 kbp per cell: 25 % of the cells hold a point, 4 % are phantoms.
//   * Groups the fast way cannot serve (no points at all, only points <= 0, strands the index has never seen, chromosomes
//     beyond the index) map to an all-ones record: the slow way's cell table knows what to do with them.
#pragma once
#include <stdint.h>
#include <algorithm>
#include <vector>

#ifdef __CUDACC__
#define D2_HD __host__ __device__ __forceinline__
#else
#define D2_HD inline
#endif

struct D2Rec { uint32_t x, y; };                     // x: cells 0..31, y[0..13]: cells 32..45, y[14..31]: absolute slot prefix
constexpr uint32_t D2_CELLS = 46;
constexpr uint32_t D2_DIV46 = 93368855u;             // ceil(2^32 / 46): umulhi(k, D2_DIV46) == k / 46 for k < 2^17
constexpr uint32_t D2_NC_BITS = 17, D2_NC_MASK = (1u << D2_NC_BITS) - 1u;
constexpr uint32_t D2_MAX_SLOTS = 1u << 18;

struct D2Params {
  uint32_t cell_w;                                   // coarse cell width, bp
  uint32_t magic, shift;                             // cell(s) = umulhi(s, magic) >> shift, exact for 0 <= s < 2^31
  uint32_t nsig;                                     // 2: '+' and '-' reads have records of their own, 1: they share (-i)
  uint32_t n_rec;                                    // records
  uint32_t n_gt;                                     // chromosome table entries = n_chrom + 1 (the last: chromosomes beyond the index)
};

D2_HD uint32_t d2_umulhi(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}
D2_HD uint32_t d2_popc64(uint64_t v) {
#ifdef __CUDA_ARCH__
  return (uint32_t)__popcll(v);
#else
  return (uint32_t)__builtin_popcountll(v);
#endif
}

// The lookup.  cc = min(chromosome id, n_chrom), sg = 0 for '+', 1 for '-' (0 when nsig == 1), s >= 1, len = e - s >= 0.
// Returns true and the slot if the read is served here; false if it has to take the slow way.
D2_HD bool d2_lookup(const D2Params &p, const uint32_t *gt, const D2Rec *recs, uint32_t cc, uint32_t sg, uint32_t s, uint32_t len,
                     uint32_t &slot) {
  const uint32_t g = gt[cc];
  const uint32_t kc = d2_umulhi(s, p.magic) >> p.shift;          // coarse cell of s
  const uint32_t rem = s - kc * p.cell_w;                        // offset of s inside it
  const uint32_t kl = kc < (g & D2_NC_MASK) ? kc : (g & D2_NC_MASK);   // beyond the chromosome's last cell: its trailing, clear cell
  const uint32_t qd = d2_umulhi(kl, D2_DIV46);
  const uint32_t r = kl - qd * D2_CELLS;
  const D2Rec rec = recs[((g >> D2_NC_BITS) + qd) * p.nsig + sg];
  const uint64_t w = ((uint64_t)(rec.y & 0x3FFFu) << 32) | rec.x;
  slot = (rec.y >> 14) + d2_popc64(w & (((uint64_t)1 << r) - 1u));
  return !((w >> r) & 1u) && rem + len < p.cell_w;
}

// ---- host side: construction -------------------------------------------------------------------------------------------------
struct D2Tables {
  D2Params p;
  std::vector<D2Rec> recs;
  std::vector<uint32_t> gtab;        // [n_chrom + 1]  first record of the chromosome (before interleaving) << 17 | last cell
  uint64_t cells_total = 0, cells_point = 0, cells_phantom = 0;      // statistics over the cells of served groups
};

// records needed at cell width w: per chromosome ceil(cells / 46) for each strand signature, plus the all-ones block
inline uint64_t d2_records_needed(const std::vector<int64_t> &cmax, uint32_t nsig, uint64_t w, bool *fits) {
  uint64_t rec = 1;
  *fits = true;
  for (int64_t mx : cmax) {
    if (mx < 1) continue;
    const uint64_t ncell = (uint64_t)mx / w + 2;
    if (ncell - 1 > D2_NC_MASK) *fits = false;
    rec += (ncell + D2_CELLS - 1) / D2_CELLS;
  }
  return rec * nsig;
}

// goff / points: the RANK engine's groups (g = chromosome * n_class + class; every non-empty group ends with an INT32_MAX sentinel).
// cls_sig[sg]: class of '+' (sg 0) and '-' (sg 1) reads, -1 if the index has no region of that strand.
// Returns false if no cell width fits max_records.
inline bool d2_build(int32_t n_chrom, int32_t n_class, const int cls_sig[2], const std::vector<int32_t> &goff,
                     const std::vector<int32_t> &points, uint64_t max_records, D2Tables &out, uint32_t force_cell_w = 0) {
  const uint32_t nsig = cls_sig[0] == cls_sig[1] ? 1u : 2u;
  if (n_chrom < 1 || points.size() >= D2_MAX_SLOTS) return false;
  // per (chromosome, signature): served?  largest point
  auto group_of = [&](int32_t c, uint32_t sg) { return cls_sig[sg] < 0 ? -1 : c * n_class + cls_sig[sg]; };
  auto served_max = [&](int32_t c, uint32_t sg) -> int64_t {
    const int g = group_of(c, sg);
    if (g < 0) return 0;
    const int32_t gb = goff[g], ge = goff[g + 1];
    if (ge - gb < 2) return 0;
    const int32_t mx = points[ge - 2];
    return mx >= 1 ? (int64_t)mx : 0;                            // only points <= 0: the general path's business
  };
  std::vector<int64_t> cmax((size_t)n_chrom, 0);
  for (int32_t c = 0; c < n_chrom; c++)
    for (uint32_t sg = 0; sg < nsig; sg++) cmax[c] = std::max(cmax[c], served_max(c, sg));
  // smallest cell width whose records fit
  uint64_t w = force_cell_w;
  if (!w) {
    uint64_t lo = 16, hi = (uint64_t)1 << 30;
    bool fits;
    if (d2_records_needed(cmax, nsig, hi, &fits) > max_records || !fits) return false;
    while (lo < hi) {
      const uint64_t mid = (lo + hi) / 2;
      if (d2_records_needed(cmax, nsig, mid, &fits) <= max_records && fits) hi = mid; else lo = mid + 1;
    }
    w = lo;
  } else {
    bool fits;
    if (d2_records_needed(cmax, nsig, w, &fits) > max_records || !fits) return false;
  }
  D2Params &p = out.p;
  p.cell_w = (uint32_t)w; p.nsig = nsig; p.n_gt = (uint32_t)n_chrom + 1;
  uint32_t l = 0;
  while (((uint64_t)1 << l) < w) l++;                            // 2^(l-1) < w <= 2^l, l >= 1 because w >= 16
  p.magic = (uint32_t)((((uint64_t)1 << (31 + l)) + w - 1) / w); // in [2^31, 2^32)
  p.shift = l - 1;
  out.gtab.assign((size_t)n_chrom + 1, 0u);                      // record block 0 = all ones, last cell 0
  out.recs.assign(nsig, D2Rec{0xFFFFFFFFu, 0x3FFFu});
  out.cells_total = out.cells_point = out.cells_phantom = 0;
  for (int32_t c = 0; c < n_chrom; c++) {
    if (cmax[c] < 1) continue;
    const uint64_t ncell = (uint64_t)cmax[c] / w + 2;
    const uint64_t nrec = (ncell + D2_CELLS - 1) / D2_CELLS;
    const uint64_t rb = out.recs.size() / nsig;
    out.gtab[c] = (uint32_t)(rb << D2_NC_BITS) | (uint32_t)(ncell - 1);
    out.recs.resize(out.recs.size() + nrec * nsig, D2Rec{0xFFFFFFFFu, 0x3FFFu});
    for (uint32_t sg = 0; sg < nsig; sg++) {
      if (served_max(c, sg) < 1) continue;                       // stays all ones
      const int g = group_of(c, sg);
      const int32_t gb = goff[g], ge = goff[g + 1] - 1;          // the group's points without the sentinel
      int32_t j = gb;
      for (uint64_t R = 0; R < nrec; R++) {
        const int64_t first = (int64_t)(R * D2_CELLS * w);
        while (j < ge && (int64_t)points[j] < first) j++;
        uint64_t map = 0;
        uint32_t debt = 0;
        int32_t t = j;
        for (uint32_t r = 0; r < D2_CELLS; r++) {
          const uint64_t k = R * D2_CELLS + r;
          if (k >= ncell) break;                                 // never looked at: the last cell is where lookups are clamped to
          const int64_t hi = (int64_t)((k + 1) * w);
          uint32_t m = 0;
          while (t < ge && (int64_t)points[t] < hi) { t++; m++; }
          out.cells_total++;
          if (m) { map |= (uint64_t)1 << r; debt += m - 1; out.cells_point++; }
          else if (debt) { map |= (uint64_t)1 << r; debt--; out.cells_phantom++; }
        }
        D2Rec &rec = out.recs[(rb + R) * nsig + sg];
        rec.x = (uint32_t)map;
        rec.y = (uint32_t)(map >> 32) | ((uint32_t)j << 14);     // j < 2^18: checked above
      }
    }
  }
  p.n_rec = (uint32_t)out.recs.size();
  return out.recs.size() <= max_records && (out.recs.size() / nsig) < ((uint64_t)1 << (32 - D2_NC_BITS));
}
