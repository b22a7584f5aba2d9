// gtb_mgpu.cu -- several GPUs behind one index, in one process (include/gtb200.h: gtb_mgpu_*).
//
// Count and coverage are sums over queries (GenomicRegionSetOverlaps::CountIndexOverlaps / CalcIndexCoverage,
// genomic_intervals.cpp:5304-5317, :5269-5285), so a batch of host-resident queries can be cut anywhere: device k gets the
// k-th contiguous slice of the batch through its own host link and its own index, and the values are added up at the end.
// Nothing is routed and nothing is exchanged between the devices but M values each; what scales is the host-to-device
// transfer, which is what bounds a single device end to end (5 bytes per read over one link).  One host thread per device and call.
#include "gtb_overlap.cuh"
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

struct gtb_mgpu {
  std::vector<gtb_ctx *> ctx;
  std::string last_error;
};

struct gtb_mgpu_index {
  gtb_mgpu *mg = nullptr;
  std::vector<gtb_index *> ix;
  int64_t n_regions = 0;
  int64_t queries_seen = 0;                       // query regions of all batches so far: the stream-order base of the next one
  std::vector<std::vector<uint64_t>> partial;     // per device, finish
};

namespace {
// runs f(k) for every device on a thread of its own; the first non-OK status (lowest device) is the result
template <class F>
int each_device(gtb_mgpu *mg, F f) {
  const int n = (int)mg->ctx.size();
  std::vector<int> rc((size_t)n, GTB_OK);
  std::vector<std::thread> th;
  for (int k = 1; k < n; k++) th.emplace_back([&, k] { rc[(size_t)k] = f(k); });
  rc[0] = f(0);
  for (auto &t : th) t.join();
  for (int k = 0; k < n; k++)
    if (rc[(size_t)k] != GTB_OK) { mg->last_error = gtb_ctx_last_error(mg->ctx[(size_t)k]); return rc[(size_t)k]; }
  return GTB_OK;
}
}  // namespace

extern "C" int gtb_mgpu_create(int n_devices, const int *devices, gtb_mgpu **out) {
  if (!out || n_devices < 1 || n_devices > 64) return GTB_ERR_ARG;
  *out = nullptr;
  gtb_mgpu *mg = new gtb_mgpu();
  for (int k = 0; k < n_devices; k++) {
    gtb_ctx *c = nullptr;
    const int rc = gtb_ctx_create(devices ? devices[k] : k, &c);
    if (rc != GTB_OK) { gtb_mgpu_destroy(mg); return rc; }
    mg->ctx.push_back(c);
  }
  *out = mg;
  return GTB_OK;
}

extern "C" void gtb_mgpu_destroy(gtb_mgpu *mg) {
  if (!mg) return;
  for (gtb_ctx *c : mg->ctx) gtb_ctx_destroy(c);
  delete mg;
}

extern "C" int gtb_mgpu_device_count(const gtb_mgpu *mg) { return mg ? (int)mg->ctx.size() : 0; }
extern "C" gtb_ctx *gtb_mgpu_ctx(gtb_mgpu *mg, int k) { return mg && k >= 0 && k < (int)mg->ctx.size() ? mg->ctx[(size_t)k] : nullptr; }
extern "C" const char *gtb_mgpu_last_error(const gtb_mgpu *mg) { return mg ? mg->last_error.c_str() : ""; }

extern "C" int gtb_mgpu_index_create(gtb_mgpu *mg, const gtb_set *regions, int op, unsigned flags, gtb_mgpu_index **out, int64_t *err_index) {
  if (!mg || !regions || !out) return GTB_ERR_ARG;
  *out = nullptr;
  gtb_mgpu_index *mi = new gtb_mgpu_index();
  mi->mg = mg; mi->n_regions = regions->n_regions;
  mi->ix.assign(mg->ctx.size(), nullptr);
  mi->partial.resize(mg->ctx.size());
  std::vector<int64_t> err(mg->ctx.size(), -1);
  const int rc = each_device(mg, [&](int k) { return gtb_index_create(mg->ctx[(size_t)k], regions, op, flags, &mi->ix[(size_t)k], &err[(size_t)k]); });
  if (rc != GTB_OK) {
    if (err_index) *err_index = err[0];                                  // (every device sees the same region set)
    gtb_mgpu_index_destroy(mi);
    return rc;
  }
  *out = mi;
  return GTB_OK;
}

extern "C" void gtb_mgpu_index_destroy(gtb_mgpu_index *mi) {
  if (!mi) return;
  for (gtb_index *ix : mi->ix) if (ix) gtb_index_destroy(ix);
  delete mi;
}

extern "C" int gtb_mgpu_index_reset(gtb_mgpu_index *mi) {
  if (!mi) return GTB_ERR_ARG;
  mi->queries_seen = 0;
  return each_device(mi->mg, [&](int k) { return gtb_index_reset(mi->ix[(size_t)k]); });
}

extern "C" int gtb_mgpu_index_add_queries(gtb_mgpu_index *mi, const gtb_set *q) {
  if (!mi || !q) return GTB_ERR_ARG;
  if (q->n_regions < 0 || q->n_intervals < 0) return GTB_ERR_ARG;
  if (q->n_regions == 0) return GTB_OK;
  if (!q->region_offset && q->n_regions != q->n_intervals && (q->n_intervals % q->n_regions != 0 || q->n_intervals < 2 * q->n_regions)) return GTB_ERR_ARG;
  const int64_t uniform_k = q->region_offset ? 0 : q->n_intervals / q->n_regions;      // regions of k intervals each, no offsets (gtb_set)
  const int n = (int)mi->ix.size();
  const int64_t base = mi->queries_seen;
  const int rc = each_device(mi->mg, [&](int k) {
    // slice k: regions [r0, r1); slices start on multiples of 16 regions so that single-interval slices keep the 16-byte
    // alignment the engines' vector loads want
    const int64_t per = ((q->n_regions + n - 1) / n + 15) & ~(int64_t)15;
    const int64_t r0 = std::min(q->n_regions, per * k), r1 = std::min(q->n_regions, per * (k + 1));
    if (r1 <= r0) return (int)GTB_OK;
    gtb_set s = *q;
    std::vector<int64_t> off;
    int64_t i0 = r0 * std::max<int64_t>(uniform_k, 1), i1 = r1 * std::max<int64_t>(uniform_k, 1);
    if (q->region_offset) {
      i0 = q->region_offset[r0]; i1 = q->region_offset[r1];
      off.resize((size_t)(r1 - r0) + 1);
      for (int64_t r = r0; r <= r1; r++) off[(size_t)(r - r0)] = q->region_offset[r] - i0;    // the slice's own CSR, from 0
      s.region_offset = off.data();
    }
    s.n_regions = r1 - r0; s.n_intervals = i1 - i0;
    s.chrom = q->chrom + i0; s.start = q->start + i0; s.stop = q->stop + i0; s.strand = q->strand + i0;
    s.weight = q->weight ? q->weight + r0 : nullptr;
    gtb_index *ix = mi->ix[(size_t)k];
    ix->queries_seen = base + r0;                                        // errors are reported by stream-order index over all slices
    return gtb_index_add_queries(ix, &s, GTB_MEM_HOST);
  });
  mi->queries_seen = base + q->n_regions;
  return rc;
}

extern "C" int gtb_mgpu_index_add_packed(gtb_mgpu_index *mi, const gtb_packed_reads *p) {
  if (!mi || !p || p->n < 0) return GTB_ERR_ARG;
  if (p->n == 0) return GTB_OK;
  const int n = (int)mi->ix.size();
  const int64_t base = mi->queries_seen;
  const int rc = each_device(mi->mg, [&](int k) {
    const int64_t per = ((p->n + n - 1) / n + 15) & ~(int64_t)15;
    const int64_t r0 = std::min(p->n, per * k), r1 = std::min(p->n, per * (k + 1));
    if (r1 <= r0) return (int)GTB_OK;
    gtb_packed_reads s = *p;
    s.n = r1 - r0; s.start = p->start + r0; s.meta = p->meta + r0;
    gtb_index *ix = mi->ix[(size_t)k];
    ix->queries_seen = base + r0;
    return gtb_index_add_packed(ix, &s, GTB_MEM_HOST);
  });
  mi->queries_seen = base + p->n;
  return rc;
}

extern "C" int gtb_mgpu_index_finish(gtb_mgpu_index *mi, uint64_t *out, int64_t *err_index) {
  if (!mi || !out) return GTB_ERR_ARG;
  const int n = (int)mi->ix.size();
  std::vector<int> rcs((size_t)n, GTB_OK);
  std::vector<int64_t> err((size_t)n, -1);
  // every device finishes; a query error on one of them must not keep the others from being heard (the earliest one counts)
  each_device(mi->mg, [&](int k) {
    mi->partial[(size_t)k].resize((size_t)std::max<int64_t>(mi->n_regions, 1));
    rcs[(size_t)k] = gtb_index_finish(mi->ix[(size_t)k], mi->partial[(size_t)k].data(), GTB_MEM_HOST, &err[(size_t)k]);
    return (int)GTB_OK;
  });
  int rc = GTB_OK;
  int64_t first = -1;
  for (int k = 0; k < n; k++) {
    if (rcs[(size_t)k] == GTB_OK) continue;
    const bool query_error = err[(size_t)k] >= 0;
    if (rc == GTB_OK || (query_error && (first < 0 || err[(size_t)k] < first))) {
      rc = rcs[(size_t)k]; first = query_error ? err[(size_t)k] : first;
      mi->mg->last_error = gtb_ctx_last_error(mi->mg->ctx[(size_t)k]);
    }
  }
  if (err_index) *err_index = first;
  if (rc != GTB_OK) return rc;
  for (int64_t r = 0; r < mi->n_regions; r++) {
    uint64_t v = 0;
    for (int k = 0; k < n; k++) v += mi->partial[(size_t)k][(size_t)r];
    out[r] = v;
  }
  return GTB_OK;
}
