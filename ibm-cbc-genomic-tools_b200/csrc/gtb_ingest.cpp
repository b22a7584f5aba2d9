// gtb_ingest.cpp -- host side of the host->device path of query batches (plain C++, compiled by g++).
//
// The C ABI takes queries as 13 bytes per interval (int32 chrom, start, stop + int8 strand).  Over PCIe that is the
// whole cost of an end-to-end call: the kernels need ~1 ms for 100 M reads, the copy ~24 ms.  Almost every real
// batch is far more regular than the layout allows -- chromosome ids are small, reads are short, strands are '+'/'-' --
// so a pool of host threads re-encodes each chunk into 5, 6 or 8 bytes per interval while the previous chunk is on the wire:
//     start (int32, as is)   +   meta of 1, 2 or 4 bytes -- the widest: (stop - start) | chrom << 16 | (strand == '-') << 30
// and a small device kernel expands it again next to the engine (gtb_overlap.cu: unpack_kernel).  A chunk holding
// anything the packed form cannot express (chrom >= 16384, stop - start outside [0, 65535], a strand byte other than
// '+'/'-') is sent in the plain layout instead, so results never depend on this path.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

struct gtb_ingest {
  std::vector<std::thread> workers;
  std::mutex mu;
  std::condition_variable cv_work, cv_done;
  std::function<void(int)> job;          // job(part) for part in [0, parts)
  int parts = 0, next_part = 0, running = 0;
  uint64_t generation = 0;
  bool stop = false;
};

static void worker_main(gtb_ingest *p) {
  uint64_t seen = 0;
  for (;;) {
    std::unique_lock<std::mutex> lk(p->mu);
    p->cv_work.wait(lk, [&] { return p->stop || (p->generation != seen && p->next_part < p->parts); });
    if (p->stop) return;
    while (p->next_part < p->parts) {
      const int part = p->next_part++;
      p->running++;
      lk.unlock();
      p->job(part);
      lk.lock();
      p->running--;
    }
    seen = p->generation;
    if (p->running == 0) p->cv_done.notify_all();
  }
}

extern "C++" gtb_ingest *gtb_ingest_create(int threads) {
  if (threads <= 0) return nullptr;
  gtb_ingest *p = new gtb_ingest();
  for (int i = 0; i < threads; i++) p->workers.emplace_back(worker_main, p);
  return p;
}

extern "C++" void gtb_ingest_destroy(gtb_ingest *p) {
  if (!p) return;
  {
    std::lock_guard<std::mutex> lk(p->mu);
    p->stop = true;
  }
  p->cv_work.notify_all();
  for (auto &t : p->workers) t.join();
  delete p;
}

extern "C++" int gtb_ingest_threads(const gtb_ingest *p) { return p ? (int)p->workers.size() : 0; }

static void parallel_for(gtb_ingest *p, int parts, std::function<void(int)> fn) {
  std::unique_lock<std::mutex> lk(p->mu);
  p->job = std::move(fn);
  p->parts = parts; p->next_part = 0; p->running = 0;
  p->generation++;
  p->cv_work.notify_all();
  p->cv_done.wait(lk, [&] { return p->next_part >= p->parts && p->running == 0; });
}

// one contiguous slice; returns non-zero if some query does not fit the packed form
#if defined(__GNUC__) && defined(__x86_64__)
__attribute__((target_clones("avx512f", "avx2", "default")))
#endif
static uint32_t pack_slice(const int32_t *__restrict__ chrom, const int32_t *__restrict__ start, const int32_t *__restrict__ stop,
                           const int8_t *__restrict__ strand, int64_t n, uint32_t *__restrict__ meta, int32_t *__restrict__ start_out) {
  uint32_t bad = 0;
  for (int64_t i = 0; i < n; i++) {
    // all 32-bit lanes so that the loop vectorises: with stop >= start the difference fits an unsigned 32-bit value exactly
    const uint32_t c = (uint32_t)chrom[i];
    const uint32_t d = (uint32_t)stop[i] - (uint32_t)start[i];
    const uint32_t sb = (uint32_t)(uint8_t)strand[i];
    const uint32_t minus = sb == (uint32_t)'-' ? 1u : 0u;
    const uint32_t plus = sb == (uint32_t)'+' ? 1u : 0u;
    bad |= (c >= 16384u ? 1u : 0u) | (stop[i] < start[i] ? 1u : 0u) | (d >= 65536u ? 1u : 0u) | ((plus | minus) ^ 1u);
    meta[i] = (d & 0xFFFFu) | (c << 16) | (minus << 30);
  }
  if (start_out) memcpy(start_out, start, (size_t)n * sizeof(int32_t));
  return bad;
}

// Packs n queries with the pool; start_out may be null (the caller then copies `start` from where it lies).
// Returns 1 if every query fit.
extern "C++" int gtb_ingest_pack(gtb_ingest *p, const int32_t *chrom, const int32_t *start, const int32_t *stop, const int8_t *strand,
                                 int64_t n, uint32_t *meta, int32_t *start_out) {
  if (n <= 0) return 1;
  const int threads = (int)p->workers.size();
  const int64_t grain = 64 * 1024;
  const int parts = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)threads * 4, (n + grain - 1) / grain));
  std::atomic<uint32_t> bad{0};
  parallel_for(p, parts, [&](int part) {
    const int64_t lo = n * part / parts, hi = n * (part + 1) / parts;
    const uint32_t b = pack_slice(chrom + lo, start + lo, stop + lo, strand + lo, hi - lo, meta + lo, start_out ? start_out + lo : nullptr);
    if (b) bad.fetch_or(b, std::memory_order_relaxed);
  });
  return bad.load() == 0;
}

// The narrower forms.  Two bytes per interval next to the start: stop - start < 256, chrom < 128
//     meta16 = (stop - start) | chrom << 8 | (strand == '-') << 15
// One byte: every interval of the chunk has the length of the first (what a sequencing run looks like), chrom < 128
//     meta8 = chrom | (strand == '-') << 7
#if defined(__GNUC__) && defined(__x86_64__)
__attribute__((target_clones("avx512f", "avx2", "default")))
#endif
static uint32_t pack_slice16(const int32_t *__restrict__ chrom, const int32_t *__restrict__ start, const int32_t *__restrict__ stop,
                             const int8_t *__restrict__ strand, int64_t n, uint16_t *__restrict__ meta) {
  uint32_t bad = 0;
  for (int64_t i = 0; i < n; i++) {
    const uint32_t c = (uint32_t)chrom[i];
    const uint32_t d = (uint32_t)stop[i] - (uint32_t)start[i];
    const uint32_t sb = (uint32_t)(uint8_t)strand[i];
    const uint32_t minus = sb == (uint32_t)'-' ? 1u : 0u;
    const uint32_t plus = sb == (uint32_t)'+' ? 1u : 0u;
    bad |= (c >= 128u ? 1u : 0u) | (stop[i] < start[i] ? 1u : 0u) | (d >= 256u ? 1u : 0u) | ((plus | minus) ^ 1u);
    meta[i] = (uint16_t)((d & 0xFFu) | ((c & 0x7Fu) << 8) | (minus << 15));
  }
  return bad;
}

#if defined(__GNUC__) && defined(__x86_64__)
__attribute__((target_clones("avx512f", "avx2", "default")))
#endif
static uint32_t pack_slice8(const int32_t *__restrict__ chrom, const int32_t *__restrict__ start, const int32_t *__restrict__ stop,
                            const int8_t *__restrict__ strand, int64_t n, uint32_t len0, uint8_t *__restrict__ meta) {
  uint32_t bad = 0;
  for (int64_t i = 0; i < n; i++) {
    const uint32_t c = (uint32_t)chrom[i];
    const uint32_t d = (uint32_t)stop[i] - (uint32_t)start[i];
    const uint32_t sb = (uint32_t)(uint8_t)strand[i];
    const uint32_t minus = sb == (uint32_t)'-' ? 1u : 0u;
    const uint32_t plus = sb == (uint32_t)'+' ? 1u : 0u;
    bad |= (c >= 128u ? 1u : 0u) | (d != len0 ? 1u : 0u) | ((plus | minus) ^ 1u);
    meta[i] = (uint8_t)((c & 0x7Fu) | (minus << 7));
  }
  return bad;
}

// Packs n queries into `width` bytes of meta per interval (1, 2 or 4; see above) with the pool; for width 1 *len0 receives the
// common stop - start.  Returns 1 if every query fit.
extern "C++" int gtb_ingest_pack_width(gtb_ingest *p, int width, const int32_t *chrom, const int32_t *start, const int32_t *stop,
                                       const int8_t *strand, int64_t n, void *meta, int32_t *start_out, uint32_t *len0) {
  if (n <= 0) return 1;
  if (width == 4) return gtb_ingest_pack(p, chrom, start, stop, strand, n, (uint32_t *)meta, start_out);
  const int threads = (int)p->workers.size();
  const int64_t grain = 64 * 1024;
  const int parts = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)threads * 4, (n + grain - 1) / grain));
  const uint32_t common = (uint32_t)stop[0] - (uint32_t)start[0];
  if (width == 1 && (stop[0] < start[0])) return 0;
  if (len0) *len0 = common;
  std::atomic<uint32_t> bad{0};
  parallel_for(p, parts, [&](int part) {
    const int64_t lo = n * part / parts, hi = n * (part + 1) / parts;
    const uint32_t b = width == 2 ? pack_slice16(chrom + lo, start + lo, stop + lo, strand + lo, hi - lo, (uint16_t *)meta + lo)
                                  : pack_slice8(chrom + lo, start + lo, stop + lo, strand + lo, hi - lo, common, (uint8_t *)meta + lo);
    if (start_out) memcpy(start_out + lo, start + lo, (size_t)(hi - lo) * sizeof(int32_t));
    if (b) bad.fetch_or(b, std::memory_order_relaxed);
  });
  return bad.load() == 0;
}

