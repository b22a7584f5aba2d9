// gtb_bucket.cu -- fast path: partition queries by genome bucket, rank in shared memory.
// (placeholder until the bucket engine lands: reports "not supported" so the RANK engine serves.)
#include "gtb_overlap.cuh"

struct gtb_bucket_state { int unused; };

bool gtb_bucket_supported(const gtb_index *, const QueryView &) { return false; }
int gtb_bucket_prepare(gtb_index *) { return GTB_OK; }
int gtb_bucket_accumulate(gtb_index *ix, const QueryView &) { return gtb_fail(ix->ctx, GTB_ERR_UNSUPPORTED, "bucket engine not built"); }
int gtb_bucket_flush(gtb_index *) { return GTB_OK; }
void gtb_bucket_destroy(gtb_index *) {}
