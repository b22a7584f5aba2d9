"""ctypes view of libgtb200.so (include/gtb200.h) for tests, bench.py and the multi-GPU driver.

This is plumbing only: every number is computed by the CUDA library.  There is no CPU fallback -- if
the shared library is missing or no CUDA device is present, import / Context() raise."""
import ctypes
import json
import os
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.dirname(os.path.dirname(_HERE))
LIB_PATH = os.environ.get("GTB200_LIB") or os.path.join(PKG_ROOT, "lib", "libgtb200.so")     # override: kernel A/B experiments

MATCH_GAPS = 1 << 0
IGNORE_STRAND = 1 << 1
SORTED_RULES = 1 << 2
MEM_HOST = 0
MEM_DEVICE = 1 << 8
ENGINE_AUTO = 0
ENGINE_ENUMERATE = 1 << 16
ENGINE_RANK = 1 << 17
ENGINE_BUCKET = 1 << 19
ENGINE_DIRECT = 1 << 20
OP_COUNT = 0
OP_COVERAGE = 1

OK = 0
ERR_ARG = 1
ERR_QUERY_STOP_NONPOSITIVE = 2
ERR_QUERY_START_GT_STOP = 3
ERR_QUERY_REGION = 4
ERR_INDEX_REGION = 5
ERR_WINDOW = 6
ERR_UNSUPPORTED = 7
ERR_NO_DEVICE = 100

EXPORTS = [
    "gtb_abi_version", "gtb_ctx_create", "gtb_ctx_destroy", "gtb_ctx_set_stream", "gtb_ctx_get_stream", "gtb_ctx_synchronize",
    "gtb_ctx_last_error", "gtb_ctx_launch_count", "gtb_ctx_transfer_stats", "gtb_ctx_profile", "gtb_ctx_profile_report",
    "gtb_index_create", "gtb_index_destroy", "gtb_index_reset", "gtb_index_add_queries", "gtb_index_add_packed", "gtb_index_finish",
    "gtb_index_finish_async", "gtb_index_status", "gtb_index_query_counts", "gtb_index_query_matches",
    "gtb_overlap_count", "gtb_overlap_coverage",
    "gtb_mgpu_create", "gtb_mgpu_destroy", "gtb_mgpu_device_count", "gtb_mgpu_ctx", "gtb_mgpu_last_error", "gtb_mgpu_index_create",
    "gtb_mgpu_index_destroy", "gtb_mgpu_index_reset", "gtb_mgpu_index_add_queries", "gtb_mgpu_index_add_packed", "gtb_mgpu_index_finish",
    "gtb_scan_create", "gtb_scan_destroy", "gtb_scan_reset", "gtb_scan_add_reads", "gtb_scan_finish", "gtb_scan_fetch",
    "gtb_scan_peaks", "gtb_scan_peaks_fetch",
    "gtb_synth_reads", "gtb_synth_reads_range", "gtb_gather_u64", "gtb_sort_regions", "gtb_link_regions",
]


class GtbError(RuntimeError):
    def __init__(self, code, message, index=-1):
        super().__init__("gtb200 error %d: %s (index %d)" % (code, message, index))
        self.code = code
        self.index = index


class _Set(ctypes.Structure):
    _fields_ = [("n_regions", ctypes.c_int64), ("n_intervals", ctypes.c_int64),
                ("chrom", ctypes.c_void_p), ("start", ctypes.c_void_p), ("stop", ctypes.c_void_p),
                ("strand", ctypes.c_void_p), ("weight", ctypes.c_void_p), ("region_offset", ctypes.c_void_p)]


class _Packed(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int64), ("start", ctypes.c_void_p), ("meta", ctypes.c_void_p), ("read_len", ctypes.c_int32)]


class _ScanParams(ctypes.Structure):
    _fields_ = [("win_step", ctypes.c_int64), ("win_size", ctypes.c_int64), ("min_reads", ctypes.c_int64),
                ("op", ctypes.c_int32), ("ignore_strand", ctypes.c_int32), ("emulate_sorted", ctypes.c_int32),
                ("reserved", ctypes.c_int32)]


def load_library(path=LIB_PATH):
    if not os.path.exists(path):
        raise ImportError("libgtb200.so not built (%s): run `make -C ibm-cbc-genomic-tools_b200 lib` or "
                          "__graft_entry__.build(); there is no CPU fallback" % path)
    lib = ctypes.CDLL(path)
    vp, i64, u32, ci = ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint, ctypes.c_int
    P = ctypes.POINTER
    sig = {
        "gtb_abi_version": (ci, []),
        "gtb_ctx_create": (ci, [ci, P(vp)]),
        "gtb_ctx_destroy": (None, [vp]),
        "gtb_ctx_set_stream": (ci, [vp, vp]),
        "gtb_ctx_get_stream": (vp, [vp]),
        "gtb_ctx_synchronize": (ci, [vp]),
        "gtb_ctx_last_error": (ctypes.c_char_p, [vp]),
        "gtb_ctx_launch_count": (i64, [vp]),
        "gtb_ctx_transfer_stats": (ci, [vp, P(i64), P(i64), P(i64), P(i64)]),
        "gtb_ctx_profile": (ci, [vp, ci]),
        "gtb_ctx_profile_report": (ci, [vp, ctypes.c_char_p, ctypes.c_size_t]),
        "gtb_index_create": (ci, [vp, P(_Set), ci, u32, P(vp), P(i64)]),
        "gtb_index_destroy": (None, [vp]),
        "gtb_index_reset": (ci, [vp]),
        "gtb_index_add_queries": (ci, [vp, P(_Set), u32]),
        "gtb_index_add_packed": (ci, [vp, P(_Packed), u32]),
        "gtb_index_finish": (ci, [vp, vp, u32, P(i64)]),
        "gtb_index_finish_async": (ci, [vp, vp, u32]),
        "gtb_index_status": (ci, [vp, P(i64)]),
        "gtb_index_query_counts": (ci, [vp, P(_Set), u32, vp, u32, P(i64)]),
        "gtb_index_query_matches": (ci, [vp, P(_Set), u32, vp, ci, vp, vp, P(i64)]),
        "gtb_overlap_count": (ci, [vp, P(_Set), u32, P(_Set), u32, vp, P(i64)]),
        "gtb_overlap_coverage": (ci, [vp, P(_Set), u32, P(_Set), u32, vp, P(i64)]),
        "gtb_mgpu_create": (ci, [ci, P(ci), P(vp)]),
        "gtb_mgpu_destroy": (None, [vp]),
        "gtb_mgpu_device_count": (ci, [vp]),
        "gtb_mgpu_ctx": (vp, [vp, ci]),
        "gtb_mgpu_last_error": (ctypes.c_char_p, [vp]),
        "gtb_mgpu_index_create": (ci, [vp, P(_Set), ci, u32, P(vp), P(i64)]),
        "gtb_mgpu_index_destroy": (None, [vp]),
        "gtb_mgpu_index_reset": (ci, [vp]),
        "gtb_mgpu_index_add_queries": (ci, [vp, P(_Set)]),
        "gtb_mgpu_index_add_packed": (ci, [vp, P(_Packed)]),
        "gtb_mgpu_index_finish": (ci, [vp, vp, P(i64)]),
        "gtb_scan_create": (ci, [vp, ctypes.c_int32, vp, P(_ScanParams), P(vp)]),
        "gtb_scan_destroy": (None, [vp]),
        "gtb_scan_reset": (ci, [vp]),
        "gtb_scan_add_reads": (ci, [vp, P(_Set), u32]),
        "gtb_scan_finish": (ci, [vp, P(i64)]),
        "gtb_scan_fetch": (ci, [vp, i64, i64, vp, vp, vp, vp]),
        "gtb_gather_u64": (ci, [vp, vp, vp, i64, vp, vp]),
        "gtb_sort_regions": (ci, [vp, i64, vp, vp, vp, vp, ci, vp]),
        "gtb_link_regions": (ci, [vp, i64, vp, vp, vp, i64, P(i64), vp, vp]),
        "gtb_synth_reads": (ci, [vp, ctypes.c_uint64, i64, i64, ctypes.c_int32, ctypes.c_int32, vp, vp, vp, vp, vp]),
        "gtb_synth_reads_range": (ci, [vp, ctypes.c_uint64, i64, i64, ctypes.c_int32, ctypes.c_int32, vp, ctypes.c_uint64,
                                       ctypes.c_uint64, vp, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = load_library()
    return _lib


def _np_ptr(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def host_set(s, weight=None, offsets=None, per_region=1):
    """dict of numpy arrays -> (gtb_set with host pointers, keepalive).  per_region = k > 1 (and no offsets): regions of k intervals each"""
    chrom = np.ascontiguousarray(s["chrom"], dtype=np.int32)
    start = np.ascontiguousarray(s["start"], dtype=np.int32)
    stop = np.ascontiguousarray(s["stop"], dtype=np.int32)
    strand = np.ascontiguousarray(s["strand"], dtype=np.int8)
    w = None if weight is None else np.ascontiguousarray(weight, dtype=np.int32)
    off = None if offsets is None else np.ascontiguousarray(offsets, dtype=np.int64)
    n_reg = len(off) - 1 if off is not None else len(chrom) // per_region
    st = _Set(n_reg, len(chrom), _np_ptr(chrom), _np_ptr(start), _np_ptr(stop), _np_ptr(strand), _np_ptr(w), _np_ptr(off))
    return st, (chrom, start, stop, strand, w, off)


def device_set(t, weight=None, offsets=None, per_region=1):
    """dict of torch CUDA tensors (int32/int32/int32/int8) -> (gtb_set with device pointers, keepalive).
    per_region = k > 1 (and no offsets): regions of k intervals each"""
    import torch
    assert t["chrom"].dtype == torch.int32 and t["start"].dtype == torch.int32 and t["stop"].dtype == torch.int32
    assert t["strand"].dtype == torch.int8 and t["chrom"].is_cuda
    n = t["chrom"].numel()
    n_reg = offsets.numel() - 1 if offsets is not None else n // per_region
    st = _Set(n_reg, n, t["chrom"].data_ptr(), t["start"].data_ptr(), t["stop"].data_ptr(), t["strand"].data_ptr(),
              None if weight is None else weight.data_ptr(), None if offsets is None else offsets.data_ptr())
    return st, (t, weight, offsets)


def pinned_set(t, weight=None, offsets=None):
    """dict of pinned torch CPU tensors -> (gtb_set with host pointers, keepalive)."""
    n = t["chrom"].numel()
    n_reg = offsets.numel() - 1 if offsets is not None else n
    st = _Set(n_reg, n, t["chrom"].data_ptr(), t["start"].data_ptr(), t["stop"].data_ptr(), t["strand"].data_ptr(),
              None if weight is None else weight.data_ptr(), None if offsets is None else offsets.data_ptr())
    return st, (t, weight, offsets)


class Context:
    def __init__(self, device=0):
        self._h = ctypes.c_void_p()
        rc = lib().gtb_ctx_create(device, ctypes.byref(self._h))
        if rc != OK:
            raise GtbError(rc, "gtb_ctx_create failed (no CUDA device? there is no CPU fallback)")
        self.device = device
        self._children = weakref.WeakSet()      # indexes and scans of this context: closed with it (they hold device memory of it)

    def close(self):
        if self._h:
            for child in list(self._children):
                child.close()
            lib().gtb_ctx_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc, index=-1):
        if rc != OK:
            raise GtbError(rc, lib().gtb_ctx_last_error(self._h).decode(), index)

    def set_stream(self, cuda_stream_ptr):
        self.check(lib().gtb_ctx_set_stream(self._h, ctypes.c_void_p(cuda_stream_ptr)))

    def stream_ptr(self):
        """cudaStream_t of the context's kernels, as an integer (torch.cuda.ExternalStream takes it)"""
        return int(lib().gtb_ctx_get_stream(self._h) or 0)

    def synchronize(self):
        self.check(lib().gtb_ctx_synchronize(self._h))

    def launch_count(self):
        return int(lib().gtb_ctx_launch_count(self._h))

    def transfer_stats(self):
        """{'h2d_bytes', 'd2h_bytes', 'packed_chunks', 'raw_chunks'} since the context was created."""
        v = [ctypes.c_int64(0) for _ in range(4)]
        self.check(lib().gtb_ctx_transfer_stats(self._h, *[ctypes.byref(x) for x in v]))
        return dict(zip(("h2d_bytes", "d2h_bytes", "packed_chunks", "raw_chunks"), [int(x.value) for x in v]))

    def profile(self, enable):
        self.check(lib().gtb_ctx_profile(self._h, int(enable)))

    def profile_report(self):
        buf = ctypes.create_string_buffer(1 << 16)
        self.check(lib().gtb_ctx_profile_report(self._h, buf, len(buf)))
        return json.loads(buf.value.decode())

    # one-shot forms ------------------------------------------------------------------------------
    def _one_shot(self, fn, queries, regions, flags, qweight, qoffsets, roffsets):
        qs, k1 = host_set(queries, qweight, qoffsets)
        rs, k2 = host_set(regions, None, roffsets)
        out = np.zeros(rs.n_regions, dtype=np.uint64)
        err = ctypes.c_int64(-1)
        rc = fn(self._h, ctypes.byref(qs), MEM_HOST, ctypes.byref(rs), flags, _np_ptr(out), ctypes.byref(err))
        self.check(rc, err.value)
        return out

    def overlap_count(self, queries, regions, flags=0, qweight=None, qoffsets=None, roffsets=None):
        return self._one_shot(lib().gtb_overlap_count, queries, regions, flags, qweight, qoffsets, roffsets)

    def overlap_coverage(self, queries, regions, flags=0, qweight=None, qoffsets=None, roffsets=None):
        return self._one_shot(lib().gtb_overlap_coverage, queries, regions, flags, qweight, qoffsets, roffsets)

    def sort_regions(self, chrom_rank, start, stop, strand, by_strand=False):
        """permutation that puts regions into `genomic_regions gsort` order (device radix sort); numpy arrays in, int64 array out"""
        cr = np.ascontiguousarray(chrom_rank, dtype=np.int32); st = np.ascontiguousarray(start, dtype=np.int32)
        sp = np.ascontiguousarray(stop, dtype=np.int32); sd = np.ascontiguousarray(strand, dtype=np.int8)
        perm = np.zeros(len(cr), dtype=np.int64)
        self.check(lib().gtb_sort_regions(self._h, len(cr), _np_ptr(cr), _np_ptr(st), _np_ptr(sp), _np_ptr(sd), int(by_strand), _np_ptr(perm)))
        return perm

    def link_regions(self, group_rank, start, stop, max_difference=0):
        """(head index, linked stop) of the linked regions of a sorted stream (`genomic_regions link`); numpy arrays in and out"""
        g = np.ascontiguousarray(group_rank, dtype=np.int32); st = np.ascontiguousarray(start, dtype=np.int32)
        sp = np.ascontiguousarray(stop, dtype=np.int32)
        head = np.zeros(max(len(g), 1), dtype=np.int64); lstop = np.zeros(max(len(g), 1), dtype=np.int32)
        n_linked = ctypes.c_int64(0)
        self.check(lib().gtb_link_regions(self._h, len(g), _np_ptr(g), _np_ptr(st), _np_ptr(sp), int(max_difference), ctypes.byref(n_linked),
                                          _np_ptr(head), _np_ptr(lstop)))
        return head[:n_linked.value], lstop[:n_linked.value]

    def gather_u64(self, table_ptr, index_ptr, n, out_ptr, cuda_stream_ptr=0):
        """out[k] = table[index[k]] on the device (the multi-GPU driver's scatter to file order)"""
        self.check(lib().gtb_gather_u64(self._h, ctypes.c_void_p(table_ptr), ctypes.c_void_p(index_ptr), n, ctypes.c_void_p(out_ptr),
                                        ctypes.c_void_p(cuda_stream_ptr) if cuda_stream_ptr else None))

    def synth_reads(self, seed, first, n, read_len, chrom_len, out, p_range=None):
        """out: dict of preallocated torch CUDA tensors; p_range = (p_lo, p_hi) restricts the effective positions."""
        cl = np.ascontiguousarray(chrom_len, dtype=np.int64)
        p_lo, p_hi = (0, (1 << 64) - 1) if p_range is None else p_range
        self.check(lib().gtb_synth_reads_range(self._h, seed, first, n, read_len, len(cl), _np_ptr(cl), p_lo, p_hi,
                                               out["chrom"].data_ptr(), out["start"].data_ptr(), out["stop"].data_ptr(),
                                               out["strand"].data_ptr()))


class MultiGpu:
    """Several GPUs behind one index in this process (gtb_mgpu_*): host-resident query batches are cut into one slice per device."""

    def __init__(self, devices):
        self._h = ctypes.c_void_p()
        arr = (ctypes.c_int * len(devices))(*devices)
        rc = lib().gtb_mgpu_create(len(devices), arr, ctypes.byref(self._h))
        if rc != OK:
            raise GtbError(rc, "gtb_mgpu_create failed", -1)
        self.n = len(devices)

    def check(self, rc, index=-1):
        if rc != OK:
            raise GtbError(rc, lib().gtb_mgpu_last_error(self._h).decode(), index)

    def transfer_stats(self):
        """(h2d bytes, d2h bytes, packed chunks, raw chunks) summed over the devices' contexts"""
        tot = [0, 0, 0, 0]
        for k in range(self.n):
            v = [ctypes.c_int64(0) for _ in range(4)]
            lib().gtb_ctx_transfer_stats(ctypes.c_void_p(lib().gtb_mgpu_ctx(self._h, k)), *[ctypes.byref(x) for x in v])
            tot = [a + b.value for a, b in zip(tot, v)]
        return tuple(tot)

    def close(self):
        if self._h:
            lib().gtb_mgpu_destroy(self._h)
            self._h = ctypes.c_void_p()


class MultiIndex:
    def __init__(self, mg, regions, op=OP_COUNT, flags=0, roffsets=None):
        self.mg = mg
        self._h = ctypes.c_void_p()
        rs, keep = host_set(regions, None, roffsets)
        err = ctypes.c_int64(-1)
        mg.check(lib().gtb_mgpu_index_create(mg._h, ctypes.byref(rs), op, flags, ctypes.byref(self._h), ctypes.byref(err)), err.value)
        self.n_regions = rs.n_regions

    def reset(self):
        self.mg.check(lib().gtb_mgpu_index_reset(self._h))

    def add_set(self, st):
        self.mg.check(lib().gtb_mgpu_index_add_queries(self._h, ctypes.byref(st)))

    def add_host(self, queries, weight=None, offsets=None):
        st, keep = host_set(queries, weight, offsets)
        self.add_set(st)

    def add_packed_ptr(self, n, start_ptr, meta_ptr, read_len):
        pk = _Packed(n, start_ptr, meta_ptr, read_len)
        self.mg.check(lib().gtb_mgpu_index_add_packed(self._h, ctypes.byref(pk)))

    def finish(self, out=None):
        if out is None:
            out = np.zeros(self.n_regions, dtype=np.uint64)
        err = ctypes.c_int64(-1)
        rc = lib().gtb_mgpu_index_finish(self._h, _np_ptr(out), ctypes.byref(err))
        self.mg.check(rc, err.value)
        return out

    def close(self):
        if self._h:
            lib().gtb_mgpu_index_destroy(self._h)
            self._h = ctypes.c_void_p()


class Index:
    """Device-resident index region set (gtb_index)."""

    def __init__(self, ctx, regions, op=OP_COUNT, flags=0, roffsets=None):
        self.ctx = ctx
        self._h = ctypes.c_void_p()
        rs, keep = host_set(regions, None, roffsets)
        err = ctypes.c_int64(-1)
        rc = lib().gtb_index_create(ctx._h, ctypes.byref(rs), op, flags, ctypes.byref(self._h), ctypes.byref(err))
        ctx.check(rc, err.value)
        self.n_regions = rs.n_regions
        ctx._children.add(self)

    def close(self):
        if self._h:
            lib().gtb_index_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        self.ctx.check(lib().gtb_index_reset(self._h))

    def add_set(self, st, mem):
        self.ctx.check(lib().gtb_index_add_queries(self._h, ctypes.byref(st), mem))

    def add_host(self, queries, weight=None, offsets=None, per_region=1):
        st, keep = host_set(queries, weight, offsets, per_region)
        self.add_set(st, MEM_HOST)
        self.ctx.synchronize()          # numpy temporaries in `keep` must outlive the async copies

    def add_device(self, tensors, weight=None, offsets=None, per_region=1):
        st, keep = device_set(tensors, weight, offsets, per_region)
        self.add_set(st, MEM_DEVICE)

    def add_packed_ptr(self, n, start_ptr, meta_ptr, read_len, mem):
        """reads in the packed form of gtb200.h (int32 start + uint8 chromosome | strand bit, one length); host pointers are
        copied inside the call"""
        pk = _Packed(n, start_ptr, meta_ptr, read_len)
        self.ctx.check(lib().gtb_index_add_packed(self._h, ctypes.byref(pk), mem))

    def query_counts(self, queries, weight=None, offsets=None):
        """per-QUERY overlap counts (uint32, one per query region) for host arrays: the dual of add + finish"""
        st, keep = host_set(queries, weight, offsets)
        out = np.zeros(st.n_regions, dtype=np.uint32)
        err = ctypes.c_int64(-1)
        rc = lib().gtb_index_query_counts(self._h, ctypes.byref(st), MEM_HOST, _np_ptr(out), MEM_HOST, ctypes.byref(err))
        self.ctx.check(rc, err.value)
        return out

    def query_matches(self, queries, offsets=None, bin_bits=None):
        """(match_offset, matches): for every query region of the host arrays the index regions it overlaps, in the order the
        reference's GetOverlap / NextOverlap walk hands them out (bin_bits: the -B list of the Unsorted class)"""
        counts = self.query_counts(queries, offsets=offsets)
        st, keep = host_set(queries, None, offsets)
        off = np.concatenate([[0], np.cumsum(counts.astype(np.int64))]).astype(np.int64)
        matches = np.zeros(max(int(off[-1]), 1), dtype=np.int32)
        bits = None if bin_bits is None else np.ascontiguousarray(bin_bits, dtype=np.int32)
        err = ctypes.c_int64(-1)
        rc = lib().gtb_index_query_matches(self._h, ctypes.byref(st), MEM_HOST, None if bits is None else _np_ptr(bits), 0 if bits is None else len(bits),
                                           _np_ptr(off), _np_ptr(matches), ctypes.byref(err))
        self.ctx.check(rc, err.value)
        return off, matches[:int(off[-1])]

    def finish(self, out=None):
        if out is None:
            out = np.zeros(self.n_regions, dtype=np.uint64)
        err = ctypes.c_int64(-1)
        rc = lib().gtb_index_finish(self._h, _np_ptr(out), MEM_HOST, ctypes.byref(err))
        self.ctx.check(rc, err.value)
        return out

    def finish_ptr(self, ptr, mem):
        err = ctypes.c_int64(-1)
        rc = lib().gtb_index_finish(self._h, ctypes.c_void_p(ptr), mem, ctypes.byref(err))
        self.ctx.check(rc, err.value)

    def finish_async_ptr(self, ptr):
        """enqueue the finish (values to device memory at ptr, valid in stream order); status() reports errors later"""
        self.ctx.check(lib().gtb_index_finish_async(self._h, ctypes.c_void_p(ptr), MEM_DEVICE))

    def status(self):
        err = ctypes.c_int64(-1)
        rc = lib().gtb_index_status(self._h, ctypes.byref(err))
        self.ctx.check(rc, err.value)


class Scan:
    """Device-resident micro-window histogram (gtb_scan)."""

    def __init__(self, ctx, bound, win_step, win_size, op="1", ignore_strand=False, min_reads=10, emulate_sorted=False):
        self.ctx = ctx
        self._h = ctypes.c_void_p()
        b = np.ascontiguousarray(bound, dtype=np.int64)
        prm = _ScanParams(win_step, win_size, min_reads, ord(op), int(ignore_strand), int(emulate_sorted), 0)
        ctx.check(lib().gtb_scan_create(ctx._h, len(b), _np_ptr(b), ctypes.byref(prm), ctypes.byref(self._h)))
        ctx._children.add(self)

    def close(self):
        if self._h:
            lib().gtb_scan_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        self.ctx.check(lib().gtb_scan_reset(self._h))

    def add_set(self, st, mem):
        self.ctx.check(lib().gtb_scan_add_reads(self._h, ctypes.byref(st), mem))

    def add_host(self, reads, weight=None, offsets=None):
        st, keep = host_set(reads, weight, offsets)
        self.add_set(st, MEM_HOST)
        self.ctx.synchronize()

    def add_device(self, tensors, weight=None, offsets=None):
        st, keep = device_set(tensors, weight, offsets)
        self.add_set(st, MEM_DEVICE)

    def finish(self):
        n = ctypes.c_int64(0)
        self.ctx.check(lib().gtb_scan_finish(self._h, ctypes.byref(n)))
        return int(n.value)

    def fetch(self, first=0, count=None, n_total=None):
        if count is None:
            count = n_total - first
        out = {"chrom": np.zeros(count, dtype=np.int32), "strand": np.zeros(count, dtype=np.int8),
               "win": np.zeros(count, dtype=np.int64), "value": np.zeros(count, dtype=np.int64)}
        self.ctx.check(lib().gtb_scan_fetch(self._h, first, count, _np_ptr(out["chrom"]), _np_ptr(out["strand"]),
                                            _np_ptr(out["win"]), _np_ptr(out["value"])))
        return out
