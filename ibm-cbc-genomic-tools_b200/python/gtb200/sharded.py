"""Genome-sharded multi-GPU driver for overlap count / coverage (north-star decomposition).

One process per GPU (torch.distributed; NCCL on the GPUs, gloo in the CPU tests).  The genome is cut into
G contiguous (chromosome, coordinate) ranges balanced by read mass.  Every index region is OWNED by exactly one
shard -- the one whose range contains the region's span start -- and that shard computes the region's final
value, so the merge is a GATHER of per-region values, not a reduction.  A shard therefore needs every query
that can overlap a region it owns: the queries intersecting [range start, max stop of the owned regions].
Queries reaching across a range boundary are replicated to both neighbours (`route`).  There is no data-path
collective besides the final gather (plus a gather of the (index, code) pair of the first fatal query, so that
error behaviour matches the single-GPU engine, genomic_intervals.cpp:5740-5741).

Nothing here computes counts: the per-shard engine is the CUDA library (gtb200.Index).  Tests may inject another
engine factory to exercise the planning / routing / gather logic without a GPU.
"""
import numpy as np

from . import ERR_INDEX_REGION, MEM_DEVICE, OK, OP_COUNT, Context, GtbError, Index

_MARGIN = 2          # empty positions between chromosomes on the linear axis


class ShardPlan:
    """Cut points of the genome and region ownership.

    regions  : dict chrom/start/stop/strand (+ roffsets for multi-interval regions), file order
    n_shards : G
    read_hist: optional (bin_width, {chrom id: int64 array of read counts per bin}) -- a coarse histogram of
               read START positions used to balance the shards by read count; without it the cut points
               balance genome length (exact for uniform reads)
    """

    def __init__(self, regions, n_shards, roffsets=None, read_hist=None, chrom_extent=None):
        chrom = np.asarray(regions["chrom"], dtype=np.int64)
        start = np.asarray(regions["start"], dtype=np.int64)
        stop = np.asarray(regions["stop"], dtype=np.int64)
        off = np.arange(len(chrom) + 1, dtype=np.int64) if roffsets is None else np.asarray(roffsets, dtype=np.int64)
        self.n_regions = len(off) - 1
        self.n_shards = int(n_shards)
        first, last = off[:-1], np.maximum(off[1:] - 1, off[:-1])
        has = off[1:] > off[:-1]
        r_chrom = np.where(has, chrom[np.minimum(first, max(len(chrom) - 1, 0))], 0) if len(chrom) else np.zeros(self.n_regions, np.int64)
        r_start = np.where(has, start[np.minimum(first, max(len(chrom) - 1, 0))], 1) if len(chrom) else np.ones(self.n_regions, np.int64)
        r_stop = np.where(has, stop[np.minimum(last, max(len(chrom) - 1, 0))], 0) if len(chrom) else np.zeros(self.n_regions, np.int64)
        n_chrom = int(r_chrom.max()) + 1 if self.n_regions else 1
        # linear axis: chromosome c occupies [cum[c], cum[c] + extent[c] + 1]; position p of c maps to cum[c] + clip(p, 0, extent[c] + 1)
        extent = np.zeros(n_chrom, dtype=np.int64)
        if self.n_regions:
            np.maximum.at(extent, r_chrom, np.maximum(r_stop, 0))
        if chrom_extent is not None:
            ce = np.asarray(chrom_extent, dtype=np.int64)
            if len(ce) > n_chrom:
                extent = np.concatenate([extent, np.zeros(len(ce) - n_chrom, dtype=np.int64)])
                n_chrom = len(ce)
            extent[:len(ce)] = np.maximum(extent[:len(ce)], ce)
        self.n_chrom = n_chrom
        self.extent = extent
        self.cum = np.concatenate([[0], np.cumsum(extent + 1 + _MARGIN)]).astype(np.int64)
        total = int(self.cum[-1])
        # ---- cut points
        if read_hist is None:
            cuts = [(total * s) // self.n_shards for s in range(self.n_shards + 1)]
        else:
            width, per_chrom = read_hist
            pos, mass = [], []
            for c in range(n_chrom):
                h = np.asarray(per_chrom.get(c, []), dtype=np.int64)
                if len(h) == 0:
                    continue
                p = self.cum[c] + np.minimum((np.arange(len(h), dtype=np.int64) + 1) * width, extent[c] + 1)   # right edge of each bin
                pos.append(p); mass.append(h)
            if pos:
                pos = np.concatenate(pos); mass = np.concatenate(mass)
                cs = np.cumsum(mass)
                tot = int(cs[-1]) if len(cs) else 0
            else:
                tot = 0
            cuts = [0]
            for s in range(1, self.n_shards):
                if tot == 0:
                    cuts.append((total * s) // self.n_shards)
                else:
                    k = int(np.searchsorted(cs, (tot * s + self.n_shards - 1) // self.n_shards, side="left"))
                    cuts.append(int(pos[min(k, len(pos) - 1)]))
            cuts.append(total)
            for s in range(1, len(cuts)):
                cuts[s] = max(cuts[s], cuts[s - 1])
        cuts[0], cuts[-1] = -(1 << 62), 1 << 62          # the outer shards are open-ended
        self.lo = np.array(cuts[:-1], dtype=np.int64)     # shard s owns linear positions [lo[s], lo[s+1])
        # ---- ownership: the shard containing the span start
        self.region_lin_start = self.lin(r_chrom, r_start)
        self.region_lin_stop = np.maximum(self.lin(r_chrom, r_stop), self.region_lin_start)
        self.owner = (np.searchsorted(self.lo, self.region_lin_start, side="right") - 1).astype(np.int64)
        self.owned = [np.nonzero(self.owner == s)[0] for s in range(self.n_shards)]          # file order within a shard
        self.max_owned = max((len(o) for o in self.owned), default=0)
        # what a shard must see: [lo[s], reach[s]]
        self.reach = np.array([int(self.region_lin_stop[o].max()) if len(o) else int(self.lo[s]) - 1 for s, o in enumerate(self.owned)],
                              dtype=np.int64)
        self._off = off

    def lin(self, chrom, pos):
        chrom = np.asarray(chrom, dtype=np.int64)
        pos = np.asarray(pos, dtype=np.int64)
        known = (chrom >= 0) & (chrom < self.n_chrom)
        c = np.where(known, chrom, 0)
        v = self.cum[c] + np.clip(pos, 0, self.extent[c] + 1)
        return np.where(known, v, -1)                     # unknown chromosomes overlap nothing; park them in shard 0

    def cut_positions(self):
        """Cut points as (chromosome id, 1-based position) pairs, for generators that work per chromosome."""
        out = []
        for s in range(1, self.n_shards):
            c = int(np.searchsorted(self.cum, self.lo[s], side="right") - 1)
            c = min(max(c, 0), self.n_chrom - 1)
            out.append((c, int(self.lo[s] - self.cum[c])))
        return out

    def subset(self, regions, shard, roffsets=None):
        """The index regions a shard owns (file order kept) as (regions dict, roffsets or None)."""
        ids = self.owned[shard]
        off = self._off
        if roffsets is None:
            return {k: np.ascontiguousarray(np.asarray(regions[k])[ids]) for k in ("chrom", "start", "stop", "strand")}, None
        lens = off[ids + 1] - off[ids]
        new_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        take = np.concatenate([np.arange(off[i], off[i + 1]) for i in ids]) if len(ids) else np.zeros(0, dtype=np.int64)
        return {k: np.ascontiguousarray(np.asarray(regions[k])[take]) for k in ("chrom", "start", "stop", "strand")}, new_off

    def route(self, queries, shard, qoffsets=None):
        """Indices (into the query REGION list) of the queries shard `shard` must process."""
        chrom = np.asarray(queries["chrom"], dtype=np.int64)
        start = np.asarray(queries["start"], dtype=np.int64)
        stop = np.asarray(queries["stop"], dtype=np.int64)
        if qoffsets is not None:
            off = np.asarray(qoffsets, dtype=np.int64)
            first, last = off[:-1], off[1:] - 1
            chrom, start, stop = chrom[first], start[first], stop[last]
        invalid = (stop <= 0) | (start > stop)            # potentially fatal: every shard must see them (chromosome presence decides)
        q_lo = self.lin(chrom, start)
        q_hi = np.maximum(self.lin(chrom, stop), q_lo)
        need = (q_hi >= self.lo[shard]) & (q_lo <= self.reach[shard])
        return np.nonzero(need | invalid)[0]


def _take_queries(queries, ids, qweight=None, qoffsets=None):
    if qoffsets is None:
        sub = {k: np.ascontiguousarray(np.asarray(queries[k])[ids]) for k in ("chrom", "start", "stop", "strand")}
        return sub, (None if qweight is None else np.ascontiguousarray(np.asarray(qweight)[ids])), None
    off = np.asarray(qoffsets, dtype=np.int64)
    lens = off[ids + 1] - off[ids]
    new_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    take = np.concatenate([np.arange(off[i], off[i + 1]) for i in ids]) if len(ids) else np.zeros(0, dtype=np.int64)
    sub = {k: np.ascontiguousarray(np.asarray(queries[k])[take]) for k in ("chrom", "start", "stop", "strand")}
    return sub, (None if qweight is None else np.ascontiguousarray(np.asarray(qweight)[ids])), new_off


def first_malformed_region(regions, roffsets):
    """Index of the first multi-interval region that is not same-chromosome / same-strand / start-sorted /
    non-overlapping (GenomicRegion::IsCompatibleSortedAndNonoverlapping, genomic_intervals.cpp:1153-1161), or -1.
    Checked on every rank before sharding so that all ranks fail alike (the engine would report a shard-local index)."""
    if roffsets is None:
        return -1
    off = np.asarray(roffsets, dtype=np.int64)
    chrom, start, stop, strand = (np.asarray(regions[k]) for k in ("chrom", "start", "stop", "strand"))
    n = len(chrom)
    if n < 2:
        return -1
    same_region = np.ones(n - 1, dtype=bool)
    same_region[off[1:-1][(off[1:-1] > 0) & (off[1:-1] < n)] - 1] = False          # pairs (i, i+1) straddling a region boundary
    bad = same_region & ((chrom[1:] != chrom[:-1]) | (strand[1:] != strand[:-1]) | (start[1:] < start[:-1]) | (start[1:] <= stop[:-1]))
    if not bad.any():
        return -1
    i = int(np.nonzero(bad)[0][0]) + 1
    return int(np.searchsorted(off, i, side="right") - 1)


class CudaShardEngine:
    """The per-shard engine of the product: a gtb200.Index on this rank's GPU."""

    def __init__(self, regions, roffsets, op, flags, device):
        self.ctx = Context(device)
        self.index = Index(self.ctx, regions, op, flags, roffsets=roffsets)

    def add(self, queries, weight, offsets):
        self.index.add_host(queries, weight=weight, offsets=offsets)

    def finish(self):
        """-> (status, local index of the first fatal query or -1, values)"""
        try:
            return OK, -1, self.index.finish()
        except GtbError as e:
            return e.code, e.index, np.zeros(self.index.n_regions, dtype=np.uint64)

    def close(self):
        self.index.close()
        self.ctx.close()


class ShardedOverlap:
    """count / coverage over G ranks.  Every rank calls the same methods with the same index regions; queries may be
    given in full on every rank (each keeps what `route` assigns to it) or pre-routed with add_routed()."""

    def __init__(self, regions, op=OP_COUNT, flags=0, roffsets=None, read_hist=None, chrom_extent=None, group=None,
                 engine_factory=None, device=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        bad = first_malformed_region(regions, roffsets)
        if bad >= 0:
            raise GtbError(ERR_INDEX_REGION, "index regions should be compatible, sorted and non-overlapping!", bad)
        self.plan = ShardPlan(regions, self.world, roffsets=roffsets, read_hist=read_hist, chrom_extent=chrom_extent)
        sub, sub_off = self.plan.subset(regions, self.rank, roffsets)
        if engine_factory is None:
            dev = self.rank if device is None else device
            engine_factory = lambda r, o, op_, fl: CudaShardEngine(r, o, op_, fl, dev)
        self.engine = engine_factory(sub, sub_off, op, flags)
        self.global_ids = []          # per added batch: global stream index of every local query
        self.seen = 0

    def add(self, queries, qweight=None, qoffsets=None):
        """All ranks pass the SAME batch; this rank processes its routed part."""
        ids = self.plan.route(queries, self.rank, qoffsets)
        n = (len(qoffsets) - 1) if qoffsets is not None else len(queries["chrom"])
        sub, w, off = _take_queries(queries, ids, qweight, qoffsets)
        self.global_ids.append(self.seen + ids)
        self.seen += n
        if len(ids):
            self.engine.add(sub, w, off)

    def add_routed(self, queries, global_index, qweight=None, qoffsets=None):
        """This rank passes only the queries it is responsible for (e.g. read from a per-range file);
        global_index[i] = position of query i in the overall stream (for error reporting)."""
        self.global_ids.append(np.asarray(global_index, dtype=np.int64))
        self.engine.add(queries, qweight, qoffsets)

    def finish(self):
        """-> uint64 values in index-file order on every rank (one all-gather); raises GtbError like the single-GPU engine."""
        import torch
        status, local_idx, vals = self.engine.finish()
        gids = np.concatenate(self.global_ids) if self.global_ids else np.zeros(0, dtype=np.int64)
        err_global = int(gids[local_idx]) if status != OK and 0 <= local_idx < len(gids) else (1 << 62)
        pad = self.plan.max_owned
        send = np.zeros(pad + 2, dtype=np.int64)
        send[:len(vals)] = np.asarray(vals, dtype=np.uint64).view(np.int64)
        send[pad] = err_global
        send[pad + 1] = status
        backend = self.dist.get_backend(self.group) if self.dist.is_initialized() else "none"
        dev = "cuda" if backend == "nccl" else "cpu"
        t = torch.from_numpy(send).to(dev)
        if self.world > 1:
            out = torch.empty(self.world * (pad + 2), dtype=torch.int64, device=dev)
            self.dist.all_gather_into_tensor(out, t, group=self.group)            # THE collective
            table = out.cpu().numpy().reshape(self.world, pad + 2)
        else:
            table = send[None, :]
        first = min(range(self.world), key=lambda s: (int(table[s, pad]), s))
        if int(table[first, pad + 1]) != OK and int(table[first, pad]) < (1 << 62):
            raise GtbError(int(table[first, pad + 1]), "fatal query (reported by shard %d)" % first, int(table[first, pad]))
        result = np.zeros(self.plan.n_regions, dtype=np.uint64)
        for s in range(self.world):
            ids = self.plan.owned[s]
            result[ids] = table[s, :len(ids)].view(np.uint64)
        return result

    def close(self):
        self.engine.close()


# ---------------------------------------------------------------------------------------------------------------
# device-resident form (bench.py, pipelines that keep the reads in HBM): same plan, same gather, torch tensors
# ---------------------------------------------------------------------------------------------------------------
def route_mask_torch(plan, t, shard):
    """Boolean CUDA tensor: which single-interval reads of the device set `t` shard `shard` must process (== ShardPlan.route)."""
    import torch
    dev = t["chrom"].device
    cum = torch.from_numpy(plan.cum).to(dev)
    ext = torch.from_numpy(plan.extent).to(dev)
    c = t["chrom"].long()
    known = (c >= 0) & (c < plan.n_chrom)
    cc = torch.where(known, c, torch.zeros_like(c))
    s, e = t["start"].long(), t["stop"].long()
    q_lo = torch.where(known, cum[cc] + torch.minimum(torch.clamp(s, min=0), ext[cc] + 1), torch.full_like(c, -1))
    q_hi = torch.maximum(torch.where(known, cum[cc] + torch.minimum(torch.clamp(e, min=0), ext[cc] + 1), torch.full_like(c, -1)), q_lo)
    invalid = (e <= 0) | (s > e)
    return ((q_hi >= int(plan.lo[shard])) & (q_lo <= int(plan.reach[shard]))) | invalid


class ShardedDeviceOverlap:
    """One rank of the genome-sharded count/coverage with device-resident reads.  step() = reset, stream the local
    reads (own range + replicated boundary reads), finalise the owned regions, ONE all-gather, scatter to file order."""

    def __init__(self, ctx, regions, plan, op=OP_COUNT, flags=0, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        assert plan.n_shards == self.world
        self.plan, self.ctx = plan, ctx
        sub, _ = plan.subset(regions, self.rank, None)
        self.index = Index(ctx, sub, op, flags)
        pad = max(plan.max_owned, 1)
        self.pad = pad
        self.vals = torch.zeros(pad, dtype=torch.int64, device="cuda")
        self.table = torch.zeros(self.world * pad, dtype=torch.int64, device="cuda")
        src = np.concatenate([s * pad + np.arange(len(plan.owned[s])) for s in range(self.world)]) if plan.n_regions else np.zeros(0, np.int64)
        dst = np.concatenate(plan.owned) if plan.n_regions else np.zeros(0, np.int64)
        order = np.argsort(dst, kind="stable")
        self.src_in_file_order = torch.from_numpy(src[order].astype(np.int64)).cuda()      # result[k] = table[src_in_file_order[k]]
        self.result = torch.zeros(plan.n_regions, dtype=torch.int64, device="cuda")

    def step(self, dset, mem, defer_status=False):
        """One pass over the local reads (`dset`: one gtb_set or a list of them -- batches of a stream too long for one call).
        With defer_status the host never waits: the engine's kernels, the collective and the scatter are ordered by stream
        events, and the caller asks for the engine's verdict later with check() (a loop that steps many times and looks at the
        result at the end)."""
        torch = self.torch
        sets = dset if isinstance(dset, (list, tuple)) else [dset]

        def add(st):
            if isinstance(st, dict):                                           # reads in the packed form: {"n", "start", "meta", "read_len"} (pointers)
                self.index.add_packed_ptr(st["n"], st["start"], st["meta"], st["read_len"], mem)
            else:
                self.index.add_set(st, mem)
        if not defer_status:
            self.index.reset()
            for st in sets:
                add(st)
            self.index.finish_ptr(self.vals.data_ptr(), MEM_DEVICE)            # owned values stay on the device; waits, raises GtbError
        else:
            # the library's kernels run on the context's stream, the collective and the scatter on torch's
            lib_stream = torch.cuda.ExternalStream(self.ctx.stream_ptr()) if self.ctx.stream_ptr() else torch.cuda.default_stream()
            lib_stream.wait_stream(torch.cuda.current_stream())               # last step's readers of vals are done before it is rewritten
            self.index.reset()
            for st in sets:
                add(st)
            self.index.finish_async_ptr(self.vals.data_ptr())
            torch.cuda.current_stream().wait_stream(lib_stream)
        cur = torch.cuda.current_stream().cuda_stream or 1           # 0 is torch's name for the legacy default stream: cudaStreamLegacy
        if self.world > 1:
            self.dist.all_gather_into_tensor(self.table, self.vals, group=self.group)        # THE collective
            self.ctx.gather_u64(self.table.data_ptr(), self.src_in_file_order.data_ptr(), self.plan.n_regions, self.result.data_ptr(), cur)
        else:
            self.ctx.gather_u64(self.vals.data_ptr(), self.src_in_file_order.data_ptr(), self.plan.n_regions, self.result.data_ptr(), cur)
        return self.result

    def check(self):
        """the engine's status after steps with defer_status (waits for the stream; raises GtbError like the single-GPU engine)"""
        self.index.status()

    def close(self):
        self.index.close()


# ---------------------------------------------------------------------------------------------------------------
# window counts (genomic_scans counts) over G ranks: SURVEY.md section 8e, third bullet
# ---------------------------------------------------------------------------------------------------------------
class ScanShardPlan:
    """The micro-window grid of a genome, cut into G contiguous pieces of equal size (or of equal read mass, given a
    histogram).  Micro-windows are numbered along the chromosomes in id order, `win_step` bp each; window k of a chromosome
    sums micro-windows k-1 .. k-1+combine-1 (genomic_intervals.cpp:5058-5075), so the rank that owns micro-window m0 = k-1
    needs the reads of micro-windows [m0, m0 + combine): a HALO of combine - 1 micro-windows to the right of its piece,
    never across a chromosome end.  A read is routed by the micro-window of its point (start or centre); reads in a halo go to
    two ranks.  Every rank computes the windows whose FIRST micro-window it owns; there is no exchange besides the final gather."""

    def __init__(self, bound, win_step, win_size, n_shards, micro_mass=None):
        self.bound = np.asarray(bound, dtype=np.int64)
        self.step, self.combine, self.n_shards = int(win_step), int(win_size // win_step), int(n_shards)
        self.n_micro = np.where(self.bound >= 0, self.bound // self.step, 0).astype(np.int64)
        self.cum = np.concatenate([[0], np.cumsum(self.n_micro)]).astype(np.int64)         # first global micro-window of each chromosome
        total = int(self.cum[-1])
        if micro_mass is None:
            cuts = [(total * s) // self.n_shards for s in range(self.n_shards + 1)]
        else:                                                                              # (positions, cumulative read counts) on the global grid
            pos, cs = micro_mass
            tot = int(cs[-1]) if len(cs) else 0
            cuts = [0] + [int(pos[min(int(np.searchsorted(cs, (tot * s) // self.n_shards)), len(pos) - 1)]) if tot else (total * s) // self.n_shards
                          for s in range(1, self.n_shards)] + [total]
            for s in range(1, len(cuts)):
                cuts[s] = max(cuts[s], cuts[s - 1])
        self.cuts = np.array(cuts, dtype=np.int64)                                           # rank s owns global micro-windows [cuts[s], cuts[s+1])

    def micro_of(self, reads, op="1"):
        """Global micro-window of every read's point, -1 if it counts nowhere (genomic_intervals.cpp:5040-5049)."""
        chrom = np.asarray(reads["chrom"], dtype=np.int64)
        s, e = np.asarray(reads["start"], dtype=np.int64), np.asarray(reads["stop"], dtype=np.int64)
        known = (chrom >= 0) & (chrom < len(self.bound))
        c = np.where(known, chrom, 0)
        pos = s if op == "1" else s + (e - s) // 2
        w = (pos - 1) // self.step
        ok = known & (self.bound[c] >= 0) & ~((s > e) | (e <= 0)) & (pos >= 1) & (w < self.n_micro[c])
        return np.where(ok, self.cum[c] + w, -1)

    def route(self, reads, shard, op="1"):
        """Indices of the reads rank `shard` must histogram: those in its piece and in the halo right of it (same chromosome)."""
        m = self.micro_of(reads, op)
        lo, hi = int(self.cuts[shard]), int(self.cuts[shard + 1])
        if hi <= lo:
            return np.zeros(0, dtype=np.int64)
        c_last = int(np.searchsorted(self.cum, hi - 1, side="right") - 1)                   # chromosome of the piece's last micro-window
        halo_hi = min(hi + self.combine - 1, int(self.cum[c_last + 1]))
        return np.nonzero((m >= lo) & (m < halo_hi))[0]

    def owns(self, chrom, win, shard):
        """Whether rank `shard` emits window `win` (1-based) of chromosome `chrom`: it owns the window's first micro-window."""
        g = self.cum[np.asarray(chrom, dtype=np.int64)] + np.asarray(win, dtype=np.int64) - 1
        g = np.minimum(g, max(int(self.cum[-1]) - 1, 0))      # (the spurious window of a trailing chromosome without micro-windows: the last rank's)
        return (g >= self.cuts[shard]) & (g < self.cuts[shard + 1])


class CudaScanEngine:
    """The per-shard engine of the product: a gtb200.Scan on this rank's GPU."""

    def __init__(self, bound, win_step, win_size, op, ignore_strand, min_reads, device):
        from . import Scan
        self.ctx = Context(device)
        self.scan = Scan(self.ctx, bound, win_step, win_size, op, ignore_strand, min_reads)

    def add(self, reads):
        self.scan.add_host(reads)

    def finish(self):
        n = self.scan.finish()
        return self.scan.fetch(0, n)

    def close(self):
        self.scan.close()
        self.ctx.close()


class ShardedScan:
    """genomic_scans counts over G ranks.  Every rank passes the same read batches to add() (each keeps what the plan routes
    to it); finish() returns the qualifying windows in the reference's order (chromosome id, '+' before '-', window) on every
    rank after ONE all-gather.  Unsorted-scanner semantics, the reference's spurious window on chromosomes shorter than one
    window included: every rank's engine emits it, the rank that owns the chromosome's first micro-window keeps it."""

    def __init__(self, bound, win_step, win_size, op="1", ignore_strand=False, min_reads=10, micro_mass=None, group=None,
                 engine_factory=None, device=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.op = op
        self.plan = ScanShardPlan(bound, win_step, win_size, self.world, micro_mass)
        if engine_factory is None:
            dev = self.rank if device is None else device
            engine_factory = lambda *a: CudaScanEngine(*a, dev)
        self.engine = engine_factory(np.asarray(bound, dtype=np.int64), win_step, win_size, op, ignore_strand, min_reads)

    def add(self, reads):
        ids = self.plan.route(reads, self.rank, self.op)
        if len(ids):
            self.engine.add({k: np.ascontiguousarray(np.asarray(reads[k])[ids]) for k in ("chrom", "start", "stop", "strand")})

    def finish(self):
        import torch
        got = self.engine.finish()
        keep = self.plan.owns(got["chrom"], got["win"], self.rank) if len(got["win"]) else np.zeros(0, dtype=bool)
        mine = np.stack([got["chrom"][keep].astype(np.int64), (got["strand"][keep] != ord("+")).astype(np.int64),
                         got["win"][keep].astype(np.int64), got["value"][keep].astype(np.int64)], 1) if keep.any() else np.zeros((0, 4), np.int64)
        if self.world > 1:
            backend = self.dist.get_backend(self.group)
            dev = "cuda" if backend == "nccl" else "cpu"
            sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(self.world)]
            self.dist.all_gather(sizes, torch.tensor([len(mine)], dtype=torch.int64, device=dev), group=self.group)
            pad = max(int(s.item()) for s in sizes)
            buf = torch.zeros((max(pad, 1), 4), dtype=torch.int64, device=dev)
            buf[:len(mine)] = torch.from_numpy(mine).to(dev)
            parts = [torch.zeros_like(buf) for _ in range(self.world)]
            self.dist.all_gather(parts, buf, group=self.group)                              # THE collective
            allw = np.concatenate([p.cpu().numpy()[:int(s.item())] for p, s in zip(parts, sizes)], 0)
        else:
            allw = mine
        order = np.lexsort((allw[:, 2], allw[:, 1], allw[:, 0]))                           # chromosome, strand, window: the reference's order
        allw = allw[order]
        return {"chrom": allw[:, 0].astype(np.int32), "strand": np.where(allw[:, 1] == 0, ord("+"), ord("-")).astype(np.int8),
                "win": allw[:, 2], "value": allw[:, 3]}

    def close(self):
        self.engine.close()
