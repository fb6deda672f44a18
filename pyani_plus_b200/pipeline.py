"""The whole hot path over one fixed base stream as a replayable step.

``SourmashStep`` strings the stages together for a stream whose geometry is known up front (the
plan): K1 hash -> finalize (fused with the all-gather when there are several GPUs) -> K2 intersect
-> ANI.  Nothing inside a step reads the device back, so the same launches can be recorded once
into a CUDA graph (``capture``) and replayed with ONE launch per step -- what matters when a step
is short (many GPUs, small runs) and the host would otherwise be the bottleneck.  The status word
(bucket / segment overflow) is read once after the step; an overflow means "re-plan with more slack
and run again", exactly as in the eager path.

Reference scope: this is ``prepare_genomes`` + ``compute_sourmash_tile`` for a whole run
(``pyani_plus/methods/sourmash.py:34-84,147-206``) without the files in between.
"""

from __future__ import annotations

from . import engine as _engine


class SourmashStep:
    """One pass of the hot path for a planned stream on this rank."""

    def __init__(self, eng, plan, bufs: dict, tab: dict, k: int, *, world: int = 1, rank: int = 0,  # noqa: ANN001, PLR0913
                 gather=None, size_hint: int | None = None, h_ascii=None, k2_method: str = "auto",  # noqa: ANN001
                 host_threads: int = 0) -> None:
        """``gather`` is a ``multi_gpu.SymmetricGather`` (fused finalize + all-gather) or None (NCCL
        all-gather when ``world > 1``); ``h_ascii`` the pinned ASCII stream for host-input steps;
        ``k2_method`` as in ``Engine.intersect`` -- ``"auto"`` is resolved ONCE, on the first step, from
        that step's data (one extra checked intersect), and then kept (a captured graph cannot branch)."""
        self.k2_method = k2_method
        self.eng, self.plan, self.bufs, self.tab, self.k = eng, plan, bufs, tab, k
        self.world, self.rank, self.gather = world, rank, gather
        self.size_hint = plan.sketch_size_hint() if size_hint is None else size_hint
        self.h_ascii = h_ascii
        self.host_threads = host_threads  # host-input steps: threads of the ingest pool to use (0 = all)
        self.out: dict = {}
        self._graphs: dict = {}
        self._pinned: dict = {}

    # ------------------------------------------------------------------ the launches of one step
    def enqueue(self, *, from_host: bool = False, marks: list | None = None, to_host: bool = False) -> None:
        """Enqueue every launch of one step on the current stream (no host synchronisation).

        ``marks``: three CUDA events recorded after the sketch, gather and intersect stages.
        ``to_host``: also enqueue the device->host copies of identity / cov_query / sketch sizes into
        pinned buffers (``out["identity_host"]`` ...), valid after the stream is synchronised.
        """
        from . import multi_gpu  # noqa: PLC0415

        eng, plan, bufs, tab, k = self.eng, self.plan, self.bufs, self.tab, self.k
        if from_host:  # ingest pipeline: host threads pack, packed chunks cross PCIe, K1 follows chunk by chunk
            eng.sketch_host(self.h_ascii, plan, bufs, tab, k, finalize=self.gather is None, threads=self.host_threads)
        elif self.gather is not None:
            eng.hash_packed(plan, bufs, tab, k)
        else:
            eng.sketch_packed(plan, bufs, tab, k)
        if marks is not None:
            marks[0].record()
        if self.gather is not None:
            all_rows, all_counts = self.gather.gather(eng, plan, tab)
        else:
            all_rows, all_counts = multi_gpu.all_gather_tables(tab["table"], tab["counts"], self.world)
        table = _engine.SketchTable(all_rows, all_counts, k, plan.scaled)
        if marks is not None:
            marks[1].record()
        if self.k2_method == "auto":  # first step only: let the engine choose from this data, then stick to it
            eng.intersect(table, rank=self.rank, world=self.world, max_count=self.size_hint, method="auto")
            self.k2_method = eng.last_intersect_method
            if self.world > 1:  # all ranks must take the same path (their partial matrices are summed)
                import torch.distributed as dist  # noqa: PLC0415

                flag = eng.torch.tensor([1 if self.k2_method == "index" else 0], device=eng.device)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
                self.k2_method = "index" if int(flag.item()) else "probe"
        # K2's shared memory is sized from the genome lengths (the kernel verifies it): no read-back
        ov = eng.intersect(table, rank=self.rank, world=self.world, max_count=self.size_hint, check=False,
                           method=self.k2_method)
        if marks is not None:
            marks[2].record()
        # With several ranks ``ov`` is this rank's PARTIAL matrix (the pairs / hash range it owns; the ranks'
        # matrices sum to the whole), so identity / cov_query are partial too: NaN -- which in a complete
        # matrix means "no common hash" -- also marks every pair another rank computed.  Callers that want
        # one complete result sum the integer matrices first (``multi_gpu.combine_partial``) and derive ANI
        # from the sum (``run.intersect_sharded`` + ``engine.ani_host``); never sum the float matrices.
        ident, cov = eng.ani_device(ov, table)
        self.out.update(table=table, ov=ov, identity=ident, cov_query=cov, partial=self.world > 1)
        if to_host:
            torch = eng.torch
            for name, t in (("identity", ident), ("cov_query", cov), ("counts", table.counts)):
                buf = self._pinned.get(name)
                if buf is None or buf.shape != t.shape:
                    buf = self._pinned[name] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                buf.copy_(t, non_blocking=True)
                self.out[name + "_host"] = buf

    def finish(self) -> None:
        """Synchronise and turn the step's status bits into errors."""
        st = self.eng.check_status()
        if st & _engine.ST_BUCKET_OVERFLOW:
            msg = "a sketch bucket overflowed: re-plan the stream with more slack"
            raise _engine.EngineError(msg)
        if st & _engine.ST_SEGMENT_OVERFLOW:
            msg = "a sketch is larger than size_hint: re-plan the step with a larger hint"
            raise _engine.EngineError(msg)

    def run(self, *, from_host: bool = False, to_host: bool = False) -> dict:
        """One eager step, checked; returns ``out`` (device tensors, plus pinned copies if asked)."""
        self.enqueue(from_host=from_host, to_host=to_host)
        self.finish()
        return self.out

    # ------------------------------------------------------------------ CUDA graph form
    def capture(self, *, from_host: bool = False, to_host: bool = False) -> bool:
        """Record one step into a CUDA graph (after one eager warm-up step).  Returns False when the
        capture failed (the eager path stays usable); with several ranks every rank must call this
        and agree on the outcome before replaying (see ``bench.py``)."""
        torch = self.eng.torch
        key = (from_host, to_host)
        if from_host:  # the host threads pack inside the step: a replayed graph would skip that work
            return False
        self.run(from_host=from_host, to_host=to_host)  # sets kernel attributes, allocates pinned buffers
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        self.eng.freeze_workspace()  # the captured kernels hold its address
        try:
            with torch.cuda.graph(graph):
                self.enqueue(from_host=from_host, to_host=to_host)
        except Exception:  # noqa: BLE001
            torch.cuda.synchronize()
            return False
        self._graphs[key] = (graph, dict(self.out))
        return True

    def replay(self, *, from_host: bool = False, to_host: bool = False) -> dict:
        """Launch the captured step (one graph launch); ``finish()`` must follow before results are read."""
        graph, out = self._graphs[(from_host, to_host)]
        graph.replay()
        self.out = out
        return out
