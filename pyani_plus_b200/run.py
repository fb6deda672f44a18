"""Whole runs of the sourmash path on one or several GPUs: the package's entry for everything that is more
than one kernel call.

One process per GPU (``torchrun`` sets RANK / LOCAL_RANK / WORLD_SIZE); ``torch.distributed`` is the plumbing.
The path has exactly one exchange step (``multi_gpu``): rank r sketches a contiguous slice of the genomes,
the sketches are exchanged once (finalize fused with the all-gather over peer-mapped memory, or NCCL),
every rank intersects the part of the pair work it owns, and the partial count matrices are summed on
rank 0, which alone talks to the database.

* ``DistContext``     who am I, plus the few collectives the host logic needs (broadcast of a job, MAX);
* ``make_step``       a ``pipeline.SourmashStep`` for a planned base stream with everything the ranks must agree
                      on agreed (row stride is the planner's job: ``agree_row_stride``; size hint; exchange);
                      ``bench.py`` runs exactly this;
* ``all_vs_all_files`` FASTA files -> (sorted MD5s, counts, identity, cov_query) for the CLI
                      (``private_cli.compute_sourmash_distributed``), with the ``.sig`` cache honoured.

Reference scope: ``prepare_genomes`` + ``compute_sourmash_tile`` for a whole run
(``pyani_plus/methods/sourmash.py:34-84,147-206``), driven in the reference by ``public_cli.run_method``
(:206-329) through snakemake and ``private_cli.compute_column`` (:757-973).
"""

from __future__ import annotations

import logging
import os
from dataclasses import dataclass
from pathlib import Path

import numpy as np

from pyani_plus_b200 import multi_gpu

JOB_STOP = "stop"


@dataclass
class DistContext:
    """Rank / world of this process and the host-side collectives of a run."""

    world: int = 1
    rank: int = 0
    local_rank: int = 0
    owns_group: bool = False

    @classmethod
    def from_env(cls, *, backend: str | None = None) -> "DistContext":
        """Read torchrun's environment; with WORLD_SIZE > 1 join (or create) the default process group:
        ``nccl`` when CUDA is there, ``gloo`` otherwise (the CPU tests)."""
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if world <= 1:
            return cls()
        import torch  # noqa: PLC0415
        import torch.distributed as dist  # noqa: PLC0415

        rank, local_rank = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
        owns = False
        if not dist.is_initialized():
            use_cuda = torch.cuda.is_available()
            if use_cuda:
                torch.cuda.set_device(local_rank)
            kwargs = {"device_id": torch.device("cuda", local_rank)} if use_cuda else {}
            dist.init_process_group(backend or ("nccl" if use_cuda else "gloo"), **kwargs)
            owns = True
        return cls(world=world, rank=rank, local_rank=local_rank, owns_group=owns)

    def close(self) -> None:
        if self.world > 1 and self.owns_group:
            import torch.distributed as dist  # noqa: PLC0415

            if dist.is_initialized():
                dist.destroy_process_group()

    def barrier(self) -> None:
        if self.world > 1:
            import torch.distributed as dist  # noqa: PLC0415

            dist.barrier()

    def broadcast_object(self, obj, src: int = 0):  # noqa: ANN001, ANN201
        """Pickle-broadcast a small host object (a job description) from ``src``."""
        if self.world == 1:
            return obj
        import torch.distributed as dist  # noqa: PLC0415

        box = [obj if self.rank == src else None]
        dist.broadcast_object_list(box, src=src)
        return box[0]

    def max_int(self, value: int, device=None) -> int:  # noqa: ANN001
        return multi_gpu.agree_max(value, self.world, device)


def agree_row_stride(eng, tile_off: np.ndarray, scaled: int, ctx: DistContext, slack: float = 1.0):  # noqa: ANN001, ANN201
    """Plan this rank's stream with the row stride ALL ranks need (each rank's own plan follows its own
    genome lengths; the exchange needs one stride)."""
    plan = eng.plan_stream(tile_off, scaled, slack)
    if ctx.world == 1:
        return plan
    stride = ctx.max_int(plan.row_stride, eng.device)
    return plan if stride == plan.row_stride else eng.plan_stream(tile_off, scaled, slack, row_stride=stride)


def make_step(eng, plan, bufs: dict, tab: dict, k: int, ctx: DistContext, *, h_ascii=None,  # noqa: ANN001, PLR0913
              k2_method: str = "auto", nccl_gather: bool = False, host_threads: int = 0,
              logger: logging.Logger | None = None):  # noqa: ANN201
    """The hot path over a planned stream as a replayable ``pipeline.SourmashStep`` on this rank.

    Agrees the sketch-size hint over the ranks and sets the exchange up: the finalize fused with the
    all-gather through peer-mapped symmetric memory, unless ``nccl_gather`` asks for the plain NCCL
    all-gather (a failure to set symmetric memory up is an error otherwise).  Returns (step, exchange name).
    """
    from pyani_plus_b200 import pipeline  # noqa: PLC0415

    size_hint = ctx.max_int(plan.sketch_size_hint(), eng.device)
    gather = None
    exchange = None
    if ctx.world > 1:
        per_rank = plan.n_genomes  # every rank plans the same number of rows (dummies pad the last slices)
        multi_gpu.assert_same_shape(tab["table"])
        if not nccl_gather:
            gather = multi_gpu.SymmetricGather.create(per_rank, plan.row_stride, ctx.world, ctx.rank, eng.device,
                                                      logger=logger)
        exchange = ("finalize fused with the all-gather (peer-memory stores over NVLink)" if gather is not None
                    else "NCCL all-gather of sketch rows")
    step = pipeline.SourmashStep(eng, plan, bufs, tab, k, world=ctx.world, rank=ctx.rank, gather=gather,
                                 size_hint=size_hint, h_ascii=h_ascii, k2_method=k2_method,
                                 host_threads=host_threads)
    return step, exchange


def exchange_rows(eng, local, ctx: DistContext, n_total: int):  # noqa: ANN001, ANN201
    """All-gather compact sketch tables (``engine.SketchTable`` with ``per_rank`` rows of one agreed stride)
    and drop the padding rows: the table of all ``n_total`` genomes in slice order, on every rank."""
    from pyani_plus_b200 import engine  # noqa: PLC0415

    if ctx.world == 1:
        return local
    multi_gpu.assert_same_shape(local.rows)
    rows, counts = multi_gpu.all_gather_tables(local.rows, local.counts, ctx.world, sizes_in_last_slot=False)
    real = eng.torch.from_numpy(multi_gpu.real_rows(n_total, ctx.world)).to(rows.device)
    return engine.SketchTable(rows[real].contiguous(), counts[real].contiguous(), local.k, local.scaled)


def intersect_sharded(eng, table, ctx: DistContext, *, method: str = "auto"):  # noqa: ANN001, ANN201
    """All-vs-all counts with the pair work sharded over the ranks; the sum lands on rank 0 (other ranks get
    their partial matrix back).  All ranks take the same K2 form."""
    if ctx.world == 1:
        return eng.intersect(table, method=method)
    max_count = ctx.max_int(int(table.counts.max().item()) if table.n else 0, eng.device)
    if method == "auto":
        pairs = table.n * (table.n - 1) / 2 / ctx.world
        method = "index" if pairs * 2 * max_count * eng.COST_PROBE_PER_ELEMENT > 3e-4 and table.n >= 2 else "probe"
    ov = eng.intersect(table, rank=ctx.rank, world=ctx.world, max_count=max_count, method=method)
    return multi_gpu.combine_partial(ov, ctx.world, dst=0)


def all_vs_all_files(logger: logging.Logger, ctx: DistContext, entries: list[tuple[str, str]], ksize: int,  # noqa: PLR0913
                     scaled: int, sig_cache: Path, *, k2_method: str = "auto"):  # noqa: ANN201
    """Sketch (or load from the cache) and intersect the genomes ``entries`` = [(md5, FASTA path)] over all
    ranks.  Every rank calls this with the same arguments.  Rank 0 returns (md5s sorted, sketch sizes,
    ov uint32 [n, n], identity, cov_query float64 [n, n] with NaN = no common hash); other ranks None.
    """
    from pyani_plus_b200 import engine  # noqa: PLC0415
    from pyani_plus_b200.methods import sourmash  # noqa: PLC0415

    entries = sorted(entries)
    n = len(entries)
    eng = sourmash.get_engine()
    begin, end, per_rank = multi_gpu.slice_for_rank(n, ctx.rank, ctx.world)
    mine = entries[begin:end]
    msg = f"rank {ctx.rank}/{ctx.world}: genomes {begin}..{end} of {n}"
    logger.debug(msg)
    sketches = sourmash.sketches_for(logger, mine, ksize, scaled, sig_cache)  # cache, else FASTA -> GPU -> cache
    local_max = max((len(s) for s in sketches), default=0)
    stride = ctx.max_int(max(16, (local_max + 15) // 16 * 16), eng.device)
    local = eng.table_from_host(sketches, ksize, scaled, stride=stride, rows=per_rank if ctx.world > 1 else None)
    table = exchange_rows(eng, local, ctx, n)
    ov = intersect_sharded(eng, table, ctx, method=k2_method)
    if ctx.rank != 0:
        return None
    counts = table.counts.cpu().numpy()
    ov_h = ov.cpu().numpy().astype(np.uint32, copy=False)
    identity, cov_query = engine.ani_host(ov_h, counts, counts, ksize)
    return [md5 for md5, _ in entries], counts, ov_h, identity, cov_query
