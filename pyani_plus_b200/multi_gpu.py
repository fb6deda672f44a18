"""Multi-GPU form of the sourmash path: one process per GPU, ``torch.distributed`` for the plumbing.

The path shards naturally and has exactly one exchange step:

1. *sketch*    genomes are independent: rank r sketches the contiguous slice ``slice_for_rank`` of the
               genome list (K1, no communication);
2. *exchange*  one all-gather of the fixed-stride sketch rows and of their sizes -- NCCL over
               NVLink/NVSwitch on GPUs, gloo in the CPU tests (the only collective on the data path;
               ~80 KB per 5 Mb genome at scaled=1000, i.e. 0.8 GB for 10,000 genomes);
3. *intersect* every rank holds all sketches; the K2 work items (query row x value cell x block of
               subject columns) are dealt round-robin by item id (``item % world == rank`` inside the
               kernel), so the ranks' overlap matrices are disjoint and SUM to the full matrix;
4. *results*   each rank copies its own partial matrix device->host; ``combine_partial`` (an
               all-reduce / reduce to rank 0) is only needed when one process wants the whole matrix.

Every function takes plain tensors so that the same code runs under gloo on the CPU for the tests
(the CUDA kernels themselves obviously need a GPU).
"""

from __future__ import annotations

import numpy as np


def slice_for_rank(n_items: int, rank: int, world: int) -> tuple[int, int, int]:
    """Contiguous slice [begin, end) of ``n_items`` owned by ``rank`` and the padded per-rank count.

    Every rank gets ``per_rank = ceil(n / world)`` slots so that the all-gather is a plain fixed-size
    one; the real items are spread as evenly as possible (the first ``n % world`` ranks own one more
    than the others) and the unused slots of a rank hold empty dummy genomes.
    """
    if world < 1 or not 0 <= rank < world:
        msg = f"bad rank/world {rank}/{world}"
        raise ValueError(msg)
    per_rank = -(-n_items // world) if n_items else 0
    base, extra = divmod(n_items, world)
    begin = rank * base + min(rank, extra)
    end = begin + base + (1 if rank < extra else 0)
    return begin, end, per_rank


def all_gather_tables(rows, counts, world: int, *, sizes_in_last_slot: bool = True):  # noqa: ANN001, ANN201
    """All-gather this rank's ``[per_rank, stride]`` sketch rows and ``[per_rank]`` sizes.

    Returns ``([world * per_rank, stride] rows, [world * per_rank] counts)``; row ``r * per_rank + i``
    is genome ``i`` of rank ``r``.  With ``world == 1`` the inputs are returned unchanged.

    The sketch-finalize kernel stores every row's size in the row's last slot, so ONE collective
    moves sketches and sizes together (``sizes_in_last_slot``); otherwise the sizes are gathered by
    a second, tiny collective.  Every rank must pass the same shape: callers establish that once, when
    they set the exchange up (``assert_same_shape`` -- it synchronises, so it cannot live in here, where a
    CUDA graph may be capturing).
    """
    if world == 1:
        return rows, counts
    import torch  # noqa: PLC0415
    import torch.distributed as dist  # noqa: PLC0415

    all_rows = torch.empty((world * rows.shape[0], rows.shape[1]), dtype=rows.dtype, device=rows.device)
    dist.all_gather_into_tensor(all_rows, rows.contiguous())
    if sizes_in_last_slot:
        all_counts = all_rows[:, -1].to(counts.dtype).contiguous()
    else:
        all_counts = torch.empty(world * counts.shape[0], dtype=counts.dtype, device=counts.device)
        dist.all_gather_into_tensor(all_counts, counts.contiguous())
    return all_rows, all_counts


def agree_max(value: int, world: int, device=None) -> int:  # noqa: ANN001
    """MAX of an integer over the ranks (row strides, size hints: every rank must use the same one)."""
    if world == 1:
        return int(value)
    import torch  # noqa: PLC0415
    import torch.distributed as dist  # noqa: PLC0415

    t = torch.tensor([int(value)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(t.item())


def assert_same_shape(rows) -> None:  # noqa: ANN001
    """Every rank must hand the exchange rows of ONE shape: a rank-local stride (each rank plans its own
    genome slice) would make peers write / read rows at the wrong offsets."""
    import torch  # noqa: PLC0415
    import torch.distributed as dist  # noqa: PLC0415

    shape = torch.tensor([rows.shape[0], rows.shape[1], -rows.shape[0], -rows.shape[1]], dtype=torch.int64,
                         device=rows.device)
    dist.all_reduce(shape, op=dist.ReduceOp.MAX)
    lo = (-shape[2:]).tolist()
    hi = shape[:2].tolist()
    if lo != hi:
        msg = (f"ranks disagree on the sketch-table shape (rows x stride between {lo} and {hi}): agree on the "
               "row stride first (run.agree_row_stride / Engine.plan_stream(row_stride=...))")
        raise ValueError(msg)


class SymmetricGather:
    """Gathered sketch table in peer-mapped (symmetric) memory: the fused finalize + all-gather.

    Every rank allocates the same ``[world * per_rank, stride]`` table with
    ``torch.distributed._symmetric_memory`` and exchanges the peer pointers once.  Per step,
    ``Engine.finalize_gather`` makes each rank's finalize kernel store its sorted sketches directly
    into all ranks' tables over NVLink, bracketed by two symmetric-memory barriers (before: nobody is
    still reading the table from the previous step; after: all stores have landed).  No NCCL kernel and
    no second pass over the sketches.  ``available()`` is False when symmetric memory cannot be set up
    (then ``all_gather_tables`` = NCCL is used; that is a transport choice, not a CPU fallback).
    """

    def __init__(self, per_rank: int, stride: int, world: int, rank: int, device) -> None:  # noqa: ANN001
        import torch  # noqa: PLC0415
        import torch.distributed as dist  # noqa: PLC0415
        import torch.distributed._symmetric_memory as symm  # noqa: PLC0415

        self.world, self.rank, self.per_rank, self.stride = world, rank, per_rank, stride
        self.rows = symm.empty((world * per_rank, stride), dtype=torch.int64, device=device)
        self.handle = symm.rendezvous(self.rows, dist.group.WORLD)
        self.peer_ptrs = [int(p) for p in self.handle.buffer_ptrs]
        if len(self.peer_ptrs) != world or self.peer_ptrs[rank] != self.rows.data_ptr():
            msg = "symmetric memory rendezvous returned unexpected pointers"
            raise RuntimeError(msg)

    @classmethod
    def create(cls, per_rank: int, stride: int, world: int, rank: int, device, *,  # noqa: ANN001, ANN206
               allow_nccl: bool = False, logger=None):  # noqa: ANN001
        """The gather object.  All ranks must pass the same ``per_rank`` / ``stride`` (checked).

        A failure to set symmetric memory up is an error: it is logged and re-raised, unless the caller
        explicitly accepts the plain NCCL all-gather as the exchange (``allow_nccl``, what
        ``--nccl-gather`` asks for), in which case None is returned on EVERY rank if any rank failed."""
        import torch  # noqa: PLC0415
        import torch.distributed as dist  # noqa: PLC0415

        if agree_max(stride, world, device) != stride or agree_max(-stride, world, device) != -stride or \
                agree_max(per_rank, world, device) != per_rank:
            msg = f"rank {rank}: per_rank={per_rank} / stride={stride} differ between ranks"
            raise ValueError(msg)
        err = None
        obj = None
        try:
            obj = cls(per_rank, stride, world, rank, device)
        except Exception as exc:  # noqa: BLE001
            err = exc
            if logger is not None:
                logger.warning("symmetric-memory gather unavailable on rank %d: %s", rank, exc)
        ok = torch.tensor([0 if err else 1], device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()):
            return obj
        if allow_nccl:
            return None
        msg = "symmetric-memory gather could not be set up on every rank (pass nccl_gather to use NCCL)"
        raise RuntimeError(msg) from err

    def gather(self, eng, plan, tab: dict):  # noqa: ANN001, ANN201
        """Finalize this rank's rows and scatter them to all ranks; returns (all_rows, all_counts)."""
        import torch  # noqa: PLC0415

        self.handle.barrier(channel=0)  # nobody still reads the table of the previous step
        eng.finalize_gather(plan, tab, self.peer_ptrs, self.rank, self.per_rank)
        self.handle.barrier(channel=1)  # every rank's stores have landed
        return self.rows, self.rows[:, -1].to(torch.int32).contiguous()


def item_owner(item_id: int, world: int) -> int:
    """Rank that processes K2 work item ``item_id`` (mirrors ``id % world`` in intersect_kernel)."""
    return item_id % world


def combine_partial(ov, world: int, *, dst: int | None = None):  # noqa: ANN001, ANN201
    """Sum the ranks' disjoint partial overlap matrices (all-reduce, or reduce to ``dst``)."""
    if world == 1:
        return ov
    import torch.distributed as dist  # noqa: PLC0415

    if dst is None:
        dist.all_reduce(ov, op=dist.ReduceOp.SUM)
    else:
        dist.reduce(ov, dst=dst, op=dist.ReduceOp.SUM)
    return ov


def real_rows(n_items: int, world: int) -> np.ndarray:
    """Indices into the gathered table of the real (non-dummy) genomes, in original order."""
    _, _, per_rank = slice_for_rank(n_items, 0, world)
    idx = []
    for r in range(world):
        begin, end, _ = slice_for_rank(n_items, r, world)
        idx.extend(r * per_rank + i for i in range(end - begin))
    return np.asarray(idx, dtype=np.int64)
