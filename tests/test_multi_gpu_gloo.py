"""Host-side logic of the multi-GPU path under gloo, world_size 2, on the CPU.

The CUDA kernels need a GPU; what is tested here is the plumbing around them: genome slicing, the
one all-gather of fixed-stride sketch rows, round-robin ownership of K2 work items (partials are
disjoint and sum to the whole), and the optional reduction.  Sketches / intersections are produced
by the oracle so that the test is about the distribution logic only.
"""

from __future__ import annotations

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pyani_plus_b200 import multi_gpu

N, LENGTH, K, SCALED, SEED = 7, 60_000, 31, 50, 20261017


def test_slices_cover_everything() -> None:
    for n in (0, 1, 7, 8, 100, 101):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                b, e, per = multi_gpu.slice_for_rank(n, r, world)
                assert 0 <= e - b <= per and per * world >= n
                seen.extend(range(b, e))
            assert seen == list(range(n))
            idx = multi_gpu.real_rows(n, world)
            assert len(idx) == n and len(set(idx.tolist())) == n
    with pytest.raises(ValueError, match="bad rank/world"):
        multi_gpu.slice_for_rank(4, 2, 2)
    assert [multi_gpu.item_owner(i, 3) for i in range(5)] == [0, 1, 2, 0, 1]


def _worker(rank: int, world: int, port: int, out: dict) -> None:
    from oracle import oracle

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        begin, end, per = multi_gpu.slice_for_rank(N, rank, world)
        stride = 2048
        rows = torch.zeros((per, stride), dtype=torch.int64)
        counts = torch.zeros(per, dtype=torch.int32)
        for i, g in enumerate(range(begin, end)):  # "K1" on this rank's slice
            h = oracle.sketch_records([oracle.synth_genome(SEED, g, LENGTH)], K, SCALED)
            rows[i, : len(h)] = torch.from_numpy(h.view(np.int64))
            counts[i] = len(h)
            rows[i, -1] = len(h)  # what the finalize kernel leaves in the last slot
        all_rows, all_counts = multi_gpu.all_gather_tables(rows, counts, world)  # the one exchange step
        _, counts2 = multi_gpu.all_gather_tables(rows, counts, world, sizes_in_last_slot=False)
        assert (all_counts == counts2).all()
        assert all_rows.shape == (world * per, stride) and all_counts.shape == (world * per,)
        n_rows = world * per
        sk = [all_rows[r, : all_counts[r]].numpy().view(np.uint64) for r in range(n_rows)]
        ov = torch.zeros((n_rows, n_rows), dtype=torch.int32)
        item = 0
        for i in range(n_rows):  # "K2": items dealt round-robin, exactly one owner each
            for j in range(i + 1, n_rows):
                if multi_gpu.item_owner(item, world) == rank:
                    c = oracle.intersect(sk[i], sk[j])
                    ov[i, j] = ov[j, i] = c
                item += 1
        if rank == 0:
            for i in range(n_rows):
                ov[i, i] = int(all_counts[i])
        partial_nonzero = int((ov != 0).sum())
        whole = multi_gpu.combine_partial(ov.clone(), world)
        if rank == 0:
            out["whole"] = whole.numpy()
            out["counts"] = all_counts.numpy()
            out["partial_nonzero"] = partial_nonzero
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_pipeline() -> None:
    from oracle import oracle

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    with mp.get_context("spawn").Manager() as manager:
        out = manager.dict()
        mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
        whole, counts = out["whole"], out["counts"]
        partial_nonzero = out["partial_nonzero"]
    idx = multi_gpu.real_rows(N, 2)
    want_h, want_c = oracle.synth_sketch_batch(SEED, 0, N, LENGTH, K, SCALED)
    want = oracle.intersect_all(want_h, want_c)
    assert (counts[idx] == want_c).all()
    assert (whole[np.ix_(idx, idx)] == want).all()
    dummy = sorted(set(range(len(counts))) - set(idx.tolist()))
    assert dummy and (counts[dummy] == 0).all() and (whole[dummy] == 0).all()
    assert 0 < partial_nonzero < int((whole != 0).sum())  # each rank really held only a part


def _worker_agreement(rank: int, world: int, port: int, out: dict) -> None:
    """run.DistContext + the agreement helpers: what ranks with DIFFERENT genome lengths must settle first."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world), RANK=str(rank),
                      LOCAL_RANK=str(rank))
    from pyani_plus_b200 import run as run_mod

    ctx = run_mod.DistContext.from_env()  # gloo here (no CUDA)
    try:
        assert (ctx.world, ctx.rank) == (world, rank)
        # each rank planned its own slice: strides differ (longer genomes on rank 1) -> agree on the maximum
        my_stride = 1024 * (3 + rank)
        assert ctx.max_int(my_stride) == 1024 * (2 + world)
        assert multi_gpu.agree_max(7 - rank, world) == 7
        # handing rows of different strides to the exchange is refused on EVERY rank, not silently corrupted
        rows = torch.zeros((4, my_stride), dtype=torch.int64)
        try:
            multi_gpu.assert_same_shape(rows)
            refused = False
        except ValueError as err:
            refused = "disagree on the sketch-table shape" in str(err)
        agreed = torch.zeros((4, ctx.max_int(my_stride)), dtype=torch.int64)
        multi_gpu.assert_same_shape(agreed)  # no error
        # the job hand-over of the CLI: rank 0 broadcasts a description (or "stop")
        job = ctx.broadcast_object({"entries": [("md5", "/x.fna")], "ksize": 31} if rank == 0 else None)
        stop = ctx.broadcast_object(run_mod.JOB_STOP if rank == 0 else None)
        ctx.barrier()
        out[rank] = (refused, job, stop)
    finally:
        ctx.close()


def test_ranks_agree_on_stride_and_job() -> None:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    with mp.get_context("spawn").Manager() as manager:
        out = manager.dict()
        mp.spawn(_worker_agreement, args=(2, port, out), nprocs=2, join=True)
        got = dict(out)
    for rank in (0, 1):
        refused, job, stop = got[rank]
        assert refused
        assert job == {"entries": [("md5", "/x.fna")], "ksize": 31}
        assert stop == "stop"


def test_single_process_context_is_trivial(monkeypatch: pytest.MonkeyPatch) -> None:
    from pyani_plus_b200 import run as run_mod

    monkeypatch.delenv("WORLD_SIZE", raising=False)
    ctx = run_mod.DistContext.from_env()
    assert (ctx.world, ctx.rank) == (1, 0)
    assert ctx.broadcast_object({"a": 1}) == {"a": 1} and ctx.max_int(5) == 5
    ctx.barrier()
    ctx.close()
