"""torchrun worker for tests/test_multi_gpu_cuda.py: the N-GPU path against the single-GPU result.

Each rank sketches its slice, the sketches are exchanged both ways (fused finalize + peer-memory
scatter, and finalize + NCCL all-gather), K2 runs sharded, and rank 0 checks everything against a
single-GPU computation of the same genomes.
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pyani_plus_b200 import engine, multi_gpu, stream as pstream  # noqa: E402

SEED, N, LENGTH, K, SCALED = 20261017, 21, 300_000, 31, 100
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
eng = engine.Engine(int(os.environ["LOCAL_RANK"]))
g0, g1, per_rank = multi_gpu.slice_for_rank(N, rank, world)
tiles_per = LENGTH // pstream.TILE + 1
tile_off = np.zeros(per_rank + 1, dtype=np.int64)
for i in range(per_rank):
    tile_off[i + 1] = tile_off[i] + (tiles_per if i < g1 - g0 else 1)
plan = eng.plan_stream(tile_off, SCALED)
d_ascii = torch.full((plan.n_bases,), pstream.PAD, dtype=torch.uint8, device=eng.device)
gen, _ = eng.synth_ascii_stream(SEED, g0, g1 - g0, LENGTH)
d_ascii[: (g1 - g0) * tiles_per * pstream.TILE] = gen[: (g1 - g0) * tiles_per * pstream.TILE]
bufs = eng.alloc_stream_buffers(plan)
eng.pack(d_ascii, plan, bufs)

results = {}
fused = multi_gpu.SymmetricGather.create(per_rank, plan.row_stride, world, rank, eng.device, allow_nccl=True)
ok = torch.tensor([1 if fused is not None else 0], device=eng.device)
modes = ["nccl"] + (["fused"] * 2 if int(ok.item()) else [])  # fused twice: the table is re-used
for mode in modes:
    tab = eng.alloc_table(plan)
    if mode == "fused":
        eng.hash_packed(plan, bufs, tab, K)
        rows, counts = fused.gather(eng, plan, tab)
    else:
        eng.sketch_packed(plan, bufs, tab, K)
        rows, counts = multi_gpu.all_gather_tables(tab["table"], tab["counts"], world)
    assert eng.check_status() == 0
    table = engine.SketchTable(rows, counts, K, SCALED)
    part = eng.intersect(table, rank=rank, world=world)
    whole = multi_gpu.combine_partial(part.clone(), world)
    results[mode] = (table.to_host(), whole.cpu().numpy(), int((part != 0).sum().item()))

# the same step through pipeline.SourmashStep, eagerly and as a replayed CUDA graph
from pyani_plus_b200 import pipeline  # noqa: E402

tab = eng.alloc_table(plan)
stepper = pipeline.SourmashStep(eng, plan, bufs, tab, K, world=world, rank=rank,
                                gather=fused if int(ok.item()) else None)
out = stepper.run()
results["step_eager"] = (out["table"].to_host(), multi_gpu.combine_partial(out["ov"].clone(), world).cpu().numpy(),
                         int((out["ov"] != 0).sum().item()))
captured = torch.tensor([1 if stepper.capture() else 0], device=eng.device)
dist.all_reduce(captured, op=dist.ReduceOp.MIN)
if int(captured.item()):
    for _ in range(3):
        out = stepper.replay()
        stepper.finish()
    results["step_graph"] = (out["table"].to_host(),
                             multi_gpu.combine_partial(out["ov"].clone(), world).cpu().numpy(),
                             int((out["ov"] != 0).sum().item()))

# ---- genomes of very different lengths: every rank plans its own slice, so the row strides differ and the
# ranks must agree on one before the exchange (run.agree_row_stride); both exchanges, through run.make_step
from oracle import oracle  # noqa: E402  (test infrastructure: the genomes, and the expected sketches)
from pyani_plus_b200 import run as run_mod  # noqa: E402

ctx = run_mod.DistContext.from_env()
N2 = 9
lengths = [150_000 + 420_000 * g for g in range(N2)]  # rank 0 gets the short ones, the last rank the long ones
h0, h1, per2 = multi_gpu.slice_for_rank(N2, rank, world)
mine = [np.frombuffer(oracle.synth_genome(SEED, g, lengths[g]), dtype=np.uint8) for g in range(h0, h1)]
toff2 = pstream.plan_tiles([len(x) for x in mine] + [0] * (per2 - len(mine)))
own_stride = eng.plan_stream(toff2, SCALED).row_stride
plan2 = run_mod.agree_row_stride(eng, toff2, SCALED, ctx)
strides = [None] * world
dist.all_gather_object(strides, own_stride)
h_ascii2 = torch.empty(plan2.n_bases, dtype=torch.uint8)
pstream.fill_ascii_stream(h_ascii2.numpy(), toff2, [[x] for x in mine] + [[]] * (per2 - len(mine)))
uneven = {}
for nccl in (False, True):
    bufs2 = eng.alloc_stream_buffers(plan2, host_packed=True)
    tab2 = eng.alloc_table(plan2)
    step2, exchange2 = run_mod.make_step(eng, plan2, bufs2, tab2, K, ctx, h_ascii=h_ascii2, nccl_gather=nccl)
    out2 = step2.run(from_host=True)
    uneven[exchange2] = (out2["table"].to_host(), multi_gpu.combine_partial(out2["ov"].clone(), world).cpu().numpy())

if rank == 0:
    assert len(set(strides)) > 1, f"test needs ranks whose own strides differ, got {strides}"
    assert len(uneven) == 2, list(uneven)
    idx2 = multi_gpu.real_rows(N2, world)
    want = [oracle.sketch_records([oracle.synth_genome(SEED, g, lengths[g])], K, SCALED) for g in range(N2)]
    for name, (sk, ov) in uneven.items():
        for g, r in enumerate(idx2):
            assert sk[r].tolist() == want[g].tolist(), (name, g)
        for a, ra in enumerate(idx2):
            for b, rb in enumerate(idx2):
                assert ov[ra, rb] == oracle.intersect(want[a], want[b]), (name, a, b)

if rank == 0:
    full, full_off = eng.synth_ascii_stream(SEED, 0, N, LENGTH)
    ref = eng.sketch_ascii_stream(full, full_off, K, SCALED, from_host=False)
    ref_sk = ref.to_host()
    ref_ov = eng.intersect(ref).cpu().numpy()
    idx = multi_gpu.real_rows(N, world)
    for mode, (sk, ov, nonzero) in results.items():
        for g, r in enumerate(idx):
            assert sk[r].tolist() == ref_sk[g].tolist(), (mode, g)
        assert (ov[np.ix_(idx, idx)] == ref_ov).all(), mode
        assert 0 < nonzero < int((ov != 0).sum()), mode  # this rank really computed only a part
    print("MULTI_GPU_OK modes=" + ",".join(results), flush=True)
dist.barrier()
dist.destroy_process_group()
