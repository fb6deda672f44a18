import os, sys, time
from pathlib import Path
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pyani_plus_b200 import engine, multi_gpu, stream as pstream
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
eng = engine.Engine(int(os.environ["LOCAL_RANK"]))
N, LENGTH, K, SCALED = 100, 5_000_000, 31, 1000
g0, g1, per_rank = multi_gpu.slice_for_rank(N, rank, world)
tiles_per = LENGTH // pstream.TILE + 1
tile_off = np.arange(per_rank + 1, dtype=np.int64) * tiles_per
plan = eng.plan_stream(tile_off, SCALED)
gen, _ = eng.synth_ascii_stream(20261017, g0, per_rank, LENGTH)
bufs = eng.alloc_stream_buffers(plan); eng.pack(gen, plan, bufs); del gen
tab = eng.alloc_table(plan)
fused = multi_gpu.SymmetricGather(per_rank, plan.row_stride, world, rank, eng.device)
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    ts=[]
    for _ in range(reps):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
    return round(float(np.median(ts)),4)
eng.hash_packed(plan, bufs, tab, K); torch.cuda.synchronize()
res = {}
res["barrier"] = timeit(lambda: fused.handle.barrier(channel=0))
res["hash"] = timeit(lambda: eng.hash_packed(plan, bufs, tab, K))
def fin():
    eng.hash_packed(plan, bufs, tab, K); eng.finalize_gather(plan, tab, fused.peer_ptrs, rank, per_rank)
res["hash+finalize_gather"] = timeit(fin)
def fin2():
    eng.hash_packed(plan, bufs, tab, K); eng.finalize(plan, tab)
res["hash+finalize_local"] = timeit(fin2)
def fin3():
    eng.hash_packed(plan, bufs, tab, K); eng.finalize_gather(plan, tab, [fused.peer_ptrs[rank]]*1, 0, per_rank)
res["hash+finalize_gather_selfonly"] = timeit(fin3)
def g():
    eng.hash_packed(plan, bufs, tab, K); fused.gather(eng, plan, tab)
res["hash+gather()"] = timeit(g)
def n():
    eng.sketch_packed(plan, bufs, tab, K); multi_gpu.all_gather_tables(tab["table"], tab["counts"], world)
res["sketch+nccl"] = timeit(n)
print(rank, res, flush=True)
dist.barrier(); dist.destroy_process_group()
