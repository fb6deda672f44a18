"""Time the K1 hash kernel (and K2) alone on a BASELINE workload, for kernel-variant A/B runs.

    PANIB200_LIB=tools/variants_x.so python tools/time_k1.py [config2] [reps]
"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from pyani_plus_b200 import engine, stream as pstream  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "config2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
n, length, k, scaled, _ = bench.WORKLOADS[workload]
eng = engine.Engine(0)
d_ascii, tile_off = eng.synth_ascii_stream(bench.SEED, 0, n, length)
plan = eng.plan_stream(tile_off, scaled)
bufs = eng.alloc_stream_buffers(plan)
eng.pack(d_ascii, plan, bufs)
del d_ascii
tab = eng.alloc_table(plan)
sk = eng._sketch_args(plan, bufs, tab, k, 42)
hash_args = sk[:12] + (sk[13], sk[14], sk[15])
flush = torch.empty(256 << 20, dtype=torch.uint8, device=eng.device)


def timeit(fn):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


t_hash = timeit(lambda: engine._check(eng.lib.panib_sketch_hash_only(*hash_args)))
t_fin = timeit(lambda: engine._check(eng.lib.panib_sketch_finalize(
    tab["table"].data_ptr(), plan.row_stride, plan.n_genomes, plan.d_nb.data_ptr(), tab["counts"].data_ptr(),
    tab["flags"].data_ptr(), 0, eng._stream())))
eng.sketch_packed(plan, bufs, tab, k)
assert eng.check_status() == 0
table = engine.SketchTable(tab["table"], tab["counts"], k, scaled)
chk = int(table.counts.sum().item())
mc = int(length / scaled * 1.3) + 64
real_max = int(table.counts.max().item())
t_k2 = timeit(lambda: eng.intersect(table, method="probe", check=False, max_count=real_max))
ov = eng.intersect(table, method="probe")
idx_note = ""
try:
    t_idx = timeit(lambda: eng.intersect(table, method="index", check=False, max_count=real_max))
    ov_idx = eng.intersect(table, method="index")
    same = bool((ov_idx == ov).all().item())
    eng.intersect(table, method="auto")
    idx_note = (f" | K2 index {t_idx[0]:.3f} ms = {n*(n-1)/2/t_idx[0]/1e3:.2f} Mpairs/s equal={same} "
                f"auto->{eng.last_intersect_method} est={eng.last_intersect_estimates}")
except Exception as exc:  # noqa: BLE001
    idx_note = f" | K2 index failed: {exc}"
print(f"{workload} hash {t_hash[0]:.3f} ms (min {t_hash[1]:.3f}) = {n*length/t_hash[0]/1e6:.1f} Gbp/s | "
      f"finalize {t_fin[0]:.3f} ms | K2 {t_k2[0]:.3f} ms = {n*(n-1)/2/t_k2[0]/1e3:.2f} Mpairs/s | "
      f"sum(counts)={chk} sum(ov)={int(ov.to(torch.int64).sum().item())}{idx_note}")
