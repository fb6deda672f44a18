"""Turn ncu output (gpurun_out/*.ncu-rep, launches_*.csv) into the small text summaries kept in profiles/.

    python tools/summarise_ncu.py r01            # reads gpurun_out/{launches,prof_k1,prof_k2}_r01*, writes profiles/r01_*.md
"""
from __future__ import annotations

import collections
import csv
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "gpurun_out"
PROF = ROOT / "profiles"

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__cycles_elapsed.max", "sm__cycles_active.avg",
]


def raw_page(rep: Path) -> list[dict[str, tuple[str, str]]]:
    txt = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True,
                         check=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    return [{h: (u, v) for h, u, v in zip(hdr, units, r, strict=False)} for r in rows[2:]]


def summarise_report(rep: Path, title: str) -> str:
    lines = [f"## {title}", "", f"source: `{rep.name}` (ncu --set full --clock-control none)", ""]
    for i, m in enumerate(raw_page(rep)):
        lines.append(f"### launch {i}: `{m.get('Kernel Name', ('', '?'))[1][:110]}`")
        lines.append("")
        lines.append("| metric | value | unit |")
        lines.append("|---|---|---|")
        for k in KEYS:
            if k in m:
                lines.append(f"| {k} | {m[k][1]} | {m[k][0]} |")
        stalls = sorted(
            ((float(v[1] or 0), k) for k, v in m.items() if "issue_stalled" in k and k.endswith("per_issue_active.ratio")),
            reverse=True)
        lines.append("")
        lines.append("warp stall reasons (cycles per issued instruction): "
                     + ", ".join(f"{k.split('issue_stalled_')[1].split('_per_issue')[0]} {x:.2f}" for x, k in stalls[:8]))
        lines.append("")
    return "\n".join(lines)


def summarise_launches(path: Path) -> str:
    rows = list(csv.reader(path.read_text().splitlines()))
    hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hdr_i]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    mu = hdr.index("Metric Unit")
    tot: dict[str, float] = collections.defaultdict(float)
    cnt: dict[str, int] = collections.Counter()
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
    for r in rows[hdr_i + 1:]:
        if len(r) <= mv:
            continue
        try:
            t = float(r[mv].replace(",", "")) * scale.get(r[mu], 1.0)
        except ValueError:
            continue
        name = r[kn].split("(")[0][:90]
        tot[name] += t
        cnt[name] += 1
    total = sum(tot.values()) or 1.0
    lines = [f"## launch list ({path.name}): every kernel launch of the profiled command, summed by kernel", "",
             "(ncu `--metrics gpu__time_duration.sum --clock-control none`; per-launch times are cold-cache and "
             "serialised, so compare SHARES)", "", "| kernel | launches | total us | share | us / launch |",
             "|---|---|---|---|---|"]
    for name, t in sorted(tot.items(), key=lambda kv: -kv[1]):
        lines.append(f"| `{name}` | {cnt[name]} | {t:.1f} | {100 * t / total:.1f}% | {t / cnt[name]:.1f} |")
    step = {k: v for k, v in tot.items() if "panib::" in k and "synth" not in k and "pack_ascii" not in k}
    if step:
        st = sum(step.values())
        top = max(step, key=step.get)
        lines += ["", f"Kernels of the device-resident step only (panib kernels without the input generator and the "
                  f"e2e pack): `{top}` = {100 * step[top] / st:.1f} % of their {st / 1e3:.2f} ms; the rest of the list "
                  "is input generation, the L2 flush fills between steps and the e2e (host-input) passes, whose K1 "
                  "runs as 16-32 chunk launches overlapped with the H2D copies."]
    return "\n".join(lines) + "\n"


def main() -> None:
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    PROF.mkdir(exist_ok=True)
    launches = OUT / f"launches_{tag}.csv"
    if launches.is_file():
        (PROF / f"{tag}_launches.md").write_text(summarise_launches(launches))
    launches_idx = OUT / f"launches_idx_{tag}.csv"
    if launches_idx.is_file():
        (PROF / f"{tag}_launches_k2_index.md").write_text(
            "`python tools/time_k1.py config3 1`: K1, the probing K2 and the inverted-index K2 (CUB radix sort + "
            "scan + `index_*` kernels) on 1,000 genomes\n\n" + summarise_launches(launches_idx))
    r1 = tag == "r01"  # round 1 profiled configs[1] (100 genomes); later rounds the bench default, configs[2]
    for kind, title in (("k1", "K1 sketch_hash_kernel (configs[1]: 100 x 5 Mb)" if r1 else
                         "K1 sketch_hash_kernel<31> (configs[2]: 1,000 x 5 Mb, one launch = 5 Gbp)"),
                        ("k2", "K2 intersect_kernel (config 2: 4,950 pairs)"),
                        ("k2c3", "K2 intersect_kernel (config 3: 1,000 genomes, 499,500 pairs)"),
                        ("k2probe", "K2 intersect_kernel, probing form (configs[2]: 1,000 genomes, 499,500 pairs)"),
                        ("k2idx", "K2 index_dense_kernel, AND+POPC over the bit matrix (config 3: 1,000 genomes)" if r1
                         else "K2 inverted-index form, every kernel of its steps (configs[2]: 1,000 genomes)")):
        rep = OUT / f"prof_{kind}_{tag}.ncu-rep"
        if rep.is_file():
            (PROF / f"{tag}_{kind}_ncu.md").write_text(summarise_report(rep, title) + "\n")
    print("wrote", sorted(p.name for p in PROF.glob(f"{tag}_*")))


if __name__ == "__main__":
    main()
