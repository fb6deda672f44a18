#!/usr/bin/env python
"""Turn ``ncu --set full`` captures into profiles/ncu_rNN.json, the file bench.py reads its ncu-derived
roofline fields from (per-launch DRAM traffic, executed warp instructions, pipe utilisation).

    python tools/ncu_to_json.py r02 config3 k1=gpurun_out/prof_k1_r02.ncu-rep k2_index=gpurun_out/prof_k2idx_r02.ncu-rep \
        [--sha gpurun_out/source_sha_r02.txt] [--bases 5000000000]

Every capture carries ``source_sha`` = bench.kernel_source_sha(<kernel>) of the CUDA sources that kernel was
built from (taken from --sha when the capture script recorded them on the GPU box -- a JSON object
{"all": ..., "k1": ..., ...} -- else computed from the working tree).  bench.py prints ``traffic: null``
whenever the tree's sources of that kernel no longer match the hash.
"""
from __future__ import annotations

import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def raw_rows(rep: Path) -> list[dict[str, str]]:
    txt = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True,
                         check=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[0]
    return [dict(zip(hdr, r, strict=False)) for r in rows[2:]]


def f(row: dict, key: str) -> float | None:
    v = row.get(key)
    try:
        return float(v.replace(",", "")) if v not in (None, "") else None
    except ValueError:
        return None


def units(rep: Path) -> dict[str, str]:
    txt = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True,
                         check=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    return dict(zip(rows[0], rows[1], strict=False))


SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3,
         "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}


def summarise(rep: Path) -> dict:
    """Sum over the launches in the report (one logical kernel call may be several launches)."""
    rows = raw_rows(rep)
    un = units(rep)
    out = {"launches": len(rows), "kernels": sorted({r.get("Kernel Name", "?")[:80] for r in rows})}
    # one logical call = one launch of every distinct kernel in the report: average the launches of a kernel
    # (a capture window may hold several calls), then sum over the kernels
    per_kernel: dict[str, dict[str, list[float]]] = {}
    for r in rows:
        acc = per_kernel.setdefault(r.get("Kernel Name", "?"), {"dram_bytes": [], "inst_executed": [], "duration_ms": []})
        dram = 0.0
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            v = f(r, key)
            if v is not None:
                dram += v * SCALE.get(un.get(key, "byte"), 1.0)
        acc["dram_bytes"].append(dram)
        v = f(r, "smsp__inst_executed.sum")
        if v is not None:
            acc["inst_executed"].append(v)
        v = f(r, "gpu__time_duration.sum")
        if v is not None:
            acc["duration_ms"].append(v * SCALE.get(un.get("gpu__time_duration.sum", "ms"), 1.0))
    tot = {name: sum(sum(a[name]) / len(a[name]) for a in per_kernel.values() if a[name])
           for name in ("dram_bytes", "inst_executed", "duration_ms")}
    out["per_kernel"] = {k[:60]: {"launches_in_report": len(a["duration_ms"]),
                                  "duration_ms": sum(a["duration_ms"]) / max(1, len(a["duration_ms"])),
                                  "dram_bytes": sum(a["dram_bytes"]) / max(1, len(a["dram_bytes"]))}
                         for k, a in per_kernel.items()}
    out.update(tot)
    big = max(rows, key=lambda r: f(r, "gpu__time_duration.sum") or 0.0)
    out["pipe_busy"] = {
        "of_launch": big.get("Kernel Name", "?")[:80],
        "alu": f(big, "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
        "fma_heavy": f(big, "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        "issue": f(big, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "lsu": f(big, "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
        "xu": f(big, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
        "dram": f(big, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "l1tex": f(big, "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
        "warps_active_per_cycle": f(big, "sm__warps_active.avg.per_cycle_active"),
    }
    stalls = sorted(((float(v or 0), k.split("issue_stalled_")[1].split("_per_issue")[0]) for k, v in big.items()
                     if "issue_stalled" in k and k.endswith("per_issue_active.ratio")), reverse=True)
    out["top_stalls"] = {name: round(val, 3) for val, name in stalls[:6]}
    return out


def main() -> None:
    tag, workload = sys.argv[1], sys.argv[2]
    sha, bases = None, None
    caps = {}
    args = sys.argv[3:]
    i = 0
    while i < len(args):
        if args[i] == "--sha":
            txt = Path(args[i + 1]).read_text().strip()
            sha = json.loads(txt) if txt.startswith("{") else {"all": txt}
            i += 2
        elif args[i] == "--bases":
            bases = int(args[i + 1])
            i += 2
        else:
            key, rep = args[i].split("=", 1)
            caps[key] = summarise(Path(rep))
            caps[key]["report"] = Path(rep).name
            i += 1
    n, length, *_ = bench.WORKLOADS[workload]
    for key in caps:
        if key == "k1":
            caps[key]["bases"] = bases or n * length
        if key in bench.KERNEL_SOURCES:
            caps[key]["source_sha"] = (sha or {}).get(key) or bench.kernel_source_sha(key)
    out_path = ROOT / "profiles" / f"ncu_{tag}.json"
    data = json.loads(out_path.read_text()) if out_path.is_file() else {"captures": {}}
    data["source_sha"] = (sha or {}).get("all") or bench.kernel_source_sha()
    data["how"] = "ncu --set full --clock-control none, one call of each kernel after warm-up (tools/profile_r2.sh)"
    data["captures"].setdefault(workload, {}).update(caps)
    out_path.write_text(json.dumps(data, indent=1) + "\n")
    print(json.dumps(data["captures"][workload], indent=1))


if __name__ == "__main__":
    main()
