#!/bin/bash
# End-of-round check on one B200, most important first (every step has its own timeout):
#   gpurun --timeout 1500 -- 'bash tools/gpu_check.sh final3 [profile]'
# 1 GPU parity suite   2 smoke()   3 bench.py, default workload (configs[2])   4 the reference arm
# 5 configs[1] and configs[4] bench lines   6 drop-in CLI wall time at 1,000 genomes
# Everything lands in gpurun_out/<tag>_*; copy what is to be judged into profiles/.
# (Kernel profiles are a separate call: tools/profile_r2.sh + tools/ncu_to_json.py.)
set -u
OUT=gpurun_out
TAG=${1:-final3}
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_box.log 2>&1
lscpu | grep -E "^CPU\(s\)|Model name|NUMA node\(s\)" >> $OUT/${TAG}_box.log
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_tests.log 2>&1; echo "rc=$?" >> $OUT/${TAG}_tests.log
tail -3 $OUT/${TAG}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "rc=$?" >> $OUT/${TAG}_smoke.log
tail -2 $OUT/${TAG}_smoke.log
# optional second argument "profile": re-take the ncu captures of the bench step first (a kernel source changed)
if [ "${2:-}" = profile ]; then
  bash tools/profile_r2.sh r02 config3 > $OUT/${TAG}_profile.log 2>&1
  python tools/ncu_to_json.py r02 config3 k1=$OUT/prof_k1_r02.ncu-rep k2_index=$OUT/prof_k2idx_r02.ncu-rep \
      k2_probe=$OUT/prof_k2probe_r02.ncu-rep --sha $OUT/source_sha_r02.txt > $OUT/${TAG}_ncu_to_json.log 2>&1
  cp profiles/ncu_r02.json $OUT/ncu_r02.json   # the bench lines below then carry the measured fields
fi
timeout 600 python bench.py > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err; echo "rc=$?" >> $OUT/${TAG}_bench_default.err
timeout 600 python bench.py --impl reference > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err; echo "rc=$?" >> $OUT/${TAG}_bench_reference.err
for w in config2 config5; do
  timeout 400 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_$w.json 2> $OUT/${TAG}_bench_$w.err; echo "rc=$?" >> $OUT/${TAG}_bench_$w.err
done
python - <<PY
import json
for v in ("default", "reference", "config2", "config5"):
    try:
        d = json.loads(open("$OUT/${TAG}_bench_%s.json" % v).read().strip().splitlines()[-1])
        e = d.get("e2e") or {}
        r = d.get("roofline") or {}
        print(v, "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e ms", e.get("ms_per_step"), e.get("ingest"),
              "roof", r.get("frac"), "traffic", r.get("traffic"), "parity", (d.get("parity") or {}).get("ok"))
    except Exception as exc:
        print(v, "failed", exc)
PY
timeout 500 python tools/cli_walltime.py 1000 5000000 1000 cold,warm > $OUT/${TAG}_cli_1000.log 2>&1; echo "rc=$?" >> $OUT/${TAG}_cli_1000.log
grep -E "wall|rc=" $OUT/${TAG}_cli_1000.log
