#!/bin/bash
# A/B runs of the N = 1 bench under environment switches (one gpurun call):  profiles/run_ab.sh <tag> "<ENV=..>" "<ENV=..>" ...
tag=$1; shift
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -q -x > $out/t_$tag.log 2>&1
tail -4 $out/t_$tag.log
i=0
for envs in "$@"; do
  i=$((i+1))
  echo "== [$i] $envs"
  env $envs timeout 300 python bench.py --steps ${AB_STEPS:-100} --warmup 5 --no-cpu-baseline --no-exact > $out/bench_${tag}_$i.json 2> $out/bench_${tag}_$i.err || tail -5 $out/bench_${tag}_$i.err
  python - $out/bench_${tag}_$i.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("ms/step", round(d["ms_per_step"], 4), "value", round(d["value"] / 1e6, 2), "e2e", round(d["e2e"]["value"] / 1e6, 2), d["e2e"].get("with_i64_ids"))
    print({k: round(v["us_per_step"], 1) for k, v in d["kernels_in_step"].items()})
    print({k: round(v["ms_per_launch"] * 1e3, 1) for k, v in d["kernels"].items()})
    print("loss", d["loss_first_steps"])
except Exception as e:
    print("no result", e)
PY
done
