#!/bin/bash
# N-GPU bench line(s) (one gpurun --gpus N call):  profiles/run_n.sh <tag> <N> [tests] [extra bench args...]
tag=$1; n=$2; shift 2
out=gpurun_out
mkdir -p $out
if [ "$1" = "tests" ]; then
  shift
  timeout 600 python -m pytest tests -m gpu -q -x > $out/t_$tag.log 2>&1
  tail -4 $out/t_$tag.log
fi
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 100 --warmup 5 "$@" > $out/bench_${tag}_n$n.json 2> $out/bench_${tag}_n$n.err || tail -20 $out/bench_${tag}_n$n.err
python - $out/bench_${tag}_n$n.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("n", d["n_gpus"], "ms/step", round(d["ms_per_step"], 4), "value", round(d["value"] / 1e6, 2), "e2e", round(d["e2e"]["value"] / 1e6, 2))
    print({k: round(v["us_per_step"], 1) for k, v in d["kernels_in_step"].items()})
    print("loss", d["loss_first_steps"], d["loss_step0_rank0"])
except Exception as e:
    print("no result", e)
PY
