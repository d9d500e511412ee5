#!/bin/bash
# The last GPU call of round 2 (one B200, ~12 min):  profiles/run_final.sh <tag>
#   1. the GPU tests closest to the last change (tower / models / precision), then -- at the very end, with whatever
#      time is left -- the rest of the suite;
#   2. the bench line of this commit, and the same with the first block's preparation left in line (CTR_PREPARE_EARLY=0);
#   3. the ncu launch list of the bench command, `ncu --set full` of every libctr_b200 kernel of one eager step (raw-page
#      CSV made on the box), per-instruction stall sampling of the sweep and the lookup;
#   4. the CUPTI kernel timeline of three graph replays.
# Everything lands in gpurun_out/; profiles/summarise_ncu.py / summarise_launches.py turn the CSVs into the committed files.
tag=${1:-r2_final}
out=gpurun_out
mkdir -p $out
t0=$(date +%s)
stamp() { echo "== [$(( $(date +%s) - t0 )) s] $*"; }

stamp "tests near the change"
timeout 300 python -m pytest tests/test_gpu_tower.py tests/test_gpu_models.py tests/test_gpu_precision.py -m gpu -q -x \
    > $out/t_near_$tag.log 2>&1
tail -3 $out/t_near_$tag.log

stamp "bench (this commit)"
timeout 300 python bench.py --steps 100 --warmup 5 > $out/bench_${tag}_n1.json 2> $out/bench_${tag}_n1.err || tail -5 $out/bench_${tag}_n1.err
stamp "bench (CTR_PREPARE_EARLY=0)"
CTR_PREPARE_EARLY=0 timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-exact \
    > $out/bench_${tag}_noprep.json 2> $out/bench_${tag}_noprep.err || tail -5 $out/bench_${tag}_noprep.err
python - $out/bench_${tag}_n1.json $out/bench_${tag}_noprep.json <<'PY'
import json, sys
for p in sys.argv[1:]:
    try:
        d = json.load(open(p))
        print(p, "ms/step", round(d["ms_per_step"], 4), "value", round(d["value"] / 1e6, 2), "e2e", round(d["e2e"]["value"] / 1e6, 2),
              "roofline", round(d["roofline"]["frac"], 3), "loss", d.get("loss_first_steps"))
        print({k: round(v["us_per_step"], 1) for k, v in d["kernels_in_step"].items()})
    except Exception as e:
        print(p, "no result", e)
PY

stamp "ncu launch list"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-exact --no-graph > $out/ncu_ll_$tag.log 2>&1
stamp "ncu --set full, one eager step"
K='regex:emb_|radix_|linear_tf32|wgrad_|bn_|col_|head_|dense_adagrad|reset_counters'
timeout 300 ncu --set full --clock-control none -k "$K" -s 200 -c 60 -o $out/${tag}_kernels_full -f \
    python bench.py --roofline-only --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_full_$tag.log 2>&1
ncu -i $out/${tag}_kernels_full.ncu-rep --page raw --csv --print-units base > $out/${tag}_kernels_full.raw.csv 2>/dev/null
rm -f $out/${tag}_kernels_full.ncu-rep
stamp "ncu source-level capture of the sweep and the lookup"
timeout 150 ncu --set full --clock-control none --import-source on -k 'regex:emb_bwd_sweep|emb_pool_fwd' -s 8 -c 2 \
    -o $out/${tag}_emb_source -f python bench.py --roofline-only --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_src_$tag.log 2>&1
ncu -i $out/${tag}_emb_source.ncu-rep --page source --csv > $out/${tag}_emb_source.source.csv 2>/dev/null
ncu -i $out/${tag}_emb_source.ncu-rep --page raw --csv --print-units base > $out/${tag}_emb_source.raw.csv 2>/dev/null

stamp "kernel timeline"
timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-exact --trace $out/trace_${tag}_n1.txt > $out/trace_$tag.log 2>&1

stamp "the rest of the GPU tests"
timeout 420 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_tower.py --deselect tests/test_gpu_models.py \
    --deselect tests/test_gpu_precision.py > $out/t_rest_$tag.log 2>&1
tail -3 $out/t_rest_$tag.log
stamp done
ls -la $out | tail -25
