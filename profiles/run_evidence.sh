#!/bin/bash
# One GPU call's worth of evidence for the N = 1 headline path (run on the B200 box through gpurun):
#   profiles/run_evidence.sh <tag> [tests]
# GPU tests (optional), the plain bench line, the ncu launch list of the same command, an `ncu --set full` capture of
# every libctr_b200 kernel of one eager training step (raw-page CSV made on the box), and a source-level capture of
# the embedding kernels.  Everything lands in gpurun_out/; profiles/summarise_ncu.py turns the CSV into the committed summary.
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
if [ "$2" = "tests" ]; then
  timeout 900 python -m pytest tests -m gpu -q > $out/t_full_$tag.log 2>&1
  tail -3 $out/t_full_$tag.log
fi
python - > $out/h2d_$tag.txt 2>&1 <<'EOF'
import torch, time
for mb in (1, 9, 17, 64):
    h = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
    d = torch.empty(mb << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print(f"pinned H2D {mb} MiB: {20 * (mb << 20) / (e0.elapsed_time(e1) * 1e-3) / 1e9:.1f} GB/s")
EOF
cat $out/h2d_$tag.txt
timeout 600 python bench.py --steps 100 --warmup 5 > $out/bench_${tag}_n1.json 2> $out/bench_${tag}_n1.err || exit 1
cat $out/bench_${tag}_n1.json
# launch list of the bench command (graph replays are expanded into their kernels by ncu)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-exact --no-graph > $out/ncu_ll_$tag.log 2>&1
# --set full of one eager step (the third of kernels_in_step's), every kernel of the library
K='regex:emb_|radix_|linear_tf32|wgrad_|bn_|col_|head_|dense_adagrad|reset_counters'
timeout 900 ncu --set full --clock-control none -k "$K" -s 200 -c 60 -o $out/${tag}_kernels_full -f \
    python bench.py --roofline-only --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_full_$tag.log 2>&1
ncu -i $out/${tag}_kernels_full.ncu-rep --page raw --csv --print-units base > $out/${tag}_kernels_full.raw.csv 2>/dev/null
rm -f $out/${tag}_kernels_full.ncu-rep
# source-level (stall sampling) capture of the embedding kernels only
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:emb_bwd_sweep|radix_seg|emb_pool_fwd|emb_keygen' -s 20 -c 6 \
    -o $out/${tag}_emb_source -f python bench.py --roofline-only --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_src_$tag.log 2>&1
ls -la $out | tail -12
