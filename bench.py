#!/usr/bin/env python
"""Headline benchmark: train samples/s of a Criteo-shape DeepFM (BASELINE.json configs[1]:
26 sparse + 13 dense, emb dim 16, batch 65536) on N B200s, plus the embedding kernels'
achieved HBM GB/s against the measured roofline, the end-to-end number through the public
module API with host buffers, and the reference's CPU path timed beside it.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference        # the reference arithmetic on the host cores

One JSON line on stdout (rank 0).  A "step" is forward + backward + optimizer update of one
batch of synthetic Zipf(1.05) ids / N(0,1) dense / Bernoulli(0.25) labels.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# Criteo-like cardinalities: 3 tables >= 1e7, the rest 1e1 .. 2e6
CRITEO_VOCABS = [10_000_000, 10_000_000, 10_000_000, 2_000_000, 1_000_000, 1_000_000, 500_000, 300_000,
                 100_000, 100_000, 50_000, 20_000, 10_000, 10_000, 5_000, 5_000, 2_000, 2_000, 1_000, 1_000,
                 1_000, 500, 100, 50, 20, 10]
NUM_DENSE = 13
EMB_DIM = 16
HIDDEN = [256, 128, 64]
ZIPF_A = 1.05
LR = 0.01


def feat_configs(vocabs=CRITEO_VOCABS, dim=EMB_DIM, num_dense=NUM_DENSE):
    fc = [{"name": f"C{i + 1}", "type": "sparse", "num_embeddings": v, "emb_dim": dim} for i, v in enumerate(vocabs)]
    fc += [{"name": f"I{i + 1}", "type": "dense"} for i in range(num_dense)]
    return fc


def zipf_ids(gen, n, V, a=ZIPF_A):
    """Zipf(a)-distributed ranks in [0, V) by inverting the continuous CDF (bounded Zipf)."""
    u = torch.rand(n, generator=gen, dtype=torch.float64)
    s = 1.0 - a
    x = ((V ** s - 1.0) * u + 1.0) ** (1.0 / s)
    return (x.floor().long() - 1).clamp_(0, V - 1)


def make_batch(seed, B, vocabs=CRITEO_VOCABS, num_dense=NUM_DENSE, pin=False):
    gen = torch.Generator().manual_seed(seed)
    feats = {}
    for i, V in enumerate(vocabs):
        feats[f"C{i + 1}"] = zipf_ids(gen, B, V).reshape(B, 1)
    feats["dense_features"] = torch.randn(B, num_dense, generator=gen)
    labels = (torch.rand(B, 1, generator=gen) < 0.25).float()
    if pin:
        feats = {k: v.pin_memory() for k, v in feats.items()}
        labels = labels.pin_memory()
    return feats, labels


def batch_bytes(batch):
    feats, labels = batch
    return sum(v.numel() * v.element_size() for v in feats.values()) + labels.numel() * labels.element_size()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except (OSError, KeyError, ValueError):
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, B, budget_s):
    """The reference arithmetic (oracle.models.OracleDeepFM == nn.Embedding + dense autograd +
    dense torch.optim.Adagrad, the path torchctr/trainer.py:291-303 drives) on the host cores."""
    from oracle import models as om
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    fc = feat_configs()
    torch.manual_seed(0)
    model = om.OracleDeepFM(fc, HIDDEN).train()
    opt = torch.optim.Adagrad(model.parameters(), lr=LR)
    batches = [make_batch(100 + i, B) for i in range(2)]

    def step(i):
        opt.zero_grad()
        loss = model.training_step(batches[i % len(batches)], i)
        loss.backward()
        opt.step()
        return loss.item()

    t0 = time.perf_counter()
    step(0)
    first = time.perf_counter() - t0
    warm_done = 1
    while warm_done < warmup and (time.perf_counter() - t0) < 0.3 * budget_s:
        step(warm_done); warm_done += 1
    per = max((time.perf_counter() - t0) / warm_done, 1e-3)
    k = max(1, min(steps, int((budget_s - (time.perf_counter() - t0)) / per)))
    t1 = time.perf_counter()
    for i in range(k):
        step(warm_done + i)
    dt = time.perf_counter() - t1
    return {"value": B * k / dt, "steps": k, "warmup": warm_done, "ms_per_step": 1e3 * dt / k, "cores": threads,
            "first_step_s": first, "B": B}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = args.batch
    r = cpu_reference_run(args.steps, args.warmup, B, budget_s=150.0)
    sample = (f"{r['steps']} steps (after {r['warmup']} warm-up) of B={B} samples each, full-size tables "
              f"({sum(CRITEO_VOCABS)} rows x {EMB_DIM}), dense autograd + dense torch.optim.Adagrad, eager fp32 on CPU")
    line = {
        "impl": "reference", "metric": "train samples/s, Criteo-shape DeepFM", "value": r["value"], "unit": "samples/s",
        "n_gpus": args.gpus, "steps": r["steps"], "warmup": r["warmup"], "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(B, 1),
        "cpu_baseline": {"value": r["value"], "unit": "samples/s", "cores": r["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": r["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def workload_config(B, n_gpus):
    return {"workload": "BASELINE configs[1]: DeepFM, 26 sparse (Zipf(1.05) ids, Criteo-like cardinalities, "
                        f"{sum(CRITEO_VOCABS)} rows total) + 13 dense, emb dim 16, tower 256-128-64, Adagrad",
            "batch_per_gpu": B, "global_batch": B * n_gpus, "emb_dim": EMB_DIM, "num_sparse": len(CRITEO_VOCABS),
            "num_dense": NUM_DENSE, "table_rows": sum(CRITEO_VOCABS),
            "l2": "inputs larger than L2: 2.4 GB of tables + 2.4 GB optimizer state, 4 distinct batches cycled"}


def kernel_roofline(model, resident, B, dev, iters=12):
    """Times ctr_emb_pool_fwd / ctr_emb_bwd_plan / ctr_emb_bwd_apply (fused Adagrad) of the D=16 table group one
    launch at a time.  Algorithmic bytes per SURVEY.md 8(d): fwd 8S + 4DN + 4DB*F; update 8S + 4DB*F + 16DU."""
    from torchctr_b200 import ops
    from torchctr_b200.nn.embedding import _layout
    names = model._names
    tables = [model.embeddings[n] for n in names]
    D = EMB_DIM
    S = B * len(names)
    flush = torch.empty(1 << 30, dtype=torch.uint8, device=dev)   # > L2; its fill also hides the host launch latency
    peak, peak_src = measured_peak_hbm()
    acc = {"emb_pool_fwd": 0.0, "emb_bwd_plan": 0.0, "emb_bwd_apply": 0.0}
    uniq_total = 0
    for it in range(iters + 2):
        feats, _ = resident[it % len(resident)]
        entries = [(t, feats[n], None) for t, n in zip(tables, names)]
        cols, width, stride, dense_col = _layout(entries, NUM_DENSE)
        out = torch.empty(B, stride, device=dev)
        gout = torch.randn(B, stride, device=dev)
        specs = [ops.FeatureSpec(ids=feats[n], table=t.weight.data, num_rows=t.num_embeddings, D=D, out_col=c,
                                 state0=t.opt_state0) for t, n, c in zip(tables, names, cols)]
        fwd = ops.make_group(specs, B, out, stride, dense=feats["dense_features"], dense_col=dense_col, zero_from=width)
        bwd = ops.make_group(specs, B, gout, stride)
        ws = torch.empty(ops.emb_bwd_workspace_bytes(bwd) + 256, dtype=torch.uint8, device=dev)
        opt = ops.make_opt("adagrad", lr=LR, eps=1e-10)
        for name, fn in (("emb_pool_fwd", lambda: ops.emb_pool_fwd(fwd)), ("emb_bwd_plan", lambda: ops.emb_bwd_plan(bwd, ws, runs=False)),
                         ("emb_bwd_apply", lambda: ops.emb_bwd_apply(bwd, ws, opt))):
            flush.fill_(it & 0xff)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            if it >= 2:
                acc[name] += e0.elapsed_time(e1)
        if it >= 2:
            uniq_total += sum(int(torch.unique(feats[n]).numel()) for n in names)
    U = uniq_total / iters
    bytes_ = {"emb_pool_fwd": 8 * S + 4 * D * S + 4 * D * S, "emb_bwd_apply": 8 * S + 4 * D * S + 16 * D * U}
    kern = {}
    for name, total in acc.items():
        ms = total / iters
        kern[name] = {"ms_per_launch": ms}
        if name in bytes_:
            kern[name]["algorithmic_bytes"] = bytes_[name]
            kern[name]["achieved_GBs"] = bytes_[name] / (ms * 1e-3) / 1e9
            kern[name]["frac_of_peak"] = kern[name]["achieved_GBs"] / peak
    dom = "emb_bwd_apply" if kern["emb_bwd_apply"]["ms_per_launch"] >= kern["emb_pool_fwd"]["ms_per_launch"] else "emb_pool_fwd"
    k = kern[dom]
    roofline = {"bound": "hbm", "kernel": dom + (" (sparse gradient reduce + fused Adagrad row update, 26 tables x D=16)"
                                                 if dom == "emb_bwd_apply" else " (gather + pool + concat, 26 tables x D=16)"),
                "achieved": k["achieved_GBs"], "peak": peak, "unit": "GB/s", "frac": k["frac_of_peak"],
                "traffic": 234.1e6 if dom == "emb_bwd_apply" else 109.1e6,
                "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of this kernel in profiles/r1_final_kernels_full.summary.txt (ncu --set full)",
                "peak_source": peak_src, "algorithmic_bytes_per_launch": k["algorithmic_bytes"],
                "ms_per_launch": k["ms_per_launch"], "unique_rows_per_launch": U,
                "timing": f"CUDA events around each launch on the launching stream, 1 GiB L2 flush before each, {iters} launches"}
    return kern, roofline


def kernels_in_step(eager_step, resident, steps=6, sleep_cycles=16_000_000):
    """CUDA-event time of every libctr_b200 entry point INSIDE the real training step (caches as the step leaves
    them).  The step runs eagerly behind a device-side sleep, so the host has enqueued all of it before the GPU
    starts and the events see back-to-back kernels, not launch gaps.  -> {name: {calls_per_step, us_per_step}}"""
    from torchctr_b200 import ops
    nb = len(resident)
    with ops.KernelTimer() as kt:
        for i in range(steps):
            torch.cuda._sleep(sleep_cycles)
            eager_step(resident[i % nb], i)
            torch.cuda.synchronize()
        summ = kt.summary()
    return {k: {"calls_per_step": n / steps, "us_per_step": 1e3 * ms / steps} for k, (n, ms) in summ.items()}


# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from torchctr_b200 import ops
    from torchctr_b200.models import DeepFM

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    from torchctr_b200.nn import set_matmul_precision
    torch.backends.cuda.matmul.allow_tf32 = True        # tower GEMMs on tensor cores (TF32 in, fp32 accumulate)
    torch.backends.cudnn.allow_tf32 = True
    set_matmul_precision(args.precision)                # "tf32" (headline) | "tf32x3" (error-compensated, fp32-grade)

    torch.manual_seed(0)
    fc = feat_configs()
    if world > 1:
        from torchctr_b200.parallel import shard_model
    model = DeepFM(fc, HIDDEN)                            # identical on every rank (same seed)
    if world > 1:
        dedup = {"auto": None, "on": True, "off": False}[args.dedup]
        model = shard_model(model, None, device=dev, dedup=dedup)    # keep this rank's rows of every table, drop the rest
    model = model.to(dev).train()
    if world > 1:
        model.enable_flat_dense_grads()                   # dense gradients live in one buffer: all-reduce without cat / copies
    from torchctr_b200.optim import FusedAdagrad
    opt = FusedAdagrad(model.dense_parameters(), lr=LR)           # torch.optim.Adagrad arithmetic, one launch; tables take the fused row update
    model.bind_optimizer(opt, kind="adagrad")

    nb = 4
    host = [make_batch(1000 * rank + i, B, pin=True) for i in range(nb)]
    resident = [({k: v.to(dev) for k, v in f.items()}, l.to(dev)) for f, l in host]
    h2d = batch_bytes(host[0])

    def eager_step(batch, i):
        if world > 1:
            model.zero_dense_grads()
        else:
            opt.zero_grad(set_to_none=True)
        loss = model.training_step(batch, i)
        if world > 1:
            (loss / world).backward()                   # tables: all-to-all of gradients + owner-side fused update
            model.reduce_dense_grads()                  # tower: one flat all-reduce
        else:
            loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):                 # eager warm-up (also sizes workspaces / optimizer state)
        eager_step(resident[i % nb], i)
    launches_per_step = None
    in_step = kernels_in_step(eager_step, resident)      # at N > 1: this rank's kernels, peers' rows over NVLink included
    if args.roofline_only:                               # profiling hook: just the embedding entry points
        kern, roofline = kernel_roofline(model, resident, B, dev, iters=args.steps)
        emit({"kernels": kern, "roofline": roofline, "kernels_in_step": in_step})
        return
    graphed_holder = [None]
    if args.no_graph:
        step = eager_step
    else:
        from torchctr_b200.graph import GraphedTrainStep
        l0 = ops.kernel_launches()

        def sharded_backward(loss):                      # tables: owners pull gradients over NVLink; tower: all-reduce
            (loss / world).backward()
            model.reduce_dense_grads()
        graphed = GraphedTrainStep(model, opt, resident[0], warmup=1, backward_fn=sharded_backward if world > 1 else None,
                                   zero_grad_fn=model.zero_dense_grads if world > 1 else None)
        launches_per_step = (ops.kernel_launches() - l0) // 2      # one eager warm-up + one captured step
        graphed_holder[0] = graphed
        step = lambda batch, i: graphed(batch)           # noqa: E731

    def timed(batches, steps, read_loss):
        barrier()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        t0.record()
        prefetch = getattr(graphed_holder[0], "prefetch", None) if read_loss else None
        if prefetch is not None:
            prefetch(batches[0])
        for i in range(steps):
            loss = step(batches[i % nb], i)
            if read_loss:
                if prefetch is not None and i + 1 < steps:
                    prefetch(batches[(i + 1) % nb])      # next batch's host -> device copy runs under this step
                loss.item()                              # device -> host read of the step's result
        t1.record()
        barrier()
        wall = time.perf_counter() - w0
        ms = t0.elapsed_time(t1)
        if read_loss:
            ms = max(ms, 1e3 * wall)                    # host-inclusive for the end-to-end leg
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    t_load = time.perf_counter()                         # ~0.8 s of the same steps: the GPU is under the benchmark's load while
    i = 0                                                # nvidia-smi (100 ms period) samples clocks and throttle reasons
    while time.perf_counter() - t_load < 0.8:
        for _ in range(25):
            step(resident[i % nb], i)
            i += 1
        torch.cuda.synchronize()
    launches0 = ops.kernel_launches()
    ms = timed(resident, args.steps, read_loss=False)
    launches = ops.kernel_launches() - launches0 if launches_per_step is None else launches_per_step * args.steps
    # end to end: pinned host buffers in, loss out, every step
    for i in range(3):
        step(host[i % nb], i)
    ms_e2e = timed(host, args.steps, read_loss=True)
    clock_info = clocks.stop() if rank == 0 else None

    value = B * world * args.steps / (ms / 1e3)
    e2e = B * world * args.steps / (ms_e2e / 1e3)

    # the same step with the tower GEMMs in the exact mode (3xTF32 on the same tcgen05 kernels): the configuration the
    # 1e-5 parity tests run in, timed the same way
    exact = None
    if world == 1 and args.precision == "tf32" and not args.no_graph and not args.no_exact:
        from torchctr_b200.graph import GraphedTrainStep
        set_matmul_precision("tf32x3")
        graphed_x3 = GraphedTrainStep(model, opt, resident[0], warmup=1)
        step_tf32, step = step, (lambda batch, i: graphed_x3(batch))
        for i in range(3):
            step(resident[i % nb], i)
        ms_x3 = timed(resident, args.steps, read_loss=False)
        step = step_tf32
        set_matmul_precision(args.precision)
        exact = {"value": B * args.steps / (ms_x3 / 1e3), "unit": "samples/s", "ms_per_step": ms_x3 / args.steps,
                 "dtype": "f32 tables / interaction / BatchNorm + 3xTF32 (error-compensated, fp32-grade) tower GEMMs on tcgen05"}

    # ---- kernel roofline: each embedding entry point timed alone, CUDA events on the launching stream,
    # L2 flushed (1 GiB written) before every launch, on the step's real tensors
    kern, roofline = kernel_roofline(model, resident, B, dev) if (rank == 0 and world == 1) else ({}, None)
    if roofline is not None:
        # the same kernel inside the real step (caches as the step leaves them): CUDA events, eager step behind a device sleep
        name = roofline["kernel"].split(" ")[0] + f"_d{EMB_DIM}"
        if name in in_step:
            us = in_step[name]["us_per_step"] / max(in_step[name]["calls_per_step"], 1)
            roofline["in_step"] = {"us_per_launch": us, "achieved": roofline["algorithmic_bytes_per_launch"] / (us * 1e-6) / 1e9,
                                   "frac": roofline["algorithmic_bytes_per_launch"] / (us * 1e-6) / 1e9 / roofline["peak"]}
        roofline["random_access_ceiling"] = ("B200 random 64-byte row reads top out at 2.2 TB/s (34 G rows/s), read-modify-write of "
                                             "64-byte rows at 17 G rows/s (profiles/micro/random_access.cu): this kernel's gathers "
                                             "are 64-byte rows, so ~0.35 of the copy peak is its hardware ceiling")

    line = {
        "metric": "train samples/s, Criteo-shape DeepFM", "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 tables / interaction / BatchNorm + " + ("tf32" if args.precision == "tf32" else "3xTF32 (fp32-grade)")
                 + " tower GEMMs (tcgen05, fp32 accumulate)",
        "data": "synthetic", "config": workload_config(B, world), "exact_mode": exact,
        "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches, "cuda_graph": not args.no_graph, "kernels": kern, "kernels_in_step": in_step,
        "roofline": roofline,
        "parallelism": "single GPU" if world == 1 else f"tables row-sharded over {world} GPUs (owner = (row + table) mod P); rows read "
                       "and gradients pulled through NVLink peer mappings inside the lookup / update kernels (no all-to-all), "
                       "batch data-parallel, tower replicated + one NCCL all-reduce"
                       + ("; requester-side de-duplication: every distinct row crosses NVLink once per direction"
                          if getattr(getattr(model, "_sharded", None), "dedup", False) else ""),
        "clocks": clock_info,
        "tower_matmul": "tcgen05 kernels (ctr_linear_fwd forward / dgrad, ctr_linear_wgrad), fp32 accumulate in TMEM; precision "
                        + args.precision + " (parity: tests/test_gpu_precision.py)",
    }
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference_run(3, 1, B, budget_s=40.0)
            line["cpu_baseline"] = {
                "value": r["value"], "unit": "samples/s", "cores": r["cores"], "kind": "port",
                "sample": f"{r['steps']} steps of B={r['B']} samples, full-size tables, dense autograd + dense Adagrad "
                          f"(oracle.models.OracleDeepFM, eager fp32 CPU)"}
        emit(line)
    if world > 1:
        # the measurement is done and printed: tear down without letting a slow NCCL / IPC teardown hold the job
        sys.stdout.flush()
        torch.cuda.synchronize()
        dist.barrier()
        t = threading.Thread(target=dist.destroy_process_group, daemon=True)
        t.start()
        t.join(timeout=20)
        os._exit(0)


_REAL_STDOUT = None


def _stdout_to_stderr():
    """NCCL and friends print banners on fd 1; the contract is ONE JSON line on stdout.  Route fd 1 to stderr for the
    whole run and keep the real stdout for emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _stdout_to_stderr()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of one CUDA graph per step")
    ap.add_argument("--roofline-only", action="store_true", help="only time the embedding entry points (ncu target)")
    ap.add_argument("--precision", default="tf32", choices=["tf32", "tf32x3"], help="tower GEMM precision of the headline run")
    ap.add_argument("--no-exact", action="store_true", help="skip the extra timing of the 3xTF32 mode")
    ap.add_argument("--dedup", default="auto", choices=["auto", "on", "off"],
                    help="N > 1: fetch / send every distinct row once (auto: above 4 GPUs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
