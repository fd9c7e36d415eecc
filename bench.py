#!/usr/bin/env python
"""bench.py -- BERT-large pretraining step throughput (training_sequences_per_second, the reference's own metric:
run_pretraining.py:748) on N B200s of one node, synthetic data, bf16, LAMB.

  python bench.py --gpus 1 --steps 8 --warmup 3                      # our arm (default)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
  python bench.py --impl reference ...                               # CPU arm: the oracle restatement on the host cores

One "step" = H2D of one synthetic batch (e2e pass only) + forward + loss + backward (+ DDP bucketed NCCL allreduce)
+ FusedLAMBAMP step + zero_grad, i.e. take_training_step + take_optimizer_step of the reference driver with
gradient_accumulation_steps = 1.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "training_sequences_per_second"
_T0 = time.time()


def log(msg):
    print(f"[bench +{time.time() - _T0:7.1f}s] {msg}", file=sys.stderr, flush=True)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--seq", type=int, default=512, help="512 = phase 2 (headline), 128 = phase 1")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU micro-batch (0 = 128 @512, 512 @128 -- 65536 tokens per GPU per step; the reference ran 32 @512 / 256 @128 on 80 GB "
                         "A100s, SURVEY.md 8d; 180 GB fits 4x that and the LAMB / allreduce cost per sequence halves again vs 64)")
    ap.add_argument("--max-pred", type=int, default=0)
    ap.add_argument("--no-dropout", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-batch", type=int, default=0, help="sequences per CPU step (bounded sample)")
    ap.add_argument("--bucket-mb", type=int, default=100)
    ap.add_argument("--dynamic-mlm-gather", action="store_true", help="use torch.nonzero (host sync per step) like the reference instead of nonzero_static(batch*max_pred)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d.get("hbm_gbs"), tf_burst=d.get("bf16_tflops"), tf_sustained=d.get("bf16_tflops_sustained"), source="measured")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi sampled every 200 ms DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()          # exact PID we started
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = [r for t, r in self.rows if t0 <= t <= t1 and len(r) >= 8] or [r for _, r in self.rows if len(r) >= 8]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[1]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[4 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "power_w_max": max(float(r[3]) for r in rows),
                "samples": len(rows), "reasons": reasons}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle (reference algorithm restated for CPU, oracle/) on a bounded sample
# ------------------------------------------------------------------------------------------------
def usable_cores():
    """Host cores this process may actually use: affinity mask capped by the cgroup CPU quota (a container that
    *sees* 128 CPUs but is throttled to a fraction of them collapses under 128 spinning OpenMP threads)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            n = min(n, max(1, int(float(q) / float(per))))
    except Exception:
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                n = min(n, max(1, q // per))
        except Exception:
            pass
    return n


def cpu_reference_run(cfg, S, P, ref_batch, steps, warmup):
    threads = int(os.environ.get("DLE_CPU_THREADS", 0)) or min(usable_cores(), 64)
    os.environ["OMP_NUM_THREADS"] = str(threads)          # read by libgomp when oracle/liblamb_oracle.so first runs
    import numpy as np
    import torch
    from oracle import bert_oracle as O
    from oracle import lamb_oracle as LO
    torch.set_num_threads(threads)
    sd = {k: v.requires_grad_(True) for k, v in O.init_params(cfg, seed=42).items()}
    no_decay = ['bias', 'gamma', 'beta', 'LayerNorm']
    names = list(sd.keys())
    groups = []
    for decay, wd in ((True, 0.01), (False, 0.0)):
        ks = [k for k in names if (not any(nd in k for nd in no_decay)) == decay]
        groups.append(dict(keys=ks, params=[sd[k].detach().numpy() for k in ks], grads=None,
                           exp_avg=[np.zeros(tuple(sd[k].shape), np.float32) for k in ks],
                           exp_avg_sq=[np.zeros(tuple(sd[k].shape), np.float32) for k in ks],
                           lr=6e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=wd, step=0, bias_correction=True, grad_averaging=True))
    batches = [O.synthetic_batch(ref_batch, S, cfg["vocab_size"], P, seed=42 + i) for i in range(2)]

    def step(i):
        t_a = time.perf_counter()
        loss, *_ = O.forward_loss(sd, cfg, batches[i % 2])
        loss.backward()
        log(f"  cpu fwd+bwd {time.perf_counter() - t_a:.1f}s")
        for g in groups:
            g["grads"] = [sd[k].grad.numpy() for k in g["keys"]]
        LO.lamb_step(groups)                       # updates the numpy views of the torch parameters in place
        for k in names:
            sd[k].grad = None
        return loss.item()

    log(f"cpu arm: params built, {torch.get_num_threads()} threads")
    for i in range(warmup):
        step(i)
        log(f"cpu warm-up {i} done")
    t0 = time.perf_counter()
    for i in range(steps):
        step(i)
        log(f"cpu step {i} done")
    dt = time.perf_counter() - t0
    return dict(value=ref_batch * steps / dt, ms_per_step=1000.0 * dt / steps, cores=threads, threads=torch.get_num_threads(),
                sample=f"{threads} threads of {os.cpu_count()} visible CPUs; {steps} steps x {ref_batch} sequences (S={S}) of the same workload after {warmup} warm-up; fp32 torch-CPU "
                       f"forward/backward (oracle/bert_oracle.py) + OpenMP LAMB (oracle/lamb_oracle.c)")


def workload(args):
    from deeplearningexamples_b200 import training as T
    S = args.seq
    B = args.batch or (128 if S >= 384 else 512)
    P = args.max_pred or (80 if S >= 384 else 20)
    cfg = dict(T.BERT_LARGE)
    cfg["vocab_size"] = 30528                     # 30522 padded to a multiple of 8 (run_pretraining.py:383-384)
    return cfg, S, B, P


def config_dict(args, cfg, S, B, P, n):
    phase = "phase-2" if S >= 384 else "phase-1"
    return {"workload": f"BERT-large {phase} pretraining step seq{S} bf16 LAMB (BASELINE.json configs[{2 if S >= 384 else 1}] per-GPU shape)",
            "seq_len": S, "micro_batch_per_gpu": B, "global_batch": B * n, "max_predictions_per_seq": P,
            "gradient_accumulation_steps": 1, "dropout": 0.0 if args.no_dropout else 0.1, "parallelism": f"dp{n}",
            "attention_mask": "all ones (padded to full length)", "mlm_gather": "torch.nonzero (sync)" if args.dynamic_mlm_gather else "nonzero_static(batch*max_pred), sync-free",
            "l2_policy": "per-step working set (weights 0.67 GB + activations >10 GB) exceeds the 126 MB L2; no explicit flush"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cfg, S, B, P = workload(args)
    ref_batch = args.ref_batch or 2
    steps, warm = max(1, min(args.steps, 4)), max(0, min(args.warmup, 1))
    r = cpu_reference_run(cfg, S, P, ref_batch, steps, warm)
    n = args.gpus
    line = {"impl": "reference", "metric": METRIC, "value": round(r["value"], 4), "unit": "sequences/s", "n_gpus": n, "steps": steps,
            "warmup": warm, "ms_per_step": round(r["ms_per_step"], 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config_dict(args, cfg, S, B, P, n),
            "cpu_baseline": {"value": round(r["value"], 4), "unit": "sequences/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": round(r["value"], 4), "unit": "sequences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference run_pretraining.py asserts CUDA (run_pretraining.py:325) and cannot run on CPU; this arm times the CPU "
                    "restatement of the same step (oracle/) on the box's host cores"}
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from deeplearningexamples_b200 import _lib as L
    from deeplearningexamples_b200 import kernels as K
    from deeplearningexamples_b200 import ops
    from deeplearningexamples_b200 import training as T

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a B200: there is no CPU path")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group(backend="nccl", init_method="env://", device_id=device)
    L.load()
    cfg, S, B, P = workload(args)
    if args.no_dropout:
        cfg["hidden_dropout_prob"] = cfg["attention_probs_dropout_prob"] = 0.0
    ops.manual_seed(42 + rank)
    model, opt, scaler, sched, crit, _ = T.prepare_model_and_optimizer(cfg, device, distributed=world > 1, bucket_cap_mb=args.bucket_mb,
                                                                      seed=42, static_masked_count=None if args.dynamic_mlm_gather else B * P)
    model.train()
    host = [T.synthetic_batch(B, S, cfg["vocab_size"], P, seed=T.rank_seed(42, rank) + 100 * i, pin=True) for i in range(4)]
    dev = [{k: v.to(device) for k, v in hb.items()} for hb in host[:2]]
    stage = {k: torch.empty_like(v, device=device) for k, v in host[0].items()}
    loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()
    loss_acc = torch.zeros(1, dtype=torch.float32, device=device)
    h2d = sum(v.numel() * v.element_size() for v in host[0].values())

    def step_resident(i):
        loss = T.take_training_step(scaler, model, crit, dev[i % 2], loss_acc)
        T.take_optimizer_step(sched, opt, scaler)
        return loss

    def step_e2e(i):
        hb = host[i % 4]
        for k in stage:
            stage[k].copy_(hb[k], non_blocking=True)
        loss = T.take_training_step(scaler, model, crit, stage, loss_acc)
        T.take_optimizer_step(sched, opt, scaler)
        loss_host.copy_(loss.detach().float().reshape(1), non_blocking=True)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t_host = time.perf_counter()
        for i in range(steps):
            fn(i)
        e1.record()
        t_host = time.perf_counter() - t_host          # host time to ENQUEUE the steps (no sync inside)
        barrier()
        log(f"  host enqueue {1000 * t_host / steps:.2f} ms/step vs device {e0.elapsed_time(e1) / steps:.2f} ms/step")
        return T.max_over_ranks(e0.elapsed_time(e1), device)

    log(f"model built: B={B} S={S} world={world}")
    for i in range(max(args.warmup, 3)):
        step_resident(i)
        torch.cuda.synchronize()
        log(f"warm-up step {i} done")
    step_e2e(0)
    torch.cuda.synchronize()
    log("warm-up done")
    sampler = ClockSampler(local) if rank == 0 else None
    t_start = time.time()
    if sampler:
        sampler.start()
        time.sleep(0.3)
    n0 = L.launch_count["n"]
    torch.cuda.nvtx.range_push("timed_resident")
    torch.cuda.profiler.start()          # ncu --profile-from-start off: capture exactly the timed region (all threads)
    ms_res = timed(step_resident, args.steps)
    torch.cuda.profiler.stop()
    torch.cuda.nvtx.range_pop()
    log(f"resident pass: {ms_res / args.steps:.2f} ms/step")
    launches = L.launch_count["n"] - n0
    K.gemm_profile = []
    ms_e2e = timed(step_e2e, args.steps)
    log(f"e2e pass: {ms_e2e / args.steps:.2f} ms/step")
    prof, K.gemm_profile = K.gemm_profile, None
    t_end = time.time()
    clocks = sampler.stop(t_start, t_end) if sampler else None
    final_loss = loss_host.item()

    n = world
    value = T.global_throughput(B, n, args.steps, ms_res)
    e2e = T.global_throughput(B, n, args.steps, ms_e2e)
    pk = peaks()
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r01_gemm_traffic.json")
    if os.path.exists(tpath):                     # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu capture
        tj = json.load(open(tpath))
        traffic, traffic_src = tj["avg_dram_traffic_bytes_per_launch"], (
            "profiles/r01_gemm_traffic.json (ncu --set full, 12 forward GEMM launches of this command; the same 12 launches move "
            f"{tj.get('avg_algorithmic_bytes_per_launch', 0)} algorithmic bytes on average -- algorithmic_bytes_per_launch_avg below averages ALL timed launches)")
    gemm_ms = sum(a.elapsed_time(b) for a, b, _, _ in prof)
    gemm_flops = sum(f for _, _, f, _ in prof)
    ach = gemm_flops / (gemm_ms / 1000.0) / 1e12 if gemm_ms > 0 else 0.0
    flops_seq = T.train_flops_per_seq(cfg, S, P)
    line = {"metric": METRIC, "value": round(value, 2), "unit": "sequences/s", "n_gpus": n, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(ms_res / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": config_dict(args, cfg, S, B, P, n),
            "e2e": {"value": round(e2e, 2), "unit": "sequences/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": round(ms_e2e / args.steps, 3)},
            "gpu_launches": launches, "hbm_peak_gb": round(torch.cuda.max_memory_allocated() / 2**30, 1),
            "roofline": {"kernel": "gemm_bf16_tcgen05_kernel (all dense projections/FFN, fwd+dgrad+wgrad)", "bound": "tensor",
                         "achieved": round(ach, 1), "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": round(ach / pk["tf_sustained"], 4),
                         "peak_source": pk["source"] + " (bf16_tflops_sustained: kernel timed inside a long step)",
                         "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch_avg": int(sum(2.0 * (m * k + n * k + m * n) for _, _, _, (m, n, k, *_r) in prof) / max(len(prof), 1)),
                         "launches_timed": len(prof), "share_of_step": round(gemm_ms / ms_e2e, 4),
                         "how": "CUDA events around every GEMM launch on the launching stream during the e2e timed pass; "
                                "algorithmic flops = 2*M*N*K per launch"},
            "model_flops_utilisation": {"train_gflop_per_seq": round(flops_seq / 1e9, 1),
                                        "achieved_tflops_per_gpu": round(value / n * flops_seq / 1e12, 1),
                                        "frac_of_sustained_peak": round(value / n * flops_seq / 1e12 / pk["tf_sustained"], 4)},
            "final_loss": round(final_loss, 4)}
    if clocks is not None:
        line["clocks"] = clocks
    if rank == 0 and n == 1 and not args.no_cpu_baseline:
        del model, opt
        torch.cuda.empty_cache()
        log("cpu baseline ...")
        r = cpu_reference_run(cfg, S, P, args.ref_batch or 2, 2, 1)
        log("cpu baseline done")
        line["cpu_baseline"] = {"value": round(r["value"], 4), "unit": "sequences/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
