#!/usr/bin/env python
"""Benchmark of the BASELINE.json metric: Flow.log_prob samples/s on the 10-layer RQ-NSF, D=784, batch 2^20.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--rows R]

One "step" = one Flow.log_prob pass over one synthetic Gaussian batch of R rows per GPU (weak scaling: every rank
owns R rows and a replica of the weights; the only collective is the all-gather of per-sample log-probs).  Prints
ONE JSON line (rank 0).  See DESIGN.md section "Measurement" for what each key means.
"""
import os
os.environ.setdefault("NCCL_DEBUG", "WARN")      # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FEATURES, HIDDEN, LAYERS, BINS, BLOCKS = 784, 256, 10, 8, 2
METRIC = "Flow.log_prob samples/sec, 10-layer RQ-NSF D=784"
# algorithmic work per sample
D_ID = FEATURES // 2
M_PARAMS = 3 * BINS - 1
FLOP_FINAL_PER_ROW = 2 * HIDDEN * (D_ID * M_PARAMS)                      # final conditioner layer, one coupling
FLOP_PER_SAMPLE = LAYERS * (2 * (D_ID * HIDDEN + 2 * BLOCKS * HIDDEN * HIDDEN + HIDDEN * D_ID * M_PARAMS)
                            + 2 * FEATURES * FEATURES)


#: MMA columns issued per useful parameter column in the fused kernel at cfg 3: 24/23 rows per feature, 400/392 features
FUSED_PAD = (24.0 / 23.0) * (400.0 / 392.0)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p.get(
            "bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        time.sleep(0.05)
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        smax = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def build_flow(seed=0):
    from nflows_b200.flows import recipes
    torch.manual_seed(seed)
    return recipes.perturb_(recipes.rq_nsf(FEATURES, HIDDEN, LAYERS, num_bins=BINS, tail_bound=3.0, num_blocks=BLOCKS).eval())


def cpu_oracle_rate(flow, budget_s=12.0, chunk=2048, max_rows=1 << 15):
    """The CPU restatement of the reference path (oracle/, torch ATen fp32, all host threads) on a bounded sample."""
    from oracle import flow_oracle as O
    sd = {k: v.detach().cpu().clone() for k, v in flow.state_dict().items()}
    spec = O.nsf_spec(LAYERS, num_bins=BINS, tail_bound=3.0)
    g = torch.Generator().manual_seed(123)
    x = torch.randn(chunk, FEATURES, generator=g)
    with torch.no_grad():
        O.flow_log_prob(sd, spec, x[:256])      # warm-up
        rows, t0 = 0, time.perf_counter()
        while rows < max_rows and time.perf_counter() - t0 < budget_s:
            O.flow_log_prob(sd, spec, x)
            rows += chunk
        dt = time.perf_counter() - t0
    return rows / dt, rows, torch.get_num_threads()


def torch_cuda_rate(flow, dev, budget_s=6.0, chunk=1 << 14, max_rows=1 << 18):
    """SURVEY.md section 8(d) "reference-CUDA baseline": the same restatement of the reference path (oracle/, plain torch ops,
    true-fp32 matmuls) with weights and data on the GPU, chunked so the [chunk, d_t * 23] parameter tensor the reference
    materialises fits.  A reported context figure, like cpu_baseline; none of this package's kernels run here."""
    from oracle import flow_oracle as O
    torch.backends.cuda.matmul.allow_tf32 = False
    sd = {k: v.detach().to(dev).clone() for k, v in flow.state_dict().items()}
    spec = O.nsf_spec(LAYERS, num_bins=BINS, tail_bound=3.0)
    g = torch.Generator(device=dev).manual_seed(123)
    x = torch.randn(chunk, FEATURES, device=dev, generator=g)
    with torch.no_grad():
        O.flow_log_prob(sd, spec, x[:1024])
        torch.cuda.synchronize()
        rows, t0 = 0, time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        while rows < max_rows and time.perf_counter() - t0 < budget_s:
            O.flow_log_prob(sd, spec, x)
            rows += chunk
            torch.cuda.synchronize()
        e1.record()
        torch.cuda.synchronize()
    return rows / (e0.elapsed_time(e1) * 1e-3), rows


def spline_hbm_roofline(dev, peaks, rows=1 << 20, d_t=32, bins=8, iters=10):
    """The HBM-bound spline segment (BASELINE configs[1] shape): conditioner output [rows, d_t*(3K-1)] resident in HBM ->
    nfk_rqs_rows.  Algorithmic bytes = 4*(M+2) per transformed element + 4 per identity element copied (read+write)."""
    from nflows_b200 import _native as N
    from nflows_b200 import kernels as K
    m = 3 * bins - 1
    g = torch.Generator(device=dev).manual_seed(7)
    x = torch.randn(rows, 2 * d_t, device=dev, generator=g)
    params = torch.randn(rows, d_t * m, device=dev, generator=g)
    t_cols = torch.arange(0, 2 * d_t, 2, device=dev, dtype=torch.int32)
    id_cols = torch.arange(1, 2 * d_t, 2, device=dev, dtype=torch.int32)
    lad = torch.zeros(rows, device=dev)
    y = torch.empty_like(x)
    flags = K.new_flags(dev)
    desc = N.spline_desc(bins, "linear", 3.0, 0, 1, 0, 1, 1e-3, 1e-3, 1e-3, False, 128.0 ** 0.5)
    for _ in range(3):
        K.rqs_rows(desc, False, x, params, t_cols, id_cols, lad, flags, out=y)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        K.rqs_rows(desc, False, x, params, t_cols, id_cols, lad, flags, out=y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    nbytes = rows * (d_t * (4 * (m + 2)) + d_t * 8 + 8)
    achieved = nbytes / (ms * 1e-3) / 1e9
    return {"kernel": "rqs_rows_kernel<8,exact>", "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": achieved / peaks["hbm_gbs"], "traffic": None, "avg_launch_ms": ms,
            "workload": "rows=2^20 d_t=32 K=8 params in HBM (3.2 GB > L2)", "peak_kind": "copy bandwidth, %s" % peaks["source"]}


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port; the Python reference itself cannot travel to the GPU box)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    flow = build_flow()
    rates, rows_total = [], 0
    for i in range(args.warmup + args.steps):
        r, rows, threads = cpu_oracle_rate(flow, budget_s=args.ref_budget, max_rows=args.ref_rows)
        if i >= args.warmup:
            rates.append(r)
            rows_total += rows
    value = sum(rates) / len(rates)
    sample = "{} rows per step in chunks of 2048 (of the {}-row workload)".format(rows_total // max(1, args.steps), args.rows)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * (rows_total / max(1, args.steps)) / value, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg3 10x[ActNorm,RandPerm+LULinear,RQ-coupling(H=256,2 blocks,K=8,B=3)] D=784 log_prob",
                   "rows_per_gpu": args.rows},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_native(args):
    import torch.distributed as dist

    from nflows_b200 import _native
    from nflows_b200 import kernels as K

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()

    flow = build_flow().to(dev)
    rows = args.rows
    gen = torch.Generator(device=dev).manual_seed(1 + rank)
    x = torch.randn(rows, FEATURES, device=dev, generator=gen)
    gathered = torch.empty(world * rows, device=dev) if world > 1 else None

    def step(inp):
        lp = flow.log_prob(inp)
        if world > 1:
            dist.all_gather_into_tensor(gathered, lp)
            return gathered
        return lp

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(args.warmup):
            step(x)
        # ---- device-resident timing ----------------------------------------------------------------------------
        sampler = ClockSampler(local)
        fence()
        if rank == 0:
            sampler.start()
        launches0 = _native.launch_count()
        K.TIMELINE = [] if rank == 0 else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            out = step(x)
        e1.record()
        fence()
        timeline, K.TIMELINE = K.TIMELINE, None
        launches = _native.launch_count() - launches0
        clocks = sampler.stop() if rank == 0 else None
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms_total = float(ms.item())

        # ---- end to end: pinned host inputs -> H2D -> log_prob -> D2H of the result, every step ------------------
        host_x = torch.empty(rows, FEATURES, pin_memory=True)
        host_x.copy_(x)
        host_out = torch.empty(out.numel(), pin_memory=True)
        e2e_steps = max(1, min(args.steps, 3))
        fence()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        from nflows_b200 import sharding
        lp_dev = torch.empty(rows, device=dev)
        for _ in range(e2e_steps):
            sharding.log_prob_streamed(flow, host_x, dev, chunk_rows=args.e2e_chunk, out=lp_dev)   # H2D overlapped with the kernels
            if world > 1:
                dist.all_gather_into_tensor(gathered, lp_dev)
                host_out.copy_(gathered, non_blocking=True)
            else:
                host_out.copy_(lp_dev, non_blocking=True)
        f1.record()
        fence()
        ms2 = torch.tensor([f0.elapsed_time(f1)], device=dev)
        if world > 1:
            dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
        e2e_ms = float(ms2.item()) / e2e_steps

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_per_step = ms_total / args.steps
    value = world * rows / (ms_per_step * 1e-3)
    result = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "cfg3 10x[ActNorm,RandPerm+LULinear,RQ-coupling(H=256,2 blocks,K=8,B=3)] D=784 log_prob",
                   "rows_per_gpu": rows, "global_batch": world * rows, "parallelism": "dp%d batch-shard" % world,
                   "l2": "inputs (3.3 GB/GPU) exceed the 126 MB L2; no flush needed", "peaks": peaks["source"],
                   "arithmetic": "fp32 in / fp32 out; dense layers multiply fp16 (hi,lo) split pairs with 3 tcgen05 kind::f16 "
                                 "MMAs per product (22-bit operands) and accumulate in fp32"},
        "clocks": clocks, "gpu_launches": int(launches),
        "e2e": {"value": world * rows / (e2e_ms * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": rows * FEATURES * 4,
                "d2h_bytes_per_step": int(out.numel()) * 4, "ms_per_step": e2e_ms},
        "tflops_effective": FLOP_PER_SAMPLE * value / 1e12,
    }
    # ---- roofline of the dominant kernel (final conditioner layer + spline), timed live with CUDA events -----------
    if timeline:
        torch.cuda.synchronize()
        tags = {}
        for tag, rows_k, a, b in timeline:
            t = tags.setdefault(tag, [0.0, 0, 0])
            t[0] += a.elapsed_time(b)
            t[1] += 1
            t[2] += rows_k
        result["timeline_ms_per_step"] = {k: round(v[0] / args.steps, 3) for k, v in sorted(tags.items(), key=lambda kv: -kv[1][0])}
        top = max(tags.items(), key=lambda kv: kv[1][0])
        tag, (tms, count, rows_k) = top
        flops = FLOP_FINAL_PER_ROW * rows_k
        achieved = flops / (tms * 1e-3) / 1e12
        peak = peaks["bf16_tflops_sustained"]
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            t = json.load(open(tpath))
            if t.get("kernel") == tag and "dram_bytes_per_row" in t:
                traffic = t["dram_bytes_per_row"] * (rows_k // count)     # per launch, like `achieved`
        result["roofline"] = {"kernel": tag, "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                              "frac": achieved / peak, "traffic": traffic, "launches": count,
                              "avg_launch_ms": tms / count, "share_of_step": tms / ms_total,
                              "peak_kind": "bf16 dense sustained, %s" % peaks["source"],
                              "mma_tflops_executed": 3.0 * achieved * FUSED_PAD,
                              "note": "achieved counts ALGORITHMIC flops (2 per weight per row); the kernel executes 3 "
                                      "fp16 MMAs per algorithmic multiply-add plus tile padding (mma_tflops_executed)"}
    if not args.no_spline_roofline:
        result["roofline_spline"] = spline_hbm_roofline(dev, peaks)
    # ---- CPU baseline (oracle port) on this box's host cores ---------------------------------------------------
    if world == 1 and not args.no_cpu_baseline:
        try:
            rate, sample_rows = torch_cuda_rate(flow, dev)
            result["torch_cuda_baseline"] = {"value": rate, "unit": "samples/s", "kind": "port", "sample":
                                             "%d rows in chunks of 16384, torch eager fp32 (allow_tf32 off) on the same GPU" % sample_rows}
        except Exception as exc:     # a context figure only: never let it take the bench line down
            result["torch_cuda_baseline"] = {"unavailable": "%s: %s" % (type(exc).__name__, exc)}
        rate, sample_rows, threads = cpu_oracle_rate(flow.cpu(), budget_s=args.ref_budget, max_rows=args.ref_rows)
        result["cpu_baseline"] = {"value": rate, "unit": "samples/s", "cores": threads, "kind": "port",
                                  "sample": "%d rows of the same workload in chunks of 2048" % sample_rows}
    print(json.dumps(result))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--rows", type=int, default=1 << 20, help="rows per GPU (BASELINE: 2^20)")
    ap.add_argument("--ref-budget", type=float, default=12.0, help="seconds of CPU work per reference step")
    ap.add_argument("--ref-rows", type=int, default=1 << 15)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--block-rows", type=int, default=0, help="override config.{trunk,affine,coupling}_block_rows (experiments)")
    ap.add_argument("--no-spline-roofline", action="store_true")
    ap.add_argument("--e2e-chunk", type=int, default=1 << 17, help="rows per host->device chunk of the end-to-end leg")
    args = ap.parse_args()
    if args.block_rows:
        from nflows_b200 import config
        config.trunk_block_rows = config.affine_block_rows = config.coupling_block_rows = args.block_rows
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
