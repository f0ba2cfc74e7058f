#!/usr/bin/env python
"""Headline benchmark: latent steps x MC samples per second of the fused NMA ELBO + gradient + Adamax
step on a synthetic AR(1) series (BASELINE.json configs[4]: T = 10^8, kernel_len = 50, time-sharded).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--rows P] [--T n]

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for what each key means.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "latent steps x MC samples per second, fused NMA ELBO+gradient+Adamax step"
UNIT = "latent-step-samples/s"
K_LEN, B_DIMS, FLOWS, FW, HID = 50, 50, 3, 10, 1
THETA_TRUE = (5.0, 0.5, 3.0)


def flops_per_row(cfg):
    """Algorithmic forward MACs per row; fwd+bwd FLOP = 6 x MAC."""
    C = cfg.C
    feat = sum(cfg.Lin(i) * (cfg.Cf_in * C + 3 * C * C) for i in range(cfg.F))
    conv = sum(cfg.N(i) * cfg.K * (C + 1) * C for i in range(cfg.F))
    pw = sum(cfg.N(i) * cfg.H * C * C for i in range(cfg.F))
    head = sum((cfg.N(i) // (2 if cfg.D == 2 else 1)) * 2 * C for i in range(cfg.F))
    mac = feat + conv + pw + head
    return {"mac_fwd": mac, "flop_step": 6 * mac, "conv_mac": conv}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            d["source"] = "measured"
            return d
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# ----------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the reference step on host cores
# ----------------------------------------------------------------------------------------------
def cpu_step_time(rows, steps, warmup, threads):
    """Times the oracle (torch-CPU fp32 restatement of AR.py's train step incl. Adamax) on `rows` rows."""
    from oracle import nma_oracle as O
    from viforssms_b200.config import ar_config, param_layout
    torch.set_num_threads(threads)
    cfg = ar_config(p=rows, K=K_LEN, B=B_DIMS, F=FLOWS, H=HID, feat_window=FW, T=10 ** 8)
    layout, n = param_layout(cfg)
    g = torch.Generator().manual_seed(1)
    params = O.glorot_init(layout, n, g)
    m = torch.zeros(n); v = torch.zeros(n)
    tf = torch.randn(rows, cfg.L0, cfg.Cf, generator=g)
    tf[:, :, -1] = 1.0
    times = []
    for it in range(warmup + steps):
        eps = torch.randn(rows, cfg.L0, generator=g)
        theta = torch.tensor(THETA_TRUE).log().abs().repeat(rows, 1) * 0 + torch.tensor([4.0, 0.5, 1.0]) \
            + 0.1 * torch.randn(rows, 3, generator=g)
        t0 = time.perf_counter()
        r = O.step_reference(cfg, layout, params, eps, theta, tf)
        gn = float(r["grad_params"].norm())
        params, m, v = O.adamax_step(params, r["grad_params"], m, v, 1e-3, 0.95, clip=(2.5e8, gn))
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return float(np.median(times)), float(np.sum(times)), cfg


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    rows = args.cpu_rows
    med, total, cfg = cpu_step_time(rows, args.steps, args.warmup, threads)
    units = rows * B_DIMS
    value = units / med
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": med * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),      # the native arm's config, key for key; the bounded sample is described below
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d rows x %d steps of the same AR(1) K=50 B=50 3-flow step (torch-CPU fp32 oracle, "
                                   "the reference's TensorFlow 1.8 cannot be installed)" % (rows, args.steps)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, rows_note=None):
    c = {"workload": "AR(1) synthetic T=%d, kernel_len=50, batch_dims=50, no_flows=3, network_dims=50,50,50, "
                     "feat_window=10 (BASELINE.json configs[4])" % args.T,
         "rows_per_gpu_per_step": args.rows, "units_per_row": B_DIMS,
         "l2": "per-step working set (saved activations and tensor-core operands, ~1.1 MB per row: %.1f GB at these rows) far "
               "exceeds the 126 MB L2; no flush needed" % (args.rows * 1.075e6 / 1e9),
         "parallelism": "time-sharded series, rows sharded %d-way, NCCL gradient all-reduce" % args.gpus,
         "launch": "eager" if args.no_graph else "one CUDA graph per step",
         "conv_operands": {"bf16": "2-term bf16 split (3 products, kind::f16), fp32 accumulate",
                           "tf32": "3xTF32 split (kind::tf32), fp32 accumulate"}[args.conv_split]}
    if rows_note:
        c["note"] = rows_note
    return c



STAGE_NAMES = {0: "conv_fwd", 1: "conv_dgrad", 2: "conv_wgrad", 3: "epi_bwd", 5: "feat_bwd"}
# dram__bytes_read.sum + dram__bytes_write.sum per ROW of the flow-0 launch of each tcgen05 conv kernel, from the
# ncu --set full capture in profiles/r01_final.md (2048 rows: 301.2 / 239.0 / 1346.5 MB); scaled by rows below
NCU_DRAM_BYTES_PER_ROW = {"conv_fwd[0]": 301.2e6 / 2048, "conv_dgrad[0]": 239.0e6 / 2048, "conv_wgrad[0]": 1346.5e6 / 2048}
# the same for the bf16-split kernels (profiles/r01_bf16.md: 119.6+90.5 / 105.7+43.6 / 552.7+14.4 MB at 2048 rows)
NCU_DRAM_BYTES_PER_ROW_BF16 = {"conv_fwd[0]": 210.1e6 / 2048, "conv_dgrad[0]": 149.3e6 / 2048, "conv_wgrad[0]": 567.1e6 / 2048}


def roofline_report(stepper, peaks, ms_step):
    """Times every kernel family of the step alone (CUDA events on the launching stream, re-launched on
    the workspace the last step left) and reports the dominant one against the tensor-pipe roofline.

    The conv is a dense contraction (K*(C+1) = 2550 deep, 50 wide): the governing roofline is the tensor
    pipe, not HBM.  achieved = ALGORITHMIC flops of that launch (2 * rows * N_i * K * 51 * 50,
    DESIGN.md) / its duration.  peak = TF32 dense, taken as 1/2 of the MEASURED bf16 burst figure in
    MEASURED_PEAKS.json (the file has no TF32 number; tcgen05 kind::tf32 runs at half the kind::f16 rate)."""
    cfg = stepper.cfg
    stages = {}
    best = None
    for i in range(cfg.F):
        for st, name in STAGE_NAMES.items():
            ms = stepper.time_stage(st, i)
            stages["%s[%d]" % (name, i)] = round(ms, 4)
            if st in (0, 1, 2):
                flop = 2.0 * stepper.rows * cfg.N(i) * cfg.K * (cfg.C + 1) * cfg.C
                if best is None or ms > best[0]:
                    best = (ms, "%s[%d]" % (name, i), flop)
    stages["feat_fwd[all]"] = round(stepper.time_stage(4, 0), 4)
    ms, name, flop = best
    bf = bool(stepper.eng.bf16_split)
    # kind::f16 (the bf16 split) runs at the measured bf16 figure, kind::tf32 at half of it
    tf32_peak = (1.0 if bf else 0.5) * float(peaks["bf16_tflops"])
    achieved = flop / (ms * 1e-3) / 1e12
    conv_ms = sum(v for k, v in stages.items() if k.startswith("conv_"))
    roof = {"bound": "tensor", "kernel": name, "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s",
            "frac": achieved / tf32_peak,
            "traffic": ((NCU_DRAM_BYTES_PER_ROW_BF16 if bf else NCU_DRAM_BYTES_PER_ROW)[name] * stepper.rows
                        if (name in NCU_DRAM_BYTES_PER_ROW and stepper.eng.tensor_cores and cfg.K == 50) else None),
            "traffic_note": "dram bytes of this launch: ncu --set full at 2048 rows (profiles/%s) scaled by rows"
                            % ("r01_bf16.md" if bf else "r01_final.md"),
            "peak_source": ("bf16_tflops (%s) of MEASURED_PEAKS.json: the kernel issues kind::f16 MMAs, 3 per algorithmic MAC"
                            if bf else "0.5 x bf16_tflops (%s) of MEASURED_PEAKS.json = TF32 dense") % peaks.get("source", "measured"),
            "ms_per_launch": ms, "flop_per_launch": flop,
            "conv_share_of_step": conv_ms / ms_step if ms_step > 0 else None}
    return roof, stages

# ----------------------------------------------------------------------------------------------
# native arm
# ----------------------------------------------------------------------------------------------
def run_native(args):
    import torch.distributed as dist
    from viforssms_b200.config import ar_config, param_layout
    from viforssms_b200.trainer import ARStepper

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl native needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # NCCL announces itself on stdout when its communicator is created ("NCCL version ..."); stdout must carry exactly one
    # JSON line, so file descriptor 1 points at stderr until the communicators exist (restored before the timed region)
    saved_stdout = None
    if world > 1:
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, "launch with torchrun --nproc-per-node == --gpus"

    stepper = ARStepper(T=args.T, rows=args.rows, K=K_LEN, B=B_DIMS, F=FLOWS, H=HID, fw=FW, theta=THETA_TRUE,
                        x0=10.0, obs_std=1.0, device=dev, rank=rank, world=world, seed=1,
                        tensor_cores=7 if args.conv_split == "bf16" else 3, device_theta=args.device_theta)
    cfg = stepper.cfg
    fl = flops_per_row(cfg)
    units_step_rank = args.rows * B_DIMS
    from viforssms_b200 import lib as _lib
    L = _lib.load()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident leg ----------------
    if not args.no_graph:
        stepper.capture()          # the iteration as one CUDA graph (3 eager warm-up steps inside)
    else:
        n0 = L.nma_launch_count()
        stepper.step_resident()
        stepper.launches_per_step = int(L.nma_launch_count() - n0)
    for _ in range(args.warmup):
        stepper.step_resident()
    barrier()
    if saved_stdout is not None:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    st = torch.cuda.current_stream()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(args.steps):
        stepper.step_resident()
    e1.record(st)
    barrier()
    launches = stepper.launches_per_step * args.steps     # kernels of this library per step (counted on an eager step) x steps
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    clocks = sampler.stop() if rank == 0 else None

    # ---------------- end-to-end leg (host buffers, H2D + D2H inside the timed region) ----------------
    for _ in range(2):
        stepper.step_e2e()
    barrier()
    f0 = torch.cuda.Event(enable_timing=True); f1 = torch.cuda.Event(enable_timing=True)
    f0.record(st)
    t_host0 = time.perf_counter()
    for _ in range(args.steps):
        stepper.step_e2e()
    f1.record(st)
    barrier()
    t_host = time.perf_counter() - t_host0
    ms2 = torch.tensor([max(f0.elapsed_time(f1), t_host * 1e3)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_ms_total = float(ms2.item())

    # ---------------- dominant kernel, timed alone (rank 0) ----------------
    roof = None
    cpu_base = None
    stages = None
    if rank == 0:
        peaks = measured_peaks()
        roof, stages = roofline_report(stepper, peaks, ms_total / args.steps)
        if world == 1 and not args.no_cpu:
            threads = os.cpu_count() or 1
            med, total, _ = cpu_step_time(args.cpu_rows, args.cpu_steps, 2, threads)
            cpu_base = {"value": args.cpu_rows * B_DIMS / med, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": "%d rows x %d steps (%.1f s) of the same step on the torch-CPU fp32 oracle; the "
                                  "reference's TensorFlow 1.8 cannot be installed here" % (args.cpu_rows, args.cpu_steps, total)}

    if rank == 0:
        units_total = units_step_rank * world * args.steps
        value = units_total / (ms_total * 1e-3)
        wc = workload_config(args)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": wc, "workspace_bytes": int(stepper.eng.workspace_bytes),
            "clocks": clocks,
            "e2e": {"value": units_total / (e2e_ms_total * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": stepper.h2d_bytes, "d2h_bytes_per_step": stepper.d2h_bytes,
                    "ms_per_step": e2e_ms_total / args.steps},
            "gpu_launches": launches,
            "roofline": roof, "cpu_baseline": cpu_base, "stage_ms": stages,
            "flop_per_unit": fl["flop_step"] / B_DIMS,
            "achieved_tflops_step": fl["flop_step"] * args.rows * world * args.steps / (ms_total * 1e-3) / 1e12,
        }
        print(json.dumps(line), flush=True)
    # Leave without tearing NCCL down: the CUDA graph holds captured NCCL kernels, and destroying the communicator
    # while such a graph is alive blocks forever (seen on 8 GPUs: the line was printed, the job never exited).
    stepper.close()
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--rows", type=int, default=16384, help="rows (MC samples x subsequences) per GPU per step")
    ap.add_argument("--T", type=int, default=10 ** 8)
    ap.add_argument("--cpu-rows", type=int, default=400, help="rows of the bounded CPU sample (throughput is flat in rows: +8 %% from 100 to 1000)")
    ap.add_argument("--cpu-steps", type=int, default=50, help="steps of the cpu_baseline leg (about 10 s of CPU work on the 16-thread box)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--conv-split", default="bf16", choices=["tf32", "bf16"],
                    help="operand split of the conv GEMMs: 3xTF32 (kind::tf32) or 2-term bf16 (kind::f16, twice the rate)")
    ap.add_argument("--device-theta", action="store_true",
                    help="theta posterior through nma_theta_flow_fwd/_bwd instead of the host autograd module "
                         "(written without a GPU at hand: opt-in until tests/test_gpu_unverified.py has passed)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from the host instead of one CUDA graph per step")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
