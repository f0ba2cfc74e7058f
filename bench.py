#!/usr/bin/env python
"""Benchmark of the fused NMA ELBO + gradient + Adamax step: latent steps x MC samples per second.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
                  [--config ar_1e8|ar_default|lv_fix_theta|lv_batch|fhn|sv] [--scaling weak|strong] [--rows P] [--T n]

Headline (default): BASELINE.json configs[4] - synthetic AR(1) series, T = 10^8, kernel_len = 50, time-sharded over the
GPUs.  The other configs are the reference scripts' own shapes on synthetic series of the scripts'
own SDEs.  One JSON line on stdout (rank 0); DESIGN.md "Measurement" says what each key means.
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
    """Algorithmic forward MACs per row; fwd+bwd FLOP = 6 x MAC.  For the Lotka-Volterra models the
    feature MLP runs over the whole window and ends in a layer as wide as the flow's conv input (conv over 1 + L0-1
    channels)."""
    C = cfg.C
    lv = cfg.model in (3, 4, 5)
    if lv:
        LW = cfg.L0 - 1
        feat = sum(LW * (cfg.Cf_in * C + 2 * C * C + C * cfg.Lin(i)) for i in range(cfg.F))
        conv = sum(cfg.N(i) * cfg.K * (LW + 1) * C for i in range(cfg.F))
    else:
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
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "sm_max_mhz": 1965.0,
            "source": "fallback"}


def native_libs_loaded():
    """This repository's shared objects mapped into this process (read from /proc/self/maps)."""
    out = set()
    try:
        for line in open("/proc/self/maps"):
            path = line.rstrip("\n").split(" ")[-1]
            if path.endswith(".so") and os.path.realpath(path).startswith(os.path.realpath(ROOT)):
                out.add(os.path.relpath(os.path.realpath(path), os.path.realpath(ROOT)))
    except Exception:
        pass
    return sorted(out)


def ncu_traffic(mode, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel`, per ROW, from this round's
    `ncu --set full` capture of this very program (profiles/traffic.json, written by tools/summarize_ncu.py from the
    .ncu-rep; rows of the capture recorded there).  None when no capture of this build is committed."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        d = json.load(open(p))
        e = d[mode][kernel]
        return float(e["dram_bytes"]) / float(e["rows"]), e.get("source")
    except Exception:
        return None, None


# ----------------------------------------------------------------------------------------------
# workloads
# ----------------------------------------------------------------------------------------------
class AR1e8:
    """BASELINE.json configs[4]: AR(1), T = 10^8, generated on the device (A12/A13 scans), time-sharded."""
    name = "ar_1e8"

    def __init__(self, args, dev, rank, world):
        from viforssms_b200.trainer import ARStepper
        self.args, self.world = args, world
        rows = args.rows if args.rows else 16384
        if args.scaling == "strong":
            rows = max(rows // world, 1)
        self.rows = rows
        self.tc = 7 if args.conv_split == "bf16" else 3
        self.st = ARStepper(T=args.T, rows=rows, K=K_LEN, B=B_DIMS, F=FLOWS, H=HID, fw=FW, theta=THETA_TRUE, x0=10.0,
                            obs_std=1.0, device=dev, rank=rank, world=world, seed=1, tensor_cores=self.tc,
                            device_theta=not args.host_theta)
        self.cfg = self.st.cfg
        self.units_per_step_rank = rows * B_DIMS
        self.h2d_bytes, self.d2h_bytes = self.st.h2d_bytes, self.st.d2h_bytes

    def describe(self):
        a = self.args
        return {"workload": "AR(1) synthetic T=%d, kernel_len=50, batch_dims=50, no_flows=3, network_dims=50,50,50, "
                            "feat_window=10 (BASELINE.json configs[4])" % a.T,
                "rows_per_gpu_per_step": self.rows, "units_per_row": B_DIMS,
                "l2": "per-step working set (saved activations and tensor-core operands, ~1.1 MB per row: %.1f GB at these "
                      "rows) far exceeds the 126 MB L2; no flush needed" % (self.rows * 1.075e6 / 1e9),
                "parallelism": "time-sharded series, rows sharded %d-way, gradient all-reduce issued by the library "
                               "(NCCL, per flow, side stream)" % self.world,
                "launch": "eager" if a.no_graph else "one CUDA graph per step",
                "step": "one nma_train_step call: in-library Philox noise, theta posterior, ELBO + gradients, all-reduce, "
                        "clip + Adamax, logged means" if not a.host_theta else "host-composed (autograd theta posterior)",
                "conv_operands": self.operands()}

    def operands(self):
        return {7: "2-term bf16 split (3 products, kind::f16), fp32 accumulate",
                3: "3xTF32 split (kind::tf32), fp32 accumulate"}[self.tc]

    def dtype(self):
        return "bf16x2-split/f32acc" if self.tc == 7 else "tf32x3-split/f32acc"

    def prepare(self):
        from viforssms_b200 import lib as _lib
        L = _lib.load()
        if not self.args.no_graph:
            self.st.capture()
        else:
            n0 = L.nma_launch_count()
            self.st.step_resident()
            self.st.launches_per_step = int(L.nma_launch_count() - n0)

    @property
    def launches_per_step(self):
        return self.st.launches_per_step

    def step_resident(self):
        self.st.step_resident()

    def step_e2e(self):
        return self.st.step_e2e()

    def switch_operands(self, tc):
        """Re-run in the other operand split of the conv GEMMs (the graph is re-captured)."""
        st = self.st
        if st.graph is not None:
            torch.cuda.synchronize()
            st.graph.reset()
            st.graph = None
        st.eng.set_tensor_cores(tc)
        self.tc = tc
        self.prepare()

    STAGE_NAMES = {0: "conv_fwd", 1: "conv_dgrad", 2: "conv_wgrad", 3: "epi_bwd", 5: "feat_bwd"}

    def roofline(self, peaks, ms_step):
        """Times every kernel family of the step alone (CUDA events on the launching stream, re-launched on the
        workspace the last step left) and reports the dominant one against the tensor-pipe roofline: the conv is a dense
        contraction (K*(C+1) = 2550 deep, 50 wide), SURVEY section 8d.  achieved = ALGORITHMIC flops of that launch
        (2 * rows * N_i * K * 51 * 50) / its duration; peak = the measured bf16 figure for kind::f16 MMAs (bf16 split),
        half of it for kind::tf32."""
        st, cfg = self.st, self.cfg
        stages, best = {}, None
        for i in range(cfg.F):
            for sid, name in self.STAGE_NAMES.items():
                ms = st.time_stage(sid, i)
                stages["%s[%d]" % (name, i)] = round(ms, 4)
                if sid in (0, 1, 2):
                    flop = 2.0 * st.rows * cfg.N(i) * cfg.K * (cfg.C + 1) * cfg.C
                    if best is None or ms > best[0]:
                        best = (ms, "%s[%d]" % (name, i), flop)
        stages["feat_fwd[all]"] = round(st.time_stage(4, 0), 4)
        ms, name, flop = best
        bf = bool(st.eng.bf16_split)
        peak = (1.0 if bf else 0.5) * float(peaks["bf16_tflops"])
        achieved = flop / (ms * 1e-3) / 1e12
        conv_ms = sum(v for k, v in stages.items() if k.startswith("conv_"))
        per_row, src = ncu_traffic("bf16" if bf else "tf32", name)
        roof = {"bound": "tensor", "kernel": name, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": per_row * st.rows if per_row is not None else None,
                "traffic_note": ("dram bytes of this launch: %s, per row x rows" % src) if per_row is not None else
                                "no ncu capture of this build committed (profiles/traffic.json)",
                "peak_source": ("bf16_tflops (%s) of MEASURED_PEAKS.json: the kernel issues kind::f16 MMAs, 3 per algorithmic MAC"
                                if bf else "0.5 x bf16_tflops (%s) of MEASURED_PEAKS.json = TF32 dense") % peaks.get("source"),
                "ms_per_launch": ms, "flop_per_launch": flop,
                "conv_share_of_step": conv_ms / ms_step if ms_step > 0 else None}
        return roof, stages

    def cpu_inputs(self, rows):
        from oracle import nma_oracle as O
        from viforssms_b200.config import ar_config, param_layout
        cfg = ar_config(p=rows, K=K_LEN, B=B_DIMS, F=FLOWS, H=HID, feat_window=FW, T=10 ** 8)
        layout, n = param_layout(cfg)
        g = torch.Generator().manual_seed(1)
        params = O.glorot_init(layout, n, g)
        tf = torch.randn(rows, cfg.L0, cfg.Cf, generator=g)
        tf[:, :, -1] = 1.0
        return cfg, layout, n, params, tf, None, torch.tensor([4.0, 0.5, 1.0]), 2.5e8

    def cpu_rows(self):
        return self.args.cpu_rows or 400

    def close(self):
        self.st.close()


class FacadeWorkload:
    """configs[0..3]: the reference scripts' own shapes, driven through the drop-in VI_SSM classes (the call a user of
    the script makes: `_iteration(batch_select)` = one sess.run).  `step_resident` replays the iteration with the
    subsequence starts already on the device; `step_e2e` draws them with the script's own np.random.choice call, copies
    them in and reads the ELBO back."""

    def __init__(self, args, dev, rank, world):
        self.args, self.dev = args, dev
        if world != 1:
            raise SystemExit("--config %s is a single-GPU workload (the script's own shape); use ar_1e8 for --gpus > 1" % self.name)
        np.random.seed(1)
        self.m = self.build(dev)
        self.m.build_flow()
        self.cfg = self.m.cfg
        self.rows = self.cfg.p
        self.units_per_step_rank = self.cfg.p * self.cfg.B
        self.h2d_bytes, self.d2h_bytes = self.cfg.p * 8, 32
        self.launches_per_step = None

    def dtype(self):
        tc = int(self.m.eng._lib.nma_get_tensor_cores(self.m.eng._h))
        return "bf16x2-split/f32acc" if tc & 4 else ("tf32x3-split/f32acc" if tc & 1 else "f32")

    def describe(self):
        c = self.cfg
        return {"workload": self.title, "rows_per_gpu_per_step": c.p, "units_per_row": c.B,
                "shape": {"p": c.p, "kernel_len": c.K, "batch_dims": c.B, "flow_dims": c.D, "no_flows": c.F,
                          "hidden_1x1": c.H, "feature_channels": c.Cf, "dtheta": c.dtheta},
                "l2": "working set of %.1f MB fits the 126 MB L2; a 256 MB buffer is written between timed steps"
                      % (self.m.eng.workspace_bytes / 1e6),
                "parallelism": "single GPU",
                "launch": "eager" if os.environ.get("NMA_FACADE_GRAPH") == "0" else "one CUDA graph per step",
                "compute": self.dtype()}

    def _draw(self):
        return self.m._draw()

    def _run(self):
        self.m._main_iteration()

    def prepare(self):
        from viforssms_b200 import lib as _lib
        L = _lib.load()
        self.flush = torch.empty(64 << 20, dtype=torch.float32, device=self.dev)
        self.m.pre_train = False
        self.m.idx_dev.copy_(torch.from_numpy(np.ascontiguousarray(self._draw(), dtype=np.int64)))
        n0 = L.nma_launch_count()
        self._run()                      # eager
        self.launches_per_step = int(L.nma_launch_count() - n0)
        self._run()                      # capture + replay
        self._run()
        torch.cuda.synchronize()

    def step_resident(self):
        self.flush.zero_()               # 256 MB > L2: the next step starts cold
        self._run()

    def step_e2e(self):
        self.flush.zero_()
        idx = np.ascontiguousarray(self._draw(), dtype=np.int64)
        self.m.idx_dev.copy_(torch.from_numpy(idx))
        self._run()
        return float(self.m.scalars_dev[0].item())

    def roofline(self, peaks, ms_step):
        """Whole-step figure: these shapes (p = 1..200 rows) do not fill the machine, the step is latency-bound; achieved
        = algorithmic FLOP of the step / step time against the peak of the pipe its conv runs on."""
        fl = flops_per_row(self.cfg)
        achieved = fl["flop_step"] * self.cfg.p / (ms_step * 1e-3) / 1e12
        dt = self.dtype()
        if dt == "f32":
            clock = float(peaks.get("sm_max_mhz") or 1965.0)
            peak, src = 148 * 128 * 2 * clock * 1e6 / 1e12, "theoretical FP32 FMA (148 SMs x 128 lanes x 2 x %.0f MHz): the D=2 models run the FP32 SIMT conv; MEASURED_PEAKS.json has no FP32 figure" % clock
        elif dt.startswith("bf16"):
            peak, src = float(peaks["bf16_tflops"]), "bf16_tflops (%s) of MEASURED_PEAKS.json" % peaks.get("source")
        else:
            peak, src = 0.5 * float(peaks["bf16_tflops"]), "0.5 x bf16_tflops (%s) of MEASURED_PEAKS.json = TF32 dense" % peaks.get("source")
        # the step without the L2 flush in between (what the flush costs is not the kernels' time)
        st = torch.cuda.current_stream()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(20):
            self._run()
        e1.record(st)
        torch.cuda.synchronize()
        warm_ms = e0.elapsed_time(e1) / 20
        roof = {"bound": "tensor" if dt != "f32" else "fp32-fma", "kernel": "whole step (%d kernels, one CUDA graph)" % self.launches_per_step,
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
                "peak_source": src, "ms_per_launch": ms_step, "flop_per_launch": fl["flop_step"] * self.cfg.p,
                "ms_per_step_warm_l2": warm_ms,
                "note": "latency-bound shape: %d rows x %d kernels per step" % (self.cfg.p, self.launches_per_step)}
        return roof, {"step_cold_l2": round(ms_step, 4), "step_warm_l2": round(warm_ms, 4)}

    def cpu_rows(self):
        return self.args.cpu_rows or self.cfg.p

    def cpu_inputs(self, rows):
        """Inputs of the oracle step at this workload's shape: the windows the device gathers (so the CPU arm sees the same
        series), `rows` of them."""
        from oracle import nma_oracle as O
        from viforssms_b200.config import param_layout
        import copy
        cfg = copy.copy(self.cfg)
        cfg.p = rows
        layout, n = param_layout(cfg)
        idx = np.resize(np.ascontiguousarray(self._draw(), dtype=np.int64), rows)
        tf, mask, shift = self.m.eng.gather(torch.from_numpy(idx).to(self.dev)) if rows <= self.cfg.p else (None, None, None)
        if tf is None:
            raise SystemExit("--cpu-rows must not exceed the script's p for --config %s" % self.name)
        extra = self.cpu_extra(idx, tf.cpu(), mask.cpu(), shift.cpu())
        params = self.m.blob[:self.m.n_nma].cpu().clone()
        return cfg, layout, n, params, tf.cpu(), extra, self.theta_center(), float(self.m.grad_clip)

    def close(self):
        m = self.m
        torch.cuda.synchronize()
        for g in ([getattr(m, "_graph", None)] + list(getattr(m, "_graphs", {}).values())):
            if g is not None:
                g.reset()
        m.eng.close()


class ARDefault(FacadeWorkload):
    name = "ar_default"
    title = ("AR(1) default: python main.py hyperparameters.txt - dat/AR_obs_partial.txt (T=5000), p=50, kernel_len=50, "
             "batch_dims=50, 3 flows (BASELINE.json configs[0])")

    def build(self, dev):
        import AR as ar_mod
        d = os.path.join(ROOT, "dat")
        obs = np.loadtxt(os.path.join(d, "AR_obs_partial.txt"), np.float32)
        obs_bin = np.loadtxt(os.path.join(d, "AR_obs_binary.txt"), np.float32)
        tt = np.loadtxt(os.path.join(d, "AR_time_till.txt"), np.float32)
        self.series = (obs, obs_bin, tt)
        flow = ar_mod.ThetaFlow(3, 5, base_loc=1.5, base_scale=0.5, activation="elu")
        return ar_mod.VI_SSM(obs, 1.0, 10.0, flow, [(0., 10.0)] * 3, 5000, 50, 50, 50, [50] * 3, 3, 10, obs_bin, tt,
                             pre_train=False, device=dev)

    def _draw(self):
        return self.m._draw(False)

    def _run(self):
        self.m._run(False)

    def cpu_extra(self, idx, tf, mask, shift):
        return None

    def theta_center(self):
        return torch.tensor([4.0, 0.5, 1.0])


class FHN(FacadeWorkload):
    name = "fhn"
    title = ("FitzHugh-Nagumo SDE, RealNVP coupling flow (fitz_nag_NVP.py:452-465): synthetic Euler-Maruyama series of 10^5 "
             "steps at theta*, p=50, kernel_len=20, batch_dims=50, 3 flows, 3 hidden 1x1 + BN (BASELINE.json configs[2])")

    def build(self, dev):
        import fitz_nag_NVP as mod
        N, dt = 100000, 0.1
        obs, obs_bin, tt = mod.simulate(N, dt=dt)
        self.series = (obs, obs_bin, tt)
        flow = mod.ThetaFlow(5, 4, base_loc=0., base_scale=1., activation="elu")
        return mod.VI_SSM(obs, obs_bin, tt, np.array([2., 3.]), flow, [(0., 10.)] * 5, dt, N * dt, 50, 20, 50, [50] * 5, N,
                          3, 10, learn_rate=1e-4, pre_train=False, device=dev)

    def cpu_extra(self, idx, tf, mask, shift):
        obs_bin = self.series[1]
        B = self.cfg.B
        bf = np.stack([obs_bin[:, i:i + B] for i in idx])
        return {"bin_feed": torch.from_numpy(bf.astype(np.float32))}

    def theta_center(self):
        return torch.tensor([0.69, 1.0, 1.5, -0.69, -1.2])


class SV(FacadeWorkload):
    name = "sv"
    title = ("Stochastic volatility dense model (SV_dense.py:405-416): synthetic 1809-price series, p=200, kernel_len=50, "
             "batch_dims=52, 5 flows, 3 hidden 1x1 + BN (BASELINE.json configs[3])")

    def build(self, dev):
        import SV_dense as mod
        obs = mod.simulate(1809).astype(np.float32)[300:]
        T = obs.shape[0] - 1
        n = (T // 52) * 52                      # batch_dims must tile the series (the script's own 1508 = 29 x 52)
        obs = obs[:n + 1]
        self.series = (obs,)
        flow = mod.ThetaFlow(4, 5, base_loc=0., base_scale=1., activation="relu")
        return mod.VI_SSM(obs, -8.5, flow, [(0., 10.0)] * 4, 1.0, float(n), 200, 50, 52, [50] * 5, n, 5, 5,
                          learn_rate=1e-4, pre_train=False, device=dev)

    def cpu_extra(self, idx, tf, mask, shift):
        obs = self.series[0]
        B = self.cfg.B
        dim_one = np.stack([obs[i:i + B + 1] for i in idx])
        return {"mask": mask[:, 0], "shift": shift[:, 0], "dim_one": torch.from_numpy(dim_one.astype(np.float32))}

    def theta_center(self):
        return torch.tensor([0.001, -0.6, -2.5, -0.7])


class LVFix(FacadeWorkload):
    name = "lv_fix_theta"
    title = ("Lotka-Volterra partial-observation SDE, fixed theta (lotka_volterra_partial_batch_fix_theta.py:616-631): one "
             "synthetic 151-step series, p_val=1, kernel_len=20, batch_dims=151, 3 flows (BASELINE.json configs[1])")

    def build(self, dev):
        import lotka_volterra_partial_batch_fix_theta as mod
        obs = mod.simulate(1).astype(np.float32)
        self.series = (obs, np.ones_like(obs), np.zeros_like(obs))
        priors = mod.softplus_np_(np.array([-1.0, -6.0, -1.0, -2.0]))
        return mod.VI_SSM(obs, self.series[1], self.series[2], np.array([91., 99.], np.float32),
                          np.array([1., 1.], np.float32), priors, 0.2, 30, 1, 20, 151, [50] * 5, 151, 3, 10,
                          learn_rate=1e-3, pre_train=False, device=dev)

    def prepare(self):
        super().prepare()
        self.h2d_bytes = 8

    def cpu_extra(self, idx, tf, mask, shift):
        obs_bin = self.series[1]
        B = self.cfg.B
        bf = np.stack([obs_bin[:, i:i + B] for i in idx])
        return {"mask": mask, "shift": shift, "bin_feed": torch.from_numpy(bf.astype(np.float32))}

    def theta_center(self):
        return torch.tensor(np.log1p(np.exp([-1.0, -6.0, -1.0, -2.0])), dtype=torch.float32)


class LVBatch(FacadeWorkload):
    name = "lv_batch"
    title = ("Lotka-Volterra partial-observation SDE, learned softplus-theta (lotka_volterra_partial_batch.py:677-764, the file "
             "BASELINE.json configs[1] names): three synthetic 151-step series, p_val=3, kernel_len=20, batch_dims=151, 3 flows")

    def build(self, dev):
        import lotka_volterra_partial_batch as mod
        import lotka_volterra_partial_batch_fix_theta as fix
        obs = fix.simulate(3).astype(np.float32)
        self.series = (obs, np.ones_like(obs), np.zeros_like(obs))
        priors = [(-1.0, np.sqrt(0.1)), (-6.0, np.sqrt(0.1)), (-1.0, np.sqrt(0.1)), (-2.0, np.sqrt(0.1))]
        flow = mod.ThetaFlow(4, 4, base_loc=0., base_scale=1., activation="elu", softplus_out=True)
        return mod.VI_SSM(obs, self.series[1], self.series[2], np.array([91., 99.], np.float32), np.array([1., 1.], np.float32),
                          flow, priors, 0.2, 30, 3, 20, 151, [50] * 5, 151, 3, 10, learn_rate=1e-3, pre_train=False, device=dev)

    def cpu_extra(self, idx, tf, mask, shift):
        obs_bin = self.series[1]
        B = self.cfg.B
        bf = np.stack([obs_bin[:, i:i + B] for i in idx])
        return {"mask": mask, "shift": shift, "bin_feed": torch.from_numpy(bf.astype(np.float32))}

    def theta_center(self):
        return torch.tensor(np.log1p(np.exp([-1.0, -6.0, -1.0, -2.0])), dtype=torch.float32)


WORKLOADS = {c.name: c for c in (AR1e8, ARDefault, LVFix, LVBatch, FHN, SV)}


# ----------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the reference step on host cores
# ----------------------------------------------------------------------------------------------
def cpu_step_time(wl, rows, steps, warmup, threads):
    """Times the oracle (torch-CPU fp32 restatement of the script's train step incl. clip + Adamax) on `rows` rows of
    the workload's shape."""
    from oracle import nma_oracle as O
    torch.set_num_threads(threads)
    cfg, layout, n, params, tf, extra, theta_c, clip = wl.cpu_inputs(rows)
    g = torch.Generator().manual_seed(1)
    params = params.float()
    m = torch.zeros(n); v = torch.zeros(n)
    tf = tf.float()
    if extra is not None:
        extra = {k: t.float() for k, t in extra.items()}
    times = []
    for it in range(warmup + steps):
        eps = torch.randn(rows, cfg.L0, generator=g)
        theta = theta_c.float().repeat(rows, 1) + 0.05 * torch.randn(rows, cfg.dtheta, generator=g)
        t0 = time.perf_counter()
        r = O.step_reference(cfg, layout, params, eps, theta, tf, extra=extra)
        gn = float(r["grad_params"].norm())
        params, m, v = O.adamax_step(params, r["grad_params"], m, v, 1e-3, 0.95, clip=(clip, gn))
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return float(np.median(times)), float(np.sum(times)), cfg


class _CpuOnlyAR:
    """The AR T=10^8 workload's CPU arm needs no device."""
    name = "ar_1e8"
    cpu_inputs = AR1e8.cpu_inputs

    def __init__(self, args):
        self.args = args

    def cpu_rows(self):
        return self.args.cpu_rows or 400

    def describe(self):
        a = self.args
        rows = a.rows if a.rows else 16384
        return {"workload": "AR(1) synthetic T=%d, kernel_len=50, batch_dims=50, no_flows=3, network_dims=50,50,50, "
                            "feat_window=10 (BASELINE.json configs[4])" % a.T,
                "rows_per_gpu_per_step": rows if a.scaling == "weak" else max(rows // a.gpus, 1), "units_per_row": B_DIMS}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    if args.config == "ar_1e8":
        wl = _CpuOnlyAR(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("the reference arm of --config %s reads the workload's windows through the device gather" % args.config)
        torch.cuda.set_device(0)
        wl = WORKLOADS[args.config](args, torch.device("cuda", 0), 0, 1)
    rows = wl.cpu_rows()
    med, total, cfg = cpu_step_time(wl, rows, args.steps, args.warmup, threads)
    units = rows * cfg.B
    value = units / med
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": med * 1e3, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": wl.describe(),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d rows x %d steps of the same step (torch-CPU fp32 oracle port of the reference's graph; "
                                   "its TensorFlow 1.8 cannot be installed here)" % (rows, args.steps)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# native arm
# ----------------------------------------------------------------------------------------------
def timed(wl, steps, fn, barrier, world, dev, host_clock=False):
    st = torch.cuda.current_stream()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(st)
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    e1.record(st)
    barrier()
    host = (time.perf_counter() - t0) * 1e3
    ms = torch.tensor([max(e0.elapsed_time(e1), host) if host_clock else e0.elapsed_time(e1)], device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def run_native(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl native needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # NCCL announces itself on stdout when a communicator is created ("NCCL version ..."); stdout must carry exactly one
    # JSON line, so file descriptor 1 points at stderr until the communicators exist (restored before the timed region)
    saved_stdout = None
    if world > 1:
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, "launch with torchrun --nproc-per-node == --gpus"

    wl = WORKLOADS[args.config](args, dev, rank, world)
    cfg = wl.cfg
    fl = flops_per_row(cfg)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident leg ----------------
    wl.prepare()
    for _ in range(args.warmup):
        wl.step_resident()
    barrier()
    if saved_stdout is not None:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total = timed(wl, args.steps, wl.step_resident, barrier, world, dev)
    clocks = sampler.stop() if rank == 0 else None
    launches = wl.launches_per_step * args.steps     # kernels of this library per step (counted on an eager step) x steps

    # ---------------- end-to-end leg (host buffers, H2D + D2H inside the timed region) ----------------
    for _ in range(2):
        wl.step_e2e()
    e2e_ms_total = timed(wl, args.steps, wl.step_e2e, barrier, world, dev, host_clock=True)

    # ---------------- dominant kernel, timed alone; the other operand split; CPU baseline (rank 0) ----------------
    roof = stages = cpu_base = None
    alt = {}
    if rank == 0:
        peaks = measured_peaks()
        roof, stages = wl.roofline(peaks, ms_total / args.steps)
    if args.config == "ar_1e8" and not args.no_alt:
        other = 3 if wl.tc == 7 else 7
        main_tc = wl.tc
        wl.switch_operands(other)
        for _ in range(3):
            wl.step_resident()
        ms_alt = timed(wl, args.steps, wl.step_resident, barrier, world, dev)
        alt[wl.dtype()] = {"value": wl.units_per_step_rank * world * args.steps / (ms_alt * 1e-3), "unit": UNIT,
                           "ms_per_step": ms_alt / args.steps, "conv_operands": wl.operands()}
        wl.switch_operands(main_tc)
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        rows = wl.cpu_rows()
        med, total, ccfg = cpu_step_time(wl, rows, args.cpu_steps, 2, threads)
        cpu_base = {"value": rows * ccfg.B / med, "unit": UNIT, "cores": threads, "kind": "port",
                    "sample": "%d rows x %d steps (%.1f s) of the same step on the torch-CPU fp32 oracle; the reference's "
                              "TensorFlow 1.8 cannot be installed here" % (rows, args.cpu_steps, total)}

    if rank == 0:
        units_total = wl.units_per_step_rank * world * args.steps
        line = {
            "metric": METRIC, "value": units_total / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": wl.dtype(), "data": "synthetic", "config": wl.describe(),
            "workspace_bytes": int((wl.st if hasattr(wl, "st") else wl.m).eng.workspace_bytes),
            "clocks": clocks,
            "e2e": {"value": units_total / (e2e_ms_total * 1e-3), "unit": UNIT, "h2d_bytes_per_step": wl.h2d_bytes,
                    "d2h_bytes_per_step": wl.d2h_bytes, "ms_per_step": e2e_ms_total / args.steps},
            "gpu_launches": launches, "gpu_launches_per_step": wl.launches_per_step,
            "roofline": roof, "cpu_baseline": cpu_base, "stage_ms": stages, "other_operand_split": alt or None,
            "flop_per_unit": fl["flop_step"] / cfg.B,
            "achieved_tflops_step": fl["flop_step"] * wl.rows * world * args.steps / (ms_total * 1e-3) / 1e12,
            "native_so_loaded": native_libs_loaded(),
        }
        print(json.dumps(line), flush=True)
    # orderly teardown: graph -> library communicator -> handle (wl.close), then torch's process group, then a normal
    # interpreter exit (atexit hooks run)
    wl.close()
    barrier()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default="ar_1e8", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --rows per GPU; strong: --rows in total, split over the GPUs")
    ap.add_argument("--rows", type=int, default=None, help="ar_1e8: rows (MC samples x subsequences) per step (default 16384)")
    ap.add_argument("--T", type=int, default=10 ** 8)
    ap.add_argument("--cpu-rows", type=int, default=None, help="rows of the bounded CPU sample (ar_1e8: 400; throughput is flat in rows)")
    ap.add_argument("--cpu-steps", type=int, default=None, help="steps of the cpu_baseline leg (about 10 s of CPU work)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-alt", action="store_true", help="skip the leg in the other operand split of the conv GEMMs")
    ap.add_argument("--conv-split", default="bf16", choices=["tf32", "bf16"],
                    help="operand split of the conv GEMMs: 2-term bf16 on kind::f16 (library default for AR-type models) or 3xTF32")
    ap.add_argument("--host-theta", action="store_true",
                    help="comparison path: theta posterior as a host autograd module, torch.randn noise (~400 ATen launches per step)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from the host instead of one CUDA graph per step")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    small = args.config != "ar_1e8"
    if args.steps is None:
        args.steps = 1000 if small else 10      # small shapes: >= ~0.6 s of timed region, so the 100 ms clock sampler sees it
    if args.cpu_steps is None:
        args.cpu_steps = {"ar_1e8": 50, "ar_default": 100, "fhn": 100, "sv": 20, "lv_fix_theta": 100, "lv_batch": 40}[args.config]
    if args.no_graph and small:
        os.environ["NMA_FACADE_GRAPH"] = "0"
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
