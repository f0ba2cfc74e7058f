#!/usr/bin/env python
"""HBM-roofline figures of the genuinely streaming kernels: the A12 look-back scan and the A13
time-till kernel at T = 10^8, clip + Adamax at 10^8 variables (the step's own 432 k variables are launch-bound: reported
too), the Philox noise fill and the window gather.  achieved = ALGORITHMIC bytes / CUDA-event time; peak = hbm_gbs of
MEASURED_PEAKS.json.  One JSON object per line."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    st = torch.cuda.current_stream()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        fn()
    e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    from viforssms_b200.config import ar_config
    from viforssms_b200.engine import NMAEngine, philox_normal, scan_ar1, time_till
    from viforssms_b200 import lib as _lib, feed
    import numpy as np
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        src = "measured"
    except Exception:
        peak, src = 6650.0, "fallback"
    dev = torch.device("cuda", 0)
    L = _lib.load()
    out = []

    def line(kernel, what, nbytes, ms, note=""):
        gbs = nbytes / (ms * 1e-3) / 1e9
        d = {"kernel": kernel, "what": what, "algorithmic_bytes": nbytes, "ms": round(ms, 4), "achieved": round(gbs, 1),
             "peak": peak, "peak_source": src, "unit": "GB/s", "frac": round(gbs / peak, 4), "bound": "hbm", "note": note}
        print(json.dumps(d), flush=True)
        out.append(d)

    n = 10 ** 8
    z = torch.randn(n, dtype=torch.float64, device=dev)
    x = torch.empty(n + 1, dtype=torch.float64, device=dev)
    scratch = torch.empty((int(L.nma_scan_scratch_bytes(n)) + 7) // 8, dtype=torch.float64, device=dev)

    def scan():
        _lib.check(L.nma_scan_ar1(z.data_ptr(), x.data_ptr(), n, 10.0, 0.5, 5.0, 3.0, scratch.data_ptr(), scratch.numel() * 8,
                                  torch.cuda.current_stream().cuda_stream), "scan")
    line("k_scan_lookback<ar1>", "A12 AR(1) simulate, T=1e8, fp64 (AR_dat_gen.py:11-15): 8 B noise in + 8 B state out per step",
         16 * n, timeit(scan), "single pass, decoupled look-back; includes the memset of the tile descriptors")
    fill = torch.empty(n, dtype=torch.float64, device=dev); binary = torch.empty_like(fill); till = torch.empty_like(fill)

    def tt():
        _lib.check(L.nma_time_till(x.data_ptr(), n, 1, fill.data_ptr(), binary.data_ptr(), till.data_ptr(),
                                   torch.cuda.current_stream().cuda_stream), "time_till")
    line("k_time_till", "A13 hold-fill / indicator / time-till, T=1e8, impute=1 (AR_dat_gen.py:17-31): 8 B in, 24 B out", 32 * n, timeit(tt))
    del z, x, fill, binary, till

    eng = NMAEngine(ar_config(p=4, K=4, B=4, F=1, H=1, feat_window=2, T=100), dev)
    for nn, label in ((10 ** 8, "1e8 variables"), (432286, "the step's 432 286 variables (launch-bound)")):
        w = torch.randn(nn, device=dev); g = torch.randn(nn, device=dev); m = torch.rand(nn, device=dev); v = torch.randn(nn, device=dev)
        ms = timeit(lambda: eng.adamax_step(w, g, m, v, 1e-3, 0.95, clip=2.5e8))
        line("k_sumsq + k_adamax", "A9 global norm + clip + Adamax, %s: reads w,g,m,v and g again, writes w,m,v = 32 B per variable" % label,
             32 * nn, ms)
        del w, g, m, v
    nn = 16384 * 201
    buf = torch.empty(nn, device=dev)
    ms = timeit(lambda: L.nma_philox_normal(buf.data_ptr(), nn, 1, 0, 0, 0.0, 1.0, torch.cuda.current_stream().cuda_stream))
    line("k_philox_normal", "base noise eps [16384, 201] (AR.py:31-35), Philox4x32-10 + Box-Muller: 4 B out per value", 4 * nn, ms,
         "compute-bound on the transcendental pipe at this size, not on HBM")
    # window gather (A1) at bench rows from a resident 1e7-step series
    T = 10 ** 7
    cfg = ar_config(p=16384, T=T)
    e2 = NMAEngine(cfg, dev)
    obs = torch.randn(T + 160, device=dev)
    arrays = [obs, torch.zeros(T + 151, device=dev), torch.arange(T + 152, device=dev, dtype=torch.float32),
              torch.ones(T + 151, device=dev), torch.ones(T + 151, device=dev)]
    e2.set_series(arrays)
    idx = torch.from_numpy(np.random.RandomState(1).choice(np.arange(0, T, 50), 16384, replace=False).astype(np.int64)).to(dev)
    ms = timeit(lambda: e2.gather(idx))
    line("k_gather", "A1 window gather as a materialised feed [16384, 201, 14] + mask/shift (AR.py:267-288): 4 B out per value "
         "(the step itself never materialises it: the gather is fused into the feature kernel)", 16384 * 201 * 14 * 4, ms,
         "includes torch.empty of the outputs")
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r02_streaming.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
