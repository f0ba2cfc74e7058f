"""A few plain (eager, one launch per kernel) training steps at a fixed size; the LAST step is bracketed by
cudaProfilerStart/Stop so `ncu --profile-from-start off` sees exactly one step (tools/r02_evidence.sh)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from viforssms_b200.trainer import ARStepper  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=2048)
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--T", type=int, default=10 ** 6)
ap.add_argument("--tc", type=int, default=3, help="bit set of nma_set_tensor_cores (7 = bf16 split of the conv GEMMs)")
a = ap.parse_args()
st = ARStepper(T=a.T, rows=a.rows, device=torch.device("cuda", 0), tensor_cores=a.tc)
for _ in range(a.steps - 1):
    st.step_resident()
torch.cuda.synchronize()
torch.cuda.profiler.start()
st.step_resident()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", float(st.last_elbo.item()))
st.close()
