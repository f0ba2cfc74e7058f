"""Times single kernels of the bf16-split step under the library's tuning / diagnostic environment knobs
(NMA_WB_FLUSH, NMA_WB_WAVES, NMA_DIAG).  NMA_DIAG runs skip loads: their RESULTS are invalid, only the time is read."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from viforssms_b200.trainer import ARStepper  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
rows = int(args[0]) if args else 16384
quick = "--quick" in sys.argv          # default settings only (e.g. under NMA_DGRAD_WIDE=0/1 set by the caller)
st = ARStepper(T=10 ** 6, rows=rows, device=torch.device("cuda", 0), tensor_cores=7)
for _ in range(2):
    st.step_resident()
torch.cuda.synchronize()
NAMES = {0: "conv_fwd", 1: "conv_dgrad", 2: "conv_wgrad"}
CASES = [
    ("default", {}, (0, 1, 2)),
    ("wgrad flush=4", {"NMA_WB_FLUSH": "4"}, (2,)),
    ("DIAG wgrad no loads, no drain", {"NMA_DIAG": "3", "NMA_WB_FLUSH": "100000"}, (2,)),
    ("DIAG wgrad no loads, no drain, correction MMA as N=128", {"NMA_DIAG": "11", "NMA_WB_FLUSH": "100000"}, (2,)),
    ("DIAG wgrad no loads, no drain, no correction MMA", {"NMA_DIAG": "19", "NMA_WB_FLUSH": "100000"}, (2,)),
    ("DIAG wgrad with loads, no drain, no correction MMA", {"NMA_DIAG": "16", "NMA_WB_FLUSH": "100000"}, (2,)),
]
if quick:
    CASES = CASES[:1]
for name, env, stages in CASES:
    for k in ("NMA_WB_FLUSH", "NMA_WB_WAVES", "NMA_DIAG"):
        os.environ.pop(k, None)
    os.environ.update(env)
    os.environ["NMA_DIAG_I_KNOW_RESULTS_ARE_INVALID"] = "1" if "NMA_DIAG" in env else "0"
    print(name, " ".join("%s[0]=%.3f ms" % (NAMES[s], st.time_stage(s, 0)) for s in stages), flush=True)
st.close()
