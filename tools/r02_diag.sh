#!/bin/bash
# stage time of the TS weight gradient with parts switched off (results invalid): what paces the pipeline?
for d in ${DIAGS:-0 1 2 4 3 5 6 7}; do
  NMA_DIAG_I_KNOW_RESULTS_ARE_INVALID=1 NMA_DIAG=$d timeout 200 python - <<PY
import os, sys, torch
sys.path.insert(0, ".")
from viforssms_b200.trainer import ARStepper
st = ARStepper(T=10**7, rows=16384, device=torch.device("cuda", 0), tensor_cores=7)
st._step(st.idx_dev); torch.cuda.synchronize()
print("diag", os.environ["NMA_DIAG"], "flush", os.environ.get("NMA_WS_FLUSH"), "wgrad flow0 %.3f ms  flow1 %.3f  flow2 %.3f" % (st.time_stage(2, 0), st.time_stage(2, 1), st.time_stage(2, 2)), flush=True)
st.close()
PY
done
