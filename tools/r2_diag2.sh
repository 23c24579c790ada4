#!/bin/bash
python tools/diag_fast.py gpurun_out/diag_a.npy nomesh
python tools/diag_fast.py gpurun_out/diag_b.npy nomesh
PTB_PRODUCTION_MATH=0 python tools/diag_fast.py gpurun_out/diag_c.npy nomesh
PTB_PRODUCTION_MATH=0 python tools/diag_fast.py gpurun_out/diag_d.npy nomesh
python - <<'PY'
import numpy as np
a,b,c,d = (np.load(f"gpurun_out/diag_{k}.npy") for k in "abcd")
print("fast vs fast identical:", (a == b).all(axis=1).mean())
print("exact vs exact identical:", (c == d).all(axis=1).mean())
la, lc = a[:, :3].sum(1), c[:, :3].sum(1)
rel = np.abs(la - lc) / (1e-6 + np.abs(lc))
print("fast vs exact-TU rel diff quantiles", np.quantile(rel, [0.1, 0.5, 0.9, 0.99]))
PY
rm -f gpurun_out/diag_*.npy
