#!/bin/bash
cp cpupathtrace_b200/lib/libptb.so /tmp/libptb_base.so
for name in macro_exactflags fmad_only ftz_only div_only sqrt_only ufm_only; do
  cp variants/libptb_$name.so cpupathtrace_b200/lib/libptb.so
  python tools/diag_fast.py gpurun_out/diag_$name.npy nomesh | sed -E 's/gpurun_out.diag_//'
done
cp /tmp/libptb_base.so cpupathtrace_b200/lib/libptb.so
python - <<'PY'
import numpy as np
base = np.load("gpurun_out/diag_macro_exactflags.npy")[:, :3].sum(1)
for name in ("fmad_only", "ftz_only", "div_only", "sqrt_only", "ufm_only"):
    x = np.load(f"gpurun_out/diag_{name}.npy")[:, :3].sum(1)
    rel = np.abs(x - base) / (1e-6 + np.abs(base))
    print(name, "identical", (x == base).mean(), "rel diff quantiles 50/90/99:", np.quantile(rel, [0.5, 0.9, 0.99]), "sum ratio", x.sum() / base.sum())
PY
rm -f gpurun_out/diag_*.npy
