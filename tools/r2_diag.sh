#!/bin/bash
cp cpupathtrace_b200/lib/libptb.so /tmp/libptb_base.so
python tools/diag_fast.py gpurun_out/diag_fast.npy nomesh
cp variants/libptb_macro_exactflags.so cpupathtrace_b200/lib/libptb.so
python tools/diag_fast.py gpurun_out/diag_exact.npy nomesh
cp /tmp/libptb_base.so cpupathtrace_b200/lib/libptb.so
python - <<'PY'
import numpy as np
a = np.load("gpurun_out/diag_exact.npy"); b = np.load("gpurun_out/diag_fast.npy")
la, lb = a[:, :3].sum(1), b[:, :3].sum(1)
d = lb - la
print("samples", len(a), "identical", (d == 0).mean(), "|d|>1e-3 rel", (np.abs(d) > 1e-3 * (1e-6 + np.abs(la))).mean())
print("sum exact", la.sum(), "sum fast", lb.sum(), "sum of positive d", d[d > 0].sum(), "negative", d[d < 0].sum())
rel = np.abs(d) / (1e-6 + np.abs(la))
print("rel diff quantiles", np.quantile(rel, [0.5, 0.9, 0.99, 0.999]))
small = rel < 1e-2
print("among near-equal samples: mean signed rel diff", (d[small] / (1e-6 + la[small])).mean(), "count", small.sum())
print("among diverged samples: mean exact", la[~small].mean(), "mean fast", lb[~small].mean(), "count", (~small).sum())
order = np.argsort(-np.abs(d))[:15]
for i in order: print(i, a[i], b[i])
big = np.abs(d) > 0.01
print("fraction with |d|>0.01:", big.mean(), "mean exact there", la[big].mean(), "fast", lb[big].mean())
PY
rm -f gpurun_out/diag_fast.npy gpurun_out/diag_exact.npy
