"""Dev aid: run tests/early_fc_worker.py under several A/B switch settings and report WHERE the results differ."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cfgs = {"old": dict(ARL_STREAM_UPDATE="0", ARL_GRAPH_MB="1"), "stream": dict(ARL_STREAM_UPDATE="1", ARL_GRAPH_MB="1"),
        "graph": dict(ARL_STREAM_UPDATE="0", ARL_GRAPH_MB="512"), "early": dict(ARL_EARLY_FC="1", ARL_GRAPH_MB="1"),
        "new": dict()}
res = {}
for name, over in cfgs.items():
    path = os.path.join(ROOT, "gpurun_out", "ab_%s.npz" % name)
    env = dict(os.environ, ARL_AB_DUMP=path, **over)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "early_fc_worker.py")], env=env, capture_output=True, text=True)
    print(name, p.returncode, [l for l in p.stdout.splitlines() if l.startswith("RESULT")][-1][:200] if p.returncode == 0 else p.stderr[-500:])
    if p.returncode == 0:
        res[name] = dict(np.load(path))
base = res["old"]
for name, r in res.items():
    if name == "old":
        continue
    for k in ("params", "m", "v"):
        d = np.nonzero(base[k] != r[k])[0]
        if d.size:
            lay = base["layout"]
            tens = sorted(set(int(np.searchsorted(lay[:, 0], i, side="right") - 1) for i in d[:100000]))
            print("  %s vs old: %s differs at %d elements, tensors %s, max abs %.3e, first %s" %
                  (name, k, d.size, tens, np.abs(base[k][d] - r[k][d]).max(), d[:5]))
        else:
            print("  %s vs old: %s identical" % (name, k))
