#!/usr/bin/env python
"""Dev check (run by hand, not collected by pytest): patch-resident conv forward (ARL_PCONV>=1) against the oracle, with per-kernel CUDA-event times."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.util_gpu import make_policy, relerr, t2n
from oracle import net as onet

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    pol, flat, spec = make_policy(1, max_rows=512)
    eng = pol.engine
    rng = np.random.RandomState(0)
    obs = rng.randint(0, 256, (n, 4, 104, 80), dtype=np.uint8)
    d_obs = torch.tensor(obs).cuda()
    prob = torch.zeros(n, 4, device="cuda"); val = torch.zeros(n, device="cuda")
    eng.forward(d_obs, prob=prob, value=val)
    torch.cuda.synchronize()
    print("device_error", eng.device_error())
    m = min(n, 64)
    p_ref, v_ref = onet.forward(torch.tensor(flat), torch.tensor(obs[:m]), spec, 4, True)
    print("PCONV=%s n=%d prob relerr %.3e value relerr %.3e" % (os.environ.get("ARL_PCONV"), n, relerr(t2n(prob)[:m], p_ref.numpy()),
                                                       relerr(t2n(val)[:m], v_ref.numpy())))
    # last rows too (tile tails)
    p_ref2, v_ref2 = onet.forward(torch.tensor(flat), torch.tensor(obs[-8:]), spec, 4, True)
    print("  tail rows: prob relerr %.3e value relerr %.3e" % (relerr(t2n(prob)[-8:], p_ref2.numpy()), relerr(t2n(val)[-8:], v_ref2.numpy())))
    labels, ms = eng.profile_graph(2, None, n, reps=50)
    print({l: round(1e3 * float(t), 1) for l, t in zip(labels, ms)}, "us (back-to-back relaunch, warm)")

if __name__ == "__main__":
    main()
