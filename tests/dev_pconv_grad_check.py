#!/usr/bin/env python
"""Dev check (run by hand, not collected by pytest): per-tensor gradient error of one minibatch against the oracle (patch-resident backward)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.util_gpu import make_policy, relerr, t2n
from oracle import net as onet

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    pol, flat, spec = make_policy(1, max_rows=n)
    eng = pol.engine
    rng = np.random.RandomState(5 + n)
    N = n + 40
    obs = rng.randint(0, 256, (N, 4, 104, 80), dtype=np.uint8)
    act = rng.randint(0, 4, N).astype(np.uint8)
    adv = rng.randn(N).astype(np.float32); ret = rng.randn(N).astype(np.float32)
    oldp = rng.dirichlet(np.ones(4), N).astype(np.float32); oldv = rng.randn(N).astype(np.float32)
    idx = rng.permutation(N)[:n].astype(np.int32)
    eng.opt_configure(algo=0, clip_param=0.2, v_loss_coeff=1.0, ent_loss_coeff=0.01, update=0, learning_rate=1e-3,
                      beta1=0.9, beta2=0.999, epsilon=1e-5, rho=0.9, grad_norm_clip=-1.0)
    d = [torch.tensor(x).cuda() for x in (obs, act, adv, ret, oldv, oldp)]
    eng.bind_train_inputs(*d, valids=None)
    didx = torch.tensor(idx).cuda()
    for rep in range(2):
        eng.grad_minibatch(didx, n)
        torch.cuda.synchronize()
    print("device_error", eng.device_error())
    g = t2n(eng.grad)
    _, g_ref, _ = onet.loss_and_grad(flat, obs[idx], act[idx], adv[idx], ret[idx], oldp[idx], spec, 4, "ppo", emulate_bf16=True,
                                     v_coeff=1.0, valids=None)
    i = 0
    for k, s in enumerate(onet.param_shapes(spec, (4, 104, 80), 4)):
        m = int(np.prod(s))
        print("tensor %2d %-18s relerr %.3e  |g| %.3e |ref| %.3e" % (k, s, relerr(g[i:i + m], g_ref[i:i + m]), np.linalg.norm(g[i:i+m]), np.linalg.norm(g_ref[i:i+m])))
        i += m
    print("total relerr %.3e" % relerr(g, g_ref))

if __name__ == "__main__":
    main()
