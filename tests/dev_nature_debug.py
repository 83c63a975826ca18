"""Dev check (run by hand): where does the 84x84 pad-0 geometry go to zero on the patch-resident path?"""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from accel_rl_b200 import _lib as L
from oracle import net as onet
from accel_rl_b200.policies import AtariCnnPolicy
from accel_rl_b200.envs.atari_env import EnvSpec
from accel_rl_b200.spaces import Discrete, UintBox
pads = eval(sys.argv[1]) if len(sys.argv) > 1 else [0, 0, 0]
spec = dict(conv_filter_sizes=[8, 4, 3], conv_filters=[32, 64, 64], conv_strides=[4, 2, 1], conv_pads=pads, hidden_sizes=[512])
A, n = 6, 48
flat = onet.init_params(spec, (4, 84, 84), A, np.random.RandomState(0), np.random.RandomState(1))
pol = AtariCnnPolicy(initial_param_values=flat, max_rows=n, conv_filter_sizes=[8, 4, 3], conv_filters=[32, 64, 64],
                     conv_strides=[4, 2, 1], conv_pads=[(p, p) for p in pads], hidden_sizes=[512])
pol.initialize(EnvSpec(UintBox((4, 84, 84)), Discrete(A)))
eng = pol.engine
obs = np.random.RandomState(11).randint(0, 256, (n, 4, 84, 84), dtype=np.uint8)
prob = torch.zeros(n, A, device="cuda"); val = torch.zeros(n, device="cuda")
eng.forward(torch.tensor(obs).cuda(), prob=prob, value=val)
torch.cuda.synchronize()
print("device_error", eng.device_error(), "value[:4]", val[:4].tolist())
for layer in (201, 202, 2):
    buf = torch.zeros(2_000_000, device="cuda"); cnt = C.c_long()
    rc = eng.lib.arl_debug_activation(eng.ctx, layer, L.ptr(buf), buf.numel(), C.byref(cnt), eng._s())
    torch.cuda.synchronize()
    x = buf[:cnt.value]
    print("layer", layer, "rc", rc, "count", cnt.value, "nonzero frac %.4f" % (x != 0).float().mean().item(), "absmax %.3f" % x.abs().max().item())
from tests.util_gpu import relerr, t2n
rng = np.random.RandomState(5)
act = rng.randint(0, A, n).astype(np.uint8)
adv = rng.randn(n).astype(np.float32); ret = rng.randn(n).astype(np.float32)
oldp = rng.dirichlet(np.ones(A), n).astype(np.float32); oldv = rng.randn(n).astype(np.float32)
eng.opt_configure(algo=0, clip_param=0.2, v_loss_coeff=1.0, ent_loss_coeff=0.01, update=0, learning_rate=1e-3,
                  beta1=0.9, beta2=0.999, epsilon=1e-5, rho=0.9, grad_norm_clip=-1.0)
eng.bind_train_inputs(*[torch.tensor(x).cuda() for x in (obs, act, adv, ret, oldv, oldp)], valids=None)
eng.grad_minibatch(torch.arange(n, dtype=torch.int32, device="cuda"), n)
torch.cuda.synchronize()
g = t2n(eng.grad)
_, g_ref, _ = onet.loss_and_grad(flat, obs, act, adv, ret, oldp, spec, A, "ppo", emulate_bf16=True, v_coeff=1.0, valids=None)
i = 0
for k, s in enumerate(onet.param_shapes(spec, (4, 84, 84), A)):
    m = int(np.prod(s))
    print("tensor %2d %-18s relerr %.3e" % (k, s, relerr(g[i:i + m], g_ref[i:i + m])))
    i += m
