"""dev tool: per-role clock64 timeline of CTA 0 of the last pconv_fwd_kernel<32> launch (conv layer 0).  Needs the
-DARL_TRACE build:  ARL_LIB_PATH=/root/repo/accel_rl_b200/csrc/libaccelrl_b200_trace.so python tests/trace_persist.py"""
import ctypes as C
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util_gpu import make_policy

pol, flat, spec = make_policy(1, max_rows=512)
eng = pol.engine
obs = torch.randint(0, 256, (512, 4, 104, 80), dtype=torch.uint8, device="cuda")
prob = torch.zeros(512, 4, device="cuda"); val = torch.zeros(512, device="cuda")
for _ in range(3):
    eng.forward(obs, prob=prob, value=val)
torch.cuda.synchronize()
buf = (C.c_longlong * 4096)()
eng.lib.arl_trace_read.argtypes = [C.c_void_p, C.c_int]
eng.lib.arl_trace_read(buf, 4096)
t = np.array(buf[:4092], dtype=np.int64).reshape(-1, 6)
t0 = t[0, 0]
print("conv0 fwd, CTA 0: cycles relative to the first producer event")
print(" it | prod:empty-ok  prod:issued | mma:full-ok  mma:committed | epi:tfull-ok  epi:stored")
for it in range(16):
    r = t[it] - t0
    print("%3d | %10d %10d | %10d %10d | %10d %10d" % (it, r[0], r[1], r[2], r[3], r[4], r[5]))
