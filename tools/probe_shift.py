#!/usr/bin/env python
"""Dev probe driver: row-shifted UMMA descriptors inside a 128B-swizzled shared-memory region (see probe_shift.cu).
Writes gpurun_out/probe_shift.json."""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = C.CDLL(os.path.join(ROOT, "tools", "libprobe_shift.so"))
P_ = C.c_void_p
lib.probe_fwd.argtypes = [P_, C.c_int, P_, P_, C.c_int, C.c_int, P_, C.c_int, P_]
lib.probe_wgrad.argtypes = [P_, C.c_int, P_, C.c_int, C.c_int, P_, C.c_int, P_]


def ptr(t):
    return C.c_void_p(t.data_ptr())


def main():
    torch.manual_seed(0)
    out = {}
    P = 160
    apix = torch.randn(P, 64, device="cuda").to(torch.bfloat16)
    for name, shifts in (("k_major_pitch10_3x3", [0, 1, 2, 10, 11, 12, 20, 21, 22]), ("k_major_2x2_pitch20", [0, 1, 20, 21]),
                         ("k_major_single8", [8]), ("k_major_single3", [3])):
        T, N = len(shifts), 64
        B = torch.randn(T, N, 64, device="cuda").to(torch.bfloat16)
        sh = torch.tensor(shifts, dtype=torch.int32, device="cuda")
        ref = torch.zeros(128, N, device="cuda")
        for t, s in enumerate(shifts):
            ref += apix[s:s + 128].float() @ B[t].float().t()
        for mode in (0, 1):
            D = torch.zeros(128, N, device="cuda")
            rc = lib.probe_fwd(ptr(apix), P, ptr(B), ptr(sh), T, N, ptr(D), mode, None)
            torch.cuda.synchronize()
            out["%s/bo%d" % (name, mode)] = {"rc": rc, "max_abs_err": (D - ref).abs().max().item(),
                                            "ref_absmax": ref.abs().max().item()}
    dy = torch.randn(128, 64, device="cuda").to(torch.bfloat16)
    for (s0, s1) in ((0, 8), (0, 1), (3, 13), (10, 11), (5, 5 + 16)):
        ref = torch.cat([apix[s0:s0 + 128].float().t() @ dy.float(), apix[s1:s1 + 128].float().t() @ dy.float()], 0)
        for mode in (0, 1):
            D = torch.zeros(128, 64, device="cuda")
            rc = lib.probe_wgrad(ptr(apix), P, ptr(dy), s0, s1, ptr(D), mode, None)
            torch.cuda.synchronize()
            out["mn_major_s%d_%d/bo%d" % (s0, s1, mode)] = {
                "rc": rc, "err_atom0": (D[:64] - ref[:64]).abs().max().item(),
                "err_atom1": (D[64:] - ref[64:]).abs().max().item(), "ref_absmax": ref.abs().max().item()}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe_shift.json"), "w"), indent=1)
    for k, v in out.items():
        print(k, v)


if __name__ == "__main__":
    main()
