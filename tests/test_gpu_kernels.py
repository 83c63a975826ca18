"""Per-kernel parity: CUDA (through the C ABI) vs the oracle.  Needs a B200."""
import ctypes as C

import os

import numpy as np
import pytest
import torch

from oracle import frame as oframe, gae as ogae, net as onet, sampler as osampler
from tests.util_gpu import make_policy, relerr, t2n

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pol1():
    pol, flat, spec = make_policy(1, max_rows=160)
    yield pol, flat, spec
    pol.engine.close()


def test_frame_kernel_bit_exact(pol1, golden_dir):
    eng = pol1[0].engine
    g = np.load(golden_dir + "/frames.npz")
    stack = torch.tensor(g["stacks"]).cuda()
    eng.frame_update(torch.tensor(g["raw1"]).cuda(), torch.tensor(g["raw2"]).cuda(),
                     torch.tensor(g["reset"].astype(np.uint8)).cuda(), stack)
    assert np.array_equal(t2n(stack), g["out"])            # vs the reference's own cv2 path
    # larger random batch incl. single-plane stacks (num_img_obs=1) vs the oracle
    rng = np.random.RandomState(1)
    for planes, n in ((4, 300), (1, 65)):
        a = rng.randint(0, 256, (n, 210, 160), dtype=np.uint8)
        b = rng.randint(0, 256, (n, 210, 160), dtype=np.uint8)
        st = rng.randint(0, 256, (n, planes, 104, 80), dtype=np.uint8)
        rs = (rng.rand(n) < 0.25).astype(np.uint8)
        d = torch.tensor(st).cuda()
        eng.frame_update(torch.tensor(a).cuda(), torch.tensor(b).cuda(), torch.tensor(rs).cuda(), d)
        assert np.array_equal(t2n(d), oframe.update_obs_batch(st, a, b, rs))
    # raw_a == NULL (after _reset_obs)
    d = torch.tensor(st).cuda()
    eng.frame_update(None, torch.tensor(b).cuda(), None, d)
    assert np.array_equal(t2n(d), oframe.update_obs_batch(st, None, b, None))


def test_frame_rgb_kernel_bit_exact(pol1, golden_dir):
    """north-star mode (RGB -> gray -> 84x84 area resize -> stack + bf16 copy): CUDA == oracle, bit for bit"""
    eng = pol1[0].engine
    g = np.load(golden_dir + "/frames_rgb.npz")
    n = g["raw_a"].shape[1]
    d = torch.zeros(n, 4, 84, 84, dtype=torch.uint8, device="cuda")
    d16 = torch.zeros(n, 4, 84, 84, dtype=torch.bfloat16, device="cuda")
    for s in range(g["raw_a"].shape[0]):
        eng.frame_update_rgb(torch.tensor(g["raw_a"][s]).cuda(), torch.tensor(g["raw_b"][s]).cuda(),
                             torch.tensor(g["reset"][s]).cuda(), d, d16)
        assert np.array_equal(t2n(d), g["stacks"][s])
        assert np.array_equal(t2n(d16.float()), g["stacks"][s].astype(np.float32))
    # random frames, ragged batch, single-frame and 1-plane variants against the oracle itself
    rng = np.random.RandomState(3)
    for n, planes, with_a in ((37, 4, True), (5, 1, True), (9, 4, False)):
        a = rng.randint(0, 256, (n, 210, 160, 3), dtype=np.uint8)
        b = rng.randint(0, 256, (n, 210, 160, 3), dtype=np.uint8)
        st = rng.randint(0, 256, (n, planes, 84, 84), dtype=np.uint8)
        rs = (rng.rand(n) < 0.3).astype(np.uint8)
        d = torch.tensor(st).cuda()
        eng.frame_update_rgb(torch.tensor(a).cuda() if with_a else None, torch.tensor(b).cuda(), torch.tensor(rs).cuda(), d)
        assert np.array_equal(t2n(d), oframe.rgb_update_obs_batch(st, a if with_a else None, b, rs))


@pytest.mark.parametrize("A", [4, 6, 18])
def test_action_sampling_bit_exact(pol1, golden_dir, A):
    eng = pol1[0].engine
    g = np.load(golden_dir + "/sampling.npz")
    act = torch.zeros(257, dtype=torch.uint8, device="cuda")
    eng.sample_actions(torch.tensor(g["p%d" % A]).cuda(), torch.tensor(g["u%d" % A]).cuda(), act)
    assert np.array_equal(t2n(act), g["a%d" % A])          # vs rllab weighted_sample_n
    rng = np.random.RandomState(A)
    p = rng.dirichlet(np.ones(A) * 0.3, 20000).astype(np.float32)
    u = rng.rand(20000)
    u[:50] = 1.0 - 1e-17                                   # beyond the last cumsum -> clamp to A-1
    act = torch.zeros(20000, dtype=torch.uint8, device="cuda")
    eng.sample_actions(torch.tensor(p).cuda(), torch.tensor(u).cuda(), act)
    assert np.array_equal(t2n(act), osampler.weighted_sample_n(p, u, A))


@pytest.mark.parametrize("lam,use_valids,std", [(0.95, False, False), (1.0, False, False), (0.95, True, False),
                                                (1.0, True, True), (0.95, False, True)])
@pytest.mark.parametrize("B,T", [(7, 33), (256, 128), (3, 5), (64, 1)])
def test_gae_kernel(pol1, lam, use_valids, std, B, T):
    eng = pol1[0].engine
    rng = np.random.RandomState(B * 1000 + T)
    N = B * T
    r = rng.choice([0., 0., 1., -1.], N).astype(np.float32)
    v = rng.randn(N).astype(np.float32)
    d = rng.rand(N) < 0.08
    nr = d & (rng.rand(N) < 0.5)
    lv = rng.randn(B).astype(np.float32)
    want = ogae.process_samples(r, v, d, nr, lv, 0.99, lam, T, use_valids=use_valids, standardize_adv=std)
    dv = torch.tensor(v).cuda()
    adv = torch.zeros(N, device="cuda"); ret = torch.zeros(N, device="cuda")
    valids = torch.zeros(N, dtype=torch.int8, device="cuda") if use_valids else None
    eng.gae(torch.tensor(r).cuda(), dv, torch.tensor(d).cuda(), torch.tensor(nr).cuda() if use_valids else None,
            torch.tensor(lv).cuda(), 0.99, lam, adv, ret, valids, B, T, std)
    tol = dict(rtol=1e-4, atol=1e-4)                       # north-star: advantages within 1e-4
    np.testing.assert_allclose(t2n(adv), want[0], **tol)
    np.testing.assert_allclose(t2n(ret), want[1], **tol)
    if use_valids:
        assert np.array_equal(t2n(valids), want[2])
        np.testing.assert_allclose(t2n(dv), want[3], rtol=0, atol=0)


@pytest.mark.parametrize("nmajor", [0, 1])
@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (300, 128, 512), (1000, 192, 576), (1, 64, 128)])
def test_tcgen05_gemm_tile(pol1, M, N, K, nmajor):
    from accel_rl_b200 import _lib as L
    eng = pol1[0].engine
    torch.manual_seed(M + N + K)
    A = torch.randn(M, K).to(torch.bfloat16).cuda()
    B = torch.randn(N, K).to(torch.bfloat16).cuda()
    Bm = B.t().contiguous() if nmajor else B
    D = torch.zeros(M, N, device="cuda")
    eng.check(eng.lib.arl_test_gemm(eng.ctx, L.ptr(A), L.ptr(Bm), L.ptr(D), M, N, K, nmajor, eng._s()))
    ref = A.float() @ B.float().t()
    assert (D - ref).abs().max().item() <= 2e-3 * max(1.0, ref.abs().max().item()) * 1e-1 + 1e-3


@pytest.mark.parametrize("N", [16, 32, 64, 256])
@pytest.mark.parametrize("rows,Kp", [(64, 128), (1000, 256), (777, 192)])
def test_tcgen05_wgrad_tile(pol1, rows, Kp, N):
    from accel_rl_b200 import _lib as L
    eng = pol1[0].engine
    torch.manual_seed(rows + Kp + N)
    A = torch.randn(rows, Kp).to(torch.bfloat16).cuda()
    B = torch.randn(rows, N).to(torch.bfloat16).cuda()
    D = torch.zeros(Kp, N, device="cuda")
    eng.check(eng.lib.arl_test_wgrad(eng.ctx, L.ptr(A), L.ptr(B), L.ptr(D), rows, Kp, N, eng._s()))
    ref = A.float().t() @ B.float()
    assert (D - ref).abs().max().item() <= 1e-3 + 1e-4 * ref.abs().max().item()


@pytest.mark.parametrize("spec_id,n,A", [(1, 37, 4), (0, 37, 4), (1, 128, 6), (1, 1, 18), (0, 130, 9)])
def test_policy_forward_vs_oracle(spec_id, n, A):
    pol, flat, spec = make_policy(spec_id, max_rows=n, n_actions=A)
    try:
        obs = np.random.RandomState(n).randint(0, 256, (n, 4, 104, 80), dtype=np.uint8)
        info = pol.dist_info_value(obs)
        p_ref, v_ref = onet.forward(torch.tensor(flat), torch.tensor(obs), spec, A, emulate_bf16=True)
        # bf16-operand mirror: differences are fp32 summation order + rare 1-ulp bf16 rounding flips
        np.testing.assert_allclose(info["prob"], p_ref.numpy(), rtol=2e-3, atol=2e-5)
        np.testing.assert_allclose(info["value"], v_ref.numpy(), rtol=2e-3, atol=2e-3)
        assert abs(info["prob"].sum(axis=1) - 1).max() < 1e-5
        # against the un-rounded fp32 graph: documented looser bound for bf16 operands
        p32, v32 = onet.forward(torch.tensor(flat), torch.tensor(obs), spec, A, emulate_bf16=False)
        np.testing.assert_allclose(info["prob"], p32.numpy(), rtol=3e-2, atol=1e-3)
        np.testing.assert_allclose(info["value"], v32.numpy(), rtol=3e-2, atol=2e-2)
    finally:
        pol.engine.close()


@pytest.mark.parametrize("spec_id,algo,n,valids", [(1, "ppo", 64, False), (0, "a2c", 48, False), (1, "a2c", 33, False),
                                                   (1, "ppo", 160, False), (1, "ppo", 64, True), (1, "a2c", 48, True),
                                                   (0, "a2c", 40, True), (1, "ppo", 512, False), (1, "ppo", 512, True)])
def test_loss_and_gradient_vs_oracle(spec_id, algo, n, valids):
    """losses + flat gradient of one minibatch (up to the full PPO minibatch of 512, with and without the validity mask of
    algos/pg/util.py:49-53) against (a) the oracle graph with the tensor-core operands rounded to bf16 where the CUDA path
    rounds them, (b) the plain fp32 oracle graph (the reference's arithmetic): bounds written below"""
    pol, flat, spec = make_policy(spec_id, max_rows=n)
    eng = pol.engine
    try:
        rng = np.random.RandomState(5 + n)
        N = n + 40
        obs = rng.randint(0, 256, (N, 4, 104, 80), dtype=np.uint8)
        act = rng.randint(0, 4, N).astype(np.uint8)
        adv = rng.randn(N).astype(np.float32)
        ret = rng.randn(N).astype(np.float32)
        oldp = rng.dirichlet(np.ones(4), N).astype(np.float32)
        oldv = rng.randn(N).astype(np.float32)
        val = (rng.rand(N) < 0.7).astype(np.int8) if valids else None
        idx = rng.permutation(N)[:n].astype(np.int32)
        vc = 1.0 if algo == "ppo" else 0.25
        eng.opt_configure(algo=0 if algo == "ppo" else 1, clip_param=0.2, v_loss_coeff=vc, ent_loss_coeff=0.01, update=0,
                          learning_rate=1e-3, beta1=0.9, beta2=0.999, epsilon=1e-5, rho=0.9, grad_norm_clip=-1.0)
        d = [torch.tensor(x).cuda() for x in (obs, act, adv, ret, oldv, oldp)]
        dval = torch.tensor(val).cuda() if valids else None
        eng.bind_train_inputs(*d, valids=dval)
        didx = torch.tensor(idx).cuda()
        eng.grad_minibatch(didx, n)
        torch.cuda.synchronize()
        g = t2n(eng.grad)
        vsub = val[idx] if valids else None
        if valids:
            assert 0 < int(vsub.sum()) < n                   # the mask is exercised
        loss_ref, g_ref, _ = onet.loss_and_grad(flat, obs[idx], act[idx], adv[idx], ret[idx], oldp[idx], spec, 4, algo,
                                                emulate_bf16=True, v_coeff=vc, valids=vsub)
        loss32, g32, _ = onet.loss_and_grad(flat, obs[idx], act[idx], adv[idx], ret[idx], oldp[idx], spec, 4, algo,
                                            emulate_bf16=False, v_coeff=vc, valids=vsub)
        shapes = onet.param_shapes(spec, (4, 104, 80), 4)
        i = 0
        errs = []
        for k, s in enumerate(shapes):
            m = int(np.prod(s))
            errs.append(relerr(g[i:i + m], g_ref[i:i + m]))
            i += m
        print("PER-TENSOR spec=%d algo=%s n=%d valids=%s %s" % (spec_id, algo, n, valids, " ".join("%.1e" % e for e in errs)))
        # activation gradients are stored in bf16 between layers: their rounding noise averages out over the samples that
        # carry weight, so the per-tensor relative error falls like 1/sqrt(n_eff) — measured 1.2e-2 at 45 valid rows,
        # < 1e-2 at 64, 4.4e-3 at the PPO minibatch of 512.  Bounds: 0.12/sqrt(n_eff) per tensor, 0.10/sqrt(n_eff) overall
        n_eff = int(vsub.sum()) if valids else n
        for k, s in enumerate(shapes):
            assert errs[k] < 0.12 / np.sqrt(n_eff), "tensor %d %s: %.3e (n_eff %d)" % (k, s, errs[k], n_eff)
        assert relerr(g, g_ref) < 0.10 / np.sqrt(n_eff)
        # loss value (north-star: within 1e-4 relative of the bf16-mirrored graph ... 1e-3 abs floor)
        eng.clip_update(1.0)
        losses, norms = eng.read_logs()
        assert abs(losses[0] - loss_ref) <= 1e-4 * abs(loss_ref) + 2e-4
        assert abs(norms[0] - np.linalg.norm(g_ref)) <= 5e-3 * np.linalg.norm(g_ref)
        # (b) against the UN-rounded fp32 graph = the reference's own arithmetic: what bf16 tensor-core operands cost.
        # Measured on B200 (profiles/r2_parity_fp32.md): loss 1e-6 .. 3e-4 relative, whole-gradient error 0.4 % (n = 512) .. 9 % (n = 33)
        e_loss = abs(losses[0] - loss32) / max(abs(loss32), 1e-6)
        e_grad = relerr(g, g32)
        print("FP32-GRAPH spec=%d algo=%s n=%d valids=%s loss_rel=%.3e grad_rel=%.3e cos=%.6f" %
              (spec_id, algo, n, valids, e_loss, e_grad,
               float(np.dot(g, g32) / (np.linalg.norm(g) * np.linalg.norm(g32)))))
        # the gradient error is bf16 rounding noise of the stored activations / operands: it averages out with the
        # minibatch size (0.4 % at the PPO minibatch of 512, cosine 0.99999; up to 9 % for a 33..48-sample A2C batch)
        assert e_loss <= 5e-4 and e_grad <= (6e-3 if n >= 512 else 0.12)
    finally:
        eng.close()


def test_wide_forward_tiles_option_passes_the_oracle_comparisons():
    """ARL_FWD_WIDE=15 (off by default: measured slower, csrc/api.cu pconv_make_wide): the taps of a filter row on the N
    axis of the forward / data-gradient tiles, column groups added one row apart in the epilogue.  The forward and
    loss / gradient comparisons with the oracle must pass unchanged with it."""
    import subprocess
    import sys
    if os.environ.get("ARL_FWD_WIDE"):
        pytest.skip("already running under the option")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_kernels.py"), "-q", "-x", "-k",
                        "forward or loss_and_gradient"], env=dict(os.environ, ARL_FWD_WIDE="15"), capture_output=True, text=True,
                       timeout=900, cwd=root)
    assert r.returncode == 0 and " passed" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.parametrize("tie", [1, 2])
def test_ppo_tie_gradient_option(tie):
    """Inside the clip range the two PPO surrogates (ppo.py:47-49) are exactly equal; what T.minimum's gradient does at a
    tie depends on the Theano version: once (>= 0.9; the default here, standard PPO) or to both branches, i.e. twice the
    policy gradient there (older releases) — `ppo_tie_grad` / PPO(tie_grad=2).  Half of the rows sit inside the range (old
    probabilities = the current policy's), the others outside; the loss value is the same either way."""
    n = 96
    pol, flat, spec = make_policy(1, max_rows=n)
    eng = pol.engine
    try:
        rng = np.random.RandomState(77)
        obs = rng.randint(0, 256, (n, 4, 104, 80), dtype=np.uint8)
        act = rng.randint(0, 4, n).astype(np.uint8)
        adv = rng.randn(n).astype(np.float32)
        ret = rng.randn(n).astype(np.float32)
        oldp = rng.dirichlet(np.ones(4), n).astype(np.float32)
        p_now, _ = onet.forward(torch.tensor(flat), torch.tensor(obs), spec, 4, emulate_bf16=True)
        oldp[: n // 2] = p_now.detach().numpy()[: n // 2]
        oldv = rng.randn(n).astype(np.float32)
        grads, losses = {}, {}
        for t in (1, tie):
            eng.opt_configure(algo=0, clip_param=0.2, v_loss_coeff=1.0, ent_loss_coeff=0.01, update=0, learning_rate=1e-3,
                              beta1=0.9, beta2=0.999, epsilon=1e-5, rho=0.9, grad_norm_clip=-1.0, ppo_tie_grad=t)
            eng.bind_train_inputs(*[torch.tensor(x).cuda() for x in (obs, act, adv, ret, oldv, oldp)], valids=None)
            eng.grad_minibatch(torch.arange(n, dtype=torch.int32, device="cuda"), n)
            torch.cuda.synchronize()
            grads[t] = t2n(eng.grad).copy()
            eng.clip_update(1.0)
            losses[t] = eng.read_logs()[0][0]
            eng.set_params(flat)
            eng.reset_opt_state()
        loss_ref, g_ref, _ = onet.loss_and_grad(flat, obs, act, adv, ret, oldp, spec, 4, "ppo", emulate_bf16=True, tie_grad=tie)
        assert relerr(grads[tie], g_ref) < 0.10 / np.sqrt(n)
        assert abs(losses[tie] - loss_ref) <= 1e-4 * abs(loss_ref) + 2e-4
        if tie == 2:
            assert losses[2] == losses[1]                      # same value ...
            assert relerr(grads[2], grads[1]) > 0.05           # ... different gradient
    finally:
        eng.close()


@pytest.mark.parametrize("kind,clip", [("adam", None), ("adam", 0.5), ("rmsprop", 0.5), ("rmsprop", None)])
def test_clip_and_update_rules(kind, clip):
    pol, flat, spec = make_policy(0, max_rows=8)
    eng = pol.engine
    try:
        rng = np.random.RandomState(3)
        n = flat.size
        eng.opt_configure(algo=0, clip_param=0.2, v_loss_coeff=1.0, ent_loss_coeff=0.01, update=0 if kind == "adam" else 1,
                          learning_rate=1e-3 if kind == "adam" else 7e-4, beta1=0.9, beta2=0.999,
                          epsilon=1e-5 if kind == "adam" else 1e-6, rho=0.9, grad_norm_clip=clip if clip else -1.0)
        eng.reset_opt_state()
        opt = onet.Adam(n, 1e-3, epsilon=1e-5) if kind == "adam" else onet.RMSProp(n, 7e-4)
        p = flat.copy()
        for step in range(3):
            g = (rng.randn(n) * 0.01).astype(np.float32)
            eng.grad.copy_(torch.tensor(g))
            eng.set_lr_mult(1.0 - 0.25 * step)
            eng.clip_update(1.0)
            gc, norm = onet.total_norm_clip(g, clip)
            p = opt.step(p, gc, 1.0 - 0.25 * step)
            losses, norms = eng.read_logs()
            assert abs(norms[0] - norm) <= 1e-5 * norm
        got = eng.get_params()
        assert relerr(got - flat, p - flat) < 1e-4          # the applied update
        np.testing.assert_allclose(got, p, rtol=1e-6, atol=1e-7)
    finally:
        eng.close()


@pytest.mark.parametrize("kind", ["adam", "rmsprop"])
def test_update_kernel_vs_executed_reference_rules(kind, golden_dir):
    """CUDA update kernel against outputs of the reference's OWN update-rule code (update_methods_stats.py:11-32, :55-87
    executed under oracle/theano_shim.py -> tests/golden/update_rules.npz): six steps with a changing lr_mult and
    exact-zero gradients, the golden vector placed at the front of the flat parameter vector."""
    g = np.load(golden_dir + "/update_rules.npz")
    pol, flat, spec = make_policy(0, max_rows=8)
    eng = pol.engine
    try:
        k = g["p0"].size
        eng.opt_configure(algo=0, clip_param=0.2, v_loss_coeff=1.0, ent_loss_coeff=0.01, update=0 if kind == "adam" else 1,
                          learning_rate=1e-3 if kind == "adam" else 7e-4, beta1=0.9, beta2=0.999,
                          epsilon=1e-5 if kind == "adam" else 1e-6, rho=0.9, grad_norm_clip=-1.0)
        eng.reset_opt_state()
        p0 = flat.copy()
        p0[:k] = g["p0"]
        eng.set_params(p0)
        for t in range(len(g["grads"])):
            gr = np.zeros_like(flat)
            gr[:k] = g["grads"][t]
            eng.grad.copy_(torch.tensor(gr))
            eng.set_lr_mult(float(g["lr_mults"][t]))
            eng.clip_update(1.0)
            got = eng.get_params()
            np.testing.assert_allclose(got[:k], g[kind][t], rtol=1e-6, atol=5e-8)
            assert np.array_equal(got[k:], p0[k:])            # zero gradient, zero state: untouched
        eng.read_logs()
    finally:
        eng.close()


@pytest.mark.parametrize("kind", ["adam", "rmsprop"])
def test_optimizer_state_snapshot_resume_is_bit_exact(kind):
    """snapshot / resume of the optimizer state: parameters + (m, v, update count) moved into a
    fresh engine continue the update sequence bit for bit, and match the oracle's Adam / RMSProp after 5 steps"""
    rng = np.random.RandomState(9)
    cfg = dict(algo=0, clip_param=0.2, v_loss_coeff=1.0, ent_loss_coeff=0.01, update=0 if kind == "adam" else 1,
               learning_rate=1e-3 if kind == "adam" else 7e-4, beta1=0.9, beta2=0.999,
               epsilon=1e-5 if kind == "adam" else 1e-6, rho=0.9, grad_norm_clip=0.5)
    pol, flat, spec = make_policy(0, max_rows=8)
    n = flat.size
    grads = [(rng.randn(n) * 0.01).astype(np.float32) for _ in range(5)]

    def run(eng, gs):
        for g in gs:
            eng.grad.copy_(torch.tensor(g))
            eng.clip_update(1.0)
        eng.read_logs()
        return eng.get_params()
    eng = pol.engine
    try:
        eng.opt_configure(**cfg)
        eng.reset_opt_state()
        straight = run(eng, grads)
    finally:
        eng.close()
    pol2, _, _ = make_policy(0, max_rows=8)
    eng2 = pol2.engine
    try:
        eng2.opt_configure(**cfg)
        eng2.reset_opt_state()
        mid = run(eng2, grads[:2])
        state = eng2.get_opt_state()
        assert state["step"] == 2 and state["m"].shape == (n,) and state["v"].shape == (n,)
    finally:
        eng2.close()
    pol3, _, _ = make_policy(0, max_rows=8, seed=5)           # different initial parameters: everything comes from the snapshot
    eng3 = pol3.engine
    try:
        eng3.opt_configure(**cfg)
        eng3.reset_opt_state()
        eng3.set_params(mid)
        eng3.set_opt_state(state)
        resumed = run(eng3, grads[2:])
        assert eng3.get_opt_state()["step"] == 5
    finally:
        eng3.close()
    assert np.array_equal(resumed, straight)
    opt = onet.Adam(n, 1e-3, epsilon=1e-5) if kind == "adam" else onet.RMSProp(n, 7e-4)
    p = flat.copy()
    for g in grads:
        gc, _ = onet.total_norm_clip(g, 0.5)
        p = opt.step(p, gc, 1.0)
    np.testing.assert_allclose(straight, p, rtol=1e-6, atol=1e-7)


def test_nature_cnn_84x84_geometry_vs_oracle():
    """the classic 84x84 Nature-CNN geometry (8x8/4, 4x4/2, 3x3/1, no padding): forward and per-tensor gradients against
    the oracle.  Its 21x21 space-to-depth grid is not a multiple of 8 positions, so this exercises the im2col-gather
    tiles (gemm_tc.cuh) that serve every geometry the patch-resident path does not take."""
    from accel_rl_b200.policies import AtariCnnPolicy
    from accel_rl_b200.envs.atari_env import EnvSpec
    from accel_rl_b200.spaces import Discrete, UintBox
    spec = dict(conv_filter_sizes=[8, 4, 3], conv_filters=[32, 64, 64], conv_strides=[4, 2, 1], conv_pads=[0, 0, 0],
                hidden_sizes=[512])
    A, n = 6, 48
    flat = onet.init_params(spec, (4, 84, 84), A, np.random.RandomState(0), np.random.RandomState(1))
    flat = flat + np.float32(0.01) * np.random.RandomState(2).randn(flat.size).astype(np.float32) * (flat == 0)
    pol = AtariCnnPolicy(initial_param_values=flat, max_rows=n, conv_filter_sizes=[8, 4, 3], conv_filters=[32, 64, 64],
                         conv_strides=[4, 2, 1], conv_pads=[(0, 0), (0, 0), (0, 0)], hidden_sizes=[512])
    pol.initialize(EnvSpec(UintBox((4, 84, 84)), Discrete(A)))
    eng = pol.engine
    try:
        rng = np.random.RandomState(11)
        obs = rng.randint(0, 256, (n, 4, 84, 84), dtype=np.uint8)
        prob = torch.zeros(n, A, device="cuda"); val = torch.zeros(n, device="cuda")
        eng.forward(torch.tensor(obs).cuda(), prob=prob, value=val)
        p_ref, v_ref = onet.forward(torch.tensor(flat), torch.tensor(obs), spec, A, True)
        assert relerr(t2n(prob), p_ref.numpy()) < 2e-3 and relerr(t2n(val), v_ref.numpy()) < 5e-3
        act = rng.randint(0, A, n).astype(np.uint8)
        adv = rng.randn(n).astype(np.float32); ret = rng.randn(n).astype(np.float32)
        oldp = rng.dirichlet(np.ones(A), n).astype(np.float32); oldv = rng.randn(n).astype(np.float32)
        eng.opt_configure(algo=0, clip_param=0.2, v_loss_coeff=1.0, ent_loss_coeff=0.01, update=0, learning_rate=1e-3,
                          beta1=0.9, beta2=0.999, epsilon=1e-5, rho=0.9, grad_norm_clip=-1.0)
        eng.bind_train_inputs(*[torch.tensor(x).cuda() for x in (obs, act, adv, ret, oldv, oldp)], valids=None)
        eng.grad_minibatch(torch.arange(n, dtype=torch.int32, device="cuda"), n)
        torch.cuda.synchronize()
        g = t2n(eng.grad)
        _, g_ref, _ = onet.loss_and_grad(flat, obs, act, adv, ret, oldp, spec, A, "ppo", emulate_bf16=True, v_coeff=1.0,
                                         valids=None)
        i = 0
        for k, s in enumerate(onet.param_shapes(spec, (4, 84, 84), A)):
            m = int(np.prod(s))
            assert relerr(g[i:i + m], g_ref[i:i + m]) < 1e-2, "tensor %d %s" % (k, s)
            i += m
        assert eng.device_error() == 0
    finally:
        eng.close()
