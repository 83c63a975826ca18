"""Path-level parity on the GPU: the device sampler vs the REAL reference sampler's golden buffers,
one full optimize_policy vs the oracle learner, and the runner end to end.  Needs a B200."""
import os
import zlib

import numpy as np
import pytest
import torch

from oracle import net as onet, sampler as osampler, learner as olearner, synth_ale
from tests.golden.make_golden import RULES, POOL_FRAMES
from tests.util_gpu import make_policy, relerr, t2n

pytestmark = pytest.mark.gpu


def _make_sampler(n_parallel, envs_per, T, mid_batch_reset=True, max_path_length=27000, rules=None, env_kw=None, **kw):
    from accel_rl_b200.sampler import ActsrvAltOvrlpSampler
    from accel_rl_b200.envs import AtariEnv
    r = dict(RULES if rules is None else rules)
    r.setdefault("pool_seed", 0)
    env_args = dict(game="breakout", max_start_noops=0, synth_rules=r)
    if env_kw:
        env_args.update(env_kw)
    return ActsrvAltOvrlpSampler(EnvCls=AtariEnv, env_args=env_args, horizon=T, n_parallel=n_parallel, envs_per=envs_per,
                                 max_path_length=max_path_length, mid_batch_reset=mid_batch_reset,
                                 max_decorrelation_steps=kw.pop("max_decorrelation_steps", 0), **kw)


def _buf_np(buf):
    return dict(observations=t2n(buf.observations), rewards=t2n(buf.rewards), dones=t2n(buf.dones),
                raw_reward=t2n(buf.env_infos.raw_reward), need_reset=t2n(buf.env_infos.need_reset),
                actions=t2n(buf.actions), prob=t2n(buf.agent_infos.prob), value=t2n(buf.agent_infos.value),
                extra_observations=t2n(buf.extra_observations))


@pytest.mark.parametrize("tag,mbr", [("reset", True), ("nonreset", False), ("overlength", True)])
def test_device_sampler_reproduces_reference_sampler_buffers(golden_dir, tag, mbr):
    """Everything the policy does not influence must be bit-identical to what the reference's real
    multi-process sampler produced (tests/golden/sampler_*.npz); actions must equal weighted_sample_n of
    the device probabilities with the master's uniforms; the consumed uniforms must be the reference
    master's stream."""
    from accel_rl_b200.util.seeding import set_seed
    g = np.load(os.path.join(golden_dir, "sampler_%s.npz" % tag))
    B, T, itrs = int(g["n_envs"]), int(g["horizon"]), int(g["itrs"])
    set_seed(3)                                                        # the fixture's master seed
    sampler = _make_sampler(2, 2, T, mid_batch_reset=mbr, max_path_length=int(g["max_path_length"]))
    env_spec, sample_size, horizon, _ = sampler.initialize(seed=4, affinities=dict(), discount=0.99, need_extra_obs=True)
    assert (sample_size, horizon) == (B * T, T)
    state = np.random.get_state()
    pol, flat, spec = make_policy(1, max_rows=B)                        # fixed params: no RNG consumed
    np.random.set_state(state)
    sampler.policy_init(pol)
    try:
        for itr in range(itrs):
            buf, infos = sampler.obtain_samples(itr)
            b = _buf_np(buf)
            assert np.array_equal(sampler._uniforms_host.numpy(), g["uniforms"][itr]), "master RNG consumption differs"
            if itr == 0:
                assert np.array_equal(b["observations"], g["obs_0"])
            crc = np.array([zlib.crc32(r.tobytes()) for r in b["observations"]], dtype=np.uint32)
            assert np.array_equal(crc, g["obscrc_%d" % itr]), "observation rows differ at itr %d" % itr
            assert np.array_equal(b["extra_observations"], g["extra_%d" % itr])
            assert np.array_equal(b["rewards"], g["rew_%d" % itr])
            assert np.array_equal(b["dones"], g["done_%d" % itr])
            assert np.array_equal(b["raw_reward"], g["raw_%d" % itr])
            assert np.array_equal(b["need_reset"], g["nr_%d" % itr])
            assert b["actions"].dtype == np.uint8 and b["dones"].dtype == np.bool_
            # actions: bit-exact given the device probabilities and the master's uniforms
            u = g["uniforms"][itr]                                     # (T, B)
            prob = b["prob"].reshape(B, T, -1)
            want = np.stack([osampler.weighted_sample_n(prob[:, s], u[s], 4) for s in range(T)], axis=1).reshape(-1)
            stepped = np.ones(B * T, bool)
            assert np.array_equal(b["actions"][stepped], want[stepped])
            # prob / value of every row: the oracle net on that row's observation
            p_ref, v_ref = onet.forward(torch.tensor(flat), torch.tensor(b["observations"]), spec, 4, emulate_bf16=True)
            if mbr:   # (non-reset mode leaves stale observation rows after an env finished)
                np.testing.assert_allclose(b["prob"], p_ref.numpy(), rtol=2e-3, atol=2e-5)
                np.testing.assert_allclose(b["value"], v_ref.numpy(), rtol=2e-3, atol=2e-3)
            ti = sorted([(i.Length, float(i.Return), float(i.RawReturn), int(i.NonzeroRewards), float(i.DiscountedReturn))
                         for i in infos])
            want_ti = g["traj_%d" % itr]
            assert len(ti) == len(want_ti)
            if len(ti):
                np.testing.assert_allclose(np.array(ti, dtype=np.float64), want_ti, rtol=1e-5)
        # samples_buf contract: struct with attribute + key access, per-env segment views (row = env*T + t)
        assert buf.segs_view[1]["rewards"].data_ptr() == buf.rewards[T:2 * T].data_ptr()
        assert buf["agent_infos"]["prob"].shape == (B * T, 4) and buf.segs_view[0].agent_infos["value"].shape == (T,)
    finally:
        pol.engine.close()


def test_device_sampler_vs_oracle_larger_batch():
    """B=64, T=16, default emulator rules, several iterations: device buffers == oracle restatement."""
    from accel_rl_b200.util.seeding import set_seed
    rules = dict(synth_ale.DEFAULT_RULES, pool_frames=256, life_base=40, life_mod=17, reward_mod=23, pool_seed=0)
    set_seed(1)
    B, T = 64, 16
    sampler = _make_sampler(8, 4, T, rules=rules)
    sampler.initialize(seed=2, affinities=dict(), discount=0.99, need_extra_obs=True)
    pol, flat, spec = make_policy(1, max_rows=B)
    sampler.policy_init(pol)
    pool = synth_ale.make_pool(256, seed=0)
    orc = osampler.OracleSampler(B, T, pool, {k: v for k, v in rules.items() if k != "pool_seed"}, 4, 0.99)
    try:
        n_traj = 0
        for itr in range(5):
            buf, infos = sampler.obtain_samples(itr)
            b = _buf_np(buf)
            u = sampler._uniforms_host.numpy().copy()
            gp = b["prob"].reshape(B, T, 4); gv = b["value"].reshape(B, T)
            calls = {"k": 0}

            def policy_fn(obs):   # feed the oracle sampler the device's probabilities (same numerics)
                k = calls["k"]; calls["k"] += 1
                s, j = divmod(k, 2)
                lo, hi = j * B // 2, (j + 1) * B // 2
                return gp[lo:hi, s], gv[lo:hi, s]
            ob, oinf = orc.obtain_samples(policy_fn, u)
            for k in ("observations", "extra_observations", "rewards", "dones", "raw_reward", "need_reset", "actions"):
                assert np.array_equal(b[k], ob[k]), (k, itr)
            assert len(infos) == len(oinf)
            n_traj += len(infos)
        assert n_traj > 0
    finally:
        pol.engine.close()



@pytest.mark.parametrize("mbr", [True, False])
def test_game_mix_sampler_vs_oracle(mbr):
    """BASELINE configs[2]: the 4-game mix (env e plays game e % 4: own frame-pool slice, reward table, life clock; action
    count padded to the largest minimal set, 9).  Device buffers == the oracle's restatement, bit for bit."""
    from accel_rl_b200.util.seeding import set_seed
    rules = dict(synth_ale.DEFAULT_RULES, pool_frames=256, life_base=20, life_mod=17, reward_mod=11, pool_seed=0)
    set_seed(1)
    B, T, A = 32, 12, 9
    sampler = _make_sampler(4, 4, T, rules=rules, mid_batch_reset=mbr, env_kw=dict(game="mix4"))
    env_spec, _, _, _ = sampler.initialize(seed=2, affinities=dict(), discount=0.99, need_extra_obs=True)
    assert env_spec.action_space.n == A
    pol, flat, spec = make_policy(1, max_rows=B, n_actions=A)
    sampler.policy_init(pol)
    orules = dict({k: v for k, v in rules.items() if k != "pool_seed"}, n_games=4)
    orc = osampler.OracleSampler(B, T, synth_ale.make_pool(256, seed=0), orules, A, 0.99, mid_batch_reset=mbr)
    try:
        n_traj, rew_by_game = 0, np.zeros(4)
        for itr in range(6):
            buf, infos = sampler.obtain_samples(itr)
            b = _buf_np(buf)
            u = sampler._uniforms_host.numpy().copy()
            gp = b["prob"].reshape(B, T, A); gv = b["value"].reshape(B, T)
            calls = {"k": 0}

            def policy_fn(obs):
                k = calls["k"]; calls["k"] += 1
                s, j = divmod(k, 2)
                lo, hi = j * B // 2, (j + 1) * B // 2
                return gp[lo:hi, s], gv[lo:hi, s]
            ob, oinf = orc.obtain_samples(policy_fn, u)
            for k in ("observations", "extra_observations", "rewards", "dones", "raw_reward", "need_reset", "actions"):
                assert np.array_equal(b[k], ob[k]), (k, itr)
            assert sorted((t.Length, round(float(t.Return), 4)) for t in infos) == \
                sorted((t["Length"], round(float(t["Return"]), 4)) for t in oinf)
            n_traj += len(infos)
            rew_by_game += np.abs(b["rewards"]).reshape(B // 4, 4, T).sum(axis=(0, 2))
            assert b["actions"].max() < A
        assert n_traj > 0
        assert len(set(rew_by_game.tolist())) > 1            # the games' reward tables differ
    finally:
        pol.engine.close()


@pytest.mark.parametrize("mbr,mpl", [(True, 27000), (False, 27000), (True, 11)])
def test_decorrelated_start_vs_oracle(mbr, mpl):
    """start_envs with max_decorrelation_steps > 0 (sampler/util.py:33-55): every env takes its own number of warm-up
    steps before the first rollout (reset whenever its trajectory ends, running TrajInfo carried over).  The device
    warm-up (arl_sampler_warmup) against the oracle's restatement with the same per-env step counts: all later rollout
    buffers and trajectory records bit-exact."""
    from accel_rl_b200.util.seeding import set_seed
    rules = dict(synth_ale.DEFAULT_RULES, pool_frames=128, life_base=9, life_mod=5, reward_mod=7, pool_seed=0)
    set_seed(11)
    B, T = 16, 8
    sampler = _make_sampler(4, 2, T, mid_batch_reset=mbr, max_path_length=mpl, rules=rules, max_decorrelation_steps=60)
    sampler.initialize(seed=5, affinities=dict(), discount=0.99, need_extra_obs=True)
    pol, flat, spec = make_policy(1, max_rows=B)
    sampler.policy_init(pol)
    n_steps = sampler.decorrelation_steps
    assert n_steps.shape == (B,) and n_steps.max() < 60 and len(set(n_steps.tolist())) > B // 2
    orc = osampler.OracleSampler(B, T, synth_ale.make_pool(128, seed=0), {k: v for k, v in rules.items() if k != "pool_seed"},
                                 4, 0.99, mid_batch_reset=mbr, max_path_length=mpl)
    orc.decorrelate(n_steps)
    key = lambda t: (t["env"], t["Length"], round(float(t["Return"]), 4), t["NonzeroRewards"])
    dkey = lambda t: (t.env, t.Length, round(float(t.Return), 4), t.NonzeroRewards)
    try:
        assert np.array_equal(t2n(sampler.step_buf.obs), orc.step_obs)
        n_traj = 0
        for itr in range(3):
            buf, infos = sampler.obtain_samples(itr)
            b = _buf_np(buf)
            u = sampler._uniforms_host.numpy().copy()
            gp = b["prob"].reshape(B, T, 4); gv = b["value"].reshape(B, T)
            calls = {"k": 0}

            def policy_fn(obs):
                k = calls["k"]; calls["k"] += 1
                s, j = divmod(k, 2)
                lo, hi = j * B // 2, (j + 1) * B // 2
                return gp[lo:hi, s], gv[lo:hi, s]
            ob, oinf = orc.obtain_samples(policy_fn, u)
            for k in ("observations", "extra_observations", "rewards", "dones", "raw_reward", "need_reset"):
                assert np.array_equal(b[k], ob[k]), (k, itr)
            assert sorted(dkey(t) for t in infos) == sorted(key(t) for t in oinf)
            n_traj += len(infos)
        assert n_traj > 0
        assert pol.engine.device_error() == 0
    finally:
        pol.engine.close()


def test_full_size_c2_rollout_bit_exact_and_update_properties():
    """BASELINE.json configs[1] at FULL size: 256 envs x 128 steps, preset 1, PPO 4 epochs x 64 minibatches of 512.
    (1) every integer/byte buffer of the 1.09 GB rollout equals the oracle sampler's, bit for bit (sampled actions
    included, given the device's probabilities and the master's uniforms); (2) size-independent properties of the
    learner: the minibatch gradient is linear in its rows (g(512 rows) == mean of the two 256-row halves' gradients,
    both scaled by exact powers of two), 256 updates are taken, every loss / gradient norm is finite, the Adam count
    matches, and an iteration with lr_mult = 0 leaves the parameters untouched (idempotence)."""
    from accel_rl_b200.algos import PPO
    from accel_rl_b200.util.seeding import set_seed
    rules = dict(synth_ale.DEFAULT_RULES, pool_frames=512, life_base=60, life_mod=31, reward_mod=41, pool_seed=0)
    orules = {k: v for k, v in rules.items() if k != "pool_seed"}
    set_seed(11)
    B, T = 256, 128
    sampler = _make_sampler(32, 4, T, rules=rules)
    env_spec, sample_size, horizon, mbr = sampler.initialize(seed=12, affinities=dict(), discount=0.99, need_extra_obs=True)
    assert sample_size == B * T
    pol, flat, spec = make_policy(1)
    algo = PPO(optimizer_args=dict(minibatch_size=512, epochs=4))
    algo.initialize(pol, env_spec, sample_size, horizon, mbr)
    sampler.policy_init(pol)
    algo.set_n_itr(10)
    eng = pol.engine
    orc = osampler.OracleSampler(B, T, synth_ale.make_pool(512, seed=0), orules, 4, 0.99)
    try:
        buf, infos = sampler.obtain_samples(0)
        b = _buf_np(buf)
        u = sampler._uniforms_host.numpy().copy()
        gp = b["prob"].reshape(B, T, 4); gv = b["value"].reshape(B, T)
        calls = {"k": 0}

        def policy_fn(obs):
            k = calls["k"]; calls["k"] += 1
            s, j = divmod(k, 2)
            lo, hi = j * B // 2, (j + 1) * B // 2
            return gp[lo:hi, s], gv[lo:hi, s]
        ob, oinf = orc.obtain_samples(policy_fn, u)
        for k in ("observations", "extra_observations", "rewards", "dones", "raw_reward", "need_reset", "actions"):
            assert np.array_equal(b[k], ob[k]), k
        assert len(infos) == len(oinf) and b["dones"].any()
        np.testing.assert_allclose(b["prob"].sum(1), 1.0, atol=1e-5)
        # ---- gradient linearity at the full minibatch size ----
        opt_data, info = algo.optimize_policy(0, buf)            # binds the training inputs; a real iteration
        losses_n = len(info["GradNorm"])
        assert losses_n == 4 * (B * T // 512) == 256 and np.isfinite(info["GradNorm"]).all()
        assert eng.get_opt_state()["step"] == 256
        rows = torch.randperm(B * T, device="cuda")[:512].to(torch.int32).contiguous()
        eng.grad_minibatch(rows, 512)
        torch.cuda.synchronize()
        g_full = t2n(eng.grad).copy()
        halves = []
        for h in range(2):
            eng.grad_minibatch(rows[h * 256:(h + 1) * 256].contiguous(), 256)
            torch.cuda.synchronize()
            halves.append(t2n(eng.grad).copy())
        g_mean = 0.5 * (halves[0] + halves[1])
        assert relerr(g_full, g_mean) < 2e-4
        # ---- idempotence: a zero step size changes nothing ----
        before = pol.get_param_values()
        algo._lr_mult = 0.0
        eng.set_lr_mult(0.0)
        eng.train_minibatches(algo.optimizer._idx_dev, 512, 4)
        torch.cuda.synchronize()
        assert np.array_equal(pol.get_param_values(), before)
        assert eng.device_error() == 0
    finally:
        eng.close()


@pytest.mark.parametrize("mbr", [True, False])
def test_device_sampler_rgb_mode_vs_oracle(mbr):
    """north-star frame mode inside the sampler: RGB 210x160x3 emulator frames -> gray -> 84x84 stacks.  Device rollout
    buffers == the oracle sampler run over the same RGB pool (oracle/frame.py:rgb_*, builder-defined arithmetic), with
    resets, life losses and (mbr False) envs that stop stepping; then two PPO iterations run on those buffers."""
    from accel_rl_b200.algos import PPO
    from accel_rl_b200.util.seeding import set_seed
    rules = dict(synth_ale.DEFAULT_RULES, pool_frames=64, life_base=10, life_mod=7, reward_mod=5, pool_seed=0)
    set_seed(3)
    B, T = 16, 12
    sampler = _make_sampler(4, 2, T, mid_batch_reset=mbr, rules=rules, env_kw=dict(frame_mode="rgb"))
    env_spec, sample_size, horizon, _ = sampler.initialize(seed=4, affinities=dict(), discount=0.99, need_extra_obs=True)
    assert env_spec.observation_space.shape == (4, 84, 84)
    pol, flat, spec = make_policy(1, hw=(84, 84))
    algo = PPO(optimizer_args=dict(minibatch_size=64, epochs=1))
    algo.initialize(pol, env_spec, sample_size, horizon, mbr)
    sampler.policy_init(pol)
    algo.set_n_itr(10)
    pool = synth_ale.make_pool(64, seed=0, channels=3)
    orc = osampler.OracleSampler(B, T, pool, {k: v for k, v in rules.items() if k != "pool_seed"}, 4, 0.99,
                                 mid_batch_reset=mbr)
    try:
        for itr in range(3):
            buf, infos = sampler.obtain_samples(itr)
            b = _buf_np(buf)
            u = sampler._uniforms_host.numpy().copy()
            gp = b["prob"].reshape(B, T, 4); gv = b["value"].reshape(B, T)
            calls = {"k": 0}

            def policy_fn(obs):
                k = calls["k"]; calls["k"] += 1
                s, j = divmod(k, 2)
                lo, hi = j * B // 2, (j + 1) * B // 2
                return gp[lo:hi, s], gv[lo:hi, s]
            ob, oinf = orc.obtain_samples(policy_fn, u)
            for k in ("observations", "extra_observations", "rewards", "dones", "raw_reward", "need_reset", "actions"):
                assert np.array_equal(b[k], ob[k]), (k, itr)
            assert len(infos) == len(oinf)
            # the policy saw exactly these observations: forward of the buffered rows reproduces the stored probabilities
            if mbr:
                p_ref, _ = onet.forward(torch.tensor(pol.get_param_values()), torch.tensor(ob["observations"][:32]), spec, 4, True)
                assert relerr(b["prob"][:32], p_ref.numpy()) < 2e-3
            opt_data, info = algo.optimize_policy(itr, buf)
            assert np.isfinite(info["GradNorm"]).all()
        assert pol.engine.device_error() == 0
    finally:
        pol.engine.close()


@pytest.mark.parametrize("mbr,frame_mode", [(True, "gray"), (False, "gray"), (True, "rgb")])
def test_host_emulator_sampler_vs_oracle(mbr, frame_mode):
    """Emulators in host worker processes (two alternating groups, raw screens through pinned shared memory, pixels on
    the GPU): with the oracle's synthetic ALE inside the workers the rollout buffers and the TrajInfos equal the oracle
    sampler's (= the device-resident sampler's), incl. life losses, game overs, and envs that stop stepping."""
    from functools import partial
    from accel_rl_b200.envs import AtariEnv
    from accel_rl_b200.sampler import HostEmulatorSampler
    from accel_rl_b200.util.seeding import set_seed
    from tests import fake_ale
    rgb = frame_mode == "rgb"
    rules = dict(synth_ale.DEFAULT_RULES, pool_frames=64, life_base=9, life_mod=5, reward_mod=7)
    set_seed(1)
    B, T = 16, 10
    sampler = HostEmulatorSampler(emu_factory=partial(fake_ale.make, rules=rules, pool_seed=0, channels=3 if rgb else 1),
                                  EnvCls=AtariEnv, env_args=dict(game="breakout", max_start_noops=0, frame_mode=frame_mode),
                                  horizon=T, n_parallel=2, envs_per=4, max_path_length=27000, mid_batch_reset=mbr,
                                  max_decorrelation_steps=0)
    sampler.initialize(seed=2, affinities=dict(), discount=0.99, need_extra_obs=True)
    pol, flat, spec = make_policy(1, max_rows=B, hw=(84, 84) if rgb else (104, 80))
    pool = synth_ale.make_pool(64, seed=0, channels=3 if rgb else 1)
    orc = osampler.OracleSampler(B, T, pool, rules, 4, 0.99, mid_batch_reset=mbr)

    def key(ti):
        return (ti["env"], ti["Length"], round(float(ti["Return"]), 4), round(float(ti["RawReturn"]), 4),
                ti["NonzeroRewards"], round(float(ti["DiscountedReturn"]), 3))
    try:
        sampler.policy_init(pol)
        n_traj = 0
        for itr in range(4):
            buf, infos = sampler.obtain_samples(itr)
            b = _buf_np(buf)
            u = sampler._uniforms_host.numpy().copy()
            gp = b["prob"].reshape(B, T, 4); gv = b["value"].reshape(B, T)
            calls = {"k": 0}

            def policy_fn(obs):
                k = calls["k"]; calls["k"] += 1
                s, j = divmod(k, 2)
                lo, hi = j * B // 2, (j + 1) * B // 2
                return gp[lo:hi, s], gv[lo:hi, s]
            ob, oinf = orc.obtain_samples(policy_fn, u)
            for k in ("observations", "extra_observations", "rewards", "dones", "raw_reward", "need_reset", "actions"):
                assert np.array_equal(b[k], ob[k]), (k, itr)
            assert sorted(key(t) for t in infos) == sorted(key(t) for t in oinf)
            n_traj += len(infos)
        assert n_traj > 0 and sampler.h2d_bytes > 4 * T * B * 2 * 33600
        assert pol.engine.device_error() == 0
    finally:
        sampler.shutdown()
        pol.engine.close()


def test_eval_sampler_vs_oracle_and_training_envs_untouched():
    """AAOEvalSampler.evaluate_policy (sampler_with_eval.py:20-33, worker_with_eval.py:66-99): the evaluation envs are
    reset at the start of every evaluation, run eval_horizon steps, and return exactly the trajectories the oracle
    sampler completes on fresh envs with the same probabilities and uniforms; the training rollouts before and after
    are those of a sampler that never evaluated (its envs and step buffer are untouched)."""
    from accel_rl_b200.sampler import AAOEvalSampler
    from accel_rl_b200.envs import AtariEnv
    from accel_rl_b200.util.seeding import set_seed
    rules = dict(synth_ale.DEFAULT_RULES, pool_frames=128, life_base=9, life_mod=5, reward_mod=7, pool_seed=0)
    orules = {k: v for k, v in rules.items() if k != "pool_seed"}
    set_seed(5)
    B, T = 16, 8
    Be, eval_steps = 8, 8 * 60
    sampler = AAOEvalSampler(eval_steps, 1, EnvCls=AtariEnv,
                             env_args=dict(game="breakout", max_start_noops=0, synth_rules=rules), horizon=T,
                             n_parallel=4, envs_per=2, max_path_length=27000, mid_batch_reset=True,
                             max_decorrelation_steps=0)
    assert sampler.eval_horizon == 60 and sampler._total_n_eval_envs == Be
    sampler.initialize(seed=2, affinities=dict(), discount=0.99, need_extra_obs=True)
    pol, flat, spec = make_policy(1, max_rows=B)
    sampler.policy_init(pol)
    pool = synth_ale.make_pool(128, seed=0)
    orc = osampler.OracleSampler(B, T, pool, orules, 4, 0.99)

    def feeder(prob, val, n, horizon):
        calls = {"k": 0}

        def policy_fn(obs):
            k = calls["k"]; calls["k"] += 1
            s, j = divmod(k, 2)
            lo, hi = j * n // 2, (j + 1) * n // 2
            return prob[lo:hi, s], val[lo:hi, s]
        return policy_fn

    def key(ti):
        return (ti["env"], ti["Length"], round(float(ti["Return"]), 4), round(float(ti["RawReturn"]), 4),
                ti["NonzeroRewards"], round(float(ti["DiscountedReturn"]), 3))
    try:
        n_eval_traj = 0
        for itr in range(3):
            # --- evaluation: fresh oracle envs every time ---
            infos = sampler.evaluate_policy(itr)
            eb = sampler.eval_buf
            Te = sampler.eval_horizon
            ue = sampler._eval_uniforms_host.numpy().copy()
            ep = t2n(eb.prob).reshape(Be, Te, 4); ev = t2n(eb.value).reshape(Be, Te)
            eorc = osampler.OracleSampler(Be, Te, pool, orules, 4, 0.99)
            ob, oinf = eorc.obtain_samples(feeder(ep, ev, Be, Te), ue)
            for k in ("rewards", "dones", "raw_reward", "need_reset", "actions"):
                assert np.array_equal(t2n(getattr(eb, k)), ob[k]), ("eval", k, itr)
            assert sorted(key(t) for t in infos) == sorted(key(t) for t in oinf)
            n_eval_traj += len(infos)
            # --- training rollout continues as if no evaluation had happened ---
            buf, tinfos = sampler.obtain_samples(itr)
            b = _buf_np(buf)
            u = sampler._uniforms_host.numpy().copy()
            gp = b["prob"].reshape(B, T, 4); gv = b["value"].reshape(B, T)
            ob, oinf = orc.obtain_samples(feeder(gp, gv, B, T), u)
            for k in ("observations", "extra_observations", "rewards", "dones", "raw_reward", "need_reset", "actions"):
                assert np.array_equal(b[k], ob[k]), ("train", k, itr)
            assert len(tinfos) == len(oinf)
        assert n_eval_traj > 0
        assert pol.engine.device_error() == 0
    finally:
        pol.engine.close()


def _run_iteration(algo_name, B, T, spec_id, mbr, mb, epochs, standardize=False, itrs=2):
    from accel_rl_b200.algos import PPO, A2C
    from accel_rl_b200.util.seeding import set_seed
    rules = dict(synth_ale.DEFAULT_RULES, pool_frames=128, life_base=24, life_mod=11, reward_mod=7, pool_seed=0)
    set_seed(7)
    sampler = _make_sampler(B // 4, 2, T, mid_batch_reset=mbr, rules=rules)
    env_spec, sample_size, horizon, _ = sampler.initialize(seed=8, affinities=dict(), discount=0.99, need_extra_obs=True)
    pol, flat, spec = make_policy(spec_id)
    if algo_name == "ppo":
        algo = PPO(optimizer_args=dict(minibatch_size=mb, epochs=epochs), standardize_adv=standardize)
        opt = onet.Adam(flat.size, 1e-3, epsilon=1e-5)
    else:
        algo = A2C(standardize_adv=standardize)
        opt = onet.RMSProp(flat.size, 7e-4)
    algo.initialize(pol, env_spec, sample_size, horizon, mbr)
    sampler.policy_init(pol)
    algo.set_n_itr(10)
    out = []
    try:
        for itr in range(itrs):
            buf, _ = sampler.obtain_samples(itr)
            b = _buf_np(buf)
            rng_state = np.random.get_state()
            seen = []
            orig_optimize = algo.optimizer.optimize
            algo.optimizer.optimize = lambda inputs: (seen.append(orig_optimize(inputs)), seen[-1])[1]
            opt_data, opt_infos = algo.optimize_policy(itr, buf)
            algo.optimizer.optimize = orig_optimize
            dev_losses = [float(x) for x in np.atleast_1d(seen[0][0])]   # optimize() -> (losses, grad_norms) per minibatch
            torch.cuda.synchronize()
            new = pol.get_param_values()
            # oracle on the same buffers, same shuffles (replay the global stream)
            rng = np.random.RandomState()
            rng.set_state(rng_state)
            last_values = t2n(algo._last_values)
            flat_ref, losses, norms, od = olearner.optimize_policy(
                flat, opt, b, spec, 4, T, algo_name, rng, epochs=epochs, minibatch_size=mb, use_valids=not mbr,
                standardize_adv=standardize, last_values=last_values)
            out.append(dict(flat_old=flat, flat_new=new, flat_ref=flat_ref, losses=losses, norms=norms, od=od,
                            dev_losses=dev_losses,
                            opt_data={k: t2n(v) for k, v in opt_data.items() if torch.is_tensor(v)},
                            grad_norm=opt_infos["GradNorm"], value_after=t2n(buf.agent_infos.value)))
            flat = new            # continue from the device's parameters (no drift accumulation in the check)
            opt_sync = opt
            if algo_name == "ppo":
                opt_sync.m = t2n(pol.engine.m).copy(); opt_sync.v = t2n(pol.engine.v).copy()
            else:
                opt_sync.v = t2n(pol.engine.v).copy()
    finally:
        pol.engine.close()
    return out


@pytest.mark.parametrize("standardize", [False, True])
def test_ppo_iteration_vs_oracle(standardize):
    res = _run_iteration("ppo", B=16, T=16, spec_id=1, mbr=True, mb=64, epochs=2, standardize=standardize)
    for r in res:
        # advantages / returns: north-star 1e-4
        np.testing.assert_allclose(r["opt_data"]["advantages"], r["od"]["advantages"], rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(r["opt_data"]["returns"], r["od"]["returns"], rtol=1e-4, atol=1e-4)
        assert len(r["grad_norm"]) == len(r["norms"]) == 8
        # first minibatch starts from identical parameters: tight
        assert abs(r["grad_norm"][0] - r["norms"][0]) <= 5e-3 * r["norms"][0]
        # the per-minibatch LOSS series optimize() returns (single/ppo_optimizer.py:57-76): the first minibatch starts
        # from identical parameters (north-star 1e-4 relative, + the fp32 summation floor); the later ones see parameters
        # that have drifted by bf16-level gradient differences
        dl, rl = np.array(r["dev_losses"]), np.array(r["losses"])
        assert dl.shape == rl.shape == (8,)
        print("LOSS-SERIES standardize=%s device %s oracle %s" % (standardize, np.round(dl, 5), np.round(rl, 5)))
        assert abs(dl[0] - rl[0]) <= 1e-4 * abs(rl[0]) + 2e-4
        np.testing.assert_allclose(dl, rl, rtol=2e-2, atol=2e-3)
        np.testing.assert_allclose(r["grad_norm"], r["norms"], rtol=5e-2)
        # Adam amplifies bf16-level gradient differences on near-zero-gradient coordinates; compare the
        # accumulated update direction and size
        du, dr = r["flat_new"] - r["flat_old"], r["flat_ref"] - r["flat_old"]
        assert relerr(du, dr) < 0.15
        cos = float(np.dot(du, dr) / (np.linalg.norm(du) * np.linalg.norm(dr)))
        assert cos > 0.99


def test_operand_copies_refreshed_by_the_update_match_a_fresh_pack():
    """The update kernel rewrites the bf16 FC operand tiles (and pack_weights the conv tap tiles) after every step.  After
    a PPO iteration the trained engine's forward pass must equal, bit for bit, that of a fresh engine loaded with the
    same fp32 parameters (whose operand copies come from the independent pack_weights path)."""
    from accel_rl_b200.policies import AtariCnnPolicy, cnn_specs
    from accel_rl_b200.envs.atari_env import EnvSpec
    from accel_rl_b200.spaces import Discrete, UintBox
    res_obs = np.random.RandomState(3).randint(0, 256, (64, 4, 104, 80), dtype=np.uint8)
    from accel_rl_b200.algos import PPO
    from accel_rl_b200.util.seeding import set_seed
    rules = dict(synth_ale.DEFAULT_RULES, pool_frames=128, life_base=24, life_mod=11, reward_mod=7, pool_seed=0)
    set_seed(7)
    sampler = _make_sampler(4, 2, 16, rules=rules)
    env_spec, sample_size, horizon, mbr = sampler.initialize(seed=8, affinities=dict(), discount=0.99, need_extra_obs=True)
    pol, flat, spec = make_policy(1)
    algo = PPO(optimizer_args=dict(minibatch_size=64, epochs=2))
    algo.initialize(pol, env_spec, sample_size, horizon, mbr)
    sampler.policy_init(pol)
    algo.set_n_itr(10)
    try:
        buf, _ = sampler.obtain_samples(0)
        kl0, dv0 = algo.constraint_values(buf)               # aac_base.py:68-70 before the update: old == new policy
        assert abs(kl0) < 1e-6 and dv0 < 1e-10
        algo.optimize_policy(0, buf)
        trained = pol.get_param_values()
        assert not np.array_equal(trained, flat)
        # ... and after it: against the oracle net with the trained parameters on the same rows
        kl1, dv1 = algo.constraint_values(buf)
        b = _buf_np(buf)
        p_new, v_new = onet.forward(torch.tensor(trained), torch.tensor(b["observations"]), spec, 4, emulate_bf16=True)
        p_old = b["prob"].astype(np.float64)
        want_kl = float(np.mean(np.sum(p_old * (np.log(p_old + 1e-8) - np.log(p_new.numpy().astype(np.float64) + 1e-8)), axis=1)))
        want_dv = float(np.mean((v_new.numpy().astype(np.float64) - b["value"]) ** 2))
        assert kl1 > 0 and abs(kl1 - want_kl) <= 2e-2 * want_kl + 1e-7, (kl1, want_kl)
        assert abs(dv1 - want_dv) <= 2e-2 * want_dv + 1e-9, (dv1, want_dv)
        obs = torch.tensor(res_obs).cuda()
        p1 = torch.zeros(64, 4, device="cuda"); v1 = torch.zeros(64, device="cuda")
        pol.engine.forward(obs, prob=p1, value=v1)
        fresh = AtariCnnPolicy(initial_param_values=trained, max_rows=64, **cnn_specs[1])
        fresh.initialize(EnvSpec(UintBox((4, 104, 80)), Discrete(4)))
        try:
            p2 = torch.zeros(64, 4, device="cuda"); v2 = torch.zeros(64, device="cuda")
            fresh.engine.forward(obs, prob=p2, value=v2)
            torch.cuda.synchronize()
            assert torch.equal(p1, p2) and torch.equal(v1, v2)
            # ... and the data-gradient operand packs: the same minibatch gradient, bit for bit
            rng = np.random.RandomState(5)
            n = 64
            act = rng.randint(0, 4, n).astype(np.uint8)
            adv = rng.randn(n).astype(np.float32); ret = rng.randn(n).astype(np.float32)
            oldp = rng.dirichlet(np.ones(4), n).astype(np.float32); oldv = rng.randn(n).astype(np.float32)
            grads = []
            for e in (pol.engine, fresh.engine):
                e.opt_configure(algo=0, clip_param=0.2, v_loss_coeff=1.0, ent_loss_coeff=0.01, update=0, learning_rate=1e-3,
                                beta1=0.9, beta2=0.999, epsilon=1e-5, rho=0.9, grad_norm_clip=-1.0)
                e.bind_train_inputs(*[torch.tensor(x).cuda() for x in (res_obs, act, adv, ret, oldv, oldp)], valids=None)
                e.grad_minibatch(torch.arange(n, dtype=torch.int32, device="cuda"), n)
                torch.cuda.synchronize()
                grads.append(t2n(e.grad).copy())
            assert np.array_equal(grads[0], grads[1])
        finally:
            fresh.engine.close()
    finally:
        pol.engine.close()


def test_example_script_trains_and_logs(tmp_path):
    """scripts/example/example_train_ppo.py (the reference's example_train_ppo.py / example_train_a2c.py with the imports
    swapped): builds sampler / algo / policy / runner the reference's way — incl. its defaults
    (max_decorrelation_steps 2000, start no-ops 30) — trains and writes <log_dir>/<game>_<run_ID>/progress.csv"""
    import csv
    from accel_rl_b200.scripts.example.example_train_ppo import build_and_run
    for algo, n_steps, interval in (("ppo", 3 * 16 * 128, 16 * 128), ("a2c", 40 * 16 * 5, 10 * 16 * 5)):
        build_and_run(str(tmp_path), "breakout", algo, algo=algo, n_envs=16, n_steps=n_steps, n_sim_cores=2,
                      log_interval_steps=interval)
        rows = list(csv.DictReader(open(tmp_path / ("breakout_%s" % algo) / "progress.csv")))
        assert len(rows) >= 3
        assert all(np.isfinite(float(r["GradNormAverage"])) for r in rows) and float(rows[-1]["SamplesPerSecond"]) > 0
        assert float(rows[-1]["CumTotalSteps"]) >= n_steps


def test_sampler_reconfigure_and_log_overflow():
    """re-configuring the sampler frees the bf16 rollout mirror the captured training graph gathers from — the
    graph must be dropped with it (a stale graph would read freed memory); and an optimize() call with more minibatches
    than log slots must fail loudly instead of truncating the returned losses."""
    from accel_rl_b200.algos import PPO
    from accel_rl_b200._lib import ArlError
    from accel_rl_b200.util.seeding import set_seed
    rules = dict(synth_ale.DEFAULT_RULES, pool_frames=128, life_base=24, life_mod=11, reward_mod=7, pool_seed=0)
    set_seed(7)
    sampler = _make_sampler(4, 2, 16, rules=rules)
    env_spec, sample_size, horizon, mbr = sampler.initialize(seed=8, affinities=dict(), discount=0.99, need_extra_obs=True)
    pol, flat, spec = make_policy(1)
    algo = PPO(optimizer_args=dict(minibatch_size=64, epochs=2))
    algo.initialize(pol, env_spec, sample_size, horizon, mbr)
    sampler.policy_init(pol)
    algo.set_n_itr(10)
    try:
        buf, _ = sampler.obtain_samples(0)
        _, info0 = algo.optimize_policy(0, buf)
        sampler._configure_engine()                          # same torch buffers, new mirrors / tables inside the library
        buf, _ = sampler.obtain_samples(1)
        _, info1 = algo.optimize_policy(1, buf)
        torch.cuda.synchronize()
        assert np.isfinite(info0["GradNorm"]).all() and np.isfinite(info1["GradNorm"]).all()
        assert pol.engine.device_error() == 0
        idx = torch.zeros(5000 * 8, dtype=torch.int32, device="cuda")
        with pytest.raises(ArlError, match="log slots"):
            pol.engine.train_minibatches(idx, 8, 5000)
    finally:
        pol.engine.close()


def test_early_fc_update_is_bit_identical():
    """Without global-norm clipping (PPO's default) the FC weights are updated as soon as their gradient is final, while
    the conv gradient chain still runs (update_range_kernel).  Same arithmetic per element: parameters and optimizer
    state after two PPO iterations are bit-identical to the single end-of-minibatch update; the logged norms agree to
    fp32 rounding (the partial sums are grouped differently)."""
    import json
    import subprocess
    import sys
    res = {}
    for flag in ("0", "1"):
        env = dict(os.environ, ARL_EARLY_FC=flag)
        p = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "early_fc_worker.py")], env=env,
                           capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stderr[-2000:]
        line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")][-1]
        res[flag] = json.loads(line[len("RESULT "):])
    a, b = res["0"], res["1"]
    assert a["device_error"] == 0 and b["device_error"] == 0
    assert a["step"] == b["step"] == 2 * 2 * 4
    assert a["params"] == b["params"] and a["m"] == b["m"] and a["v"] == b["v"]
    np.testing.assert_allclose(a["norms"], b["norms"], rtol=2e-6)


def _ab_worker(env_over):
    import json
    import subprocess
    import sys
    env = dict(os.environ, **env_over)
    p = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "early_fc_worker.py")], env=env,
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")][-1]
    return json.loads(line[len("RESULT "):])


def test_stream_update_and_iteration_graph_are_bit_identical():
    """Round-2 training path (update_stream_kernel: finalisation + Adam + operand refresh + logs in one launch without
    a grid barrier; ONE CUDA graph holding every minibatch of the optimize() call) against the round-1 path
    (finalize_grads -> update_fused with barrier, one graph launch per minibatch), and with the update split over two
    streams (ARL_SPLIT_UPDATE) or not: same per-element arithmetic, so the
    parameters and optimizer state after two PPO iterations are bit-identical; the logged norms agree to fp32 rounding
    (block partials are grouped differently)."""
    old = _ab_worker(dict(ARL_STREAM_UPDATE="0", ARL_GRAPH_MB="1"))
    # ARL_SPLIT_UPDATE=1: the stream update as two grids (the FC range on a side stream beside the next minibatch's conv layers)
    for over in (dict(), dict(ARL_SPLIT_UPDATE="1"), dict(ARL_STREAM_UPDATE="1", ARL_GRAPH_MB="1"),
                 dict(ARL_STREAM_UPDATE="0", ARL_GRAPH_MB="512")):
        new = _ab_worker(over)
        assert old["device_error"] == 0 and new["device_error"] == 0
        assert old["step"] == new["step"] == 2 * 2 * 4
        assert old["params"] == new["params"] and old["m"] == new["m"] and old["v"] == new["v"], over
        np.testing.assert_allclose(old["norms"], new["norms"], rtol=2e-6)


def test_a2c_nonreset_iteration_vs_oracle():
    """A2C as shipped by the reference example: mid_batch_reset=False -> valids mask, zero_after_reset,
    valids-weighted loss means (SURVEY.md §8 a6')."""
    res = _run_iteration("a2c", B=16, T=12, spec_id=0, mbr=False, mb=None, epochs=1, itrs=3)
    saw_invalid = False
    for r in res:
        assert np.array_equal(r["opt_data"]["valids"], r["od"]["valids"])
        saw_invalid |= bool((r["od"]["valids"] == 0).any())
        np.testing.assert_allclose(r["opt_data"]["advantages"], r["od"]["advantages"], rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(r["opt_data"]["returns"], r["od"]["returns"], rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(r["value_after"], r["od"]["values"], rtol=0, atol=0)
        assert abs(r["grad_norm"] - r["norms"][0]) <= 5e-3 * r["norms"][0]
        du, dr = r["flat_new"] - r["flat_old"], r["flat_ref"] - r["flat_old"]
        assert relerr(du, dr) < 0.05
    assert saw_invalid, "test must exercise the validity mask"


def test_runner_trains_and_logs(tmp_path):
    from accel_rl_b200.algos import PPO
    from accel_rl_b200.policies import AtariCnnPolicy, cnn_specs
    from accel_rl_b200.runners import AccelRL
    from accel_rl_b200.util import logger
    rules = dict(synth_ale.DEFAULT_RULES, pool_frames=128, life_base=24, life_mod=11, reward_mod=7, pool_seed=0)
    sampler = _make_sampler(4, 4, 16, rules=rules)
    algo = PPO(optimizer_args=dict(minibatch_size=128, epochs=2), lr_schedule="linear")
    policy = AtariCnnPolicy(**cnn_specs[0])
    logger.configure(str(tmp_path), quiet=True)
    runner = AccelRL(algo=algo, policy=policy, sampler=sampler, n_steps=32 * 16 * 6, seed=0, log_interval_steps=32 * 16 * 2)
    runner.train()
    row = dict(logger.last_row)
    for k in ("Iteration", "CumTotalSteps", "Entropy", "GradNormAverage", "ParamsNorm", "NormFromInit", "SamplesPerSecond",
              "LengthAverage", "ReturnAverage"):
        assert k in row, k
    assert row["NormFromInit"] > 0 and np.isfinite(row["GradNormAverage"])
    assert os.path.exists(os.path.join(str(tmp_path), "progress.csv"))
    policy.engine.close()
    logger.configure(None)


def test_eval_runner_snapshot_and_resume(tmp_path):
    """AccelRLEval + AAOEvalSampler (runners/accel_rl.py:108-180): evaluation rows in progress.csv, a 'last' snapshot
    with parameters AND optimizer state, and a second runner resuming from it at the next iteration."""
    import joblib
    from accel_rl_b200.algos import PPO
    from accel_rl_b200.envs import AtariEnv
    from accel_rl_b200.policies import AtariCnnPolicy, cnn_specs
    from accel_rl_b200.runners import AccelRLEval
    from accel_rl_b200.sampler import AAOEvalSampler
    from accel_rl_b200.util import logger
    rules = dict(synth_ale.DEFAULT_RULES, pool_frames=128, life_base=9, life_mod=5, reward_mod=7, pool_seed=0)

    def make(resume=None):
        sampler = AAOEvalSampler(16 * 50, 2, EnvCls=AtariEnv,
                                 env_args=dict(game="breakout", max_start_noops=0, synth_rules=rules), horizon=16,
                                 n_parallel=4, envs_per=4, mid_batch_reset=True, max_decorrelation_steps=0)
        algo = PPO(optimizer_args=dict(minibatch_size=128, epochs=2), lr_schedule="linear")
        policy = AtariCnnPolicy(**cnn_specs[0])
        return AccelRLEval(algo=algo, policy=policy, sampler=sampler, n_steps=32 * 16 * 4, seed=0,
                           eval_interval_steps=32 * 16 * 2, resume_from=resume), policy
    logger.configure(str(tmp_path), snapshot_mode="last", quiet=True)
    runner, policy = make()
    runner.train()
    row = dict(logger.last_row)
    for k in ("Iteration", "CumCompletedSteps", "StepsInEval", "TrajsInEval", "LengthAverage", "ReturnAverage",
              "GradNormAverage", "CumTrainTime", "CumEvalTime", "SamplesPerSecond"):
        assert k in row, k
    assert row["TrajsInEval"] > 0 and row["StepsInEval"] > 0
    snap_path = os.path.join(str(tmp_path), "params.pkl")
    snap = joblib.load(snap_path)
    assert snap["itr"] == row["Iteration"] and "optimizer_state" in snap
    assert snap["optimizer_state"]["step"] == snap["itr"] * 2 * (32 * 16 // 128)
    assert snap["optimizer_state"]["m"].any() and snap["optimizer_state"]["v"].any()
    policy.engine.close()
    # resume: parameters and optimizer state come from the snapshot, iterations continue after it
    runner2, policy2 = make(resume=snap_path)
    n_itr = runner2.startup()
    assert runner2._start_itr == snap["itr"] + 1 <= n_itr
    assert np.array_equal(policy2.get_param_values(), snap["policy_param_values"])
    st = runner2.algo.optimizer.get_state()
    assert st["step"] == snap["optimizer_state"]["step"] and np.array_equal(st["m"], snap["optimizer_state"]["m"])
    policy2.engine.close()
    logger.configure(None)


def test_runner_rejects_mismatched_parallelism():
    from accel_rl_b200.algos import mPPO
    from accel_rl_b200.policies import AtariCnnPolicy, cnn_specs
    from accel_rl_b200.runners import AccelRL
    with pytest.raises(TypeError):
        AccelRL(algo=mPPO(), policy=AtariCnnPolicy(**cnn_specs[0]), sampler=_make_sampler(1, 1, 4), n_steps=100)


def test_host_fed_rollout_runs():
    """frame_feed='host': raw frames arrive from pinned host memory every step (H2D inside the step)."""
    from accel_rl_b200.util.seeding import set_seed
    set_seed(0)
    sampler = _make_sampler(4, 2, 8, frame_feed="host", host_ring_steps=4)
    sampler.initialize(seed=1, affinities=dict(), discount=0.99, need_extra_obs=True)
    pol, flat, spec = make_policy(0, max_rows=16)
    sampler.policy_init(pol)
    try:
        buf, _ = sampler.obtain_samples(0)
        torch.cuda.synchronize()
        obs = t2n(buf.observations)
        assert obs[1].any() and sampler.h2d_bytes >= 8 * 16 * 2 * 33600
        # newest plane of row (e, 1) is the box-downsample of max(ring frames) for step 0
        from oracle import frame as oframe
        ring = sampler._ring_host.numpy()
        want = oframe.downsample(np.maximum(ring[0, 3, 0], ring[0, 3, 1]))
        assert np.array_equal(obs[3 * 8 + 1, 3], want)
    finally:
        pol.engine.close()
