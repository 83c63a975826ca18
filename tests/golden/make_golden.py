"""Generates the golden fixtures in this directory by RUNNING THE REFERENCE'S OWN CODE
(/root/reference imported under stubs, see oracle/ref_harness.py).  Only runnable in the build
container; the committed .npz files are what travels.

    python tests/golden/make_golden.py
"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_harness as H, synth_ale  # noqa: E402

POOL_FRAMES = 64
RULES = dict(synth_ale.DEFAULT_RULES)
RULES.update(pool_frames=POOL_FRAMES, life_base=12, life_mul=3, life_mod=7, reward_mod=5)


def fake_policy_fn(obs, n_actions):
    """deterministic stand-in policy used for the sampler fixtures (numpy only)"""
    n = len(obs)
    x = obs.reshape(n, obs.shape[1], -1).astype(np.float64)
    feats = np.stack([x[:, -1, k::n_actions].mean(axis=1) for k in range(n_actions)], axis=1) / 16.0
    z = feats - feats.max(axis=1, keepdims=True)
    p = np.exp(z)
    p = (p / p.sum(axis=1, keepdims=True)).astype(np.float32)
    v = (x.mean(axis=(1, 2)) / 255.0).astype(np.float32)
    return p, v


def gen_frames():
    env_mod = H.ref("accel_rl.envs.atari_env")
    rng = np.random.RandomState(11)
    synth_ale.SynthALE.next_env_id = 0
    env = env_mod.AtariEnv(game="breakout", max_start_noops=0)
    cases = 10
    stacks = rng.randint(0, 256, (cases, 4, 104, 80), dtype=np.uint8)
    raw1 = rng.randint(0, 256, (cases, 210, 160), dtype=np.uint8)
    raw2 = rng.randint(0, 256, (cases, 210, 160), dtype=np.uint8)
    raw1[0] = 0; raw2[1] = 255; raw1[2] = 255; raw2[2] = 255; raw1[3] = raw2[3]
    reset = (rng.rand(cases) < 0.3)
    out = np.zeros_like(stacks)

    class OneFrame(object):   # ale stand-in returning a fixed screen
        def __init__(self, f): self.f = f
        def getScreenGrayscale(self, buf): buf[:] = self.f.reshape(buf.shape)
    for i in range(cases):
        env._obs = stacks[i].copy()
        env._raw_frame_1[:] = raw1[i].reshape(210, 160, 1)
        if reset[i]:
            env._reset_obs()
        env.ale = OneFrame(raw2[i])
        env._update_obs()          # the reference's own max -> crop -> cv2.resize -> concat
        out[i] = env._obs
    np.savez_compressed(os.path.join(HERE, "frames.npz"), stacks=stacks, raw1=raw1, raw2=raw2, reset=reset, out=out)


def gen_sampling():
    special = H.ref("rllab.misc.special")
    rng = np.random.RandomState(5)
    cases = {}
    for A in (4, 6, 18):
        p = rng.dirichlet(np.ones(A) * 0.5, 257).astype(np.float32)
        p[0] = 0; p[0, A - 1] = 1.0
        p[1] = 1.0 / A
        np.random.seed(77 + A)
        state_before = np.random.get_state()
        acts = special.weighted_sample_n(p, np.arange(A).astype(np.uint8))
        np.random.set_state(state_before)
        u = np.random.rand(257)
        cases["p%d" % A] = p; cases["a%d" % A] = acts; cases["u%d" % A] = u
    np.savez_compressed(os.path.join(HERE, "sampling.npz"), **cases)


def gen_gae():
    util = H.ref("accel_rl.algos.pg.util")
    rng = np.random.RandomState(9)
    B, T = 7, 33
    r = rng.choice([0., 0., 0., 1., -1.], (B, T)).astype(np.float32)
    v = rng.randn(B, T).astype(np.float32)
    d = rng.rand(B, T) < 0.1
    d[0, 0] = True; d[1, T - 1] = True; d[2] = False
    nr = d & (rng.rand(B, T) < 0.5)
    nr[3] = False
    lv = rng.randn(B).astype(np.float32)
    out = dict(r=r, v=v, d=d, nr=nr, lv=lv)
    for name, lam in (("gae", 0.95), ("ret", 1.0)):
        adv = np.zeros((B, T), np.float32); ret = np.zeros((B, T), np.float32)
        for e in range(B):
            if lam == 1.0:
                util.discount_returns(r[e], d[e], lv[e], 0.99, ret_dest=ret[e])
                adv[e] = ret[e] - v[e]
            else:
                util.gen_adv_est(r[e], v[e], d[e], lv[e], 0.99, lam, adv_dest=adv[e], ret_dest=ret[e])
        out["adv_" + name] = adv; out["ret_" + name] = ret
        # valids variant
        adv2, ret2, v2 = adv.copy(), ret.copy(), v.copy()
        valids = np.zeros((B, T), np.int8)
        for e in range(B):
            path = dict(env_infos=dict(need_reset=nr[e]), dones=d[e])
            util.update_valids(path, valids[e])
            util.zero_after_reset(adv2[e], ret2[e], v2[e], path)
        out["valids"] = valids
        out["adv_%s_valid" % name] = adv2; out["ret_%s_valid" % name] = ret2; out["v_valid"] = v2
    np.savez_compressed(os.path.join(HERE, "gae.npz"), **out)


def gen_mb_idxs():
    util = H.ref("accel_rl.optimizers.util")
    np.random.seed(123)
    rows = []
    for _ in range(3):
        rows.append(np.concatenate([b[0] for b in util.iterate_mb_idxs(16, 100, shuffle=True)]))
    np.savez_compressed(os.path.join(HERE, "mb_idxs.npz"), idx=np.stack(rows), seed=123, batch=16, length=100)


def gen_sampler(mid_batch_reset, tag, max_path_length=27000):
    disc = H.ref("accel_rl.spaces.discrete")
    ext = H.ref("rllab.misc.ext")
    n_parallel, envs_per, T, itrs = 2, 2, 6, 4
    ext.set_seed(3)
    sampler = H.make_sampler(n_parallel, envs_per, T, mid_batch_reset=mid_batch_reset, max_path_length=max_path_length)
    env_spec, sample_size, horizon, _ = H.initialize_sampler(sampler, seed=4, discount=0.99)
    A = env_spec.action_space.n

    class FakePolicy(object):
        recurrent = False
        action_space = disc.Discrete(A)
        def reset(self, n_batch=None): pass
        def reset_one(self, idx): pass
        def get_action(self, o):
            p, v = fake_policy_fn(o[None], A)
            return self.action_space.weighted_sample(p[0]), dict(prob=p[0], value=v[0])
        def get_actions(self, obs):
            p, v = fake_policy_fn(obs, A)
            return self.action_space.weighted_sample_n(p), dict(prob=p, value=v)

    sampler.policy_init(FakePolicy())
    out = dict(n_envs=2 * n_parallel * envs_per, horizon=T, itrs=itrs, max_path_length=max_path_length)
    state0 = np.random.get_state()
    for itr in range(itrs):
        buf, infos = sampler.obtain_samples(itr)
        if itr == 0:
            out["obs_%d" % itr] = buf.observations.copy()
        out["obscrc_%d" % itr] = np.array([zlib.crc32(row.tobytes()) for row in buf.observations], dtype=np.uint32)
        out["extra_%d" % itr] = buf.extra_observations.copy()
        out["rew_%d" % itr] = buf.rewards.copy()
        out["done_%d" % itr] = buf.dones.copy()
        out["raw_%d" % itr] = buf.env_infos.raw_reward.copy()
        out["nr_%d" % itr] = buf.env_infos.need_reset.copy()
        out["act_%d" % itr] = buf.actions.copy()
        out["prob_%d" % itr] = buf.agent_infos.prob.copy()
        out["val_%d" % itr] = buf.agent_infos.value.copy()
        ti = sorted([(i.Length, float(i.Return), float(i.RawReturn), int(i.NonzeroRewards), float(i.DiscountedReturn))
                     for i in infos])
        out["traj_%d" % itr] = np.array(ti, dtype=np.float64).reshape(-1, 5)
    sampler.shutdown()
    # the uniforms the master consumed: replay the stream from the state before the first obtain_samples
    np.random.set_state(state0)
    out["uniforms"] = np.random.rand(itrs, T, out["n_envs"])
    np.savez_compressed(os.path.join(HERE, "sampler_%s.npz" % tag), **out)


if __name__ == "__main__":
    pool = synth_ale.make_pool(POOL_FRAMES, seed=0)
    H.install(pool, RULES)
    gen_frames()
    gen_sampling()
    gen_gae()
    gen_mb_idxs()
    gen_sampler(True, "reset")
    gen_sampler(False, "nonreset")
    gen_sampler(True, "overlength", max_path_length=9)
    print("golden fixtures written to", HERE)
