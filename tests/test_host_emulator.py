"""Host half of the emulator-fed sampler (accel_rl_b200/hostsim): no GPU needed.

HostAtariEnv + the worker's collector logic must export, per env-step, exactly the raw frame pair, flags and scalars
from which the oracle's AtariEnv restatement (oracle/sampler.py:SynthAtariEnv, pinned to the reference's real AtariEnv
by tests/golden) builds its observations and step results."""
import ctypes as C
import multiprocessing as mp
from functools import partial

import numpy as np
import pytest

from accel_rl_b200.hostsim import worker as W
from accel_rl_b200.hostsim.atari_env import HostAtariEnv, FLAG_RESET, FLAG_SKIP, FLAG_NO_RECORD
from oracle import frame as oframe, ref_harness, sampler as osampler, synth_ale
from tests import fake_ale

RULES = dict(synth_ale.DEFAULT_RULES, pool_frames=32, life_base=9, life_mod=5, reward_mod=7)


def _apply(obs, f1, f2, flags):
    """what the device does with one exported step (frame_kernel): -> new stack"""
    if flags & FLAG_RESET:
        obs = np.zeros_like(obs)
        f1 = np.zeros_like(f2)
    return oframe.update_obs(obs, f1, f2)


@pytest.mark.parametrize("episodic", [True, False])
def test_host_atari_env_exports_what_the_oracle_env_consumes(episodic):
    pool = synth_ale.make_pool(32, seed=0)
    for e in (0, 3):
        host = HostAtariEnv(fake_ale.make(e, RULES), clip_reward=True, episodic_lives=episodic, max_start_noops=0)
        orc = osampler.SynthAtariEnv(e, pool, RULES, 4, 4, True, episodic)
        f1, f2 = np.zeros((210, 160), np.uint8), np.zeros((210, 160), np.uint8)
        obs = np.zeros((4, oframe.H, oframe.W), np.uint8)
        fl = host.reset(f2)
        assert fl == FLAG_RESET
        obs = _apply(obs, f1, f2, fl)
        assert np.array_equal(obs, orc.reset())
        rng = np.random.RandomState(e)
        saw_life = saw_over = False
        for _ in range(120):
            a = int(rng.randint(0, 4))
            r, raw, d, nr, fl = host.step(a, f1, f2)
            o2, r2, d2, info = orc.step(a)
            obs = _apply(obs, f1, f2, fl)
            assert np.array_equal(obs, o2)
            assert r == r2 and raw == info["raw_reward"] and bool(d) == bool(d2)
            assert (nr if episodic else None) == info.get("need_reset")
            saw_life |= bool(fl & FLAG_RESET)
            game_over = info.get("need_reset", d2)
            if game_over:
                saw_over = True
                fl = host.reset(f2)
                obs = _apply(obs, f1, f2, fl)
                assert np.array_equal(obs, orc.reset())
        assert saw_over and (saw_life or not episodic)


@pytest.mark.parametrize("mbr", [True, False])
def test_worker_processes_follow_the_collector_protocol(mbr):
    """two spawned workers (one per group) driven through the semaphore protocol for two batches: records and frames
    equal those of in-process HostAtariEnvs run through the reference collector rules (worker.py:25-113)"""
    ctx = mp.get_context("spawn")
    B, per, T = 4, 2, 14
    shape = (210, 160)
    fb = 210 * 160
    shared = dict(frames=ctx.RawArray(C.c_uint8, B * 2 * fb), ext=ctx.RawArray(C.c_uint8, B * W.EXT_DTYPE.itemsize),
                  act=ctx.RawArray(C.c_uint8, B), report=ctx.RawArray(C.c_int32, 2 * 2), arg=ctx.RawArray(C.c_int32, 1))
    frames, ext, act = W.views(shared, B, shape)
    cmd = ctx.Value("i", W.CMD_STEP, lock=False)
    ready = [ctx.Semaphore(0) for _ in range(2)]
    done = [ctx.Semaphore(0) for _ in range(2)]
    q = ctx.Queue()
    env_kwargs = dict(frame_skip=4, clip_reward=True, episodic_lives=True, max_start_noops=0, rgb=False)
    factory = partial(fake_ale.make, rules=RULES)
    procs = [ctx.Process(target=W.worker_main, daemon=True,
                         args=(w, w * per, (w + 1) * per, B, factory, env_kwargs, shape, shared, cmd, ready[w], done[w], q,
                               5 + w, mbr, 27000, 0.99)) for w in range(2)]
    for p in procs:
        p.start()
    try:
        ref = [HostAtariEnv(fake_ale.make(e, RULES), **env_kwargs) for e in range(B)]
        g1, g2 = np.zeros(shape, np.uint8), np.zeros(shape, np.uint8)
        for w in range(2):
            assert done[w].acquire(timeout=120)
        # what each worker reports before the first step: its emulators' action count, its warm-up length (none here)
        assert np.frombuffer(shared["report"], dtype=np.int32).tolist() == [4, 0, 4, 0]
        for e in range(B):
            assert ref[e].reset(g2) == ext[e]["flags"] == FLAG_RESET and np.array_equal(frames[e, 1], g2)
        rng = np.random.RandomState(0)
        lengths = [0] * B
        finished = []
        for batch in range(2):
            need = [False] * B
            for s in range(T):
                act[:] = rng.randint(0, 4, B).astype(np.uint8)
                cmd.value = W.CMD_STEP
                for w in range(2):
                    ready[w].release()
                for w in range(2):
                    assert done[w].acquire(timeout=120)
                for e in range(B):
                    x = ext[e]
                    if need[e]:
                        assert x["flags"] == FLAG_SKIP | FLAG_NO_RECORD
                        continue
                    r, raw, d, nr, fl = ref[e].step(int(act[e]), g1, g2)
                    lengths[e] += 1
                    if d and nr:
                        finished.append((e, lengths[e]))
                        lengths[e] = 0
                        if mbr:
                            fl = ref[e].reset(g2)
                        else:
                            need[e] = True
                            fl = FLAG_SKIP
                    rec = (x["reward"], x["raw_reward"], bool(x["done"]), bool(x["need_reset"]), x["flags"])
                    assert rec == (r, raw, bool(d), bool(nr), fl), (batch, s, e)
                    if not (fl & FLAG_SKIP):
                        assert np.array_equal(frames[e, 1], g2)
                        if not (fl & FLAG_RESET):
                            assert np.array_equal(frames[e, 0], g1)
            if not mbr:
                cmd.value = W.CMD_RESET_NEEDED
                for w in range(2):
                    ready[w].release()
                for w in range(2):
                    assert done[w].acquire(timeout=120)
                for e in range(B):
                    if need[e]:
                        assert ref[e].reset(g2) == ext[e]["flags"] == FLAG_RESET and np.array_equal(frames[e, 1], g2)
                    else:
                        assert ext[e]["flags"] == FLAG_SKIP | FLAG_NO_RECORD
        got = []
        while len(got) < len(finished):
            got.append(q.get(timeout=30))
        assert sorted((g[0], g[1]) for g in got) == sorted(finished) and len(finished) > 0
    finally:
        cmd.value = W.CMD_QUIT
        for w in range(2):
            ready[w].release()
        for p in procs:
            p.join(timeout=20)
            if p.is_alive():
                p.terminate()


def test_product_synthetic_emulator_follows_the_oracle_rules():
    """accel_rl_b200/hostsim/synth_emulator.py (host mirror of the device's synthetic emulator, used by
    bench.py --workload host_emulators) == oracle/synth_ale.py rule for rule"""
    from accel_rl_b200.hostsim import synth_emulator as se
    pool = synth_ale.make_pool(32, seed=0)
    for e in (0, 5, 77):
        em = se.SynthEmulator(e, RULES)
        buf = np.zeros((210, 160), np.uint8)
        for f in range(1, 150):
            assert em.act(0) == synth_ale.synth_reward(RULES, e, f)
            assert em.lives() == synth_ale.synth_lives(RULES, e, f)
            assert em.game_over() == (synth_ale.synth_lives(RULES, e, f) == 0)
            em.getScreenGrayscale(buf)
            assert np.array_equal(buf, pool[synth_ale.frame_index(RULES, e, f)])
        em.reset_game()
        assert em.f == 0


def test_game_mix_rules_product_emulator_vs_oracle():
    """BASELINE configs[2] "4-game mix": env e plays game e % 4 — own slice of the frame pool, own reward table and life
    clock.  The product's host emulator against the oracle's statement, and the properties that make it a mix."""
    from accel_rl_b200.hostsim import synth_emulator as se
    from accel_rl_b200.envs.atari_env import AtariEnv, GAME_MIXES, MINIMAL_ACTIONS
    rules = dict(RULES, pool_frames=32, n_games=4)
    pool = synth_ale.make_pool(32, seed=0)
    seen = {g: set() for g in range(4)}
    for e in range(12):
        em = se.SynthEmulator(e, rules)
        buf = np.zeros((210, 160), np.uint8)
        for f in range(1, 120):
            assert em.act(0) == synth_ale.synth_reward(rules, e, f)
            assert em.lives() == synth_ale.synth_lives(rules, e, f)
            em.getScreenGrayscale(buf)
            i = synth_ale.frame_index(rules, e, f)
            assert np.array_equal(buf, pool[i])
            seen[e % 4].add(i)
    for g in range(4):                                   # every game stays inside its own 8-frame slice, and uses all of it
        assert seen[g] == set(range(8 * g, 8 * g + 8))
    # life clocks differ by game: env 0 (game 0) vs env 4 (game 0) share the base; env 1 (game 1) is 17 frames longer
    base = lambda e: synth_ale.life_period(rules, e) - (e * rules["life_mul"]) % rules["life_mod"]
    assert base(4) == base(0) and base(1) == base(0) + 17 and base(3) == base(0) + 51
    # one game: the mix rules reduce to the plain ones
    assert all(synth_ale.frame_index(dict(rules, n_games=1), 3, f) == synth_ale.frame_index(RULES | dict(pool_frames=32), 3, f)
               for f in range(40))
    env = AtariEnv(game="mix4", max_start_noops=0, synth_rules=dict(pool_frames=32))
    assert env.synth_rules["n_games"] == 4
    assert env.action_space.n == max(MINIMAL_ACTIONS[g] for g in GAME_MIXES["mix4"]) == 9   # padded to the largest set
    with pytest.raises(ValueError):
        AtariEnv(game="mix4", max_start_noops=0, synth_rules=dict(pool_frames=30))


def test_profiling_worker_dumps_a_profile(tmp_path):
    """profile_pathname (sampler/base.py:32-43, sampler/util.py:10-19): the worker runs under cProfile and leaves
    <path>_sim_<rank>.prof when it quits"""
    import pstats
    ctx = mp.get_context("spawn")
    B, shape, fb = 2, (210, 160), 210 * 160
    shared = dict(frames=ctx.RawArray(C.c_uint8, B * 2 * fb), ext=ctx.RawArray(C.c_uint8, B * W.EXT_DTYPE.itemsize),
                  act=ctx.RawArray(C.c_uint8, B), report=ctx.RawArray(C.c_int32, 2 * 2), arg=ctx.RawArray(C.c_int32, 1))
    cmd = ctx.Value("i", W.CMD_STEP, lock=False)
    ready, done, q = ctx.Semaphore(0), ctx.Semaphore(0), ctx.Queue()
    env_kwargs = dict(frame_skip=4, clip_reward=True, episodic_lives=True, max_start_noops=0, rgb=False)
    path = str(tmp_path / "prof")
    p = ctx.Process(target=W.profiling_worker, daemon=True,
                    args=(path, 0, 0, B, B, partial(fake_ale.make, rules=RULES), env_kwargs, shape, shared, cmd, ready, done, q,
                          3, True, 27000, 0.99))
    p.start()
    try:
        assert done.acquire(timeout=120)
        for _ in range(3):
            cmd.value = W.CMD_STEP
            ready.release()
            assert done.acquire(timeout=120)
    finally:
        cmd.value = W.CMD_QUIT
        ready.release()
        p.join(timeout=30)
    assert p.exitcode == 0
    st = pstats.Stats(path + "_sim_0.prof")
    assert any("worker_main" in str(k) for k in st.stats)


@pytest.mark.skipif(not ref_harness.available(), reason="the reference tree only exists in the build container")
@pytest.mark.parametrize("episodic,noops", [(True, 0), (False, 0), (True, 7)])
def test_host_atari_env_vs_the_real_reference_atari_env(episodic, noops):
    """HostAtariEnv next to the REFERENCE'S OWN AtariEnv (accel_rl/envs/atari_env.py, imported unmodified through
    oracle/ref_harness.py with the synthetic ALE as `atari_py`): same emulator calls in the same order — checked through
    the emulator's frame counter — same rewards / dones / infos, and the exported frame pair + flag rebuild the
    reference's observation bit for bit (its cv2.resize included), with start no-ops drawn from the same numpy state."""
    pool = synth_ale.make_pool(32, seed=0)
    ref_harness.install(pool=pool, rules=RULES)
    env_mod = ref_harness.ref("accel_rl.envs.atari_env")
    for e in (0, 3):
        synth_ale.SynthALE.next_env_id = e
        np.random.seed(100 + e)
        renv = env_mod.AtariEnv(game="breakout", clip_reward=True, episodic_lives=episodic, max_start_noops=noops)
        assert renv.ale.env_id == e
        emu = fake_ale.make(e, RULES)
        host = HostAtariEnv(emu, clip_reward=True, episodic_lives=episodic, max_start_noops=noops)
        f1, f2 = np.zeros((210, 160), np.uint8), np.zeros((210, 160), np.uint8)
        np.random.seed(100 + e)                       # the reference ctor ends with reset(): same no-op draw
        obs = _apply(np.zeros((4, oframe.H, oframe.W), np.uint8), f1, f2, host.reset(f2))
        assert emu.f == renv.ale.f and np.array_equal(obs, renv.get_obs())
        rng = np.random.RandomState(e)
        overs = 0
        for k in range(150):
            a = int(rng.randint(0, 4))
            o2, r2, d2, info = renv.step(a)
            r, raw, d, nr, fl = host.step(a, f1, f2)
            obs = _apply(obs, f1, f2, fl)
            assert emu.f == renv.ale.f, k
            assert np.array_equal(obs, o2), k
            assert r == r2 and raw == info["raw_reward"] and bool(d) == bool(d2)
            assert (nr if episodic else None) == info.get("need_reset")
            if info.get("need_reset", d2):             # the collector's reset (worker.py:43-45)
                overs += 1
                np.random.seed(1000 + k)
                o3 = renv.reset()
                np.random.seed(1000 + k)
                obs = _apply(obs, f1, f2, host.reset(f2))
                assert emu.f == renv.ale.f and np.array_equal(obs, o3)
        assert overs > 0


@pytest.mark.parametrize("tag,mbr", [("reset", True), ("nonreset", False), ("overlength", True)])
def test_collector_replays_the_real_reference_samplers_buffers(golden_dir, tag, mbr):
    """The golden sampler fixtures were produced by the reference's REAL multi-process ActsrvAltOvrlpSampler
    (tests/golden/make_golden.py).  Feeding its recorded actions to hostsim's Collector, and doing on the CPU what the
    device does with the exported records and screens (ext_apply_kernel + frame_kernel, restated with oracle/frame.py),
    must reproduce its buffers: rewards, dones, raw rewards, need_reset, every observation row (CRC), the bootstrap
    observations and the completed trajectories — for the reset collector, the non-reset collector (envs that stop
    stepping, stale rows) and over-length cuts."""
    import os
    import zlib
    from tests.golden.make_golden import RULES as GRULES, POOL_FRAMES
    g = np.load(os.path.join(golden_dir, "sampler_%s.npz" % tag))
    B, T, itrs = int(g["n_envs"]), int(g["horizon"]), int(g["itrs"])
    frames = np.zeros((B, 2, 210, 160), np.uint8)
    ext = np.zeros(B, W.EXT_DTYPE)
    act = np.zeros(B, np.uint8)
    done_trajs = []
    rules = dict(GRULES)
    env_kwargs = dict(frame_skip=4, clip_reward=True, episodic_lives=True, max_start_noops=0, rgb=False)
    col = W.Collector(0, B, partial(fake_ale.make, rules=rules), env_kwargs, frames, ext, act, done_trajs.append, mbr,
                      int(g["max_path_length"]), 0.99)
    assert rules["pool_frames"] == POOL_FRAMES
    step_obs = np.zeros((B, 4, oframe.H, oframe.W), np.uint8)
    obs = np.zeros((B * T, 4, oframe.H, oframe.W), np.uint8)
    rew = np.zeros(B * T, np.float32); raw = np.zeros(B * T, np.float32)
    don = np.zeros(B * T, bool); nrs = np.zeros(B * T, bool)

    def ingest(s):
        for e in range(B):
            x = ext[e]
            if 0 <= s < T and not (x["flags"] & FLAG_NO_RECORD):
                rew[e * T + s], raw[e * T + s] = x["reward"], x["raw_reward"]
                don[e * T + s], nrs[e * T + s] = bool(x["done"]), bool(x["need_reset"])
            if not (x["flags"] & FLAG_SKIP):
                step_obs[e] = _apply(step_obs[e], frames[e, 0], frames[e, 1], int(x["flags"]))
                if 0 <= s and s + 1 < T:
                    obs[e * T + s + 1] = step_obs[e]
    col.start()
    ingest(-1)
    for itr in range(itrs):
        del done_trajs[:]
        for e in range(B):
            obs[e * T] = step_obs[e]
        for s in range(T):
            act[:] = g["act_%d" % itr][s::T]
            col.step()
            ingest(s)
        extra = step_obs.copy()
        if not mbr:
            col.reset_needed()
            ingest(T)
        for name, mine in (("rew", rew), ("done", don), ("raw", raw), ("nr", nrs), ("extra", extra)):
            assert np.array_equal(mine, g["%s_%d" % (name, itr)]), (name, itr)
        crc = np.array([zlib.crc32(r.tobytes()) for r in obs], dtype=np.uint32)
        assert np.array_equal(crc, g["obscrc_%d" % itr]), itr
        mine = sorted((t[1], float(t[2]), float(t[3]), int(t[4]), float(t[5])) for t in done_trajs)
        want = [tuple(r) for r in g["traj_%d" % itr]]
        assert len(mine) == len(want)
        for a, b in zip(mine, want):
            np.testing.assert_allclose(a, b, rtol=1e-6)


@pytest.mark.skipif(not ref_harness.available(), reason="the reference tree only exists in the build container")
@pytest.mark.parametrize("mpl", [27000, 7])
def test_decorrelated_start_vs_the_real_reference_start_envs(mpl):
    """start_envs with max_decorrelation_steps > 0: hostsim's Collector (start + warm_step) AND the oracle's
    OracleSampler.decorrelate next to the REFERENCE'S OWN start_envs (accel_rl/sampler/util.py:26-57, imported unmodified,
    its wall-clock fraction replaced by a fixed sequence): same first observations, same running TrajInfos."""
    pool = synth_ale.make_pool(32, seed=0)
    rules = dict(RULES, life_base=9, life_mod=5)
    ref_harness.install(pool=pool, rules=rules)
    env_mod = ref_harness.ref("accel_rl.envs.atari_env")
    util = ref_harness.ref("accel_rl.sampler.util")
    B, max_steps = 6, 50
    fracs = [0.0, 0.13, 0.5, 0.77, 0.99, 0.31]
    want_n = [int(f * max_steps) for f in fracs]
    # --- the reference ---
    synth_ale.SynthALE.next_env_id = 0
    renvs = [env_mod.AtariEnv(game="breakout", max_start_noops=0) for _ in range(B)]
    it = iter(fracs)
    saved = util.get_random_fraction
    util.get_random_fraction = lambda: next(it)
    try:
        robs, rinfos = util.start_envs(renvs, max_steps, mpl, 0.99, unique_ID=1)
    finally:
        util.get_random_fraction = saved
    # --- hostsim Collector: exported frames + flags, stacked on the CPU the way the device does ---
    frames = np.zeros((B, 2, 210, 160), np.uint8)
    ext = np.zeros(B, W.EXT_DTYPE)
    act = np.zeros(B, np.uint8)
    env_kwargs = dict(frame_skip=4, clip_reward=True, episodic_lives=True, max_start_noops=0, rgb=False)
    col = W.Collector(0, B, partial(fake_ale.make, rules=rules), env_kwargs, frames, ext, act, None, True, mpl, 0.99)
    step_obs = np.zeros((B, 4, oframe.H, oframe.W), np.uint8)

    def ingest():
        for e in range(B):
            if not (ext[e]["flags"] & FLAG_SKIP):
                step_obs[e] = _apply(step_obs[e], frames[e, 0], frames[e, 1], int(ext[e]["flags"]))
    np.random.seed(1)
    col.start(max_steps)
    col.warm_n = list(want_n)                        # (the worker draws these from its own stream)
    ingest()
    for k in range(max(want_n)):
        col.warm_step(k)
        ingest()
    for e in range(B):
        assert np.array_equal(step_obs[e], robs[e]), e
        assert col.trajs[e]["Length"] == rinfos[e].Length
    # --- the oracle restatement ---
    orc = osampler.OracleSampler(B, 4, pool, rules, 4, 0.99, max_path_length=mpl)
    orc.decorrelate(want_n)
    for e in range(B):
        assert np.array_equal(orc.step_obs[e], robs[e]), e
        t, r = orc.traj[e], rinfos[e]
        assert (t["Length"], t["NonzeroRewards"]) == (r.Length, r.NonzeroRewards)
        np.testing.assert_allclose([t["Return"], t["RawReturn"], t["DiscountedReturn"]],
                                   [r.Return, r.RawReturn, r.DiscountedReturn], rtol=1e-6)
