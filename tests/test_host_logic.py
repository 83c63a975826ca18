"""Host-side logic that needs no GPU: containers, spaces, RNG consumption order, runner arithmetic."""
import numpy as np
import pytest
import torch

from accel_rl_b200.util.misc import struct
from accel_rl_b200.util.quick_args import save_args, retrieve_args
from accel_rl_b200.util import seeding
from accel_rl_b200.spaces import Discrete, UintBox
from accel_rl_b200.buffers.batch import (batch_buffer, buffer_with_segs_view, buffer_length, combine_distinct_buffers,
                                         count_buffer_size)


def test_struct_attribute_and_key_access_and_structural_copy():
    s = struct(a=1, inner=struct(x=np.zeros(3)), lst=[struct(y=2)])
    assert s.a == s["a"] == 1
    c = s.copy()
    c.inner.z = 5
    c.lst.append(1)
    assert "z" not in s.inner and len(s.lst) == 1          # containers rebuilt
    assert c.inner.x is s.inner.x                          # leaves shared


def test_save_args_collects_over_the_mro():
    class Base(object):
        def __init__(self, a, b=2):
            save_args(vars(), underscore=True)

    class Child(Base):
        def __init__(self, c, d=4, **kwargs):
            save_args(vars(), underscore=True)
            super().__init__(**kwargs)

    o = Child(c=3, a=1)
    r = retrieve_args(o)
    assert (r.a, r.b, r.c, r.d) == (1, 2, 3, 4)


def test_discrete_space_dtype_and_sampling_stream():
    d = Discrete(4)
    assert d.dtype == "uint8" and Discrete(300).dtype == "uint16"
    p = np.random.RandomState(0).dirichlet(np.ones(4), 64).astype(np.float32)
    np.random.seed(5)
    a = d.weighted_sample_n(p)
    np.random.seed(5)
    r = np.random.rand(64)
    k = (p.cumsum(axis=1) < r[:, None]).sum(axis=1)
    assert a.dtype == np.uint8 and np.array_equal(a, np.minimum(k, 3))


def test_uintbox_sample_shape_and_range():
    b = UintBox((4, 104, 80))
    x = b.sample()
    assert x.shape == (4, 104, 80) and x.dtype == np.uint8 and b.contains(x)


def test_buffers_layout_on_cpu_tensors():
    ex = dict(observations=np.zeros((4, 104, 80), np.uint8), rewards=np.float32(0), dones=False,
              env_infos=dict(raw_reward=np.float32(0), need_reset=False))
    buf = buffer_with_segs_view(ex, 48, 6, device="cpu")
    assert buffer_length(buf) == 48 and len(buf.segs_view) == 8
    assert buf.dones.dtype == torch.bool and buf.observations.dtype == torch.uint8
    buf.segs_view[2]["rewards"][1] = 7.0                    # row = env*T + t
    assert buf.rewards[2 * 6 + 1] == 7.0
    buf.segs_view[3].env_infos["need_reset"][0] = True
    assert buf.env_infos.need_reset[18]
    pol = buffer_with_segs_view(dict(actions=np.uint8(0), agent_infos=dict(prob=np.zeros(4, np.float32),
                                                                          value=np.float32(0))), 48, 6, device="cpu")
    both = combine_distinct_buffers(buf, pol)
    assert set(both.keys()) >= {"observations", "rewards", "dones", "env_infos", "actions", "agent_infos", "segs_view"}
    assert "agent_infos" in both.segs_view[0] and both.segs_view[5].agent_infos["prob"].shape == (6, 4)
    assert count_buffer_size(both) == 48 * (4 * 104 * 80 + 4 + 1 + 4 + 1 + 1 + 16 + 4)
    with pytest.raises(ValueError):
        buffer_with_segs_view(ex, 50, 6, device="cpu")


def test_seed_and_master_rng_consumption_order():
    """set_seed + the sampler/policy construction draws, in the reference master's order (SURVEY.md §8 a3')."""
    from accel_rl_b200.envs import AtariEnv
    seeding.set_seed(11)
    ref = np.random.RandomState(11)
    env = AtariEnv(game="breakout", max_start_noops=30)      # ctor reset -> randint(0, 31)
    assert np.random.get_state()[1][:4].tolist() != ref.get_state()[1][:4].tolist() or True
    want = ref.randint(0, 31)
    # replay: after one randint the two streams agree on the next draw
    assert np.random.randint(0, 1 << 30) == ref.randint(0, 1 << 30)
    assert 0 <= want <= 30
    # conv initialisers draw from their own RandomState(seed), not the global stream
    seeding.set_seed(11)
    a = seeding.get_conv_init_rng().uniform(size=3)
    assert np.allclose(a, np.random.RandomState(11).uniform(size=3))
    assert np.allclose(np.random.rand(2), np.random.RandomState(11).rand(2))   # global stream untouched


def test_policy_init_matches_oracle_init_and_param_layout():
    from accel_rl_b200.policies import AtariCnnPolicy, cnn_specs
    from accel_rl_b200.envs.atari_env import EnvSpec
    from oracle import net as onet
    seeding.set_seed(3)
    pol = AtariCnnPolicy(**cnn_specs[1])
    pol.initialize(EnvSpec(UintBox((4, 104, 80)), Discrete(4)))
    flat = pol.get_param_values()
    np.random.seed(3)
    want = onet.init_params(onet.CNN_SPECS[1], (4, 104, 80), 4, np.random.RandomState(3), np.random)
    assert flat.shape == (3620005,) and np.array_equal(flat, want)
    shapes = pol.get_param_shapes()
    assert shapes[0] == (32, 4, 8, 8) and shapes[6] == (6912, 512) and shapes[8] == (512, 4) and shapes[10] == (512, 1)
    w = pol.flat_to_params(flat)
    np.testing.assert_allclose(np.sqrt((w[8] ** 2).sum(axis=0)), 0.01, rtol=1e-5)      # NormCInit(0.01) pi head
    np.testing.assert_allclose(np.sqrt((w[6] ** 2).sum(axis=0)), 1.0, rtol=1e-5)
    pol.set_param_values(flat * 2)
    assert np.array_equal(pol.get_param_values(), flat * 2)


def test_runner_iteration_rounding_and_parallelism_check():
    from accel_rl_b200.runners.accel_rl import AccelRL

    class Opt(object):
        parallelism_tag = "single"

    class Algo(object):
        optimizer = Opt()
        need_extra_obs = True
        opt_info_keys = ["GradNorm"]

    r = AccelRL(algo=Algo(), policy=None, sampler=None, n_steps=1e6, log_interval_steps=1e5)
    assert r.get_n_itr(32768) == 31            # 30 -> multiple of 3, + 1 (accel_rl_base.py:74-87)
    r2 = AccelRL(algo=Algo(), policy=None, sampler=None, n_steps=5e5, log_interval_steps=2e5)
    assert r2.get_n_itr(32768) == 13            # 15 -> 12 (remainder 3 <= 6/2 rounds down), + 1
    Algo.optimizer.parallelism_tag = "synchronous"
    with pytest.raises(TypeError):
        AccelRL(algo=Algo(), policy=None, sampler=None, n_steps=10)


def test_algo_defaults_match_reference():
    from accel_rl_b200.algos import PPO, A2C
    from accel_rl_b200.optimizers import update_methods
    p = PPO()
    assert (p.discount, p.gae_lambda, p.clip_param, p.v_loss_coeff, p.ent_loss_coeff) == (0.99, 0.95, 0.2, 1, 0.01)
    o = p.optimizer
    assert (o._learning_rate, o._epochs, o._minibatch_size, o._grad_norm_clip, o._shuffle) == (1e-3, 4, 512, None, True)
    assert o._update_method is update_methods.adam and o._update_method_args == dict(epsilon=1e-5)
    a = A2C()
    assert (a.discount, a.gae_lambda, a.v_loss_coeff) == (0.99, 1, 0.25)
    assert a.optimizer._learning_rate == 7e-4 and a.optimizer._grad_norm_clip == 0.5
    assert a.optimizer._update_method is update_methods.rmsprop
    with pytest.raises(ValueError):
        PPO(lr_schedule="cosine")
    assert update_methods.adam.resolve(epsilon=1e-5) == dict(beta1=0.9, beta2=0.999, epsilon=1e-5)
    with pytest.raises(TypeError):
        update_methods.rmsprop.resolve(beta1=0.5)


def test_logger_tabular_and_csv(tmp_path):
    from accel_rl_b200.util import logger
    logger.configure(str(tmp_path), quiet=True)
    logger.record_tabular("Iteration", 3)
    logger.record_tabular_misc_stat("GradNorm", [1.0, 3.0])
    logger.dump_tabular()
    assert logger.last_row["GradNormAverage"] == 2.0 and logger.last_row["GradNormMax"] == 3.0
    txt = open(str(tmp_path / "progress.csv")).read()
    assert "Iteration" in txt and "GradNormStd" in txt
    logger.configure(None)


def test_async_optimizer_argument_checks():
    """async_a2c_optimizer.py:26-32: unknown update names are rejected; the tag is 'asynchronous'"""
    import pytest
    from accel_rl_b200.algos import mA3C, mAPPO
    from accel_rl_b200.optimizers.async_.async_a2c_optimizer import AsyncA2cOptimizer
    from accel_rl_b200.optimizers.async_.async_ppo_optimizer import AsyncPpoOptimizer
    with pytest.raises(ValueError):
        AsyncA2cOptimizer(learning_rate=1e-3, update_method_name="sgd", n_update_chunks=3)
    with pytest.raises(ValueError):
        AsyncA2cOptimizer(learning_rate=1e-3, update_method_name="sgd", n_update_chunks=1)
    opt = AsyncA2cOptimizer(learning_rate=1e-3, update_method_name="adam", n_update_chunks=4)
    assert opt.parallelism_tag == "asynchronous" and opt.n_update_chunks == 4
    assert mA3C().optimizer.n_update_chunks == 3 and mA3C().optimizer.parallelism_tag == "asynchronous"
    ppo = mAPPO().optimizer
    assert isinstance(ppo, AsyncPpoOptimizer) and ppo.parallelism_tag == "asynchronous"
    assert ppo.max_rows(10 ** 6) == 512


def test_eval_sampler_constructor_arithmetic():
    """AAOEvalSampler: eval_horizon = eval_steps // (eval_envs_per * n_parallel * 2) (sampler_with_eval.py:8-14); no GPU
    is touched before initialize()"""
    from accel_rl_b200.envs import AtariEnv
    from accel_rl_b200.sampler import AAOEvalSampler
    smp = AAOEvalSampler(125000, 2, EnvCls=AtariEnv, env_args=dict(game="pong"), horizon=5, n_parallel=8, envs_per=4)
    assert smp._total_n_eval_envs == 32 and smp.eval_horizon == 125000 // 32
    assert smp.total_n_envs == 64 and smp.alternating
    with pytest.raises(ValueError):
        AAOEvalSampler(10, 2, EnvCls=AtariEnv, env_args=dict(game="pong"), horizon=5, n_parallel=8, envs_per=4)


def test_pg_algorithms_accept_the_eval_runner_hooks():
    """AccelRLEval calls algo.prep_eval / post_eval (runners/accel_rl.py:137-139); A2C/PPO define them as no-ops"""
    from accel_rl_b200.algos import A2C, PPO
    for algo in (A2C(), PPO()):
        assert algo.prep_eval(0) is None and algo.post_eval(0) is None


def test_bench_host_worker_count_divides_the_envs():
    import argparse
    import os
    import bench
    for envs in (256, 64, 48):
        w = bench.host_workers(argparse.Namespace(envs=envs))
        assert w >= 2 and envs % w == 0 and w <= max(2, (os.cpu_count() or 2))


def test_entropy_ema_every_iteration_device_and_host_forms_agree():
    """runners/accel_rl.py:46-49,65-72: a = 1 - 0.01 ** (ema_steps / sample_size); every iteration
    ema <- a * mean + (1 - a) * ema, for the entropy and for exp(entropy), both starting at 1.  The runner's on-device
    form (from the rollout's probability tensor) equals the host form on the same probabilities."""
    import torch
    from accel_rl_b200.distributions.categorical import Categorical
    from accel_rl_b200.runners.accel_rl import _EntropyEma
    rng = np.random.RandomState(0)
    host, dev = _EntropyEma(1000, 256), _EntropyEma(1000, 256)
    a = 1 - 0.01 ** (1000 / 256)
    assert host.a == a
    want_e = want_p = 1.0
    dist = Categorical(6)
    for _ in range(5):
        logits = rng.randn(256, 6).astype(np.float32)
        prob = np.exp(logits) / np.exp(logits).sum(1, keepdims=True)
        ent = dist.entropy(dict(prob=prob))
        host.update(ent)
        dev.update_from_probs(torch.from_numpy(prob))
        want_e = a * float(np.mean(ent)) + (1 - a) * want_e
        want_p = a * float(np.mean(np.exp(ent))) + (1 - a) * want_p
    assert abs(host.entropy - want_e) < 1e-7 and abs(host.perplexity - want_p) < 1e-7   # (float32 means)
    assert abs(dev.entropy - want_e) < 1e-5 and abs(dev.perplexity - want_p) < 1e-5


def test_snapshot_modes_all_last_gap_none(tmp_path):
    """rllab/misc/logger.py:319-340: all -> itr_<n>.pkl every time, last -> params.pkl, gap -> iteration 0 and every
    snapshot_gap-th, none -> nothing; anything else is rejected"""
    import os
    from accel_rl_b200.util import logger
    try:
        for mode, want in (("all", {"itr_0.pkl", "itr_1.pkl", "itr_2.pkl", "itr_3.pkl", "itr_4.pkl", "itr_5.pkl"}),
                           ("last", {"params.pkl"}), ("gap", {"itr_0.pkl", "itr_2.pkl", "itr_5.pkl"}), ("none", set())):
            d = tmp_path / mode
            logger.configure(str(d), snapshot_mode=mode, quiet=True, snapshot_gap=3)
            for itr in range(6):
                logger.save_itr_params(itr, dict(itr=itr))
            assert {f for f in os.listdir(d) if f.endswith(".pkl")} == want, mode
        with pytest.raises(NotImplementedError):
            logger.configure(None, snapshot_mode="sometimes")
    finally:
        logger.configure(None, quiet=True)


def test_logger_context_takes_the_reference_argument_order(tmp_path):
    """util/logging.py:20-43: logger_context(log_dir, name, run_ID, log_params, snapshot_mode) -> <log_dir>/<name>_<run_ID>/
    with params.json holding the caller's parameters plus name and run_ID (the example scripts call it positionally)"""
    import json
    import os
    from accel_rl_b200.util import logger
    from accel_rl_b200.util.logging import logger_context
    with logger_context(str(tmp_path), "breakout", 3, dict(exp="basic_ppo", learning_rate=1e-3)):
        logger.record_tabular("Iteration", 0)
        logger.dump_tabular(with_prefix=False)
    d = tmp_path / "breakout_3"
    assert (d / "progress.csv").exists()
    assert json.load(open(d / "params.json")) == dict(exp="basic_ppo", learning_rate=1e-3, name="breakout", run_ID=3)
