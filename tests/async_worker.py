"""Worker for the asynchronous data-parallel test (1 process on a single GPU, or 2 under torch.distributed.run; see
tests/test_gpu_multi.py / tests/test_gpu_async.py).  Checks async_push_pull_kernel — central (p, m, v) store in rank
0's HBM, chunk locks, per-learner Adam t — against the reference semantics (optimizers/async/base.py:59-104,
chunked_updates.py:53-120) restated in oracle/learner.py:async_push."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from oracle import learner as olearner
    from tests.util_gpu import make_policy

    def exchange(h):
        if world == 1:
            return [h]
        out = [None] * world
        dist.all_gather_object(out, h)
        return out

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for kind, clip, chunks in (("adam", 0.5, 3), ("rmsprop", None, 1), ("adam", None, 40)):
        pol, flat, spec = make_policy(0, max_rows=8)
        eng = pol.engine
        eng.set_params(flat)
        eps = 1e-5 if kind == "adam" else 1e-6
        eng.opt_configure(algo=1, clip_param=0.2, v_loss_coeff=0.25, ent_loss_coeff=0.01, update=0 if kind == "adam" else 1,
                          learning_rate=7e-4, beta1=0.9, beta2=0.999, epsilon=eps, rho=0.9,
                          grad_norm_clip=clip if clip else -1.0)
        eng.reset_opt_state()
        regions = eng.async_init(rank, world, chunks, exchange)
        assert regions >= chunks
        central = dict(p=flat.copy(), m=np.zeros_like(flat), v=np.zeros_like(flat))
        t_local = [0] * world
        # ---- deterministic order: ranks take turns (barrier between pushes) ----
        for step in range(3):
            for r in range(world):
                g = (np.random.RandomState(1000 * r + step).randn(flat.size) * 0.01).astype(np.float32)
                t_local[r] += 1
                want_local, norm = olearner.async_push(central, g, kind, t_local[r], 7e-4, clip=clip, epsilon=eps)
                if r == rank:
                    eng.grad.copy_(torch.tensor(g))
                    eng.async_push_pull()
                    torch.cuda.synchronize()
                    losses, norms = eng.read_logs()
                    assert abs(norms[0] - norm) <= 1e-5 * norm, (norms, norm)
                    np.testing.assert_allclose(eng.get_params(), want_local, rtol=2e-6, atol=2e-7)
                barrier()
            np.testing.assert_allclose(eng.async_read_central(0), central["p"], rtol=2e-6, atol=2e-7)
            np.testing.assert_allclose(eng.async_read_central(2), central["v"], rtol=2e-5, atol=1e-11)
            if kind == "adam":
                np.testing.assert_allclose(eng.async_read_central(1), central["m"], rtol=2e-5, atol=2e-9)   # FMA contraction on cancelling terms
            barrier()
        # ---- concurrent pushes (no ordering): must terminate, stay finite, leave every lock free ----
        for step in range(10):
            g = (np.random.RandomState(7 * rank + step).randn(flat.size) * 0.01).astype(np.float32)
            eng.grad.copy_(torch.tensor(g))
            eng.async_push_pull()
        barrier()
        cp = eng.async_read_central(0)
        assert np.isfinite(cp).all() and eng.device_error() == 0
        barrier()
        # a learner that pushes last holds exactly the central parameters
        if rank == 0:
            eng.grad.zero_()
            eng.async_push_pull()
            torch.cuda.synchronize()
            if kind == "rmsprop":   # zero gradient leaves p unchanged under RMSProp: local == central
                np.testing.assert_array_equal(eng.get_params(), eng.async_read_central(0))
        barrier()
        eng.close()

    # ---- path level: AccelRLAsync + mA3C / mAPPO for a few iterations ----
    from accel_rl_b200.algos import mA3C, mAPPO
    from accel_rl_b200.envs import AtariEnv
    from accel_rl_b200.policies import AtariCnnPolicy, cnn_specs
    from accel_rl_b200.runners import AccelRLAsync
    from accel_rl_b200.sampler import ActsrvAltOvrlpSampler
    from accel_rl_b200.util import logger
    logger.configure(None, quiet=True)
    rules = dict(pool_frames=128, life_base=24, life_mod=11, reward_mod=7)
    for Algo, args in ((mA3C, dict()), (mAPPO, dict(optimizer_args=dict(minibatch_size=128, epochs=2, n_update_chunks=4)))):
        sampler = ActsrvAltOvrlpSampler(EnvCls=AtariEnv, env_args=dict(game="breakout", max_start_noops=0, synth_rules=rules),
                                        horizon=16, n_parallel=4, envs_per=4, max_decorrelation_steps=0)
        algo = Algo(**args)
        policy = AtariCnnPolicy(**cnn_specs[1])
        runner = AccelRLAsync(algo=algo, policy=policy, sampler=sampler, n_steps=32 * 16 * 4 * world, seed=3,
                              affinities=[dict(gpu=i) for i in range(world)], log_interval_steps=32 * 16 * 2 * world)
        runner.startup()
        assert runner.parallelism_tag == "asynchronous" and algo.optimizer.parallelism_tag == "asynchronous"
        p0 = policy.get_param_values()
        for itr in range(3):
            samples, traj = runner.sampler.obtain_samples(itr)
            opt_data, info = runner.algo.optimize_policy(itr, samples)
            assert np.isfinite(info["GradNorm"]).all()
        barrier()
        assert np.linalg.norm(policy.get_param_values() - p0) > 0
        assert np.isfinite(algo.optimizer.central_shared_params).all()
        assert policy.engine.device_error() == 0
        barrier()
        policy.engine.close()
    # ---- ActsrvAltOvrlpPollSampler (poll_sampler.py:6-56): the policy is refreshed from the central store every
    #      poll_horizon rollout steps; with a refresh the rollout acts on the central parameters from that step on ----
    from accel_rl_b200.sampler import ActsrvAltOvrlpPollSampler
    T, ph = 12, 4
    smp = ActsrvAltOvrlpPollSampler(poll_horizon=ph, EnvCls=AtariEnv, env_args=dict(game="breakout", max_start_noops=0, synth_rules=rules),
                                    horizon=T, n_parallel=2, envs_per=2, max_decorrelation_steps=0)
    algo = mAPPO(optimizer_args=dict(minibatch_size=32, epochs=1, n_update_chunks=3))
    policy = AtariCnnPolicy(**cnn_specs[1])
    runner = AccelRLAsync(algo=algo, policy=policy, sampler=smp, n_steps=8 * T * 4 * world, seed=5,
                          affinities=[dict(gpu=i) for i in range(world)], log_interval_steps=8 * T * 2 * world)
    runner.startup()
    eng = policy.engine
    samples, _ = smp.obtain_samples(0)
    assert smp.n_polls == T // ph
    runner.algo.optimize_policy(0, samples)
    barrier()
    torch.cuda.synchronize()
    # make the local parameters stale on purpose; the central store keeps the trained ones
    central = eng.async_read_central(0).copy()
    stale = central + np.float32(0.05) * np.random.RandomState(3).randn(central.size).astype(np.float32)
    eng.set_params(stale)
    samples, _ = smp.obtain_samples(1)
    torch.cuda.synchronize()
    barrier()
    from oracle import net as onet
    obs = samples.observations.cpu().numpy()
    prob = samples.agent_infos.prob.cpu().numpy()
    B = smp.total_n_envs
    rows = lambda s: np.arange(B) * T + s
    if world == 1:     # (with other learners pushing meanwhile the central vector keeps moving)
        p_stale, _ = onet.forward(torch.tensor(stale), torch.tensor(obs[rows(0)]), onet.CNN_SPECS[1], 4, True)
        p_fresh, _ = onet.forward(torch.tensor(central), torch.tensor(obs[rows(ph - 1)]), onet.CNN_SPECS[1], 4, True)
        np.testing.assert_allclose(prob[rows(0)], p_stale.numpy(), rtol=2e-3, atol=2e-5)          # before the first poll
        np.testing.assert_allclose(prob[rows(ph - 1)], p_fresh.numpy(), rtol=2e-3, atol=2e-5)     # from step ph-1 on
        np.testing.assert_array_equal(eng.get_params(), central)
    assert eng.device_error() == 0
    barrier()
    eng.close()
    if rank == 0:
        print("ASYNC_OK world=%d" % world)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
