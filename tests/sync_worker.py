"""Worker for the 2-GPU synchronous data-parallel test (launched by torch.distributed.run; see
tests/test_gpu_multi.py).  Checks the fused P2P all-reduce + clip + Adam kernel against the reference semantics
(optimizers/sync/base.py:22-24, sync_ppo_optimizer.py:41-48, optimizers/util.py:63-76) restated with the oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from oracle import net as onet
    from tests.util_gpu import make_policy

    # ---- 1. kernel-level: fused allreduce + clip + Adam == oracle on the averaged gradient ----
    pol, flat, spec = make_policy(0, max_rows=8)
    eng = pol.engine

    def exchange(h):
        out = [None] * world
        dist.all_gather_object(out, h)
        return out
    eng.comm_init(rank, world, exchange)
    for kind, clip in (("adam", 0.5), ("rmsprop", None)):
        eng.set_params(flat)
        eng.opt_configure(algo=0, clip_param=0.2, v_loss_coeff=1.0, ent_loss_coeff=0.01, update=0 if kind == "adam" else 1,
                          learning_rate=1e-3, beta1=0.9, beta2=0.999, epsilon=1e-5 if kind == "adam" else 1e-6, rho=0.9,
                          grad_norm_clip=clip if clip else -1.0)
        eng.reset_opt_state()
        opt = onet.Adam(flat.size, 1e-3, epsilon=1e-5) if kind == "adam" else onet.RMSProp(flat.size, 1e-3)
        p = flat.copy()
        for step in range(3):
            g = (np.random.RandomState(100 * rank + step).randn(flat.size) * 0.01).astype(np.float32)
            eng.grad.copy_(torch.tensor(g))
            gsum = torch.tensor(g).cuda()
            dist.all_reduce(gsum)
            torch.cuda.synchronize()
            dist.barrier()
            eng.sync_allreduce_update()
            torch.cuda.synchronize()
            g_avg = (gsum / world).cpu().numpy()
            gc, norm = onet.total_norm_clip(g_avg, clip)
            p = opt.step(p, gc)
            losses, norms = eng.read_logs()
            assert abs(norms[0] - norm) <= 1e-5 * norm, (norms, norm)
            got = eng.get_params()
            np.testing.assert_allclose(got, p, rtol=2e-6, atol=2e-7)
            # every rank holds identical parameters (bitwise)
            mine = eng.params.clone()
            ref = mine.clone()
            dist.broadcast(ref, src=0)
            assert torch.equal(mine, ref), "parameters differ across ranks"
        assert eng.device_error() == 0
    eng.close()

    # ---- 2. path-level: AccelRLSync + mPPO for a few iterations ----
    from accel_rl_b200.algos import mPPO
    from accel_rl_b200.envs import AtariEnv
    from accel_rl_b200.policies import AtariCnnPolicy, cnn_specs
    from accel_rl_b200.runners import AccelRLSync
    from accel_rl_b200.sampler import ActsrvAltOvrlpSampler
    from accel_rl_b200.util import logger
    logger.configure(None, quiet=True)
    rules = dict(pool_frames=128, life_base=24, life_mod=11, reward_mod=7)
    sampler = ActsrvAltOvrlpSampler(EnvCls=AtariEnv, env_args=dict(game="breakout", max_start_noops=0, synth_rules=rules),
                                    horizon=16, n_parallel=4, envs_per=4, max_decorrelation_steps=0)
    algo = mPPO(optimizer_args=dict(minibatch_size=128, epochs=2))
    policy = AtariCnnPolicy(**cnn_specs[1])
    runner = AccelRLSync(algo=algo, policy=policy, sampler=sampler, n_steps=32 * 16 * 4 * world, seed=3,
                         affinities=[dict(gpu=i) for i in range(world)], log_interval_steps=32 * 16 * 2 * world)
    n_itr = runner.startup()
    p0 = policy.get_param_values()
    for itr in range(3):
        samples, traj = runner.sampler.obtain_samples(itr)
        opt_data, info = runner.algo.optimize_policy(itr, samples)
        assert len(info["GradNorm"]) == 8 and np.isfinite(info["GradNorm"]).all()
    torch.cuda.synchronize()
    mine = policy.engine.params.clone()
    ref = mine.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(mine, ref), "parameters diverged across ranks"
    assert np.linalg.norm(policy.get_param_values() - p0) > 0
    # ranks sample different data (seed + 100*rank): rewards/actions differ
    a = samples.actions.clone()
    b = a.clone()
    dist.broadcast(b, src=0)
    if rank == 1:
        assert not torch.equal(a, b)
    import hashlib
    digest = hashlib.sha256(np.ascontiguousarray(policy.get_param_values()).tobytes()).hexdigest()
    norms = [float(x) for x in info["GradNorm"]]
    assert policy.engine.device_error() == 0
    policy.engine.close()
    dist.barrier()
    if rank == 0:
        import json
        print("SYNC_DIGEST " + json.dumps(dict(params=digest, norms=norms)))
        print("SYNC_OK world=%d" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
