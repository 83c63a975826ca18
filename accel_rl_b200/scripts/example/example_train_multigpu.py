"""Multi-GPU learners from ONE script (reference: accel_rl/scripts/example/example_train_mppo.py, example_train_ma2c.py —
synchronous; example_train_mappo.py, example_train_a3c.py — asynchronous).

    python -m accel_rl_b200.scripts.example.example_train_multigpu LOG_DIR GAME RUN_ID --gpus 0,1,2,3 [--algo mppo|ma2c|mappo|ma3c]

The runner forks one full runner per GPU itself, like the reference's launch_workers (runners/multigpu_rl_base.py:20-45),
so this is a plain `python` command; the same script also runs under `torchrun --nproc-per-node N` (one process per GPU,
nothing forked).  `n_envs` is per learner, as in the reference examples."""
import argparse

from accel_rl_b200.algos import mA2C, mA3C, mAPPO, mPPO
from accel_rl_b200.envs import AtariEnv
from accel_rl_b200.policies import AtariCnnPolicy, cnn_specs
from accel_rl_b200.runners import AccelRLAsync, AccelRLSync
from accel_rl_b200.sampler import ActsrvAltOvrlpPollSampler, ActsrvAltOvrlpSampler
from accel_rl_b200.util.logging import logger_context

ALGOS = dict(mppo=(mPPO, AccelRLSync, 128), ma2c=(mA2C, AccelRLSync, 5), mappo=(mAPPO, AccelRLAsync, 128),
             ma3c=(mA3C, AccelRLAsync, 5))


def build_and_run(log_dir, game, run_ID, gpus, algo="mppo", learning_rate=None, n_envs=64, n_steps=1e6, cnn_spec=1,
                  n_sim_cores=8, poll_horizon=0, log_interval_steps=1e5):
    Algo, Runner, horizon = ALGOS[algo]
    env_args = dict(game=game, clip_reward=True, max_start_noops=30, episodic_lives=True)
    assert n_envs % (n_sim_cores * 2) == 0
    sampler_args = dict(EnvCls=AtariEnv, env_args=env_args, horizon=horizon, n_parallel=n_sim_cores,
                        envs_per=n_envs // (n_sim_cores * 2), mid_batch_reset=True, max_path_length=int(27e3))
    if poll_horizon and Runner is AccelRLAsync:
        # the sampling policy is refreshed from the central parameters every poll_horizon steps (poll_sampler.py:6-56)
        sampler = ActsrvAltOvrlpPollSampler(poll_horizon=poll_horizon, **sampler_args)
    else:
        sampler = ActsrvAltOvrlpSampler(**sampler_args)
    optimizer_args = dict() if learning_rate is None else dict(learning_rate=float(learning_rate))
    runner = Runner(algo=Algo(optimizer_args=optimizer_args), policy=AtariCnnPolicy(**cnn_specs[cnn_spec]), sampler=sampler,
                    n_steps=n_steps, log_interval_steps=log_interval_steps, affinities=[dict(gpu=g) for g in gpus], seed=None, use_gpu=True)
    log_params = dict(exp="basic_" + algo, cnn_spec=cnn_spec, n_envs=n_envs, learning_rate=learning_rate, n_gpus=len(gpus))
    with logger_context(log_dir, game, run_ID, log_params):
        runner.train()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("log_dir")
    ap.add_argument("game")
    ap.add_argument("run_ID")
    ap.add_argument("--gpus", default="0,1")
    ap.add_argument("--algo", default="mppo", choices=sorted(ALGOS))
    ap.add_argument("--learning-rate", type=float, default=None)
    ap.add_argument("--n-envs", type=int, default=64)
    ap.add_argument("--n-steps", type=float, default=1e6)
    ap.add_argument("--poll-horizon", type=int, default=0)
    a = ap.parse_args()
    build_and_run(a.log_dir, a.game, a.run_ID, [int(g) for g in a.gpus.split(",")], a.algo, a.learning_rate, a.n_envs,
                  a.n_steps, poll_horizon=a.poll_horizon)
