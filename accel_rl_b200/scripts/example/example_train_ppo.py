"""PPO / A2C on one GPU (reference: accel_rl/scripts/example/example_train_ppo.py, example_train_a2c.py).

    python -m accel_rl_b200.scripts.example.example_train_ppo LOG_DIR GAME RUN_ID [--algo a2c] [--n-envs 64] [--n-steps 1e6]

The objects are the reference's, built the same way; only the imports differ.  The reference's launcher
hands a run-slot / affinity code to its examples (scripts/launching, out of scope here): `--gpu` selects the device."""
import argparse

from accel_rl_b200.algos import A2C, PPO
from accel_rl_b200.envs import AtariEnv
from accel_rl_b200.policies import AtariCnnPolicy, cnn_specs
from accel_rl_b200.runners import AccelRL
from accel_rl_b200.sampler import ActsrvAltOvrlpSampler
from accel_rl_b200.util.logging import logger_context


def build_and_run(log_dir, game, run_ID, algo="ppo", learning_rate=None, n_envs=64, n_steps=1e6, gpu=0, cnn_spec=1,
                  n_sim_cores=8, log_interval_steps=1e5):
    env_args = dict(game=game, clip_reward=True, max_start_noops=30, episodic_lives=True)
    assert n_envs % (n_sim_cores * 2) == 0
    ppo = algo == "ppo"
    sampler = ActsrvAltOvrlpSampler(
        EnvCls=AtariEnv, env_args=env_args,
        horizon=128 if ppo else 5,
        n_parallel=n_sim_cores, envs_per=n_envs // (n_sim_cores * 2),
        mid_batch_reset=ppo,                       # example_train_a2c.py:50 runs without mid-batch reset (validity mask)
        max_path_length=int(27e3))
    assert sampler.total_n_envs == n_envs
    optimizer_args = dict() if learning_rate is None else dict(learning_rate=float(learning_rate))
    algo_obj = PPO(optimizer_args=optimizer_args) if ppo else A2C(optimizer_args=optimizer_args)
    policy = AtariCnnPolicy(**cnn_specs[cnn_spec])
    runner = AccelRL(algo=algo_obj, policy=policy, sampler=sampler, n_steps=n_steps, log_interval_steps=log_interval_steps,
                     affinities=dict(gpu=gpu), seed=None, use_gpu=True)
    log_params = dict(exp="basic_" + algo, cnn_spec=cnn_spec, n_envs=n_envs, learning_rate=learning_rate)
    with logger_context(log_dir, game, run_ID, log_params):
        runner.train()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("log_dir")
    ap.add_argument("game")
    ap.add_argument("run_ID")
    ap.add_argument("--algo", default="ppo", choices=["ppo", "a2c"])
    ap.add_argument("--learning-rate", type=float, default=None)
    ap.add_argument("--n-envs", type=int, default=64)
    ap.add_argument("--n-steps", type=float, default=1e6)
    ap.add_argument("--gpu", type=int, default=0)
    a = ap.parse_args()
    build_and_run(a.log_dir, a.game, a.run_ID, a.algo, a.learning_rate, a.n_envs, a.n_steps, a.gpu)
