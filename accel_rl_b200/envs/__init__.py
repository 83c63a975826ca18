from accel_rl_b200.envs.atari_env import AtariEnv, EnvSpec
