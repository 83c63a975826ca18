from accel_rl_b200.algos.base import RLAlgorithm
from accel_rl_b200.algos.pg.a2c import A2C, mA2C, mA3C
from accel_rl_b200.algos.pg.ppo import PPO, mPPO, mAPPO
