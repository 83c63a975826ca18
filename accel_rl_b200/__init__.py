"""accel_rl_b200 — B200-native drop-in for the rollout-sampler + A2C/PPO path of astooke/accel_rl.

Python host classes mirror the reference's plugin surface (Sampler / RLAlgorithm / AtariCnnPolicy /
optimizers / runners); all computation runs in hand-written sm_100a CUDA behind
the C ABI of libaccelrl_b200.so (include/accelrl_b200.h).  PyTorch supplies device memory, streams
and torch.distributed bootstrap only.
"""
__version__ = "0.1.0"
