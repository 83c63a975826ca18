from accel_rl_b200.optimizers.async_.base import BaseAsyncOptimizer


class AsyncA2cOptimizer(BaseAsyncOptimizer):
    """reference: accel_rl/optimizers/async/async_a2c_optimizer.py:15-109"""
