from accel_rl_b200.optimizers.async_.base import BaseAsyncOptimizer


class AsyncPpoOptimizer(BaseAsyncOptimizer):
    """reference: accel_rl/optimizers/async/async_ppo_optimizer.py (a SyntaxError upstream)"""
