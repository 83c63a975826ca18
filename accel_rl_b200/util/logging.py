"""logger_context (reference: accel_rl/util/logging.py:20-43)."""
import json
import os
from contextlib import contextmanager

from accel_rl_b200.util import logger


@contextmanager
def logger_context(log_dir, run_ID=0, name="run", log_params=None, snapshot_mode="none", snapshot_gap=1):
    exp_dir = os.path.join(log_dir, "%s_%s" % (name, run_ID))
    logger.configure(exp_dir, snapshot_mode=snapshot_mode, snapshot_gap=snapshot_gap)
    if log_params is not None:
        os.makedirs(exp_dir, exist_ok=True)
        with open(os.path.join(exp_dir, "params.json"), "w") as f:
            json.dump({k: (v if isinstance(v, (int, float, str, bool, type(None))) else str(v))
                       for k, v in log_params.items()}, f, indent=1)
    try:
        yield
    finally:
        logger.configure(None)
