"""logger_context (reference: accel_rl/util/logging.py:20-43)."""
import json
import os
from contextlib import contextmanager

from accel_rl_b200.util import logger


@contextmanager
def logger_context(log_dir, name="run", run_ID=0, log_params=None, snapshot_mode="none", snapshot_gap=1):
    """same positional order as the reference: logger_context(log_dir, name, run_ID, log_params, snapshot_mode);
    writes <log_dir>/<name>_<run_ID>/{progress.csv, params.json (always, with name and run_ID added), snapshots}"""
    exp_dir = os.path.join(log_dir, "%s_%s" % (name, run_ID))
    logger.configure(exp_dir, snapshot_mode=snapshot_mode, snapshot_gap=snapshot_gap)
    log_params = dict(log_params or {}, name=name, run_ID=run_ID)
    os.makedirs(exp_dir, exist_ok=True)
    with open(os.path.join(exp_dir, "params.json"), "w") as f:
        json.dump({k: (v if isinstance(v, (int, float, str, bool, type(None))) else str(v))
                   for k, v in log_params.items()}, f, indent=1)
    try:
        yield
    finally:
        logger.configure(None)
