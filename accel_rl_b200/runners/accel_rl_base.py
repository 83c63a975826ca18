"""AccelRLBase (reference: accel_rl/runners/accel_rl_base.py:15-150): startup order, n_itr rounding,
snapshot and log plumbing are the reference's; `init_policy` selects the CUDA device of this
process instead of calling theano.gpuarray.use (accel_rl_base.py:61-72)."""
import time

import numpy as np
import torch

from accel_rl_b200.runners.base import Runner
from accel_rl_b200.util import logger
from accel_rl_b200.util.misc import make_seed, nbytes_unit
from accel_rl_b200.util.quick_args import save_args
from accel_rl_b200.util.seeding import set_seed


class AccelRLBase(Runner):
    def __init__(self, algo, policy, sampler, n_steps, seed=None, affinities=None, use_gpu=True, resume_from=None):
        # resume_from (not in the reference, whose resume path is AtariCnnPolicy(initial_param_values=...) only):
        # a snapshot dict or the path of one written by save_itr_snapshot — restores parameters, optimizer state
        # (m, v, update count) and the iteration counter (so a linear lr schedule continues where it stopped)
        n_steps = int(n_steps)
        save_args(vars(), underscore=False)
        self._start_itr = 0
        if affinities is None:
            self.affinities = dict()
        if algo.optimizer.parallelism_tag != self.parallelism_tag:
            raise TypeError("Had mismatched parallelism between Runner ({}) and Optimizer: {}".format(
                self.parallelism_tag, algo.optimizer.parallelism_tag))

    def startup(self, master=True):
        if self.seed is None:
            self.seed = make_seed()
        set_seed(self.seed)
        self.select_device()
        env_spec, sample_size, horizon, mid_batch_reset = self.sampler.initialize(
            seed=self.seed + 1,
            affinities=self.affinities,
            discount=getattr(self.algo, "discount", None),
            need_extra_obs=self.algo.need_extra_obs,
        )
        self.init_policy(env_spec)
        self.algo.initialize(policy=self.policy, env_spec=env_spec, sample_size=sample_size, horizon=horizon,
                             mid_batch_reset=mid_batch_reset)
        self.sampler.policy_init(self.policy)
        if self.resume_from is not None:
            self.load_snapshot(self.resume_from)
        if master:
            n_itr = self.get_n_itr(sample_size)
            self.algo.set_n_itr(n_itr)
            self.init_logging()
            return n_itr

    def select_device(self):
        if not self.use_gpu:
            raise RuntimeError("accel_rl_b200 has no CPU path: use_gpu must be True")
        aff = self.affinities if isinstance(self.affinities, dict) else {}
        gpu = aff.get("gpu", None)
        if gpu is not None and gpu != "":
            torch.cuda.set_device(int(gpu))

    def init_policy(self, env_spec):
        self.policy.initialize(env_spec)
        flat_params = self.policy.get_param_values(trainable=True)
        logger.log("Policy trainable params -- number: {:,}   size: {:,.1f} {}".format(
            flat_params.size, *nbytes_unit(flat_params.nbytes)))

    def get_n_itr(self, sample_size):
        """iterations to run: n_steps / sample_size rounded to the nearest whole number of log intervals (ties round
        down), plus one so the last interval is logged (accel_rl_base.py:84-97)"""
        self._sample_size = sample_size
        interval = self._log_interval_itrs = max(self._log_steps // sample_size, 1)
        whole, rem = divmod(max(self.n_steps // sample_size, 1), interval)
        self._n_itr = n_itr = interval * (whole + (1 if 2 * rem > interval else 0)) + 1
        logger.log("Iterations to run: {}".format(n_itr))
        return n_itr

    def init_logging(self):
        self._opt_infos = {k: list() for k in self.algo.opt_info_keys}
        self._initial_param_vector = self.policy.get_param_values()
        self._start_time = self._last_time = time.time()

    def shutdown(self):
        logger.log("Training complete.")
        self.sampler.shutdown()

    def get_itr_snapshot(self, itr):
        # itr, cum_samples, policy_param_values: the reference's snapshot (accel_rl_base.py:108-113);
        # optimizer_state: what a bit-exact resume additionally needs
        snap = dict(itr=itr, cum_samples=itr * self._sample_size, policy_param_values=self.policy.get_param_values())
        try:
            snap["optimizer_state"] = self.algo.optimizer.get_state()
        except NotImplementedError:
            pass
        return snap

    def load_snapshot(self, snapshot):
        """restore parameters (+ optimizer state when the snapshot has it); training continues at snapshot itr + 1"""
        if isinstance(snapshot, str):
            import joblib
            snapshot = joblib.load(snapshot)
        self.policy.set_param_values(snapshot["policy_param_values"])
        if "optimizer_state" in snapshot:
            self.algo.optimizer.set_state(snapshot["optimizer_state"])
        self._start_itr = int(snapshot["itr"]) + 1
        logger.log("Resumed from the snapshot of iteration {}".format(snapshot["itr"]))

    def save_itr_snapshot(self, itr):
        logger.save_itr_params(itr, self.get_itr_snapshot(itr))

    def _log_infos(self, traj_infos=None):
        """Average/Std/Median/Min/Max of every public TrajInfo field and of every optimizer diagnostic collected since
        the last log, then the parameter norms (accel_rl_base.py:122-146)"""
        traj_infos = self._traj_infos if traj_infos is None else traj_infos
        fields = [k for k in (traj_infos[0] if traj_infos else ()) if not k.startswith("_") and k != "env"]
        for k in fields:
            logger.record_tabular_misc_stat(k, [info[k] for info in traj_infos])
        for k, values in (self._opt_infos or {}).items():
            logger.record_tabular_misc_stat(k, values)
            self._opt_infos[k] = list()
        params = self.policy.get_param_values()
        for name, vec in (("ParamsNorm", params), ("NormFromInit", params - self._initial_param_vector)):
            logger.record_tabular(name, np.sqrt(np.sum(vec ** 2)))

    @property
    def parallelism_tag(self):
        return "single"
