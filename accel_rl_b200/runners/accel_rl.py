"""Single-GPU training loops.

AccelRL      learning curve from the TRAINING trajectories (reference: accel_rl/runners/accel_rl.py:12-105)
AccelRLEval  learning curve from separate evaluation episodes run every `eval_interval_steps`
             (reference: runners/accel_rl.py:108-180; needs a sampler with evaluate_policy, e.g. AAOEvalSampler)

Both loops are `sample -> optimize -> book-keeping`; what differs is which trajectories feed the log.  The columns
written to progress.csv keep the reference's names (the visualizer and downstream scripts key on them); they are listed
once in `_Columns` instead of being spelled out statement by statement.
"""
import time
from collections import deque

import numpy as np

from accel_rl_b200.runners.accel_rl_base import AccelRLBase
from accel_rl_b200.util import logger


class _Columns(object):
    """progress.csv column names of the two runners"""
    ONLINE = ("Iteration", "CumCompletedTrajs", "CumCompletedSteps", "CumTotalSteps", "NewCompletedTrajs",
              "StepsInTrajWindow")
    ENTROPY = ("Entropy", "Perplexity")
    ONLINE_TIME = ("CumTime (s)", "SamplesPerSecond")
    EVAL = ("Iteration", "CumCompletedSteps", "StepsInEval", "TrajsInEval")
    EVAL_TIME = ("CumTrainTime", "CumEvalTime", "CumTotalTime", "SamplesPerSecond")


def _tabulate(names, values):
    for name, value in zip(names, values):
        logger.record_tabular(name, value)


class _EntropyEma(object):
    """exponential moving averages of the policy's entropy and perplexity over the sampled states; the time constant is
    `ema_steps` env-steps: a = 1 - 0.01 ** (ema_steps / sample_size)  (runners/accel_rl.py:46-49, 65-72)"""

    def __init__(self, ema_steps, sample_size):
        self.a = 1 - 0.01 ** (ema_steps / sample_size)
        self._state = [1., 1.]            # entropy, perplexity; becomes a 2-element device tensor on the first device update

    def update(self, entropies):
        self._fold(np.mean(entropies), np.mean(np.exp(entropies)))

    def update_from_probs(self, prob):
        """every iteration, like the reference — from the rollout's probabilities where they already are (HBM): the two
        means and the EMA stay on the device, nothing is copied to the host until a log line reads them"""
        import torch
        ent = -(prob * torch.log(prob + 1e-8)).sum(dim=1)           # distributions/categorical.py TINY
        cur = torch.stack((ent.mean(), ent.exp().mean())).double()
        if not torch.is_tensor(self._state):
            self._state = torch.tensor(self._state, dtype=torch.float64, device=prob.device)
        self._state = self._state + self.a * (cur - self._state)

    def _fold(self, entropy, perplexity):
        e, p = self.entropy, self.perplexity
        self._state = [e + self.a * (entropy - e), p + self.a * (perplexity - p)]

    entropy = property(lambda self: float(self._state[0]))
    perplexity = property(lambda self: float(self._state[1]))


class _TrainLoop(AccelRLBase):
    """sample -> optimize -> store, with two hooks: before_itr (evaluation runner) and after_itr (online runner)"""

    def train(self):
        n_itr = self.startup()
        for itr in range(self._start_itr, n_itr):
            with logger.prefix("itr #%d | " % itr):
                self.before_itr(itr)
                samples_data, traj_infos = self.sampler.obtain_samples(itr)
                opt_data, opt_infos = self.algo.optimize_policy(itr, samples_data)
                self.store_diagnostics(itr, samples_data, opt_data, traj_infos, opt_infos)
                self.after_itr(itr)
        self.shutdown()

    def before_itr(self, itr):
        pass

    def after_itr(self, itr):
        pass

    def _store_opt_infos(self, opt_infos):
        for key, val in opt_infos.items():
            self._opt_infos[key].extend(val if isinstance(val, list) else [val])

    def _announce(self):
        logger.log("optimizing over {} iterations".format(self._log_interval_itrs))


class AccelRL(_TrainLoop):
    """Runs RL; tracks performance online using learning trajectories"""

    def __init__(self, log_interval_steps=1e5, log_traj_window=100, log_ema_steps=None, **kwargs):
        super().__init__(**kwargs)
        self._log_steps = int(log_interval_steps)
        self._log_traj_window = int(log_traj_window)
        self._log_ema_steps = self._log_steps if log_ema_steps is None else int(log_ema_steps)

    def init_logging(self):
        self._traj_infos = deque(maxlen=self._log_traj_window)
        self._cum_completed_steps = self._cum_completed_trajs = self._new_completed_trajs = 0
        self._ema = None
        if hasattr(self.policy, "distribution"):
            self._ema = _EntropyEma(self._log_ema_steps, self._sample_size)
        self._announce()
        super().init_logging()

    def _logging_itr(self, itr):
        return (itr + 1) % self._log_interval_itrs == 0

    def store_diagnostics(self, itr, samples_data, opt_data, traj_infos, opt_infos):
        self._cum_completed_trajs += len(traj_infos)
        self._new_completed_trajs += len(traj_infos)
        self._cum_completed_steps += sum(info["Length"] for info in traj_infos)
        self._traj_infos.extend(traj_infos)
        self._store_opt_infos(opt_infos)
        # every iteration's rollout entropy is folded into the EMA (accel_rl.py:65-72)
        if self._ema is not None:
            prob = samples_data.agent_infos["prob"]
            if hasattr(prob, "is_cuda"):
                self._ema.update_from_probs(prob)
            else:
                self._ema.update(self.policy.distribution.entropy(samples_data.agent_infos))

    def after_itr(self, itr):
        if self._logging_itr(itr):
            self.log_diagnostics(itr)

    def log_diagnostics(self, itr):
        self.save_itr_snapshot(itr)
        _tabulate(_Columns.ONLINE, (itr, self._cum_completed_trajs, self._cum_completed_steps,
                                    (itr + 1) * self._sample_size, self._new_completed_trajs,
                                    sum(info["Length"] for info in self._traj_infos)))
        if self._ema is not None:
            _tabulate(_Columns.ENTROPY, (self._ema.entropy, self._ema.perplexity))
        self._log_infos()
        now = time.time()
        _tabulate(_Columns.ONLINE_TIME, (now - self._start_time,
                                         self._log_interval_itrs * self._sample_size / (now - self._last_time)))
        self._last_time = now
        logger.dump_tabular(with_prefix=False)
        self._new_completed_trajs = 0
        if itr < self._n_itr - 1:
            self._announce()


class AccelRLEval(_TrainLoop):
    """Runs RL; tracks learning performance offline using evaluation trajectories"""

    def __init__(self, eval_interval_steps=1e6, **kwargs):
        super().__init__(**kwargs)
        self._log_steps = int(eval_interval_steps)

    def init_logging(self):
        self._cum = dict(train=0., eval=0., total=0.)
        super().init_logging()

    def before_itr(self, itr):
        if itr % self._log_interval_itrs == 0:
            eval_traj_infos, eval_time = self.eval_policy(itr)
            self.log_diagnostics(itr, eval_traj_infos, eval_time)

    def eval_policy(self, itr):
        logger.log("evaluating policy...")
        t0 = time.time()
        self.algo.prep_eval(itr)
        traj_infos = self.sampler.evaluate_policy(itr)
        self.algo.post_eval(itr)
        logger.log("evaluation run complete")
        return traj_infos, time.time() - t0

    def store_diagnostics(self, itr, samples_data, opt_data, traj_infos, opt_infos):
        self._store_opt_infos(opt_infos)

    def log_diagnostics(self, itr, eval_traj_infos, eval_time):
        self.save_itr_snapshot(itr)
        if not eval_traj_infos:
            logger.log("ERROR: had no complete trajectories in eval.")
        _tabulate(_Columns.EVAL, (itr, itr * self._sample_size, sum(info["Length"] for info in eval_traj_infos),
                                  len(eval_traj_infos)))
        self._log_infos(eval_traj_infos)
        now = time.time()
        interval = now - self._last_time
        self._last_time = now
        train_time = interval - eval_time
        for key, dt in (("train", train_time), ("eval", eval_time), ("total", interval)):
            self._cum[key] += dt
        speed = float("nan") if itr == 0 else self._log_interval_itrs * self._sample_size / train_time
        _tabulate(_Columns.EVAL_TIME, (self._cum["train"], self._cum["eval"], self._cum["total"], speed))
        logger.dump_tabular(with_prefix=False)
        self._announce()
