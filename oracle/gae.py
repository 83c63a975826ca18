"""ORACLE (test infrastructure).  Advantage estimation restated in numpy.

Follows accel_rl/algos/pg/util.py:6-23 (gen_adv_est), :26-37 (discount_returns), :40-46
(zero_after_reset), :56-63 (update_valids) and algos/pg/aac_base.py:108-145 (process_samples:
per-env segments, optional standardisation with population std + 1e-6).  Arithmetic in float64
like NumPy-1.x scalar promotion, stored float32.
"""
import numpy as np


def gen_adv_est(rewards, values, dones, last_value, discount, gae_lambda):
    T = len(rewards)
    not_done = 1.0 - dones.astype(np.float64)
    vpred = np.append(values.astype(np.float64), np.float64(last_value))
    adv = np.zeros(T, np.float32)
    lastgaelam = 0.0
    for t in reversed(range(T)):
        delta = float(rewards[t]) + discount * vpred[t + 1] * not_done[t] - vpred[t]
        lastgaelam = delta + discount * gae_lambda * not_done[t] * lastgaelam
        adv[t] = lastgaelam
    ret = (adv + values).astype(np.float32)
    return adv, ret


def discount_returns(rewards, dones, last_value, discount):
    T = len(rewards)
    out = np.zeros(T, np.float32)
    ret = float(last_value)
    for t in reversed(range(T)):
        ret = float(rewards[t]) if dones[t] else ret * discount + float(rewards[t])
        out[t] = ret
    return out


def process_samples(rewards, values, dones, need_reset, last_values, discount, gae_lambda, horizon,
                    use_valids=False, standardize_adv=False):
    """all inputs flat (N,), env-major; -> adv, ret, valids (or None), values' (zeroed after reset when valids)"""
    N = len(rewards)
    B = N // horizon
    adv = np.zeros(N, np.float32)
    ret = np.zeros(N, np.float32)
    values = values.copy()
    valids = np.zeros(N, np.int8) if use_valids else None
    for e in range(B):
        sl = slice(e * horizon, (e + 1) * horizon)
        if gae_lambda == 1:
            ret[sl] = discount_returns(rewards[sl], dones[sl], last_values[e], discount)
            adv[sl] = ret[sl] - values[sl]
        else:
            adv[sl], ret[sl] = gen_adv_est(rewards[sl], values[sl], dones[sl], last_values[e], discount, gae_lambda)
    if use_valids:
        for e in range(B):
            sl = slice(e * horizon, (e + 1) * horizon)
            nr = need_reset[sl]
            if nr.any():
                t_inv = int(np.min(np.where(nr))) + 1
                valids[sl][:t_inv] = 1
                valids[sl][t_inv:] = 0
                adv[sl][t_inv:] = 0
                ret[sl][t_inv:] = 0
                values[sl][t_inv:] = 0
            else:
                valids[sl] = 1
    if standardize_adv:
        if not use_valids:
            adv[:] = (adv - adv.mean()) / (adv.std() + 1e-6)
        else:
            idx = valids.nonzero()
            a = adv[idx]
            adv[idx] = (a - a.mean()) / (a.std() + 1e-6)
    return adv, ret, valids, values
