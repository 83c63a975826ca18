"""ORACLE (test infrastructure).  One optimize_policy call restated on the CPU.

Follows accel_rl/algos/pg/aac_base.py:102-170 (process_samples + prep_opt_inputs + optimizer.optimize),
optimizers/single/ppo_optimizer.py:57-76, single/a2c_optimizer.py:45, sync: sync_ppo_optimizer.py:56-72
with optimizers/util.py:63-67 (average of per-rank gradients).
"""
import numpy as np
import torch

from oracle import gae as ogae, net as onet


def optimize_policy(flat, opt, buf, spec, n_actions, horizon, algo, rng, discount=0.99, gae_lambda=None,
                    v_coeff=None, ent_coeff=0.01, clip_param=0.2, lr_mult=1.0, epochs=4, minibatch_size=512,
                    grad_norm_clip="default", use_valids=False, standardize_adv=False, emulate_bf16=True,
                    last_values=None):
    """buf: dict of numpy arrays with the samples_buf keys.  -> (flat', losses, grad_norms, opt_data)"""
    gae_lambda = (0.95 if algo == "ppo" else 1.0) if gae_lambda is None else gae_lambda
    v_coeff = (1.0 if algo == "ppo" else 0.25) if v_coeff is None else v_coeff
    if grad_norm_clip == "default":
        grad_norm_clip = None if algo == "ppo" else 0.5
    if last_values is None:
        _, lv = onet.forward(torch.tensor(flat), torch.tensor(buf["extra_observations"]), spec, n_actions, emulate_bf16)
        last_values = lv.numpy()
    adv, ret, valids, values = ogae.process_samples(buf["rewards"], buf["value"], buf["dones"], buf["need_reset"],
                                                    last_values, discount, gae_lambda, horizon, use_valids,
                                                    standardize_adv)
    obs, act, oldp = buf["observations"], buf["actions"], buf["prob"]
    losses, norms = [], []
    kw = dict(emulate_bf16=emulate_bf16, v_coeff=v_coeff, ent_coeff=ent_coeff)
    if algo == "ppo":
        for _ in range(epochs):
            for idx in onet.iterate_mb_idxs(minibatch_size, len(obs), rng):
                loss, g, _ = onet.loss_and_grad(flat, obs[idx], act[idx], adv[idx], ret[idx], oldp[idx], spec, n_actions,
                                                "ppo", clip_param=clip_param, lr_mult=lr_mult,
                                                valids=None if valids is None else valids[idx], **kw)
                g, norm = onet.total_norm_clip(g.astype(np.float32), grad_norm_clip)
                flat = opt.step(flat, g, lr_mult)
                losses.append(loss); norms.append(norm)
    else:
        loss, g, _ = onet.loss_and_grad(flat, obs, act, adv, ret, oldp, spec, n_actions, "a2c", valids=valids, **kw)
        g, norm = onet.total_norm_clip(g.astype(np.float32), grad_norm_clip)
        flat = opt.step(flat, g, lr_mult)
        losses.append(loss); norms.append(norm)
    return flat, losses, norms, dict(advantages=adv, returns=ret, valids=valids, values=values, last_values=last_values)
