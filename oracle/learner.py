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


def async_push(central, grad, kind, t_local, lr, clip=None, beta1=0.9, beta2=0.999, epsilon=None, rho=0.9):
    """One asynchronous learner update against the central store (test oracle).

    reference: optimizers/async/async_a2c_optimizer.py:43-52,96-100 (local clip, push every chunk, copy back) and
    optimizers/async/chunked_updates.py:53-76 (rmsprop chunk), :79-120 (adam chunk, per-process t).
    central: dict(p=, m=, v=) float32 arrays updated in place; returns (new local params, pre-clip grad norm)."""
    g, norm = onet.total_norm_clip(np.asarray(grad, np.float32), clip)
    g = g.astype(np.float32)
    f = np.float32
    if kind == "adam":
        eps = f(1e-8 if epsilon is None else epsilon)
        a_t = f(lr * np.sqrt(1.0 - float(beta2) ** t_local) / (1.0 - float(beta1) ** t_local))
        central["m"][:] = f(beta1) * central["m"] + (f(1) - f(beta1)) * g
        central["v"][:] = f(beta2) * central["v"] + (f(1) - f(beta2)) * g * g
        central["p"][:] = central["p"] - a_t * central["m"] / (np.sqrt(central["v"]) + eps)
    else:
        eps = f(1e-6 if epsilon is None else epsilon)
        central["v"][:] = f(rho) * central["v"] + (f(1) - f(rho)) * g * g
        central["p"][:] = central["p"] - f(lr) * g / np.sqrt(central["v"] + eps)
    return central["p"].copy(), norm
