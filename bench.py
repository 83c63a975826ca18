#!/usr/bin/env python
"""Benchmark of the hot path: env-steps/s of full PPO iterations (rollout + GAE + 4-epoch update).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one PPO iteration over one batch of synthetic emulator frames: 256 envs/GPU x 128-step
rollout (policy forward + action sampling + env step + frame pipeline per step), bootstrap value +
GAE, then 4 epochs x 64 shuffled minibatches of 512 (forward, losses, backward, clip, Adam) —
BASELINE.json configs[1] (PPO Breakout-shaped, 256 envs, Nature-CNN preset 1 @ (4,104,80), bf16
tensor-core operands / fp32 accumulate).  `value` keeps the frame pool resident in HBM; `e2e` feeds
the raw frames of every step from pinned host memory through the same public API
(sampler.obtain_samples / algo.optimize_policy) with the H2D/D2H copies inside the timed region.
`--impl reference` times the CPU port of the reference path (oracle/) on the host cores.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# per-sample MACs of preset 1 @ (4,104,80), A=4
MACS = {"conv0": 3891200, "conv1": 3538944, "conv2": 3981312, "fc": 3538944}
# --frames rgb (north-star mode): classic Nature-CNN @ (4,84,84), no padding: 20x20x32, 9x9x64, 7x7x64, FC 3136->512
MACS_NATURE84 = {"conv0": 3276800, "conv1": 2654208, "conv2": 1806336, "fc": 1605632}
NATURE84_SPEC = dict(conv_filter_sizes=[8, 4, 3], conv_filters=[32, 64, 64], conv_strides=[4, 2, 1],
                     conv_pads=[(0, 0), (0, 0), (0, 0)], hidden_sizes=[512])
# per-sample FLOPs: (fwd, train, conv-only fwd, conv-only train)
FLOPS = {"gray": (29.906e6, 81.936e6, 22.823e6, 60.686e6), "rgb": (18.69e6, 49.52e6, 15.475e6, 39.870e6)}
FLOP_PER_ENV_STEP_PPO = 357.9e6     # 1 act-fwd + 1/128 bootstrap fwd + 4 x train (81.936 MFLOP)
FLOP_PER_ENV_STEP_CONV = 265.7e6    # conv tiles only: the north-star's "conv-tile roofline" numerator


def ncu_traffic(label, args=None):
    """DRAM bytes per launch of `label` from the committed ncu --set full capture (None when not captured, or when this run's
    launch shapes are not the captured ones: minibatch 512, preset 1, reference frames)"""
    if args is not None and (getattr(args, "algo", "ppo") != "ppo" or args.minibatch != 512 or args.spec != 1 or
                             getattr(args, "frames", "gray") != "gray"):
        return None
    for name in ("r2_traffic.json", "r1e_traffic.json"):
        try:
            return json.load(open(os.path.join(ROOT, "profiles", name)))["kernels"].get(label)
        except Exception:
            continue
    return None


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=256)
    ap.add_argument("--horizon", type=int, default=128)
    ap.add_argument("--spec", type=int, default=1)
    ap.add_argument("--minibatch", type=int, default=512)
    ap.add_argument("--epochs", type=int, default=4)
    ap.add_argument("--pool-frames", type=int, default=4096)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-steps", type=int, default=8, help="rollout steps in the bounded CPU sample")
    ap.add_argument("--algo", default="ppo", choices=["ppo", "a2c"],
                    help="a2c: BASELINE configs[2]-style A2C (one full-batch RMSProp step per rollout; use --envs 1024 --horizon 5)")
    ap.add_argument("--parallelism", default="sync", choices=["sync", "async"],
                    help="multi-GPU learner: sync (configs[3], fused all-reduce) or async (configs[4], central store + chunk locks)")
    ap.add_argument("--game", default="breakout",
                    help="names the action count of the synthetic envs (ALE minimal action sets: breakout 4, pong / "
                         "space_invaders 6, beam_rider 9, seaquest 18); BASELINE configs[4] is space_invaders; mix4 = "
                         "BASELINE configs[2]'s 4-game mix (env e plays game e % 4, actions padded to 9)")
    ap.add_argument("--poll-horizon", type=int, default=0,
                    help="async learners: ActsrvAltOvrlpPollSampler refreshing the policy from the central store every this "
                         "many rollout steps (0: plain sampler)")
    ap.add_argument("--frames", default="gray", choices=["gray", "rgb"],
                    help="gray: the reference's pipeline, 210x160 grayscale screens -> (4,104,80), cnn preset 1 (default, parity "
                         "pinned); rgb: north-star mode, 210x160x3 RGB screens -> gray -> (4,84,84), classic Nature-CNN")
    ap.add_argument("--workload", default="ppo", choices=["ppo", "frame_sweep", "host_emulators"],
                    help="ppo: the headline metric (default); frame_sweep: BASELINE configs[2] frame-kernel HBM GB/s sweep; "
                         "host_emulators: the same PPO iteration with the emulators stepped by worker processes on the "
                         "host cores (HostEmulatorSampler: raw screens over PCIe, two alternating groups)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        self.stop_flag = True
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = []
        for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5), ("sw_power_cap", 6)):
            if any(r[col].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "power_w_max": max(float(r[2]) for r in self.rows), "samples": len(self.rows)}


# =============================================================================================
# our arm
# =============================================================================================
def build_runner(args, frame_feed, rank, world):
    import torch
    from accel_rl_b200.algos import PPO, mPPO
    from accel_rl_b200.envs import AtariEnv
    from accel_rl_b200.policies import AtariCnnPolicy, cnn_specs
    from accel_rl_b200.runners import AccelRL, AccelRLSync
    from accel_rl_b200.sampler import ActsrvAltOvrlpSampler
    from accel_rl_b200.util import logger
    logger.configure(None, quiet=True)
    rules = dict(pool_frames=args.pool_frames)
    rgb = getattr(args, "frames", "gray") == "rgb"
    env_args = dict(game=getattr(args, "game", "breakout"), max_start_noops=0, synth_rules=rules,
                    frame_mode="rgb" if rgb else "gray")
    if frame_feed == "workers":
        from functools import partial
        from accel_rl_b200.hostsim import synth_emulator
        from accel_rl_b200.sampler import HostEmulatorSampler
        n_par = host_workers(args) // 2
        sampler = HostEmulatorSampler(
            emu_factory=partial(synth_emulator.make, rules=rules, channels=3 if rgb else 1), EnvCls=AtariEnv, env_args=env_args,
            horizon=args.horizon, n_parallel=n_par, envs_per=args.envs // (2 * n_par), max_path_length=27000,
            mid_batch_reset=True, max_decorrelation_steps=0)
    elif getattr(args, "poll_horizon", 0) > 0 and getattr(args, "parallelism", "sync") == "async":
        from accel_rl_b200.sampler import ActsrvAltOvrlpPollSampler
        sampler = ActsrvAltOvrlpPollSampler(
            poll_horizon=args.poll_horizon, EnvCls=AtariEnv, env_args=env_args,
            horizon=args.horizon, n_parallel=args.envs // 8, envs_per=4, max_path_length=27000, mid_batch_reset=True,
            max_decorrelation_steps=0, frame_feed=frame_feed)
    else:
        sampler = ActsrvAltOvrlpSampler(
            EnvCls=AtariEnv, env_args=env_args,
            horizon=args.horizon, n_parallel=args.envs // 8, envs_per=4, max_path_length=27000, mid_batch_reset=True,
            max_decorrelation_steps=0, frame_feed=frame_feed)
    from accel_rl_b200.algos import A2C, mA2C, mA3C, mAPPO
    from accel_rl_b200.runners import AccelRLAsync
    a2c = getattr(args, "algo", "ppo") == "a2c"
    asyn = getattr(args, "parallelism", "sync") == "async"
    opt_args = dict() if a2c else dict(minibatch_size=args.minibatch, epochs=args.epochs)
    policy = AtariCnnPolicy(**(NATURE84_SPEC if rgb else cnn_specs[args.spec]))
    n_steps = args.envs * args.horizon * 10 ** 6
    if world > 1 or asyn:
        Algo = (mA3C if a2c else mAPPO) if asyn else (mA2C if a2c else mPPO)
        Runner = AccelRLAsync if asyn else AccelRLSync
        runner = Runner(algo=Algo(optimizer_args=opt_args), policy=policy, sampler=sampler, n_steps=n_steps, seed=0,
                        affinities=[dict(gpu=torch.cuda.current_device())] * world, log_interval_steps=10 ** 12)
    else:
        runner = AccelRL(algo=(A2C if a2c else PPO)(optimizer_args=opt_args), policy=policy, sampler=sampler,
                         n_steps=n_steps, seed=0, affinities=dict(), log_interval_steps=10 ** 12)
    runner.startup()
    return runner


def host_workers(args):
    """worker processes of the host-emulator workload: the largest power of two <= host cores that divides the envs"""
    cores = os.cpu_count() or 2
    w = 2
    while w * 2 <= cores and args.envs % (w * 2) == 0:
        w *= 2
    return w


def run_host_emulators(args):
    """PPO iterations with the emulators in host worker processes (one JSON line)"""
    import torch
    torch.cuda.set_device(0)
    args.pool_frames = min(args.pool_frames, 256)      # every worker process builds its own copy of the frame pool
    runner = build_runner(args, "workers", 0, 1)
    smp = runner.sampler
    try:
        N = args.envs * args.horizon
        h0, d0 = smp.h2d_bytes, smp.d2h_bytes
        steps, warm = max(2, args.steps // 2), max(1, args.warmup // 2)
        clocks = ClockSampler(0)
        clocks.start()
        ms, launches, wall, _, itr = timed_iterations(runner, steps, warm, 1)
        clk = clocks.summary()
        phases = phase_split(runner, itr, reps=2)
        n_it = steps + warm + 2
        out = {"metric": "env-steps/sec (PPO Atari, %d envs/GPU)" % args.envs, "value": round(N / (ms * 1e-3), 1),
               "unit": "env-steps/s", "n_gpus": 1, "steps": steps, "warmup": warm, "ms_per_step": round(ms, 3),
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
               "config": {"workload": "PPO Breakout-shaped, %d envs x %d-step rollout, %d epochs x mb %d; emulators (host mirror "
                                      "of the synthetic emulator, Python) stepped by %d worker processes in two alternating "
                                      "groups on %d host cores, raw %s screens H2D from pinned shared memory every step"
                                      % (args.envs, args.horizon, args.epochs, args.minibatch, host_workers(args),
                                         os.cpu_count() or 1, "210x160x3" if args.frames == "rgb" else "210x160"),
                          "frames": args.frames, "workers": host_workers(args)},
               "clocks": clk, "gpu_launches": int(launches), "phases": phases,
               "h2d_bytes_per_step": int((smp.h2d_bytes - h0) / n_it), "d2h_bytes_per_step": int((smp.d2h_bytes - d0) / n_it)}
    finally:
        smp.shutdown()
        runner.policy.engine.close()
    _emit(out)


def timed_iterations(runner, steps, warmup, world, itr0=0):
    """-> (ms per step on the device: max over ranks, launches, host wall seconds)"""
    import torch
    import torch.distributed as dist
    eng = runner.policy.engine
    itr = itr0
    for _ in range(warmup):
        s, _ = runner.sampler.obtain_samples(itr)
        runner.algo.optimize_policy(itr, s)
        itr += 1
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = eng.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    ev0.record()
    last = None
    for _ in range(steps):
        s, _ = runner.sampler.obtain_samples(itr)
        _, info = runner.algo.optimize_policy(itr, s)      # reads losses / grad norms back (D2H)
        last = info
        itr += 1
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    wall = time.time() - t0
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms / steps, eng.launches - l0, wall, last, itr


def phase_split(runner, itr0, reps=3):
    """device time of the two halves of a step (CUDA events on the launching stream)"""
    import torch
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    roll, train = 0.0, 0.0
    itr = itr0
    for _ in range(reps):
        ev[0].record()
        s, _ = runner.sampler.obtain_samples(itr)
        ev[1].record()
        runner.algo.optimize_policy(itr, s)
        ev[2].record()
        torch.cuda.synchronize()
        roll += ev[0].elapsed_time(ev[1]); train += ev[1].elapsed_time(ev[2])
        itr += 1
    return {"rollout_ms": round(roll / reps, 3), "gae_update_ms": round(train / reps, 3)}


def kernel_breakdown(runner, args):
    """Device time of every kernel of one minibatch update and one rollout step, measured with CUDA events recorded
    INSIDE a replayed CUDA graph (same launch order and data residency as the timed region, no host gaps)
    -> {label: (ms per launch, launches per PPO iteration)}"""
    import torch
    eng = runner.policy.engine
    N = args.envs * args.horizon
    a2c = getattr(args, "algo", "ppo") == "a2c"
    mb = N if a2c else args.minibatch                   # A2C: one full-batch step per iteration
    n_mb = 1 if a2c else (N // args.minibatch) * args.epochs
    # (the profiling pass replays the minibatch a few times with an advancing cursor: 8 index slices, any valid rows)
    idx = torch.cat([torch.randperm(N, device="cuda")[:mb] for _ in range(8)]).to(torch.int32).contiguous()
    out = {}
    labels, ms = eng.profile_graph(0, idx, mb, reps=24)
    for l, t in zip(labels, ms):
        prev = out.get(l, (0.0, n_mb))
        out[l] = (prev[0] + float(t), n_mb)
    labels, ms = eng.profile_graph(1, None, 0, reps=24)
    for l, t in zip(labels, ms):
        prev = out.get("rollout/" + l, (0.0, args.horizon))
        out["rollout/" + l] = (prev[0] + float(t), args.horizon)
    eng.read_logs()
    return out


def kernel_flops(label, n, args):
    """algorithmic FLOPs of one launch processing n samples (None for non-GEMM kernels)"""
    base = label.split("/")[-1]
    macs = MACS_NATURE84 if getattr(args, "frames", "gray") == "rgb" else MACS
    for k in ("conv0", "conv1", "conv2"):
        if base.startswith(k + "_") and base.split("_")[1] in ("fwd", "wgrad", "dgrad"):
            return 2.0 * macs[k] * n
        if base == k + "_bwd":                       # data + weight gradient in one kernel (pconv_bwd_kernel)
            return 4.0 * macs[k] * n
    if base in ("fc_fwd", "fc_wgrad", "fc_dgrad"):
        return 2.0 * macs["fc"] * n
    return None


def workload_config(args, world):
    """the `config` object both arms print (the reference arm times the SAME workload on the host cores)"""
    rgb = args.frames == "rgb"
    N = args.envs * args.horizon
    a2c = getattr(args, "algo", "ppo") == "a2c"
    game = getattr(args, "game", "breakout")
    return {"workload": ("%s %s-shaped (A=%d), %d envs/GPU x %d-step rollout, %s, %s; synthetic %s emulator frames"
                         % ("A2C" if a2c else "PPO", game, action_count(game), args.envs, args.horizon,
                            "classic Nature-CNN @ (4,84,84)" if rgb else "cnn preset %d @ (4,104,80)" % args.spec,
                            "one full-batch RMSProp step" if a2c else "%d epochs x mb %d, Adam" % (args.epochs, args.minibatch),
                            "210x160x3 RGB" if rgb else "210x160 grayscale")),
            "frames": args.frames, "game": getattr(args, "game", "breakout"),
            "poll_horizon": getattr(args, "poll_horizon", 0),
            "algo": args.algo, "learner": ("single" if world == 1 and args.parallelism == "sync" else args.parallelism),
            "envs_per_gpu": args.envs, "horizon": args.horizon, "parallelism": "dp%d" % world,
            "l2_policy": "inputs larger than L2 (%.2f GB rollout buffer, %d MB frame pool)" %
                         (N * 4 * (84 * 84 if rgb else 104 * 80) / 1e9,
                          args.pool_frames * (100800 if rgb else 33600) // 2 ** 20),
            "step": "one full %s iteration" % ("A2C" if a2c else "PPO")}


def action_count(game):
    """policy action count of a --game (ALE minimal action sets; a mix pads to its largest set)"""
    from accel_rl_b200.envs.atari_env import GAME_MIXES, MINIMAL_ACTIONS
    if game in GAME_MIXES:
        return max(MINIMAL_ACTIONS[g] for g in GAME_MIXES[game])
    return MINIMAL_ACTIONS.get(game, 4)


def sync_parity(runner, rank, world):
    """Proof carried by every multi-GPU bench line (checked on the ranks that were just timed, after the timed region):
    (1) the parameter vectors of all ranks are bit-identical after the training iterations; (2) one synchronous step
    (fused P2P all-reduce + average + Adam, csrc/comm.cuh) from a known state equals the closed-form first Adam step
    (optimizers/update_methods_stats.py:55-87) on the NCCL-averaged gradient — what optimizers/sync/base.py:22-24 +
    optimizers/util.py:63-67 compute; (3) ... and leaves bit-identical parameters on every rank.  Raises on failure."""
    import torch
    import torch.distributed as dist
    eng = runner.policy.engine
    out = {}

    def identical():
        mine = eng.params.clone()
        ref = mine.clone()
        dist.broadcast(ref, src=0)
        ok = torch.tensor([1 if torch.equal(mine, ref) else 0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        return bool(ok.item())
    torch.cuda.synchronize()
    out["ranks_bit_identical_after_training"] = identical()
    cfg = eng.opt_cfg
    if int(cfg.update) == 0:
        eng.reset_opt_state()
        eng.set_lr_mult(1.0)
        eng.read_logs()
        p0 = eng.params.clone()
        g = torch.randn(eng.n_params, device="cuda", generator=torch.Generator("cuda").manual_seed(1000 + rank)) * 0.01
        eng.grad.copy_(g)
        gavg = g.clone()
        dist.all_reduce(gavg)
        gavg /= world
        torch.cuda.synchronize()
        dist.barrier()
        eng.sync_allreduce_update()
        torch.cuda.synchronize()
        b1, b2, eps, lr = float(cfg.beta1), float(cfg.beta2), float(cfg.epsilon), float(cfg.learning_rate)
        norm = float(gavg.double().norm())
        if float(cfg.grad_norm_clip) > 0:
            gavg = gavg * (min(norm, float(cfg.grad_norm_clip)) / (1e-7 + norm))
        a_t = lr * (1.0 - b2) ** 0.5 / (1.0 - b1)
        want = p0 - a_t * ((1 - b1) * gavg) / (((1 - b2) * gavg * gavg).sqrt() + eps)
        err = (eng.params - want).abs()
        tol = 2e-7 + 2e-6 * want.abs()
        _, norms = eng.read_logs()
        out["sync_step_max_abs_err_vs_adam"] = float(err.max())
        out["sync_step_within_tolerance"] = bool((err <= tol).all().item()) and abs(float(norms[0]) - norm) <= 1e-5 * norm
        out["ranks_bit_identical_after_sync_step"] = identical()
    out["world"] = world
    ok = torch.tensor([1 if all(v for k, v in out.items() if isinstance(v, bool)) else 0], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)              # one verdict for all ranks
    out["passed"] = bool(ok.item())
    if not out["passed"]:
        # reported in the JSON line (`parity.passed` false) and as a non-zero exit code after the line is printed
        print("sync data-parallel parity check FAILED on rank %d: %s" % (rank, out), file=sys.stderr)
    return out


def async_parity(runner, rank, world):
    """Proof carried by every asynchronous multi-GPU bench line (after the timed region): learners push in turn (a
    barrier between turns); after learner k's push the central (p, m, v) in rank 0's HBM must equal the closed-form
    Adam / RMSProp chunk update (optimizers/async/chunked_updates.py:53-120, per-learner step count) of the central state
    read before the push, with learner k's gradient, and learner k's local parameters must BE the central ones
    (optimizers/async/base.py:59-104 push then pull).  Raises on failure."""
    import numpy as np
    import torch
    import torch.distributed as dist
    eng = runner.policy.engine
    cfg = eng.opt_cfg
    b1, b2, eps, lr, rho = (float(cfg.beta1), float(cfg.beta2), float(cfg.epsilon), float(cfg.learning_rate), float(cfg.rho))
    clip = float(cfg.grad_norm_clip)
    eng.set_lr_mult(1.0)
    eng.read_logs()
    worst = 0.0
    ok = True
    for k in range(world):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        if rank == k:
            p0, m0, v0 = (eng.async_read_central(i).astype(np.float64) for i in range(3))
            t = eng.get_opt_step() + 1
            g = (np.random.RandomState(500 + k).randn(eng.n_params) * 0.01).astype(np.float32)
            eng.grad.copy_(torch.from_numpy(g))
            eng.async_push_pull()
            torch.cuda.synchronize()
            g = g.astype(np.float64)
            norm = np.sqrt((g * g).sum())
            if clip > 0:
                g = g * (min(norm, clip) / (1e-7 + norm))
            if int(cfg.update) == 0:
                m1 = b1 * m0 + (1 - b1) * g
                v1 = b2 * v0 + (1 - b2) * g * g
                want = p0 - lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t) * m1 / (np.sqrt(v1) + eps)
            else:
                v1 = rho * v0 + (1 - rho) * g * g
                want = p0 - lr * g / np.sqrt(v1 + eps)
            got = eng.async_read_central(0).astype(np.float64)
            err = np.abs(got - want)
            worst = max(worst, float(err.max()))
            ok = ok and bool((err <= 2e-7 + 2e-6 * np.abs(want)).all())
            ok = ok and bool(np.array_equal(eng.get_params(), got.astype(np.float32)))
        if world > 1:
            dist.barrier()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    w = torch.tensor([worst], device="cuda")
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        dist.all_reduce(w, op=dist.ReduceOp.MAX)
    out = {"world": world, "turn_by_turn_push_matches_chunk_update": bool(flag.item()), "max_abs_err": float(w.item())}
    if not out["turn_by_turn_push_matches_chunk_update"]:
        raise SystemExit("async data-parallel parity check FAILED: %s" % out)
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pk = peaks()
    runner = build_runner(args, "device", rank, world)
    N = args.envs * args.horizon
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms_step, launches, wall, last, itr = timed_iterations(runner, args.steps, args.warmup, world)
    clk = clocks.summary() if rank == 0 else None
    value = world * N / (ms_step * 1e-3)

    result = None
    timeline = None
    if world > 1 and args.parallelism == "sync":
        try:
            timeline = runner.policy.engine.comm_trace(reset=True)       # accumulated over warm-up + timed iterations
        except Exception:
            timeline = None
    phases = phase_split(runner, itr)
    parity = sync_parity(runner, rank, world) if (world > 1 and args.parallelism == "sync") else None
    if args.parallelism == "async":
        parity = async_parity(runner, rank, world)
    if rank == 0:
        # (per-kernel timing replays single kernels of THIS rank's context after everything else was measured and
        # checked; under the asynchronous learner the update is a cross-GPU kernel and is left out)
        bd = kernel_breakdown(runner, args) if (args.spec == 1 and args.parallelism == "sync") else {}
        if world > 1:
            # the synchronous learner's update is the cross-GPU pair sync_fc_kernel / sync_tail_kernel (csrc/comm.cuh),
            # timed on the device in `sync_timeline`; the local update kernel is not on this path
            bd.pop("clip_update", None)
        # dominant kernel by time per PPO iteration
        roof = None
        kernels = []
        total_ms = sum(t * c for t, c in bd.values())
        for label, (t, count) in sorted(bd.items(), key=lambda kv: -kv[1][0] * kv[1][1]):
            n = args.envs if label.startswith("rollout/") else (N if args.algo == "a2c" else args.minibatch)
            fl = kernel_flops(label, n, args)
            kernels.append({"kernel": label, "ms": round(t, 5), "launches_per_step": count,
                            "share": round(t * count / total_ms, 4) if total_ms else None,
                            "tflops": round(fl / (t * 1e-3) / 1e12, 2) if fl else None})
        # `roofline`: the kernel with the largest share of the step.  GEMM kernels are held against the tensor peak
        # (flops = 2 * MACs of the launch); the update kernel against HBM with its algorithmic 28 B per parameter (read p,
        # m, v, g; write p, m, v — DESIGN.md §3).  `roofline_tensor`: the largest GEMM kernel, when that is not the same.
        roof_tensor = None
        traffic_note = "DRAM bytes per launch (read + write), ncu --set full with cold caches (profiles/r2_full.md)"
        if kernels and kernels[0]["kernel"] == "clip_update":
            k = kernels[0]
            nbytes = 28.0 * runner.policy.engine.n_params
            gbs = nbytes / (k["ms"] * 1e-3) / 1e9
            roof = {"bound": "hbm", "kernel": "clip_update", "achieved": round(gbs, 1), "peak": pk["hbm"], "unit": "GB/s",
                    "frac": round(gbs / pk["hbm"], 4), "traffic": ncu_traffic("clip_update", args), "traffic_note": traffic_note,
                    "algorithmic_bytes_per_launch": nbytes,
                    "peak_source": pk["src"] + " (copy bandwidth; the kernel is timed alone, its 58 MB of optimiser state "
                                               "and parameters partly stay in the 126 MB L2 between launches)",
                    "share_of_step": k["share"]}
        for k in kernels:
            if k["tflops"] is not None:
                roof_tensor = {"bound": "tensor", "kernel": k["kernel"], "achieved": k["tflops"], "peak": pk["tf_sustained"],
                        "unit": "TFLOP/s", "frac": round(k["tflops"] / pk["tf_sustained"], 4),
                        "traffic": ncu_traffic(k["kernel"], args),
                        "traffic_note": traffic_note,
                        "peak_source": pk["src"] + " (sustained bf16, kernel timed inside a long step)",
                        "share_of_step": k["share"]}
                roof_t = roof_tensor
                if k["kernel"] in ("conv1_wgrad", "conv2_wgrad"):
                    # these two kernels are launched on a capped grid so the data-gradient chain runs beside them
                    # (csrc/api.cu: wgrad_ctas, ARL_WGRAD_CTAS); `achieved` is the whole-GPU figure of that launch
                    ctas = int(os.environ.get("ARL_WGRAD_CTAS", "48"))
                    roof_t["note"] = ("launched on %d of 148 SMs by design, concurrent with the data-gradient kernels: %.0f TFLOP/s "
                                      "per occupied-SM share; tensor pipe 41-51 %% active in profiles/r2_full.md"
                                      % (ctas, k["tflops"] * 148.0 / ctas))
                break
        if roof is None:
            roof, roof_tensor = roof_tensor, None
        # per-env-step model FLOPs: PPO = fwd + fwd/T + epochs*train; A2C = fwd + fwd/T + train
        fwd, train, cfwd, ctrain = FLOPS[args.frames]
        rgb = args.frames == "rgb"
        k_train = args.epochs if args.algo == "ppo" else 1
        flop_step = (fwd * (1 + 1.0 / args.horizon) + k_train * train) if args.spec == 1 else None
        flop_conv = (cfwd * (1 + 1.0 / args.horizon) + k_train * ctrain) if args.spec == 1 else None
        result = {
            "metric": "env-steps/sec (%s Atari, %d envs/GPU)" % (args.algo.upper(), args.envs), "value": round(value, 1),
            "unit": "env-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args, world),
            "clocks": clk,
            "gpu_launches": int(launches),
            "model_flops_per_env_step": flop_step,
            "tensor_roofline_frac_whole_step": round(value / world * flop_step / (pk["tf_sustained"] * 1e12), 4)
            if flop_step else None,
            "conv_tile_roofline_frac": round(value / world * flop_conv / (pk["tf_sustained"] * 1e12), 4) if flop_conv else None,
            "roofline": roof,
            "roofline_tensor": roof_tensor,
            "kernels_note": "per-kernel ms (and the roofline `achieved` figures): measured live in this run, after the timed "
                            "region, by re-launching each kernel node of the product's captured graphs alone (24 chained "
                            "replays, CUDA events, warm caches); `share` = that time x launches per step / the sum over all "
                            "kernels, i.e. a share of summed kernel time, not of the timed region (whose kernels overlap on three "
                            "streams; profiles/r2_timeline.md has the in-graph timeline)",
            "kernels": kernels,
            "host_wall_s": round(wall, 3),
            "phases": phases,
        }
        if parity is not None:
            result["parity"] = parity
        if timeline is not None:
            result["sync_timeline"] = timeline
    # ---- e2e: raw frames from pinned host memory every step, results read back ----
    if not args.no_e2e:
        runner.policy.engine.close()
        del runner
        torch.cuda.empty_cache()
        r2 = build_runner(args, "host", rank, world)
        smp = r2.sampler
        h0, d0 = smp.h2d_bytes, smp.d2h_bytes
        steps2 = max(2, args.steps // 2)
        ms2, _, _, _, _ = timed_iterations(r2, steps2, max(1, args.warmup // 2), world)
        n_it = steps2 + max(1, args.warmup // 2)
        idx_bytes = args.epochs * (N // args.minibatch) * args.minibatch * 4
        log_bytes = 2 * 4 * args.epochs * (N // args.minibatch)
        if rank == 0:
            result["e2e"] = {"value": round(world * N / (ms2 * 1e-3), 1), "unit": "env-steps/s",
                             "h2d_bytes_per_step": int((smp.h2d_bytes - h0) / n_it + idx_bytes),
                             "d2h_bytes_per_step": int((smp.d2h_bytes - d0) / n_it + log_bytes),
                             "ms_per_step": round(ms2, 3),
                             "note": "raw emulator frames (2 x %s u8 per env-step) H2D from pinned memory on a copy "
                                     "stream, actions D2H every step, losses/grad norms D2H every iteration"
                                     % ("210x160x3" if args.frames == "rgb" else "210x160")}
        r2.policy.engine.close()
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        result["cpu_baseline"] = cpu_port(args, steps=1)
    if rank == 0:
        _emit(result)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if parity is not None and parity.get("passed") is False:
        raise SystemExit(3)


# =============================================================================================
# CPU port of the reference path (oracle/): --impl reference and the cpu_baseline leg
# =============================================================================================
def ref_n_parallel(envs, cores):
    """simulator processes per alternating group: the largest divisor of envs/2 that is <= (cores - 1) // 2
    (SURVEY.md 8d / BASELINE.md 3: 2*n_parallel workers + the master)"""
    cap = max(1, (cores - 1) // 2)
    return max(d for d in range(1, cap + 1) if (envs // 2) % d == 0)


def cpu_port(args, steps=1, warmup=0):
    """The reference's CPU implementation of the path, restated (oracle/): the MULTI-PROCESS sampler structure of
    ActsrvAltOvrlpSampler (oracle/mp_sampler.py: 2*n_parallel simulator processes in two alternating groups, shared
    buffers, semaphores; the master serves actions from an fp32 torch-CPU policy) + GAE + PPO epochs on all host threads,
    on a BOUNDED sample of the workload per step: `--cpu-sample-steps` rollout steps of all envs, trained for the same
    epochs x minibatch ratio.  -> per-step records (what was actually timed)"""
    import numpy as np
    import torch
    from oracle import net as onet, mp_sampler, learner as olearner, synth_ale
    cores = os.cpu_count() or 1
    B, Ts = args.envs, args.cpu_sample_steps
    rgb = getattr(args, "frames", "gray") == "rgb"
    spec = dict(NATURE84_SPEC, conv_pads=[0, 0, 0]) if rgb else onet.CNN_SPECS[args.spec]
    from accel_rl_b200.envs.atari_env import GAME_MIXES
    game = getattr(args, "game", "breakout")
    algo = getattr(args, "algo", "ppo")
    A = action_count(game)
    rules = dict(synth_ale.DEFAULT_RULES, pool_frames=256, n_games=len(GAME_MIXES.get(game, ("",))))
    pool = synth_ale.make_pool(256, seed=0, channels=3) if rgb else synth_ale.make_pool(256, seed=0)
    flat = onet.init_params(spec, (4, 84, 84) if rgb else (4, 104, 80), A, np.random.RandomState(0),
                            np.random.RandomState(1))
    n_par = ref_n_parallel(B, cores)
    master_threads = max(1, cores - 2 * n_par)
    Ts = min(Ts, args.horizon)
    smp = mp_sampler.MpOracleSampler(n_par, B // (2 * n_par), Ts, pool, rules, A, 0.99)
    opt = onet.RMSProp(flat.size) if algo == "a2c" else onet.Adam(flat.size, 1e-3, epsilon=1e-5)
    rng = np.random.RandomState(0)

    def policy_fn(obs):
        with torch.no_grad():
            p, v = onet.forward(torch.from_numpy(flat), torch.from_numpy(np.ascontiguousarray(obs)), spec, A)
        return p.numpy(), v.numpy()

    mb = min(args.minibatch, B * Ts)
    times = []
    try:
        for it in range(warmup + steps):
            torch.set_num_threads(master_threads)        # the simulator processes own the other cores while sampling
            t0 = time.time()
            buf, _ = smp.obtain_samples(policy_fn, rng.rand(Ts, B))
            t1 = time.time()
            torch.set_num_threads(cores)
            flat, _, _, _ = olearner.optimize_policy(flat, opt, buf, spec, A, Ts, algo, rng, epochs=args.epochs,
                                                     minibatch_size=mb, emulate_bf16=False)
            t2 = time.time()
            if it >= warmup:
                times.append((t1 - t0, t2 - t1))
    finally:
        smp.shutdown()
    samp = float(np.mean([a for a, _ in times]))
    learn = float(np.mean([b for _, b in times]))
    n = B * Ts
    return {"value": round(n / (samp + learn), 1), "unit": "env-steps/s", "cores": cores, "kind": "port",
            "sample": "per step: %d envs x %d rollout steps (%d env-steps) through %d simulator processes (two alternating "
                      "groups of %d) + master, then GAE + %s on them; oracle/ port: Theano replaced by an fp32 "
                      "torch-CPU restatement (%d threads serving actions, %d for the learner), ALE by the synthetic emulator"
                      % (B, Ts, n, 2 * n_par, n_par,
                         "one full-batch RMSProp step" if algo == "a2c" else "%d epochs x mb %d" % (args.epochs, mb),
                         master_threads, cores),
            "sampler_env_steps_per_s": round(n / samp, 1), "learner_env_steps_per_s": round(n / learn, 1),
            "seconds_per_sample": round(samp + learn, 3), "sample_env_steps": n, "sim_processes": 2 * n_par}


def run_frame_sweep(args):
    """BASELINE.json configs[2]: frame-kernel achieved HBM GB/s over env counts, reference mode (2 x 210x160 gray ->
    104x80 stack, the rollout's frame_kernel incl. its bf16 mirror) and north-star mode (2 x 210x160x3 RGB -> 84x84).
    Algorithmic bytes per env-step: 75 520 B / 208 656 B.  Prints one JSON line."""
    import numpy as np
    import torch
    from accel_rl_b200.policies import AtariCnnPolicy, cnn_specs
    torch.cuda.set_device(0)
    pk = peaks()
    out = {"metric": "frame-kernel achieved HBM GB/s", "unit": "GB/s", "peak": pk["hbm"], "peak_source": pk["src"],
           "bytes_per_env_step": {"reference_mode": 75520, "north_star_rgb_mode": 208656}, "game": args.game,
           "sweep": [], "rgb_sweep": []}
    for B in (256, 512, 1024, 2048, 4096):
        a = argparse.Namespace(**vars(args))
        a.envs, a.horizon, a.minibatch, a.epochs = B, 2, 512, 1
        runner = build_runner(a, "device", 0, 1)
        s, _ = runner.sampler.obtain_samples(0)
        labels, ms = runner.policy.engine.profile_graph(1, None, 0, reps=20)
        t = float(sum(m for l, m in zip(labels, ms) if l == "frame"))
        gbs = B * 75520 / (t * 1e-3) / 1e9
        out["sweep"].append({"envs": B, "us": round(t * 1e3, 2), "gbs": round(gbs, 1), "frac": round(gbs / pk["hbm"], 4)})
        runner.policy.engine.close()
        del runner
        torch.cuda.empty_cache()
    pol = AtariCnnPolicy(**cnn_specs[1])
    from accel_rl_b200.envs.atari_env import EnvSpec
    from accel_rl_b200.spaces import Discrete, UintBox
    pol.initialize(EnvSpec(UintBox((4, 104, 80)), Discrete(4)))
    eng = pol.engine
    for B in (256, 512, 1024, 2048):
        nset = max(1, int(np.ceil(400e6 / (B * 201600))))      # rotate input sets so reads come from HBM, not L2
        ra = [torch.randint(0, 256, (B, 210, 160, 3), dtype=torch.uint8, device="cuda") for _ in range(nset)]
        rb = [torch.randint(0, 256, (B, 210, 160, 3), dtype=torch.uint8, device="cuda") for _ in range(nset)]
        st = torch.zeros(B, 4, 84, 84, dtype=torch.uint8, device="cuda")
        st16 = torch.zeros(B, 4, 84, 84, dtype=torch.bfloat16, device="cuda")
        for i in range(3):
            eng.frame_update_rgb(ra[i % nset], rb[i % nset], None, st, st16)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        ev0.record()
        for i in range(reps):
            eng.frame_update_rgb(ra[i % nset], rb[i % nset], None, st, st16)
        ev1.record()
        torch.cuda.synchronize()
        t = ev0.elapsed_time(ev1) / reps
        gbs = B * 208656 / (t * 1e-3) / 1e9
        out["rgb_sweep"].append({"envs": B, "us": round(t * 1e3, 2), "gbs": round(gbs, 1), "frac": round(gbs / pk["hbm"], 4)})
        del ra, rb
        torch.cuda.empty_cache()
    eng.close()
    _emit(out)


def run_reference(args):
    """reference arm: the reference's own CPU implementation of the path (oracle port, see cpu_port) on the host cores, on
    OUR arm's metric / unit / config.  One "step" = one bounded sample of the workload; ms_per_step is what was actually
    timed per step, value = env-steps of the sample / that time (a rate: no extrapolation enters it)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.time()
    r = cpu_port(args, steps=max(1, args.steps), warmup=max(0, args.warmup))
    out = {"impl": "reference", "metric": "env-steps/sec (%s Atari, %d envs/GPU)" % (args.algo.upper(), args.envs),
           "value": r["value"], "unit": "env-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": round(1e3 * r["seconds_per_sample"], 1), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": workload_config(args, args.gpus),
           "cpu_baseline": r,
           "e2e": {"value": r["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0,
           "wall_s": round(time.time() - t0, 1)}
    _emit(out)


def _emit(obj):
    """the ONE JSON line goes to the real stdout; everything else any library prints was routed to stderr"""
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


if __name__ == "__main__":
    # keep stdout clean for the single JSON line (NCCL / torchrun banners go to stderr)
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    a = parse()
    if a.workload == "frame_sweep":
        run_frame_sweep(a)
    elif a.workload == "host_emulators":
        run_host_emulators(a)
    elif a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
