"""Engine: one libaccelrl_b200 context per process/GPU, shared by the policy, sampler and optimizer.

PyTorch is used for device memory (tensors as containers) and the current CUDA stream only; every
computation goes through the C ABI (include/accelrl_b200.h).  No CPU fallback exists.
"""
import ctypes as C

import numpy as np
import torch

from accel_rl_b200 import _lib as L


class Engine(object):
    def __init__(self, conv_filters, conv_filter_sizes, conv_strides, conv_pads, hidden_sizes, n_actions,
                 obs_shape, pixel_scale=255., max_rows=512, device=None):
        if len(hidden_sizes) != 1:
            raise NotImplementedError("libaccelrl_b200 supports exactly one hidden FC layer (cnn_specs 0/1)")
        if not torch.cuda.is_available():
            raise RuntimeError("accel_rl_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = L.load()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        pads = [p[0] if isinstance(p, (tuple, list)) else p for p in conv_pads]
        for p in conv_pads:
            if isinstance(p, (tuple, list)) and p[0] != p[1]:
                raise NotImplementedError("asymmetric conv padding is not supported")
        cfg = L.NetCfg()
        cfg.n_conv = len(conv_filters)
        for i in range(cfg.n_conv):
            cfg.conv_filters[i] = int(conv_filters[i])
            cfg.conv_sizes[i] = int(conv_filter_sizes[i])
            cfg.conv_strides[i] = int(conv_strides[i])
            cfg.conv_pads[i] = int(pads[i])
        cfg.hidden = int(hidden_sizes[0])
        cfg.n_actions = int(n_actions)
        cfg.in_c, cfg.in_h, cfg.in_w = (int(x) for x in obs_shape)
        cfg.pixel_scale = float(pixel_scale)
        cfg.max_rows = int(max_rows)
        self.cfg = cfg
        self.max_rows = int(max_rows)
        self.n_actions = int(n_actions)
        self.obs_shape = tuple(int(x) for x in obs_shape)
        self.ctx = C.c_void_p()
        with torch.cuda.device(self.device):
            rc = self.lib.arl_create(C.byref(cfg), C.byref(self.ctx))
        if rc != 0:
            msg = self.lib.arl_last_error(None)
            raise L.ArlError("arl_create failed (%d): %s" % (rc, msg.decode() if msg else "?"))
        self.n_params = int(self.lib.arl_param_count(self.ctx))
        offs = (C.c_long * 32)()
        sizes = (C.c_long * 32)()
        n = self.lib.arl_param_layout(self.ctx, offs, sizes, 32)
        self.layout = [(int(offs[i]), int(sizes[i])) for i in range(n)]
        # params, m, v in ONE allocation (what the update kernel reads and rewrites every minibatch: the library can ask the
        # L2 to keep that range resident between updates, csrc/api.cu l2_persist_setup); the gradient is a stream
        npad = (self.n_params + 63) // 64 * 64
        self._state = torch.zeros(4 * npad, dtype=torch.float32, device=self.device)
        self.params = self._state[:self.n_params]
        self.m = self._state[npad:npad + self.n_params]
        self.v = self._state[2 * npad:2 * npad + self.n_params]
        self.grad = self._state[3 * npad:3 * npad + self.n_params]
        self._keep = {}
        self._sym = None
        self._dp_mode = None            # "synchronous" / "asynchronous" once comm_init / async_init ran
        self.bind()

    # ---- plumbing ----------------------------------------------------------------------------
    def _s(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def check(self, rc):
        L.check(self.ctx, rc)

    def bind(self):
        self.check(self.lib.arl_bind_params(self.ctx, L.ptr(self.params), L.ptr(self.grad), L.ptr(self.m),
                                            L.ptr(self.v)))

    def close(self):
        if self.ctx:
            self.lib.arl_destroy(self.ctx)
            self.ctx = None

    def device_error(self):
        return int(self.lib.arl_device_error(self.ctx))

    @property
    def launches(self):
        return int(self.lib.arl_kernel_launches(self.ctx))

    # ---- parameters --------------------------------------------------------------------------
    def set_params(self, flat):
        flat = np.ascontiguousarray(np.asarray(flat, dtype=np.float32)).reshape(-1)
        if flat.size != self.n_params:
            raise ValueError("expected %d parameters, got %d" % (self.n_params, flat.size))
        self.params.copy_(torch.from_numpy(flat))
        self.pack()

    def get_params(self):
        return self.params.detach().cpu().numpy().copy()

    def pack(self):
        self.check(self.lib.arl_pack_weights(self.ctx, self._s()))

    # ---- policy ------------------------------------------------------------------------------
    def forward(self, obs, n=None, idx=None, out_rows=None, prob=None, value=None, uniforms=None, actions=None):
        n = int(obs.shape[0] if n is None else n)
        self.check(self.lib.arl_policy_forward(self.ctx, L.ptr(obs), L.ptr(idx), n, L.ptr(out_rows), L.ptr(prob),
                                               L.ptr(value), L.ptr(uniforms), L.ptr(actions), self._s()))

    def sample_actions(self, prob, uniforms, actions):
        n, a = int(prob.shape[0]), int(prob.shape[1])
        self.check(self.lib.arl_sample_actions(self.ctx, L.ptr(prob), L.ptr(uniforms), L.ptr(actions), n, a, self._s()))

    def frame_update(self, raw_a, raw_b, reset_mask, stack):
        n, planes = int(stack.shape[0]), int(stack.shape[1])
        self.check(self.lib.arl_frame_update(self.ctx, L.ptr(raw_a), L.ptr(raw_b), L.ptr(reset_mask), L.ptr(stack), n,
                                             planes, self._s()))

    def frame_update_rgb(self, raw_a, raw_b, reset_mask, stack, stack_bf16=None):
        """north-star frame mode: RGB (n,210,160,3) pairs -> (n,planes,84,84) u8 stack (+ bf16 copy)"""
        n, planes = int(stack.shape[0]), int(stack.shape[1])
        self.check(self.lib.arl_frame_update_rgb(self.ctx, L.ptr(raw_a), L.ptr(raw_b), L.ptr(reset_mask), L.ptr(stack),
                                                 L.ptr(stack_bf16), n, planes, self._s()))

    # ---- sampler -----------------------------------------------------------------------------
    def sampler_select(self, slot):
        """0: training sampler (default), 1: evaluation sampler"""
        self._sampler_slot = int(slot)
        self.check(self.lib.arl_sampler_select(self.ctx, int(slot)))

    def sampler_configure(self, cfg, keep):
        self._keep["sampler%d" % getattr(self, "_sampler_slot", 0)] = keep
        self.check(self.lib.arl_sampler_configure(self.ctx, C.byref(cfg)))

    def sampler_reset(self):
        self.check(self.lib.arl_sampler_reset(self.ctx, self._s()))

    def sampler_warmup(self, n_steps):
        """start_envs decorrelation: env e takes n_steps[e] warm-up steps (int32 device tensor [n_envs])"""
        mx = int(n_steps.max().item()) if n_steps.numel() else 0
        if mx > 0:
            self.check(self.lib.arl_sampler_warmup(self.ctx, L.ptr(n_steps), mx, self._s()))

    def rollout_run(self):
        self.check(self.lib.arl_rollout_run(self.ctx, self._s()))

    def rollout_begin(self):
        self.check(self.lib.arl_rollout_begin(self.ctx, self._s()))

    def rollout_step(self, s, staging=None):
        self.check(self.lib.arl_rollout_step(self.ctx, int(s), L.ptr(staging), self._s()))

    def rollout_end(self):
        self.check(self.lib.arl_rollout_end(self.ctx, self._s()))

    # external-emulator feed: forward + sample, then file what the host workers produced and run the frame pipeline
    def rollout_serve(self, s, e0=0, n=-1):
        self.check(self.lib.arl_rollout_serve(self.ctx, int(s), int(e0), int(n), self._s()))

    def rollout_ingest(self, s, staging, ext, e0=0, n=-1):
        self.check(self.lib.arl_rollout_ingest(self.ctx, int(s), int(e0), int(n), L.ptr(staging), L.ptr(ext), self._s()))

    def copy_async(self, dst_ptr, src_ptr, nbytes, to_device, stream=None):
        st = self._s() if stream is None else C.c_void_p(stream.cuda_stream)
        self.check(self.lib.arl_copy_async(self.ctx, C.c_void_p(dst_ptr), C.c_void_p(src_ptr), int(nbytes),
                                           1 if to_device else 0, st))

    def traj_read(self, cap):
        n = C.c_int()
        env = np.zeros(cap, np.int32); ln = np.zeros(cap, np.int32); nz = np.zeros(cap, np.int32)
        ret = np.zeros(cap, np.float32); raw = np.zeros(cap, np.float32); disc = np.zeros(cap, np.float32)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        self.check(self.lib.arl_traj_read(self.ctx, C.byref(n), p(env), p(ln), p(ret), p(raw), p(nz), p(disc), cap,
                                          self._s()))
        k = n.value
        return env[:k], ln[:k], ret[:k], raw[:k], nz[:k], disc[:k]

    # ---- advantages --------------------------------------------------------------------------
    def gae(self, rewards, values, dones, need_reset, last_values, discount, gae_lambda, adv, ret, valids, n_envs,
            horizon, standardize):
        self.check(self.lib.arl_gae(self.ctx, L.ptr(rewards), L.ptr(values), L.ptr(dones), L.ptr(need_reset),
                                    L.ptr(last_values), float(discount), float(gae_lambda), L.ptr(adv), L.ptr(ret),
                                    L.ptr(valids), int(n_envs), int(horizon), int(bool(standardize)), self._s()))

    # ---- learner -----------------------------------------------------------------------------
    def opt_configure(self, **kw):
        cfg = L.OptCfg(**kw)
        self.opt_cfg = cfg
        self.check(self.lib.arl_opt_configure(self.ctx, C.byref(cfg)))

    def bind_train_inputs(self, obs, actions, adv, ret, old_value, old_prob, valids=None):
        self._keep["train"] = (obs, actions, adv, ret, old_value, old_prob, valids)
        self.check(self.lib.arl_bind_train_inputs(self.ctx, L.ptr(obs), L.ptr(actions), L.ptr(adv), L.ptr(ret),
                                                  L.ptr(old_value), L.ptr(old_prob), L.ptr(valids), int(obs.shape[0])))

    def set_lr_mult(self, lr_mult):
        self.check(self.lib.arl_set_lr_mult(self.ctx, float(lr_mult), self._s()))

    def grad_minibatch(self, idx, mb_size):
        self.check(self.lib.arl_grad_minibatch(self.ctx, L.ptr(idx), int(mb_size), self._s()))

    def clip_update(self, gscale=1.0):
        self.check(self.lib.arl_clip_update(self.ctx, float(gscale), self._s()))

    def train_minibatches(self, idx, mb_size, count, sync=False):
        self._keep["idx"] = idx
        fn = {False: self.lib.arl_train_minibatches, True: self.lib.arl_train_minibatches_sync, "sync": self.lib.arl_train_minibatches_sync,
              "async": self.lib.arl_train_minibatches_async}[sync]
        self.check(fn(self.ctx, L.ptr(idx), int(mb_size), int(count), self._s()))

    def read_logs(self, cap=4096):
        loss = np.zeros(cap, np.float32)
        norm = np.zeros(cap, np.float32)
        n = C.c_int()
        self.check(self.lib.arl_read_logs(self.ctx, loss.ctypes.data_as(C.c_void_p), norm.ctypes.data_as(C.c_void_p),
                                          cap, C.byref(n), self._s()))
        return loss[:n.value].copy(), norm[:n.value].copy()

    def reset_opt_state(self):
        self.check(self.lib.arl_reset_opt_state(self.ctx, self._s()))

    def get_opt_step(self):
        """the optimizer's update count (Adam t) on this learner"""
        t = C.c_int()
        self.check(self.lib.arl_opt_step_get(self.ctx, C.byref(t), self._s()))
        return int(t.value)

    def get_opt_state(self):
        """-> dict(m, v, step): first/second moment vectors (RMSProp keeps its accumulator in v) and the update count"""
        if self._dp_mode is not None:
            raise NotImplementedError("optimizer-state snapshots cover the single-GPU learner (the %s learner keeps m "
                                      "and v %s)" % (self._dp_mode, "sharded across ranks" if self._dp_mode == "synchronous"
                                                     else "in the central store"))
        t = C.c_int()
        self.check(self.lib.arl_opt_step_get(self.ctx, C.byref(t), self._s()))
        return dict(m=self.m.detach().cpu().numpy().copy(), v=self.v.detach().cpu().numpy().copy(), step=int(t.value))

    def set_opt_state(self, state):
        if self._dp_mode is not None:
            raise NotImplementedError("optimizer-state restore covers the single-GPU learner")
        for k in ("m", "v"):
            a = np.ascontiguousarray(np.asarray(state[k], dtype=np.float32)).reshape(-1)
            if a.size != self.n_params:
                raise ValueError("optimizer state '%s' has %d entries, expected %d" % (k, a.size, self.n_params))
            getattr(self, k).copy_(torch.from_numpy(a))
        self.check(self.lib.arl_opt_step_set(self.ctx, int(state["step"]), self._s()))

    # ---- profiling ---------------------------------------------------------------------------
    def profile_begin(self):
        self.check(self.lib.arl_profile_begin(self.ctx, self._s()))

    def profile_graph(self, kind, idx=None, mb_size=0, reps=20, cap=256):
        names = C.create_string_buffer(cap * 24)
        ms = np.zeros(cap, np.float32)
        n = C.c_int()
        self.check(self.lib.arl_profile_graph(self.ctx, int(kind), L.ptr(idx), int(mb_size), int(reps), names, len(names),
                                              ms.ctypes.data_as(C.c_void_p), cap, C.byref(n), self._s()))
        labels = names.value.decode().split(";")[:n.value]
        return labels, ms[:n.value].copy()

    def profile_timeline(self, kind, idx, mb_size, cap=96):
        """-> [(label, completion time in us after the minibatch graph's first node)] over all streams"""
        names = C.create_string_buffer(cap * 24)
        us = np.zeros(cap, np.float32)
        n = C.c_int()
        self.check(self.lib.arl_profile_timeline(self.ctx, int(kind), L.ptr(idx), int(mb_size), names, len(names),
                                                 us.ctypes.data_as(C.c_void_p), cap, C.byref(n), self._s()))
        labels = names.value.decode().split(";")[:n.value]
        return list(zip(labels, [float(x) for x in us[:n.value]]))

    def profile_end(self, cap=4096):
        names = C.create_string_buffer(cap * 24)
        ms = np.zeros(cap, np.float32)
        n = C.c_int()
        self.check(self.lib.arl_profile_end(self.ctx, names, len(names), ms.ctypes.data_as(C.c_void_p), cap,
                                            C.byref(n), self._s()))
        labels = names.value.decode().split(";")[:n.value]
        return labels, ms[:n.value].copy()

    # ---- sync data parallel ------------------------------------------------------------------
    def comm_init(self, rank, world, exchange):
        """exchange(handle_bytes) -> list of every rank's handle bytes (e.g. torch.distributed all_gather)."""
        h = (C.c_uint8 * L.IPC_HANDLE_BYTES)()
        self._dp_mode = "synchronous"
        self.check(self.lib.arl_comm_local_init(self.ctx, int(rank), int(world), h))
        g, p = C.c_void_p(), C.c_void_p()
        self.check(self.lib.arl_comm_buffers(self.ctx, C.byref(g), C.byref(p)))
        # re-home params/grad into the symmetric allocation peers can address
        new_params = _wrap_device_f32(p.value, self.n_params, self.device)
        new_grad = _wrap_device_f32(g.value, self.n_params, self.device)
        new_params.copy_(self.params)
        new_grad.zero_()
        self.params, self.grad = new_params, new_grad
        self.bind()
        handles = exchange(bytes(h))
        blob = b"".join(handles)
        buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        self.check(self.lib.arl_comm_connect(self.ctx, buf))
        self.pack()

    def comm_trace(self, reset=True):
        """device timeline of the overlapped synchronous step (microseconds per step), see arl_comm_trace"""
        out = (C.c_double * 8)()
        self.check(self.lib.arl_comm_trace(self.ctx, out, 1 if reset else 0, self._s()))
        keys = ("fc_wait_peers_us", "fc_reduce_update_publish_us", "tail_wait_peers_us", "tail_average_update_us",
                "slack_fc_end_to_tail_start_us", "steps")
        return {k: round(float(out[i]), 2) for i, k in enumerate(keys)}

    def sync_allreduce_update(self):
        self.check(self.lib.arl_sync_allreduce_update(self.ctx, self._s()))

    # ---- async data parallel -----------------------------------------------------------------
    def async_init(self, rank, world, n_update_chunks, exchange):
        """exchange(handle_bytes) -> list of every rank's handle bytes; rank 0's entry is the central store."""
        h = (C.c_uint8 * L.IPC_HANDLE_BYTES)()
        self._dp_mode = "asynchronous"
        self.check(self.lib.arl_async_local_init(self.ctx, int(rank), int(world), int(n_update_chunks), h))
        handles = exchange(bytes(h))
        buf = (C.c_uint8 * L.IPC_HANDLE_BYTES).from_buffer_copy(handles[0])
        self.check(self.lib.arl_async_connect(self.ctx, buf))
        return int(self.lib.arl_async_regions(self.ctx))

    def async_push_pull(self):
        self.check(self.lib.arl_async_push_pull(self.ctx, self._s()))

    def async_pull(self):
        """central parameters -> local parameters + operand copies (poll sampler)"""
        self.check(self.lib.arl_async_pull(self.ctx, self._s()))

    def async_read_central(self, which=0):
        out = np.zeros(self.n_params, np.float32)
        self.check(self.lib.arl_async_read_central(self.ctx, int(which), out.ctypes.data_as(C.c_void_p), self.n_params,
                                                   self._s()))
        return out

    def comm_barrier(self):
        self.check(self.lib.arl_comm_barrier(self.ctx, self._s()))


class _RawDeviceArray(object):
    """__cuda_array_interface__ holder so torch can view memory owned by the shared library."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = dict(shape=(n,), typestr="<f4", data=(int(ptr), False), version=3, strides=None)


def _wrap_device_f32(ptr, n, device):
    return torch.as_tensor(_RawDeviceArray(ptr, n), device=device)
