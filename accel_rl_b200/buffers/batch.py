"""Rollout buffers as dicts of device tensors (reference: accel_rl/buffers/batch.py:15-89).

Same structure as the reference — a `struct` of arrays with leading dimension N, row = env*T + t,
plus `segs_view`, the list of per-env views — but the leaves are torch CUDA tensors (allocated
once, overwritten in place every iteration) instead of mp.RawArray-backed numpy arrays: the
environments live on the device, so no host process ever touches them."""
import numpy as np
import torch

from accel_rl_b200.util.misc import struct

_TORCH_DTYPES = {"uint8": torch.uint8, "bool": torch.bool, "float32": torch.float32, "float64": torch.float64,
                 "int8": torch.int8, "int32": torch.int32, "int64": torch.int64}


def build_array(value, length, device):
    v = np.asarray(value)
    if v.dtype == object:
        raise TypeError("Unsupported buffer example data type (values must cast under np.asarray())")
    return torch.zeros((length,) + v.shape, dtype=_TORCH_DTYPES[str(v.dtype)], device=device)


def batch_buffer(example, length, device="cuda"):
    if isinstance(example, dict):
        buf = struct()
        for k, v in example.items():
            buf[k] = batch_buffer(v, length, device)
        return buf
    return build_array(example, length, device)


def buffer_length(buf):
    length = None
    for k, v in buf.items():
        if k == "segs_view" or k.startswith("extra"):
            continue
        n = buffer_length(v) if isinstance(v, dict) else len(v)
        if n is None:
            continue
        if length is None:
            length = n
        elif n != length:
            raise RuntimeError("Different lengths in buffer: {}".format(k))
    return length


def _segment(buf, i, n):
    seg = struct()
    for k, v in buf.items():
        if k == "segs_view" or k.startswith("extra"):
            continue
        seg[k] = _segment(v, i, n) if isinstance(v, dict) else v[i:i + n]
    return seg


def view_segments(buf, segment_length):
    length = buffer_length(buf)
    if length % segment_length != 0:
        raise ValueError("Buffer length ({}) not divisible by requested segment_length ({})".format(
            length, segment_length))
    return [_segment(buf, i, segment_length) for i in range(0, length, segment_length)]


def buffer_with_segs_view(examples, length, segment_length, device="cuda"):
    buf = batch_buffer(examples, length, device)
    buf.segs_view = view_segments(buf, segment_length)
    return buf


def combine_distinct_buffers(buffer_1, buffer_2):
    buf = buffer_1.copy()
    other = buffer_2.copy()
    if "segs_view" in buf and "segs_view" in other:
        segs_2 = other.pop("segs_view")
        assert len(buf.segs_view) == len(segs_2)
        for s1, s2 in zip(buf.segs_view, segs_2):
            s1.update(s2)
    buf.update(other)
    return buf


def count_buffer_size(buf):
    size = 0
    for k, v in buf.items():
        if k == "segs_view":
            continue
        size += count_buffer_size(v) if isinstance(v, dict) else v.numel() * v.element_size()
    return size
