"""Small helpers with the reference's names (accel_rl/util/misc.py:3-31, :42)."""
import time


class struct(dict):
    """dict whose keys are also attributes (reference: util/misc.py:3-6)."""

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.__dict__ = self

    def copy(self):
        """Structural copy: nested struct/dict/list containers are rebuilt, leaves (arrays, tensors)
        are shared (reference: util/misc.py:8-31)."""
        return struct(**{k: _copy_containers(v) for k, v in self.items()})


def _copy_containers(obj):
    if isinstance(obj, struct):
        return obj.copy()
    if isinstance(obj, dict):
        return {k: _copy_containers(v) for k, v in obj.items()}
    if isinstance(obj, list):
        return [_copy_containers(v) for v in obj]
    return obj


def nbytes_unit(nbytes):
    unit = "B"
    for unit in ("KB", "MB", "GB"):
        nbytes /= 1024.
        if nbytes < 1000:
            break
    return nbytes, unit


def make_seed():
    """A seed in [0, 10000) from clock jitter (reference: util/misc.py:42)."""
    t = time.perf_counter_ns()
    time.sleep((t % 997) * 1e-6)
    return int((time.perf_counter_ns() ^ (t >> 7)) % 10000)
