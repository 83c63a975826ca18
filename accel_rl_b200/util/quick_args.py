"""Constructor-argument capture (reference: accel_rl/util/quick_args.py:10-25).

save_args(vars()) inside an __init__ stores every constructor argument named anywhere in the
class's MRO as an attribute of self (optionally `_`-prefixed); retrieve_args(obj) returns the
attributes with leading underscores stripped."""
import inspect

from accel_rl_b200.util.misc import struct


def save_args(values, underscore=False):
    self = values["self"]
    prefix = "_" if underscore else ""
    names = []
    for cls in type(self).__mro__:
        init = cls.__dict__.get("__init__")
        if init is not None:
            try:
                names += list(inspect.getfullargspec(init).args[1:])
            except TypeError:
                pass
    for name in names:
        if name in values:
            setattr(self, prefix + name, values[name])


def retrieve_args(obj, bunch=True):
    args = {k.lstrip("_"): v for k, v in vars(obj).items()}
    return struct(**args) if bunch else args
