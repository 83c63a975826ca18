"""Update-rule descriptors standing in for lasagne.updates.{adam, rmsprop} in optimizer_args
(spec: accel_rl/optimizers/update_methods_stats.py:11-32 rmsprop, :55-87 adam).  The arithmetic is
update_kernel / sync_allreduce_update_kernel (csrc/kernels.cuh, csrc/comm.cuh)."""


class _UpdateMethod(object):
    def __init__(self, name, kind, defaults):
        self.name, self.kind, self.defaults = name, kind, dict(defaults)

    def resolve(self, **overrides):
        args = dict(self.defaults)
        unknown = set(overrides) - set(args) - {"learning_rate"}
        if unknown:
            raise TypeError("%s() got unexpected arguments %s" % (self.name, sorted(unknown)))
        args.update({k: v for k, v in overrides.items() if k != "learning_rate"})
        return args

    def __repr__(self):
        return "<update method %s>" % self.name


adam = _UpdateMethod("adam", 0, dict(beta1=0.9, beta2=0.999, epsilon=1e-8))
rmsprop = _UpdateMethod("rmsprop", 1, dict(rho=0.9, epsilon=1e-6))

BY_NAME = {"adam": adam, "rmsprop": rmsprop}
