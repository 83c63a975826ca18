from accel_rl_b200.sampler.base import Sampler, BaseMbSampler
from accel_rl_b200.sampler.device_sampler import ActsrvAltOvrlpSampler, DeviceSampler, TrajInfo
from accel_rl_b200.sampler.device_sampler_with_eval import AAOEvalSampler
