from accel_rl_b200.spaces.discrete import Discrete
from accel_rl_b200.spaces.uintbox import UintBox
