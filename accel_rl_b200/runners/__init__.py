from accel_rl_b200.runners.accel_rl import AccelRL, AccelRLEval
from accel_rl_b200.runners.multigpu_rl import AccelRLSync, AccelRLAsync, AccelRLEvalSync, AccelRLEvalAsync
