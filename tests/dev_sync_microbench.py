#!/usr/bin/env python
"""Dev tool, run by hand (2+ GPUs under torch.distributed.run): device time of the fused all-reduce + update step alone."""
import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    from tests.util_gpu import make_policy
    pol, flat, spec = make_policy(1, max_rows=512)
    eng = pol.engine
    def exchange(h):
        out = [None] * world
        dist.all_gather_object(out, h)
        return out
    eng.comm_init(rank, world, exchange)
    eng.opt_configure(algo=0, clip_param=0.2, v_loss_coeff=1.0, ent_loss_coeff=0.01, update=0, learning_rate=1e-3, beta1=0.9,
                      beta2=0.999, epsilon=1e-5, rho=0.9, grad_norm_clip=-1.0)
    eng.reset_opt_state()
    eng.grad.normal_()
    for name, fn in (("sync_allreduce_update+pack", eng.sync_allreduce_update), ("xgpu_barrier", eng.comm_barrier)):
        for _ in range(5):
            fn()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 50
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        if rank == 0:
            print("%s: %.1f us per call (world %d)" % (name, 1e3 * e0.elapsed_time(e1) / reps, world))
    eng.close()
    dist.barrier(); dist.destroy_process_group()

if __name__ == "__main__":
    main()
