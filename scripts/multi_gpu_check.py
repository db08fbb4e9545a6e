"""torchrun worker: N-rank point-partitioned LM must reproduce the single-GPU trajectory (1e-9 per iteration)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphite_b200 import binding, synthetic  # noqa: E402
from graphite_b200.distributed import partition_by_point  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
case = sys.argv[1] if len(sys.argv) > 1 else "trafalgar-257"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 30
mode = sys.argv[3] if len(sys.argv) > 3 else "auto"  # schur_mode of the N-rank run (the single-GPU run is matrix-free)
prob = synthetic.make_named(case)
ctx = binding.Context(local)
uid = [binding.Context.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
ctx.comm_init(world, rank, uid[0])
part = partition_by_point(prob, world, rank)
P = binding.problem_from_bal(ctx, part, "f64-f64", partition=True)
traj, res = P.lm(iterations=iters, schur_mode=mode)
cams, pts = P.get_vertices()
ok = True
if rank == 0:
    ctx1 = binding.Context(local)
    P1 = binding.problem_from_bal(ctx1, prob, "f64-f64")
    t1, r1 = P1.lm(iterations=iters)
    c1, _ = P1.get_vertices()
    rel = np.abs(traj[:, 1] - t1[:, 1]) / t1[:, 1]
    same_decisions = np.array_equal(traj[:, 0] == traj[:, 1], t1[:, 0] == t1[:, 1]) and np.array_equal(traj[:, 3], t1[:, 3])
    cam_rel = np.abs(cams - c1).max() / np.abs(c1).max()
    print(f"exchange mode {P.info()['exchange_mode']} (2 = peer memory, 1 = NCCL)")
    print(f"world {world}: max rel cost diff {rel.max():.2e}, cameras rel {cam_rel:.2e}, same decisions {same_decisions}, "
          f"seconds {res['seconds_total']:.4f} vs single {r1['seconds_total']:.4f}")
    ok = rel.max() <= 1e-9 and same_decisions and cam_rel <= 1e-7
    P1.close(); ctx1.close()
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.broadcast(flag, src=0)
# cameras are replicated: every rank must hold the same values bit for bit
c = torch.from_numpy(cams).cuda()
cmax, cmin = c.clone(), c.clone()
dist.all_reduce(cmax, op=dist.ReduceOp.MAX)
dist.all_reduce(cmin, op=dist.ReduceOp.MIN)
same = bool(torch.equal(cmax, cmin))
if rank == 0:
    print("replicated cameras identical across ranks:", same)
    if ok and same and int(flag.item()) == 1:
        print("MULTI_GPU_OK")
P.close(); ctx.close()
dist.destroy_process_group()
sys.exit(0 if (ok and same) else 1)
