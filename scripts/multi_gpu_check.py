"""Two or more ranks (torchrun, one per GPU): the generator loops of the other ranks must end up with the TRAINER rank's weights
(stream-ordered ncclBroadcast between two waves) and rank 0's replay must receive every rank's rows.
torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/multi_gpu_check.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from rebel_b200 import rela
from rebel_b200.models import Net2, flatten_state_dict

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
D, F, K = 1, 4, 512


def net(seed):
    torch.manual_seed(seed)
    return Net2(num_faces=F, num_dice=D, n_hidden=256, n_layers=2, use_layer_norm=True)


ids = [rela.comm_unique_id()] if rank == 0 else [None]
dist.broadcast_object_list(ids, src=0)
comm = rela.Comm(ids[0], rank, world, local)
rela.set_generator_comm(comm, 0)
locker = rela.ModelLocker([torch.jit.script(net(100 + rank))], f"cuda:{local}")     # every rank starts from DIFFERENT weights
replay = rela.ValuePrioritizedReplay(capacity=1 << 16, seed=1, alpha=1.0, beta=1.0, prefetch=0, use_priority=False, compressed_values=False)
cfg = rela.RecursiveSolvingParams()
cfg.num_dice, cfg.num_faces, cfg.random_action_prob, cfg.sample_leaf, cfg.concurrent_games = D, F, 0.25, True, K
cfg.subgame_params.num_iters, cfg.subgame_params.max_depth, cfg.subgame_params.linear_update = 64, 2, True
loop = rela.create_cfr_thread(locker, replay, cfg, 1000 * rank)
ctx = rela.Context(); ctx.push_env_thread(loop); ctx.start()


def wait(cond, what, limit=120):
    t0 = time.time()
    while not cond():
        if ctx.error() or time.time() - t0 > limit:
            raise RuntimeError(f"rank {rank}: {what}: {ctx.error()}")
        time.sleep(0.002)


want = [float(np.asarray(flatten_state_dict(net(100).state_dict()), dtype=np.float64).sum())]
wait(lambda: loop.waves >= 6, "first waves")
ok = abs(loop.weights_checksum - want[0]) < 1e-9 * max(1.0, abs(want[0]))
print(f"rank {rank}: after {loop.waves} waves weights v{loop.weights_version} sum {loop.weights_checksum:.9f} (trainer's {want[0]:.9f}) {'OK' if ok else 'MISMATCH'}", flush=True)
assert ok
for step in range(3):                                    # the trainer moves on; the followers must too
    new = net(200 + step)
    if rank == 0:
        locker.update_model(new)
    w = float(np.asarray(flatten_state_dict(new.state_dict()), dtype=np.float64).sum())
    wait(lambda: abs(loop.weights_checksum - w) < 1e-9 * max(1.0, abs(w)), f"weights of update {step}")
    print(f"rank {rank}: update {step} arrived at wave {loop.waves} as v{loop.weights_version}", flush=True)
    dist.barrier()                                       # (a follower receives the NEWEST weights: do not let the trainer run ahead of this check)
if rank == 0:
    wait(lambda: replay.num_add() >= 8 * 2 * K * world, "rows of all ranks")
    n = replay.size()
    print(f"rank 0: replay holds {n} rows after {loop.waves} waves ({replay.num_add()} added; {2 * K * world} per wave from {world} ranks)", flush=True)
dist.barrier()
ctx.terminate()
while not ctx.terminated():
    time.sleep(0.01)
print(f"rank {rank}: loops left together after {loop.waves} waves", flush=True)
rela.set_generator_comm(None, 0)
dist.barrier()
dist.destroy_process_group()
