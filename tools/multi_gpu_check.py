"""torchrun --nproc-per-node N tools/multi_gpu_check.py
(1) vMAP mode: objects sharded across ranks == the same objects trained on one GPU (no collective).
(2) iMAP mode: rays sharded, counts + gradients all-reduced over NCCL == single-GPU full batch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from vmap_b200 import synth as vo
from vmap_b200.dist import ReplicatedStep, shard_objects
from vmap_b200.ensemble import VmapEnsemble

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


# ---- (1) object sharding -------------------------------------------------------------------
B, R, S = 4 * world, 240, 10
params = vo.init_params(B, 32, seed=1)
batches = [vo.synthetic_batch(B, R, S, seed=10 + i) for i in range(5)]
lo, hi = shard_objects(B, world, rank)
ens = VmapEnsemble(hi - lo, hidden=32, scale=2.0, device=dev)
ens.load_stacked({k: v[lo:hi] for k, v in params.items()})
for b in batches:
    ens.step({k: v[lo:hi].to(dev) for k, v in b.items()})
mine = ens.params.clone()
gathered = [torch.empty_like(mine) for _ in range(world)]
dist.all_gather(gathered, mine)
if rank == 0:
    full = VmapEnsemble(B, hidden=32, scale=2.0, device=dev)
    full.load_stacked(params)
    for b in batches:
        full.step({k: v.to(dev) for k, v in b.items()})
    err = rel(torch.cat(gathered), full.params)
    print(f"[vMAP sharded x{world}] params rel-L2 vs single GPU: {err:.2e}")
    assert err < 2e-3        # fp16-operand path; only the atomics' summation order differs

# ---- (2) iMAP: replicated H=256 model, rays sharded, NCCL all-reduce ----------------------------
R2, S2 = 64 * world, 14
p1 = vo.init_params(1, 256, seed=3)
full_b = vo.synthetic_batch(1, R2, S2, seed=5, n_cam2surf=5)
e = VmapEnsemble(1, hidden=256, scale=5.0, device=dev, impl="fp32")
e.load_stacked(p1)
stepper = ReplicatedStep(e)
sl = slice(rank * 64, (rank + 1) * 64)
local_b = {k: v[:, sl].contiguous().to(dev) for k, v in full_b.items()}
losses = [float(stepper.step(local_b)) for _ in range(3)]
if rank == 0:
    ref = VmapEnsemble(1, hidden=256, scale=5.0, device=dev, impl="fp32")
    ref.load_stacked(p1)
    fb = {k: v.to(dev) for k, v in full_b.items()}
    ref_losses = [float(ref.step(fb)) for _ in range(3)]
    err = rel(e.params, ref.params)
    print(f"[iMAP replicated x{world}] losses {losses} ref {ref_losses} params rel-L2 {err:.2e}")
    assert err < 1e-5 and all(abs(a - b) < 1e-4 * abs(b) for a, b in zip(losses, ref_losses))

# ---- (3) BASELINE configs[4]: iMAP H=256, 4800 rays x 32 samples replicated, layer-wise tensor-core path ----------
R3, S3 = 4800, 32
assert R3 % world == 0
p3 = vo.init_params(1, 256, seed=4)
full3 = vo.synthetic_batch(1, R3, S3, seed=6, n_cam2surf=5)
e3 = VmapEnsemble(1, hidden=256, scale=5.0, device=dev, impl="layerwise")
e3.load_stacked(p3)
st3 = ReplicatedStep(e3)
n_loc = R3 // world
loc3 = {k: v[:, rank * n_loc:(rank + 1) * n_loc].contiguous().to(dev) for k, v in full3.items()}
l3 = [float(st3.step(loc3)) for _ in range(3)]
torch.cuda.synchronize(); dist.barrier()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(20):
    st3.step(loc3)
ev1.record(); torch.cuda.synchronize()
t = torch.tensor([ev0.elapsed_time(ev1) / 20], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    ref3 = VmapEnsemble(1, hidden=256, scale=5.0, device=dev, impl="layerwise")
    ref3.load_stacked(p3)
    fb3 = {k: v.to(dev) for k, v in full3.items()}
    r3 = [float(ref3.step(fb3)) for _ in range(3)]
    ev0.record()
    for _ in range(20):
        ref3.step(fb3)
    ev1.record(); torch.cuda.synchronize()
    t1 = ev0.elapsed_time(ev1) / 20
    print(f"[iMAP cfg4 layer-wise x{world}] losses {l3} single-GPU {r3}")
    print(f"[iMAP cfg4 layer-wise x{world}] {float(t):.3f} ms/step ({R3 / float(t) / 1e3:.2f} M rays/s, max over ranks, incl. NCCL "
          f"all-reduce of counts + 1.28 MB of gradients) vs single GPU full batch {t1:.3f} ms/step ({R3 / t1 / 1e3:.2f} M rays/s)")
    assert all(abs(a - b) < 3e-3 * abs(b) for a, b in zip(l3, r3))
    print("multi-GPU checks OK")
dist.barrier()
dist.destroy_process_group()
