"""BASELINE configs[2] workload for ncu: 50 objects, hidden 32, depth-guided sampling ON (K3: two launches per frame),
1200 rays x 10 samples per object per step, 20 steps per frame.  Prints CUDA-event times; run under
    ncu --set full --clock-control none -k regex:k_sample -s 4 -c 2 ...      (both sampler passes)
    ncu --set full --clock-control none -k regex:k_step_fused -s 30 -c 1 ... (the step kernel at 50 objects)"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vmap_b200 import synth as vo
from vmap_b200.ensemble import VmapEnsemble
from vmap_b200.sampler import BatchedSampler, KeyframeSet

dev = torch.device("cuda:0")
B, KF, W, H = int(os.environ.get("B", 50)), 6, 1200, 680
N_FRAMES, N_PIX, N_ITER = 1000, 24, 20
g = torch.Generator(device=dev).manual_seed(0)
objs = []
for b in range(B):
    rgbs = torch.randint(0, 256, (KF, W, H, 4), generator=g, dtype=torch.uint8, device=dev)
    rgbs[..., 3] = (torch.rand(KF, W, H, generator=g, device=dev) * 3).to(torch.uint8).clamp(0, 2)
    depth = torch.rand(KF, W, H, generator=g, device=dev) * 4 + 0.5
    depth[torch.rand(KF, W, H, generator=g, device=dev) < 0.1] = 0
    twc = torch.eye(4, device=dev).repeat(KF, 1, 1)
    twc[:, :3, 3] = torch.rand(KF, 3, generator=g, device=dev) - 0.5
    u0 = torch.randint(0, W - 300, (KF,), generator=g, device=dev).float()
    v0 = torch.randint(0, H - 250, (KF,), generator=g, device=dev).float()
    objs.append(KeyframeSet(rgbs, depth, twc, torch.stack([u0, u0 + 300, v0, v0 + 250], 1), KF, [KF - 2, KF - 1]))
u = (torch.arange(W, device=dev) - 599.5) / 600.0
v = (torch.arange(H, device=dev) - 339.5) / 600.0
rays = torch.ones(W, H, 3, device=dev)
rays[:, :, 0] = u[:, None]
rays[:, :, 1] = v
smp = BatchedSampler(dev, n_bins_cam2surface=1, n_bins=9)
ens = VmapEnsemble(B, hidden=32, scale=2.0, device=dev)
ens.load_stacked(vo.init_params(B, 32, seed=0))
R = N_FRAMES * N_PIX // N_ITER
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
t_s = t_t = 0.0
n = 4
for rep in range(2 + n):
    ev[0].record()
    batch = smp.sample(objs, N_FRAMES, N_PIX, rays, seed=1, offset=rep)
    ev[1].record()
    for it in range(N_ITER):
        ens.step({k: x[:, it * R:(it + 1) * R] for k, x in batch.items()})
    ev[2].record()
    torch.cuda.synchronize()
    if rep >= 2:
        t_s += ev[0].elapsed_time(ev[1]); t_t += ev[1].elapsed_time(ev[2])
ens.check_status()
print(f"configs[2]: {B} objects x {R} rays x 10 samples: sampler {t_s / n:.3f} ms per frame ({B * N_FRAMES * N_PIX / (t_s / n) / 1e3:.1f} M rays/s), "
      f"step {t_t / n / N_ITER * 1e3:.1f} us (eager launches) -> {B * R / (t_t / n / N_ITER) / 1e3:.1f} M rays/s")
