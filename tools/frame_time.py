"""Frame-level timing with Replica-sized keyframe buffers: batched sampler (K3) + the 20 optimisation
steps of one frame (train.py:195-326), shipped vMAP shape (120 rays/step) and the BASELINE shape
(1200 rays/step).  Writes one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vmap_b200 import synth as vo
from oracle import sampler_oracle as so
from vmap_b200.ensemble import VmapEnsemble
from vmap_b200.sampler import BatchedSampler, KeyframeSet

dev = torch.device("cuda:0")
B, KF, W, H = int(os.environ.get("B", 20)), 20, 1200, 680
g = torch.Generator().manual_seed(0)
objs = []
for b in range(B):
    rgbs = torch.randint(0, 256, (KF, W, H, 4), generator=g, dtype=torch.uint8)
    rgbs[..., 3] = (torch.rand(KF, W, H, generator=g) * 3).to(torch.uint8).clamp(0, 2)
    depth = torch.rand(KF, W, H, generator=g) * 4 + 0.5
    depth[torch.rand(KF, W, H, generator=g) < 0.1] = 0
    twc = torch.eye(4).repeat(KF, 1, 1); twc[:, :3, 3] = torch.rand(KF, 3, generator=g) - 0.5
    u0 = torch.randint(0, W - 300, (KF,), generator=g).float(); v0 = torch.randint(0, H - 250, (KF,), generator=g).float()
    bbox = torch.stack([u0, u0 + 300, v0, v0 + 250], 1)
    objs.append(KeyframeSet(rgbs.to(dev), depth.to(dev), twc.to(dev), bbox.to(dev), KF, [KF - 2, KF - 1]))
rays = so.camera_ray_dirs(W, H, 600.0, 600.0, 599.5, 339.5).to(dev)
smp = BatchedSampler(dev, n_bins_cam2surface=1, n_bins=9)
ens = VmapEnsemble(B, hidden=32, scale=2.0, device=dev)
ens.load_stacked(vo.init_params(B, 32, seed=0))
out = {}
for name, n_frames, n_pix, n_iter in (("shipped_120rays", 100, 24, 20), ("baseline_1200rays", 1000, 24, 20)):
    R = n_frames * n_pix // n_iter
    for rep in range(3):
        batch = smp.sample(objs, n_frames, n_pix, rays, seed=1, offset=rep)
        for it in range(n_iter):
            ens.step({k: v[:, it * R:(it + 1) * R] for k, v in batch.items()})
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    n = 10
    t_s = t_t = 0.0
    for rep in range(n):
        e[0].record()
        batch = smp.sample(objs, n_frames, n_pix, rays, seed=2, offset=rep)
        e[1].record()
        for it in range(n_iter):
            ens.step({k: v[:, it * R:(it + 1) * R] for k, v in batch.items()})
        e[2].record()
        torch.cuda.synchronize()
        t_s += e[0].elapsed_time(e[1]); t_t += e[1].elapsed_time(e[2])
    out[name] = {"objects": B, "rays_per_object_per_frame": n_frames * n_pix, "sampler_ms": t_s / n, "train_20_steps_ms": t_t / n,
                 "sampler_Mrays_per_s": B * n_frames * n_pix / (t_s / n) / 1e3, "frame_ms": (t_s + t_t) / n}
# the same frame as ONE captured CUDA graph (vmap_b200/frame.py)
from vmap_b200.frame import FrameLoop
for name, n_frames, n_pix, n_iter in (("shipped_120rays", 100, 24, 20), ("baseline_1200rays", 1000, 24, 20)):
    ens2 = VmapEnsemble(B, hidden=32, scale=2.0, device=dev)        # fresh weights: random targets diverge after ~1000 steps
    ens2.load_stacked(vo.init_params(B, 32, seed=0))
    fl = FrameLoop(ens2, smp, n_frames, n_pix, n_iter, rays, seed=3)
    fl.set_objects(objs)
    fl.capture()
    for _ in range(3): fl.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fl.set_objects(objs)          # per-frame host work: refill the pinned tables
        fl.run()
    e1.record(); torch.cuda.synchronize()
    out[name]["frame_graph_ms"] = e0.elapsed_time(e1) / 10
    ens2.check_status()
# CPU sampler baseline: oracle restatement of vmap.py:319-459 for ONE object (the reference loops over objects)
torch.set_num_threads(16)
o0 = objs[0]
cpu = {k: getattr(o0, k).cpu() for k in ("rgbs_batch", "depth_batch", "t_wc_batch", "bbox")}
cfg = so.SamplerCfg()
t0 = time.perf_counter()
for rep in range(5):
    rnd = so.draw_randoms_reference_order(None, KF, [KF - 2, KF - 1], 100, 24, cpu["bbox"], cpu["rgbs_batch"], cpu["depth_batch"], cfg)
    so.sample_from_randoms(rnd, cpu["rgbs_batch"], cpu["depth_batch"], cpu["t_wc_batch"], cpu["bbox"], rays.cpu(), cfg)
out["cpu_sampler_ms_per_object_2400rays"] = (time.perf_counter() - t0) / 5 * 1e3
print(json.dumps(out))
