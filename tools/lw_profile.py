"""Launch list of the layer-wise (hidden 64/128/256) step at a small shape, for `ncu --metrics gpu__time_duration.sum`:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/lw_launches.csv \
        python tools/lw_profile.py [hidden rays samples]
Prints the graph-replay time per step as well (CUDA events)."""
import sys
import torch
sys.path.insert(0, ".")
from vmap_b200 import synth as vo
from vmap_b200.ensemble import VmapEnsemble

H = int(sys.argv[1]) if len(sys.argv) > 1 else 256
R = int(sys.argv[2]) if len(sys.argv) > 2 else 600
S = int(sys.argv[3]) if len(sys.argv) > 3 else 32
dev = torch.device("cuda:0")
ens = VmapEnsemble(1, hidden=H, scale=5.0, device=dev)
ens.load_stacked(vo.init_params(1, H, seed=77))
batch = {k: v.to(dev) for k, v in vo.synthetic_batch(1, R, S, seed=900, n_cam2surf=5).items()}
for _ in range(3):
    ens.step(batch)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    ens.forward_backward(batch, fuse_adam=True)
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    g.replay()
e1.record()
torch.cuda.synchronize()
print(f"hidden {H}, {R} rays x {S} samples: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us/step (graph replay)")
torch.cuda.profiler.start() if False else None
ens.step(batch)          # one eager step last: the launch list's tail is exactly one step
torch.cuda.synchronize()
