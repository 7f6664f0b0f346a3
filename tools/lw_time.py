"""Timing of the layer-wise wide-model step (eager and CUDA-graph) -- dev tool."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vmap_b200 import synth as vo
from vmap_b200.ensemble import VmapEnsemble
H, R, S = int(os.environ.get("H", 256)), int(os.environ.get("R", 4800)), int(os.environ.get("S", 32))
ens = VmapEnsemble(1, hidden=H, scale=5.0, impl="layerwise")
ens.load_stacked(vo.init_params(1, H, seed=1))
db = {k: v.cuda() for k, v in vo.synthetic_batch(1, R, S, seed=2, n_cam2surf=5).items()}
for _ in range(3): ens.step(db)
torch.cuda.synchronize()
def timeit(fn, n=10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
flop = 6 * (4 * H * H + 220 * H + 63) * R * S
t = timeit(lambda: ens.step(db))
print(f"eager step {t:.3f} ms  ({flop / t / 1e9:.0f} TFLOP/s)")
if os.environ.get("GRAPH", "1") == "1":
    g = ens.capture_step(db)
    for _ in range(2): g.replay()
    t = timeit(g.replay)
    print(f"graph step {t:.3f} ms  ({flop / t / 1e9:.0f} TFLOP/s)")
ens.check_status()
