"""Quick CUDA-event timing of the fused step at BASELINE cfg 2 (dev tool, not bench.py)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vmap_b200 import synth as vo
from vmap_b200.ensemble import VmapEnsemble

B, R, S = int(os.environ.get("B", 20)), int(os.environ.get("R", 1200)), int(os.environ.get("S", 10))
params = vo.init_params(B, 32, seed=0)
batches = [{k: v.cuda() for k, v in vo.synthetic_batch(B, R, S, seed=i).items()} for i in range(4)]
for impl in sys.argv[1:] or ["fp32", "umma"]:
    ens = VmapEnsemble(B, hidden=32, scale=2.0, impl=impl)
    ens.load_stacked(params)
    for i in range(5):
        ens.step(batches[i % 4])
    torch.cuda.synchronize()
    for what in ("step", "fwdbwd", "adam"):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 50
        e0.record()
        for i in range(n):
            if what == "step": ens.step(batches[i % 4])
            elif what == "fwdbwd": ens.forward_backward(batches[i % 4])
            else: ens.adam_step(guard_loss=False)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"{impl:5s} {what:7s} {ms*1000:9.1f} us/iter  -> {B*R/ms*1e3/1e6:8.2f} Mrays/s")
    ens.check_status()
