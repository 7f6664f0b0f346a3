"""Per-stage cycle breakdown of k_step_fused (needs the -DVMB_TRACE build: make -C vmap_b200/csrc trace).
Usage: VMB_LIB=vmap_b200/libvmap_b200_trace.so python tools/trace_fused.py

Stamps per tile (thread 0 of each group): tile start, E0 (PE forward) done, then for each of the 12 MMA stages
[operands handed to the issuer warp, accumulator ready] and the end of the stage's
epilogue; 'render' = heads + volume render + losses + ray gradients in registers; 'pe_bwd'; 'dB issue'."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vmap_b200 import synth as vo
from vmap_b200 import _lib
from vmap_b200.ensemble import VmapEnsemble

B, R, S = 20, 1200, 10
ens = VmapEnsemble(B, hidden=32, scale=2.0, impl="umma")
ens.load_stacked(vo.init_params(B, 32, seed=0))
batches = [{k: v.cuda() for k, v in vo.synthetic_batch(B, R, S, seed=i).items()} for i in range(3)]
for i in range(3):
    ens.step(batches[i])
torch.cuda.synchronize()
L = _lib.lib()
L.vmb_trace_clear()
ens.step(batches[0])
torch.cuda.synchronize()
buf = np.zeros((4, 256), dtype=np.int64)
L.vmb_trace_read(buf.ctypes.data_as(C.POINTER(C.c_longlong)))
names = ["E0 (PE fwd)"]
for st in range(5):
    names += [f"s{st} handoff", f"s{st} wait", f"s{st} epi"]
names += ["s5 handoff", "s5 wait", "render"]
for st in range(6, 11):
    names += [f"s{st} handoff", f"s{st} wait", f"s{st} epi"]
names += ["s11 handoff", "s11 wait", "pe_bwd", "dB handoff"]
per = len(names) + 1          # stamps per tile
for g in range(2):
    t = buf[g]
    n = int((t != 0).sum())
    print(f"--- group {g}: {n} stamps; per-tile stamps = {per}")
    for tile in range(min(4, n // per)):
        seg = t[tile * per:(tile + 1) * per + 1]
        d = np.diff(seg[:per])
        nxt = int(seg[per] - seg[0]) if seg[per] else int(seg[per - 1] - seg[0])
        print(f"tile {tile}: start->next start {nxt} cycles")
        print("   " + ", ".join(f"{nm} {int(x)}" for nm, x in zip(names, d)))
        grp = {"PE fwd": d[0], "fwd stages": d[1:18].sum(), "render": d[18], "bwd stages": d[19:36].sum(), "PE bwd": d[36], "dB handoff": d[37]}
        print("   totals: " + ", ".join(f"{k} {int(v)}" for k, v in grp.items()))
t = buf[3]
print("coarse (cycles from kernel entry): setup+counts done %d, weights landed %d, tiles done %d, flush done %d, grid barrier passed %d, end %d" %
      tuple(int(t[i] - t[199]) for i in (200, 201, 202, 203, 205, 204)))
if os.environ.get("TRACE_E0"):
    t2 = buf[2]
    print("--- PE-forward phase of group 0, thread 0 (hsel = 0: three 4-direction iterations): cycles")
    for tile in range(4):
        seg = t2[16 * tile:16 * tile + 8]
        if seg[0] == 0:
            break
        d = np.diff(seg[seg != 0])
        print(f"tile {tile}: prefetch issue, [ladder, stores(+wait on the previous tile's deferred MMAs at iteration 0)] x3:", d.tolist())
names2 = {210: "barriers initialised", 211: "TMEM allocated + first weight copy issued", 212: "own mask counts done", 213: "all counts done (sync)",
          200: "wgrad accumulators zeroed (setup done)", 201: "first weights landed", 202: "last segment's tiles done", 203: "its gradient row flushed",
          205: "grid barrier passed", 206: "finish: start", 207: "finish: slices reduced + updated", 208: "finish: fenced", 204: "kernel end"}
print("prologue / epilogue detail (cycles from kernel entry):")
for k in (210, 211, 212, 213, 200, 201, 202, 203, 205, 206, 207, 208, 204):
    if t[k]:
        print(f"  {names2[k]:48s} {int(t[k] - t[199])}")
