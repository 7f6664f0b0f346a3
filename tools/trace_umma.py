"""Per-stage cycle breakdown of k_step_umma (needs the -DVMB_TRACE build: make -C vmap_b200/csrc trace).
Usage: VMB_LIB=vmap_b200/libvmap_b200_trace.so python tools/trace_umma.py"""
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
ens.forward_backward(batches[0])
torch.cuda.synchronize()
buf = np.zeros((4, 256), dtype=np.int64)
L.vmb_trace_read(buf.ctypes.data_as(C.POINTER(C.c_longlong)))
names = ["tile start", "E0 done(PE)"]
for st in range(6):
    names += [f"st{st} done", f"epi{st} done"]
names[-1] = "heads done"   # after st5 sync comes the heads block
# actual order in the kernel (per tile): start, [sync,epi]x5, sync(st5), heads..., see below
for g in range(2):
    t = buf[g]
    n = int((t != 0).sum())
    print(f"--- group {g}: {n} stamps; per-tile stamps = 30")
    per = int(os.environ.get('TRACE_PER', 30))
    for tile in range(min(3, n // per)):
        seg = t[tile * per:(tile + 1) * per + 1]
        d = np.diff(seg)
        print(f"tile {tile}: total {seg[-1] - seg[0] if seg[-1] else 0} cycles; deltas:", d.tolist())
for g in range(2):
    t = buf[2 + g]
    n = int((t != 0).sum())
    st = t[:n].reshape(-1, 2)
    print(f"--- issuer {g}: per stage (issue cycles), first 24 stages:", (st[:24, 1] - st[:24, 0]).tolist())
    print(f"    gaps between stage starts:", np.diff(st[:25, 0]).tolist())

t = buf[1]
print("coarse (cycles from kernel entry): setup done %d, weights landed %d, tiles done %d, flush done %d, end %d" % tuple(int(t[i] - t[199]) for i in (200, 201, 202, 203, 204)))
print("group0 first tile start %d, last stamp %d" % (int(buf[0][0] - t[199]), int(buf[0][(buf[0] != 0).sum() - 1] - t[199])))
