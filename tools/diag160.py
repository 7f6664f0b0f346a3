import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vmap_b200 import synth as vo
from vmap_b200.ensemble import VmapEnsemble
n_obj, R, S = 160, 240, 10
params = vo.init_params(n_obj, 32, seed=2)
b = vo.synthetic_batch(n_obj, R, S, seed=3)
db = {k: v.cuda() for k, v in b.items()}
def mk(impl):
    e = VmapEnsemble(n_obj, hidden=32, scale=2.0, impl=impl); e.load_stacked(params); return e
a, u = mk("fp32"), mk("umma")
a.forward_backward(db); u.forward_backward(db)
torch.cuda.synchronize()
for k in vo.ALL_KEYS:
    ga, gu = a.view(k, a.grads).flatten(1), u.view(k, u.grads).flatten(1)
    err = (ga - gu).norm(dim=1) / (ga.norm(dim=1) + 1e-20)
    w = int(err.argmax())
    print(f"{k:24s} worst obj {w:3d} err {float(err[w]):.3e} |g_a| {float(ga[w].norm()):.3e} |g_u| {float(gu[w].norm()):.3e} median {float(err.median()):.2e} n>0.05: {int((err>0.05).sum())}")
k = "color_linear.0.weight"
ga, gu = a.view(k, a.grads), u.view(k, u.grads)
err = (ga - gu).flatten(1).norm(dim=1) / (ga.flatten(1).norm(dim=1) + 1e-20)
w = int(err.argmax())
print("worst object", w, "loss terms fp32", a.loss_terms[w].tolist(), "umma", u.loss_terms[w].tolist())
print("mask counts", int((b['sem'][w] != 0).sum()), int((b['sem'][w] != 2).sum()), int((b['mask_depth'][w] & (b['sem'][w] != 0)).sum()))
d = (ga[w] - gu[w])
print("err by out row (first 8):", (d.norm(dim=1) / (ga[w].norm(dim=1) + 1e-20))[:8].tolist())
print("err by in col block: h part", float(d[:, :32].norm() / ga[w][:, :32].norm()), "emb2 part", float(d[:, 32:].norm() / ga[w][:, 32:].norm()))
print("top errors:", sorted(err.tolist(), reverse=True)[:8])
# same data run twice with umma: determinism / race check
u2 = mk("umma"); u2.forward_backward(db); torch.cuda.synchronize()
print("umma run-to-run max rel diff", float(((u.grads - u2.grads).flatten(1).norm(dim=1) / (u.grads.flatten(1).norm(dim=1) + 1e-20)).max()))
