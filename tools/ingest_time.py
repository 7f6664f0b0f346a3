"""Timing of the frame ingest (K4) and of sampling from the shared keyframe store vs per-object copies,
Replica frame size (1200 x 680), 20 visible objects.  Writes one JSON line.  Dev / profiling tool."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import ingest_oracle as io
from oracle import sampler_oracle as so
from vmap_b200.keyframes import FrameStore
from vmap_b200.sampler import BatchedSampler, KeyframeSet, KeyframeTables

dev = torch.device("cuda:0")
W, H, KF, B = 1200, 680, 20, int(os.environ.get("B", 20))
BG = [5, 12, 30, 31, 40, 60, 92, 93, 95, 97, 98, 79]
st = FrameStore(W, H, 64, dev, max_id=4096)
rng = np.random.default_rng(0)
frames = []
for f in range(KF):
    inst, cls = io.synthetic_instance_frame(W, H, 40, seed=f)
    frames.append((torch.from_numpy(rng.integers(0, 255, (W, H, 3), dtype=np.uint8)).to(dev),
                   torch.from_numpy(rng.random((W, H), dtype=np.float32) * 4 + 0.5).to(dev),
                   torch.from_numpy(inst).to(dev), torch.from_numpy(cls).to(dev), inst, cls))


def ev_time(fn, n=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


out = {}
rgb, depth, inst_d, cls_d, inst_np, cls_np = frames[0]
T = torch.eye(4, device=dev)
slots = []
def gpu_ingest():
    s, _, _ = st.ingest(rgb, depth, inst_d, T, cls=cls_d, background_cls=BG, bbox_scale=0.2)
    st.release(s)
out["gpu_ingest_ms"] = ev_time(gpu_ingest)
# kernels only: pre-built argument struct, back-to-back C calls (no Python tensor work in the loop)
import ctypes as C
from vmap_b200 import _lib
a = _lib.IngestArgs()
P = lambda t: C.c_void_p(t.data_ptr())
bgt = torch.zeros(100, dtype=torch.uint8); bgt[BG] = 1; bgt = bgt.to(dev)
a.width, a.height, a.inst, a.cls, a.max_id = W, H, P(inst_d), P(cls_d), st.max_id
a.bbox_scale, a.min_extent, a.bg_class, a.n_class = 0.2, 10, P(bgt), 100
a.stats, a.bbox, a.rgb, a.depth = P(st.stats), P(st.bbox), P(rgb), P(depth)
a.dst_rgbx, a.dst_depth, a.dst_inst = P(st.rgbx[63]), P(st.depth[63]), P(st.inst[63])
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
out["gpu_ingest_kernels_ms"] = ev_time(lambda: st.lib.vmb_ingest_frame(st._handle, C.byref(a), stream), n=50)
px = W * H
out["gpu_ingest_algorithmic_bytes"] = px * (4 + 4) + px * (4 + 3 + 4) + px * (4 + 4 + 4)
out["gpu_ingest_GBps"] = out["gpu_ingest_algorithmic_bytes"] / out["gpu_ingest_kernels_ms"] / 1e6
t0 = time.perf_counter()
for _ in range(3): bd, obj = io.replica_frame(inst_np, cls_np, set(BG), 0.2)
out["cpu_loader_loop_ms"] = (time.perf_counter() - t0) / 3 * 1e3
vis = [i for i in sorted(bd) if i != 0][:B]
out["visible_objects"] = len(vis)

# the reference's per-frame work on the data device (train.py:121-141): one state image + 3 full-frame copies per object
bufs = [(torch.empty(KF, W, H, 4, dtype=torch.uint8, device=dev), torch.empty(KF, W, H, device=dev)) for _ in vis]
def ref_style_append():
    for (rgbs, deps), oid in zip(bufs, vis):
        state = torch.zeros_like(inst_d, dtype=torch.uint8)
        state[inst_d == oid] = 1
        state[inst_d == -1] = 2
        rgbs[3, :, :, 0:3] = rgb
        rgbs[3, :, :, 3:4] = state[..., None]
        deps[3] = depth
out["reference_style_append_on_gpu_ms"] = ev_time(ref_style_append, n=10)
out["per_object_copies_MB"] = len(vis) * KF * px * 8 / 1e6
out["shared_store_MB"] = KF * st.bytes_per_frame / 1e6

# sampling: per-object copies vs shared store (same draws)
for f in range(KF):
    r, d, i_d, c_d, _, _ = frames[f]
    slots.append(st.ingest(r, d, i_d, T, cls=c_d, background_cls=BG)[0])
kf_slot = np.tile(np.array(slots, dtype=np.int32), (len(vis), 1))
kf_bbox = np.zeros((len(vis), KF, 4), dtype=np.float32)
u0 = rng.integers(0, W - 300, (len(vis), KF)); v0 = rng.integers(0, H - 250, (len(vis), KF))
kf_bbox[..., 0], kf_bbox[..., 1], kf_bbox[..., 2], kf_bbox[..., 3] = u0, u0 + 300, v0, v0 + 250
tables = KeyframeTables(kf_slot, kf_bbox, vis, [KF] * len(vis), [[KF - 2, KF - 1]] * len(vis))
sets = []
for b, ((rgbs, deps), oid) in enumerate(zip(bufs, vis)):
    for k, s in enumerate(slots):
        rgbs[k, :, :, :3] = st.rgbx[s, :, :, :3]
        rgbs[k, :, :, 3] = (st.inst[s] == oid).to(torch.uint8) + 2 * (st.inst[s] == -1).to(torch.uint8)
        deps[k] = st.depth[s]
    sets.append(KeyframeSet(rgbs, deps, st.t_wc[slots].contiguous(), torch.from_numpy(kf_bbox[b]).to(dev), KF, [KF - 2, KF - 1]))
rays = so.camera_ray_dirs(W, H, 600.0, 600.0, 599.5, 339.5).to(dev)
smp = BatchedSampler(dev, n_bins_cam2surface=1, n_bins=9)
a = {k: v.clone() for k, v in smp.sample(sets, 100, 24, rays, seed=5, offset=1).items()}
s = smp.sample_store(st, tables, 100, 24, rays, seed=5, offset=1)
out["store_vs_copies_identical"] = all(torch.equal(a[k], s[k]) for k in a)
out["sample_copies_ms"] = ev_time(lambda: smp.sample(sets, 100, 24, rays, seed=5, offset=1))
out["sample_store_ms"] = ev_time(lambda: smp.sample_store(st, tables, 100, 24, rays, seed=5, offset=1))
print(json.dumps(out))
