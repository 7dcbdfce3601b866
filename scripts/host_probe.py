"""Host enqueue time vs device time of one batched 4-view fwd+bwd step (is the step launch-bound?)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussianip_b200 import synthetic, renderer, _lib, rasterizer

P, RES, V = 1_000_000, 1024, 4
dev = torch.device("cuda", 0)
cl = synthetic.make_cloud(P, 0, 0).to(dev)
cl.active_sh_degree = 0
leaves = (cl.xyz, cl.features_dc, cl.features_rest, cl.scaling, cl.rotation, cl.opacity)
for t in leaves:
    t.requires_grad_(True)
cams = synthetic.ahds_cameras(V, RES, RES, seed=1, device=dev)
bg = torch.zeros(3, device=dev)
g = torch.Generator().manual_seed(2)
wc = torch.randn(V * 3 * RES * RES, generator=g).to(dev)
wd = torch.randn(V * RES * RES, generator=g).to(dev)
wa = torch.randn(V * RES * RES, generator=g).to(dev)


def step():
    out = renderer.render_views(cams, cl, None, bg)
    loss = torch.dot(out["render"].reshape(-1), wc) + torch.dot(out["depth_3dgs"].reshape(-1), wd) + \
        torch.dot(out["alpha_3dgs"].reshape(-1), wa)
    loss.backward()
    for t in leaves:
        t.grad = None


for ms_flag in (True, False):
    rasterizer.set_multistream(ms_flag)
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    N = 8
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    host = []
    e0.record()
    for _ in range(N):
        t0 = time.perf_counter()
        step()
        host.append(time.perf_counter() - t0)
    e1.record()
    torch.cuda.synchronize()
    # isolated: sync before each step, so the host time is pure enqueue and the device time pure execution
    iso_h, iso_d = [], []
    for _ in range(N):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); a.record(); step(); b.record(); iso_h.append(time.perf_counter() - t0)
        torch.cuda.synchronize(); iso_d.append(a.elapsed_time(b))
    print(f"multistream={ms_flag}: back-to-back {e0.elapsed_time(e1) / N:.3f} ms/step (host enqueue {sum(host) / N * 1e3:.3f} ms); "
          f"isolated device {sum(iso_d) / N:.3f} ms, host enqueue {sum(iso_h) / N * 1e3:.3f} ms")

if os.environ.get("HOST_PROFILE"):
    import cProfile, pstats
    rasterizer.set_multistream(True)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(20):
        step()
    pr.disable()
    torch.cuda.synchronize()
    st = pstats.Stats(pr)
    st.sort_stats("cumulative").print_stats(45)
    st.sort_stats("tottime").print_stats(25)
