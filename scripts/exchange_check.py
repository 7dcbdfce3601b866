"""torchrun --nproc-per-node N scripts/exchange_check.py [--points P]: fused in-kernel gradient exchange
(NVLS multimem.red from the backward kernel) vs the NCCL all-reduce of the bucket — equality and timing."""
import argparse, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import torch.distributed as dist
from gaussianip_b200 import multiview, renderer, synthetic, _lib

ap = argparse.ArgumentParser()
ap.add_argument("--points", type=int, default=1_000_000)
ap.add_argument("--res", type=int, default=1024)
ap.add_argument("--views", type=int, default=4)
ap.add_argument("--sh", type=int, default=0)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--algo", default="auto")
ap.add_argument("--precomp", action="store_true", help="precomputed colours + 3D covariances (other gradient rows)")
a = ap.parse_args()
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)

cloud = synthetic.make_cloud(a.points, a.sh, 0)
names = ("xyz", "features_dc", "features_rest", "scaling", "rotation", "opacity")


class Model:
    active_sh_degree = a.sh

    def __init__(self, p):
        self.p = p
    get_xyz = property(lambda s: s.p["xyz"])
    get_features = property(lambda s: torch.cat((s.p["features_dc"], s.p["features_rest"]), dim=1))
    get_opacity = property(lambda s: torch.sigmoid(s.p["opacity"]))
    get_scaling = property(lambda s: torch.exp(s.p["scaling"]))
    get_rotation = property(lambda s: torch.nn.functional.normalize(s.p["rotation"]))

    def get_covariance(self, scaling_modifier=1.0):
        """[P,6] upper triangle of R S S^T R^T (gaussian_model.py:24-29), differentiable."""
        q = self.get_rotation
        r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
        R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                         2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                         2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=1).view(-1, 3, 3)
        L = R * (scaling_modifier * self.get_scaling)[:, None, :]
        S = L @ L.transpose(1, 2)
        return torch.stack([S[:, 0, 0], S[:, 0, 1], S[:, 0, 2], S[:, 1, 1], S[:, 1, 2], S[:, 2, 2]], dim=1)


n_total = a.views * world
cams = synthetic.ahds_cameras(n_total, a.res, a.res, seed=1, device=dev)
g = torch.Generator().manual_seed(7)
wc, wd, wa = (torch.randn(n_total, c, a.res, a.res, generator=g).to(dev) for c in (3, 1, 1))
bg = torch.zeros(3, device=dev)


def run(fused):
    params = {k: getattr(cloud, k).clone().to(dev).requires_grad_(True) for k in names}
    model = Model(params)
    vp = multiview.ViewParallel(params, a.points, fused_exchange=fused, exchange_algorithm=a.algo)

    class Pipe:
        compute_cov3D_python = a.precomp
        convert_SHs_python = False

    def render_views_fn(views, vsp, exchange=None):
        override = torch.sigmoid(params["features_dc"][:, 0, :]) if a.precomp else None
        return renderer.render_views([cams[v] for v in views], model, Pipe, bg, override_color=override,
                                     screenspace_points=vsp, exchange=exchange)

    def loss_fn(views, out):
        idx = torch.tensor(views, device=dev)
        return (out["render"] * wc[idx]).sum() + (out["depth_3dgs"] * wd[idx]).sum() + (out["alpha_3dgs"] * wa[idx]).sum()

    for _ in range(3):
        out = vp.step_batched(n_total, render_views_fn, loss_fn)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _lib.profile_enable(True)
    e0.record()
    for _ in range(a.steps):
        out = vp.step_batched(n_total, render_views_fn, loss_fn)
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    prof = _lib.profile_read()
    _lib.profile_enable(False)
    if rank == 0:
        print("fused" if fused else "nccl", {k: round(m / c * 1e3, 1) for k, (m, c) in prof.items() if c}, flush=True)
    ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params.values()]
                     + [out["viewspace_grad"].reshape(-1)]).clone()
    return out, flat, float(ms.item())


out_n, flat_n, ms_n = run(False)
try:
    out_f, flat_f, ms_f = run(True)
except Exception as ex:          # e.g. no NVLS multicast on this box
    if rank == 0:
        print("fused exchange unavailable:", repr(ex)[:300])
    dist.destroy_process_group()
    sys.exit(0)
scale = flat_n.abs().max().item()
err = (flat_f - flat_n).abs().max().item()
same_radii = bool(torch.equal(out_n["radii"], out_f["radii"]))
# replicas must agree bit for bit among themselves (every copy received the same adds ... in possibly different order)
chk = flat_f.double().sum().reshape(1).clone()
lo, hi = chk.clone(), chk.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
if rank == 0:
    print({"algo": a.algo, "world": world, "points": a.points, "views_per_rank": a.views, "nccl_ms_per_step": round(ms_n, 3),
           "fused_ms_per_step": round(ms_f, 3), "max_abs_err": err, "grad_scale": scale, "rel": err / scale,
           "radii_equal": same_radii, "loss_nccl": out_n["loss"].item(), "loss_fused": out_f["loss"].item(),
           "replica_checksum_spread": float((hi - lo).item())})
dist.destroy_process_group()
