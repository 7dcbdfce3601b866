"""CPU: the C-ABI library loads and exports every symbol include/gsb.h declares; host-side
logic that needs no device (layout arithmetic, argument errors, view sharding)."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from gaussianip_b200 import _lib, build
    build.build()                      # nvcc cross-compiles sm_100a without a GPU
    return _lib.load()


def header_functions():
    src = open(os.path.join(ROOT, "include", "gsb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gsb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_header_symbol(lib):
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"libgsb.so does not export {n}"
    from gaussianip_b200 import _lib
    assert set(names) == set(_lib.EXPORTS), set(names) ^ set(_lib.EXPORTS)


def test_ctypes_prototypes_match_the_header():
    """Every prototype in _lib.py has as many arguments as the declaration in include/gsb.h, pointers where the
    header has pointers, and the struct mirrors have the header's size (guards against silent ABI drift)."""
    from gaussianip_b200 import _lib
    src = open(os.path.join(ROOT, "include", "gsb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = dict(re.findall(r"\b(gsb_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S))
    assert set(decls) == set(_lib.EXPORTS)
    for name, params in decls.items():
        params = " ".join(params.split())
        args = [] if params in ("", "void") else [a.strip() for a in params.split(",")]
        restype, argtypes = _lib._PROTOS[name]
        assert len(args) == len(argtypes), f"{name}: header has {len(args)} parameters, ctypes {len(argtypes)}"
        for a, t in zip(args, argtypes):
            is_ptr = "*" in a
            t_is_ptr = t is C.c_void_p or t is C.c_char_p or issubclass(t, C._Pointer)
            assert is_ptr == t_is_ptr, f"{name}: '{a}' vs {t}"
            if not is_ptr:
                want = {"int": C.c_int, "long long": C.c_longlong, "size_t": C.c_size_t, "float": C.c_float,
                        "double": C.c_double}[" ".join(a.split()[:-1]).replace("const ", "")]
                assert t is want, f"{name}: '{a}' vs {t}"
    # struct GsbSettings: 9 x 4-byte scalars, padding to 8, 4 pointers
    assert C.sizeof(_lib.GsbSettings) == 40 + 5 * C.sizeof(C.c_void_p)       # 10 x 4-byte scalars, 5 pointers
    m = re.search(r"typedef struct GsbSettings \{(.*?)\} GsbSettings;", src, flags=re.S)
    fields = re.findall(r"(\w+)\s*(?:,|;)", re.sub(r"\b(int32_t|float|const)\b|\*", " ", m.group(1)))
    assert fields == [f[0] for f in _lib.GsbSettings._fields_], fields
    m = re.search(r"typedef struct GsbLayout \{(.*?)\} GsbLayout;", src, flags=re.S)
    names = [n.strip() for grp in re.findall(r"size_t\s+([^;]+);", m.group(1)) for n in grp.split(",")]
    assert names == [f[0] for f in _lib.GsbLayout._fields_]


def test_abi_version_and_strerror(lib):
    from gaussianip_b200 import _lib
    import re
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "gsb.h")).read()
    assert lib.gsb_abi_version() == _lib.ABI_VERSION == int(re.search(r"#define GSB_ABI_VERSION (\d+)", hdr).group(1))
    assert lib.gsb_strerror(0) == b"ok"
    assert b"invalid" in lib.gsb_strerror(-1)


def test_layout_is_consistent(lib):
    from gaussianip_b200 import _lib
    L = _lib.layout(1000, 100, 150, 5000)
    offs = [(n, getattr(L, n)) for n in _lib._LAYOUT_FIELDS if n.startswith("off_")]
    for n, o in offs:
        assert o % 256 == 0, n
    saved = [o for n, o in offs if n in ("off_geom", "off_clamped", "off_counts", "off_point_list", "off_ranges",
                                         "off_n_contrib", "off_final_T", "off_tile_order")]
    assert saved == sorted(saved) and max(saved) < L.saved_bytes
    assert L.off_clamped - L.off_geom >= 1000 * 48
    assert L.off_ranges - L.off_point_list >= 5000 * 4
    T = ((150 + 15) // 16) * ((100 + 15) // 16)
    assert L.off_n_contrib - L.off_ranges >= T * 8
    assert L.scratch_bytes > L.off_ggrad + 1000 * 48 - 1
    # the emission kernel's scan chain (u64: ticket, flag, one word per 1024-Gaussian block and per group of 32
    # blocks) and the sort's status words (blocks + groups of 32, 256 words each, per pass) fit their regions
    for P, D in ((1, 1), (1000, 5000), (1_000_000, 2_200_000), (3_000_000, 1000), (5000, 9_000_000)):
        Lb = _lib.layout(P, 64, 64, D)
        blocks = (P + 1023) // 1024
        assert Lb.off_hist - Lb.off_blocksums >= 8 * (2 + blocks + (blocks + 31) // 32), (P, D)
        n = max(P, D)
        nb = (n + 2047) // 2048
        assert Lb.off_tkeys0 - Lb.off_hist >= 4 * (8 * 256 + 256 + 8 * (nb + (nb + 31) // 32) * 256), (P, D)
    # growth is monotone in every argument
    assert _lib.layout(2000, 100, 150, 5000).saved_bytes > L.saved_bytes
    assert _lib.layout(1000, 100, 150, 9000).scratch_bytes > L.scratch_bytes


def test_invalid_arguments_return_codes(lib):
    from gaussianip_b200 import _lib
    L = _lib.GsbLayout()
    assert lib.gsb_layout(-1, 10, 10, 10, C.byref(L)) == _lib.GSB_E_INVALID
    assert lib.gsb_layout(10, 0, 10, 10, C.byref(L)) == _lib.GSB_E_INVALID
    assert lib.gsb_layout(10, 10, 10, 1 << 40, C.byref(L)) == _lib.GSB_E_UNSUPPORTED
    s = _lib.GsbSettings()            # all-null settings
    assert lib.gsb_render_fwd(C.byref(s), 1, None, 0, None, None, None, None) == _lib.GSB_E_INVALID
    assert lib.gsb_radix_sort_pairs_u32(-5, None, None, None, None, 32, None, None) == _lib.GSB_E_INVALID
    with pytest.raises(_lib.GsbError):
        _lib.check(-1, "unit test")


def test_rasterizer_argument_errors_without_gpu():
    from gaussianip_b200 import rasterizer as R
    rs = R.GaussianRasterizationSettings(32, 32, 0.5, 0.5, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 0,
                                         torch.zeros(3), False, False)
    r = R.GaussianRasterizer(rs)
    z = lambda *s: torch.zeros(*s)
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r(means3D=z(4, 3), means2D=z(4, 3), opacities=z(4, 1), scales=z(4, 3), rotations=z(4, 4))
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r(means3D=z(4, 3), means2D=z(4, 3), opacities=z(4, 1), shs=z(4, 1, 3), colors_precomp=z(4, 3),
          scales=z(4, 3), rotations=z(4, 4))
    with pytest.raises(Exception, match="scale/rotation pair"):
        r(means3D=z(4, 3), means2D=z(4, 3), opacities=z(4, 1), colors_precomp=z(4, 3), scales=z(4, 3),
          rotations=z(4, 4), cov3D_precomp=z(4, 6))
    # CPU tensors are refused loudly: there is no CPU fallback
    with pytest.raises(ValueError, match="CUDA"):
        r(means3D=z(4, 3), means2D=z(4, 3), opacities=z(4, 1), colors_precomp=z(4, 3), scales=z(4, 3),
          rotations=z(4, 4))
    assert R.GaussianRasterizationSettings._fields == (
        "image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix", "projmatrix",
        "sh_degree", "campos", "prefiltered", "debug")


def test_import_shim_resolves_reference_import():
    import diff_gaussian_rasterization as d
    from gaussianip_b200 import rasterizer as R
    assert d.GaussianRasterizer is R.GaussianRasterizer
    assert d.GaussianRasterizationSettings is R.GaussianRasterizationSettings


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gaussianip_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("CPU oracle", "").replace("the oracle", ""), fn


def test_shard_views():
    from gaussianip_b200.multiview import shard_views
    for world in (1, 2, 4, 8):
        got = sorted(v for r in range(world) for v in shard_views(64, r, world))
        assert got == list(range(64))
        sizes = {len(shard_views(64, r, world)) for r in range(world)}
        assert sizes == {64 // world}
    assert shard_views(4, 5, 8) == []
    with pytest.raises(ValueError):
        shard_views(4, 2, 2)


def test_synthetic_workload_shapes():
    from gaussianip_b200 import synthetic
    cl = synthetic.make_cloud(5000, 3, 0)
    assert cl.xyz.shape == (5000, 3) and cl.features_rest.shape == (5000, 15, 3) and cl.xyz.is_contiguous()
    ext = (cl.xyz.max(0).values - cl.xyz.min(0).values)
    assert abs(float(ext.max()) - synthetic.BODY_EXTENT) < 0.05       # poser.py:808-821 normalisation
    assert int(ext.argmax()) == 2                                     # z-up
    cams = synthetic.ahds_cameras(4, 64, 64)
    assert len(cams) == 4 and cams[0].world_view_transform.shape == (4, 4)
    assert len(synthetic.vcr_cameras(64, 32, 32)) == 64
    assert len(synthetic.playback_cameras(136, 32, 32)) == 136


def test_batched_camera_builder_equals_camera():
    import math
    import numpy as np
    from gaussianip_b200.cameras import Camera, cameras_from_c2w, look_at_c2w, orbit_position
    c2ws = [look_at_c2w(orbit_position(az, el, d)) for az, el, d in ((10, 5, 1.4), (-100, -20, 1.6), (170, 29, 1.3))]
    fovys = [0.8, 1.1, 0.7]
    batch = cameras_from_c2w(c2ws, fovys, 96, 128, device="cpu")
    for c2w, fovy, b in zip(c2ws, fovys, batch):
        ref = Camera(c2w, fovy, 96, 128, data_device="cpu")
        assert b.FoVx == ref.FoVx and b.tanfovx == ref.tanfovx and b.image_width == 128
        for name in ("world_view_transform", "projection_matrix", "full_proj_transform", "camera_center"):
            assert torch.equal(getattr(b, name), getattr(ref, name)), name


def test_camera_block_rows_equal_camera_objects():
    """cameras.CameraBlock (fixed-address camera rows for captured steps) holds exactly what Camera(c2w, fovy, H, W)
    computes, including the device copy of the intrinsics the kernels read when a step is replayed from a CUDA
    graph; update() rewrites the same tensors in place."""
    import math
    import numpy as np
    from gaussianip_b200.cameras import Camera, CameraBlock, look_at_c2w, orbit_position
    H, W = 96, 160
    block = CameraBlock(3, H, W, device="cpu")
    ptrs = [(c.world_view_transform.data_ptr(), c.full_proj_transform.data_ptr(), c.tanfov_dev.data_ptr())
            for c in block.cameras]
    for seed in (0, 1):
        rng = np.random.default_rng(seed)
        c2ws = [look_at_c2w(orbit_position(rng.uniform(-180, 180), rng.uniform(-30, 30), rng.uniform(1.3, 1.7)))
                for _ in range(3)]
        fovs = [math.radians(rng.uniform(40, 70)) for _ in range(3)]
        cams = block.update(c2ws, fovs)
        assert [(c.world_view_transform.data_ptr(), c.full_proj_transform.data_ptr(), c.tanfov_dev.data_ptr())
                for c in cams] == ptrs                                   # same addresses after every update
        for cam, c2w, fovy in zip(cams, c2ws, fovs):
            ref = Camera(c2w, fovy, H, W, data_device="cpu")
            assert torch.equal(cam.world_view_transform, ref.world_view_transform)
            assert torch.equal(cam.full_proj_transform, ref.full_proj_transform)
            assert torch.equal(cam.projection_matrix, ref.projection_matrix)
            assert torch.equal(cam.camera_center, ref.camera_center)
            assert cam.FoVx == ref.FoVx and cam.FoVy == ref.FoVy
            want = torch.tensor([math.tan(ref.FoVx * 0.5), math.tan(ref.FoVy * 0.5)], dtype=torch.float32)
            assert torch.equal(cam.tanfov_dev, want)
    with pytest.raises(ValueError):
        block.update(c2ws[:2], fovs[:2])


def test_speculation_capture_hands_out_pinned_count_rows():
    """graph.CapturedStep gives every captured forward its own row of a pinned pool (allocated before the capture)."""
    from gaussianip_b200 import rasterizer as R
    pool = torch.zeros(3, 8, dtype=torch.int32)
    spec = R.speculation(capture=True, counts_pool=pool)
    a, b = spec.take_counts(), spec.take_counts()
    assert a.data_ptr() == pool[0].data_ptr() and b.data_ptr() == pool[1].data_ptr()
    c = spec.take_counts()
    assert c.shape == (8,) and c.data_ptr() == pool[2].data_ptr()
    with pytest.raises(RuntimeError):
        spec.take_counts()
    with R.speculation():
        with pytest.raises(RuntimeError):
            R.speculation().__enter__()                                   # contexts do not nest
