"""CPU tests: the C-ABI library loads and exports every symbol include/nerf_b200.h declares, and the host-side
logic of the Python mirror (no compute calls, no GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import nerf_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def tn():
    import torch_nerf_b200 as mod

    if not (os.path.exists(mod._lib.LIB_PATH) and os.path.exists(mod._lib.SELFTEST_LIB_PATH)):
        import importlib.util

        spec = importlib.util.spec_from_file_location("nerf_build", os.path.join(ROOT, "torch-nerf_b200", "build.py"))
        b = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(b)
        b.build(verbose=False)
    return mod


def header_symbols(name="nerf_b200.h"):
    text = open(os.path.join(ROOT, "include", name)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nerf_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(tn):
    lib = ctypes.CDLL(tn._lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/nerf_b200.h but not exported"
    assert set(tn._lib.EXPORTED_SYMBOLS) == set(syms), "ctypes prototypes out of sync with the header"
    assert tn._lib.load().nerf_version() >= 100


def test_debug_header_symbols_live_in_their_libraries(tn):
    """include/nerf_b200_debug.h: the timeline hooks are exported by the main library, the tcgen05 self tests and
    micro-benchmarks only by libnerf_b200_selftest.so (they are not part of the product library)."""
    syms = set(header_symbols("nerf_b200_debug.h"))
    assert syms == set(tn._lib.DEBUG_SYMBOLS) | set(tn._lib.SELFTEST_SYMBOLS)
    main = ctypes.CDLL(tn._lib.LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    st = tn._lib.load_selftest()
    for s in tn._lib.DEBUG_SYMBOLS:
        assert hasattr(main, s)
    for s in tn._lib.SELFTEST_SYMBOLS:
        assert hasattr(st, s) and not hasattr(main, s), s


def test_struct_layouts_match_header(tn):
    assert ctypes.sizeof(tn._lib.CameraStruct) == 4 * (4 + 9 + 3 + 3) + 4 * 3
    assert ctypes.sizeof(tn._lib.MlpDims) == 12


def test_make_bins_matches_torch_linspace(tn):
    for near, far, s in ((2.0, 6.0, 64), (0.0, 1.0, 64), (2.0, 6.0, 128), (0.5, 3.5, 32)):
        bins, step = tn.make_bins(near, far, s)
        assert torch.equal(bins, torch.linspace(near, far, s + 1)[:-1])
        assert step == (far - near) / s
        ob, ostep = orc.create_t_bins(near, far, s)
        assert np.array_equal(bins.numpy(), ob) and step == ostep


def test_camera_pack_and_errors(tn):
    c2w = orc.pose_spherical(30.0, -30.0, 4.0)
    cam = tn.PerspectiveCamera({"f_x": 1111.1, "f_y": 1111.1, "img_width": 800, "img_height": 800}, torch.from_numpy(c2w), 2.0, 6.0)
    assert cam.img_width == 800 and cam.img_height == 800 and cam.focal_lengths == (1111.1, 1111.1)
    assert torch.equal(cam.intrinsic, torch.from_numpy(orc.make_intrinsic(1111.1, 1111.1, 800, 800)))
    p = cam.pack(False)
    assert p.cx == 400.0 and p.cy == 400.0 and abs(p.fx - 1111.1) < 1e-3 and p.project_to_ndc == 0
    assert np.allclose(np.array(list(p.rot)).reshape(3, 3), c2w[:3, :3]) and np.allclose(list(p.trans), c2w[:3, 3])
    p = cam.pack(True)
    assert p.project_to_ndc == 1 and abs(p.ndc_sx + 2 * 1111.1 / 800) < 1e-5 and p.ndc_two_near == 4.0
    with pytest.raises(ValueError):
        tn.PerspectiveCamera([1, 2, 3], torch.from_numpy(c2w), 2.0, 6.0)
    with pytest.raises(ValueError):
        tn.PerspectiveCamera(torch.eye(3), torch.from_numpy(c2w), 2.0, 6.0)
    bad = tn.PerspectiveCamera({"f_x": 10.0, "f_y": 11.0, "img_width": 8, "img_height": 8}, torch.from_numpy(c2w), 2.0, 6.0)
    with pytest.raises(ValueError):
        bad.pack(True)
    k = torch.from_numpy(orc.make_intrinsic(50.0, 50.0, 20, 10))
    cam2 = tn.PerspectiveCamera(k, torch.from_numpy(c2w), 0.0, 1.0)
    assert (cam2.img_width, cam2.img_height) == (20, 10)


def test_network_state_dict_layout(tn):
    net = tn.NeRF(63, 27)
    sd = net.state_dict()
    params = orc.init_nerf_params()
    assert list(sd.keys()) == list(params.keys())
    for k in sd:
        assert tuple(sd[k].shape) == params[k].shape
    assert sum(p.numel() for p in net.parameters()) == 595844
    assert net.supports_bf16() and not tn.NeRF(39, 15, 64).supports_bf16()


def test_no_cpu_fallback(tn):
    """The product path refuses CPU tensors instead of silently computing on the host."""
    with pytest.raises(RuntimeError):
        tn.PositionalEncoder(3, 10, True).encode(torch.zeros(4, 3))
    with pytest.raises(RuntimeError):
        tn.QuadratureIntegrator().integrate_along_rays(torch.zeros(2, 4), torch.zeros(2, 4, 3), torch.ones(2, 4))
    with pytest.raises(RuntimeError):
        tn.NeRF(63, 27)(torch.zeros(2, 63), torch.zeros(2, 27))


def test_renderer_argument_errors_and_screen_coords(tn):
    c2w = orc.pose_spherical(30.0, -30.0, 4.0)
    cam = tn.PerspectiveCamera({"f_x": 10.0, "f_y": 10.0, "img_width": 4, "img_height": 3}, torch.from_numpy(c2w), 2.0, 6.0)
    ren = tn.VolumeRenderer(tn.QuadratureIntegrator(), tn.StratifiedSampler(), cam)
    assert np.array_equal(ren.screen_coords.numpy(), orc.screen_coords(3, 4))
    with pytest.raises(ValueError):
        ren.render_scene(None, 4.0, 64, False, 0)
    with pytest.raises(ValueError):
        ren.render_scene(None, 4, (64, 128), False, 0)
    with pytest.raises(ValueError):
        ren.render_scene(None, 4, (64, 128, 3), False, 0, pixel_indices=torch.arange(4))
    with pytest.raises(ValueError):
        tn.PrimitiveCube("not a module")
    with pytest.raises(ValueError):
        tn.PrimitiveBase(encoders=[1])
