"""CPU tests of host-side pieces: the ray oracle's known answers and .bin codec, the synthetic generator's
determinism, MetricsAccumulator semantics (depth_metrics.h:259-304)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

from conftest import ROOT

RAYS_SO = os.path.join(ROOT, "oracle", "liboracle_rays.so")


@pytest.fixture(scope="module")
def rays_lib():
    if not os.path.exists(RAYS_SO):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "rays"])
    L = C.CDLL(RAYS_SO)
    L.oracle_rays_hw3.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.oracle_rays_3hw.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.oracle_rays_to_world.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    return L


def _rays(L, K, H, W, planar=False):
    K = np.ascontiguousarray(K, dtype=np.float32)
    out = np.empty((3, H, W) if planar else (H * W, 3), dtype=np.float32)
    (L.oracle_rays_3hw if planar else L.oracle_rays_hw3)(K.ctypes.data, H, W, out.ctypes.data)
    return out


def test_ray_oracle_known_answers(rays_lib):
    H, W = 9, 13
    K = np.array([[100.0, 0, 6.0], [0, 120.0, 4.0], [0, 0, 1]], dtype=np.float32)
    r = _rays(rays_lib, K, H, W)
    assert np.allclose(np.linalg.norm(r, axis=1), 1.0, atol=1e-6)           # unit rays
    centre = r[4 * W + 6]
    assert np.allclose(centre, [0, 0, 1])                                   # principal point looks down +z
    # K^-1 [u v 1]: x/z = (u - cx)/fx, y/z = (v - cy)/fy   (ray_direction_computer.cpp:47-48)
    u, v = 11, 2
    rr = r[v * W + u]
    assert abs(rr[0] / rr[2] - (u - 6.0) / 100.0) < 1e-6 and abs(rr[1] / rr[2] - (v - 4.0) / 120.0) < 1e-6
    planar = _rays(rays_lib, K, H, W, planar=True)
    assert np.array_equal(planar.reshape(3, -1).T, r)                       # the two layouts hold the same rays
    # rotation by identity pose leaves rays unchanged up to re-normalisation
    pose = np.eye(4, dtype=np.float32)
    out = np.empty_like(r)
    rays_lib.oracle_rays_to_world(r.ctypes.data, r.shape[0], pose.ctypes.data, out.ctypes.data)
    assert np.allclose(out, r, atol=1e-7)


def test_ray_bin_codec_roundtrip(pkg, rays_lib, tmp_path):
    H, W = 6, 10
    K = np.array([[50.0, 0, 4.5], [0, 55.0, 2.5], [0, 0, 1]], dtype=np.float32)
    r = _rays(rays_lib, K, H, W)
    f = str(tmp_path / "rays.bin")
    assert pkg.save_ray_directions(r, H, W, f)
    raw = open(f, "rb").read()
    assert len(raw) == 8 + H * W * 3 * 4                                    # int32 H, int32 W, floats
    assert np.frombuffer(raw[:8], dtype="<i4").tolist() == [H, W]
    back, h, w = pkg.load_ray_directions(f)
    assert (h, w) == (H, W) and np.array_equal(back, r)
    # the C oracle's writer (fwrite of the same layout) is byte-identical
    f2 = str(tmp_path / "rays_c.bin")
    rays_lib.oracle_rays_save.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int]
    assert rays_lib.oracle_rays_save(f2.encode(), r.ctypes.data, H, W) == 1
    assert open(f2, "rb").read() == raw
    assert not pkg.save_ray_directions(r[:-1], H, W, f)                     # dimension mismatch -> false


def test_synth_is_deterministic_and_shaped(pkg):
    a = pkg.synth.make_batch(3, 24, 32, seed=5)
    b = pkg.synth.make_batch(3, 24, 32, seed=5)
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert a["pred"].shape == (3, 1, 24, 32) and a["rgb"].shape == (3, 3, 24, 32) and a["K"].shape == (3, 3, 3)
    holes = float((a["gt"] == 0).float().mean())
    assert 0.05 < holes < 0.25
    assert float(a["pred"].min()) >= 0.05 and float(a["pred"].max()) <= 9.99
    assert float(a["rgb"].min()) >= 0 and float(a["rgb"].max()) <= 1
    c = pkg.synth.make_batch(3, 24, 32, seed=6)
    assert not torch.equal(a["pred"], c["pred"])


def test_metrics_accumulator_semantics(oracle):
    acc = oracle.MetricsAccumulator()
    assert acc.average() == {} and acc.count() == 0
    acc.update({"abs_rel": 0.2, "rmse": 1.0})
    acc.update({"abs_rel": 0.4, "rmse": 3.0})
    avg = acc.average()
    assert acc.count() == 2 and abs(avg["abs_rel"] - 0.3) < 1e-6 and abs(avg["rmse"] - 2.0) < 1e-6
    acc.reset()
    assert acc.count() == 0 and acc.average() == {}
