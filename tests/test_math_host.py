"""CPU: the shared physics header (navsim_math.h, host build) against libm / CPython, and
the C-ABI library's exported symbols.  No GPU needed."""
import ctypes
import math
import os
import random
import re

import numpy as np

from navbot_ppo_b200 import _capi
from oracle import binding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ulps(a, b):
    if a == b:
        return 0.0
    return abs(a - b) / math.ulp(max(abs(a), abs(b), 1e-300))


def test_sincos_close_to_libm():
    L = binding.lib()
    rng = random.Random(1)
    worst = 0.0
    for _ in range(200000):
        x = rng.uniform(-4.0, 4.0)
        worst = max(worst, _ulps(L.oracle_nv_sin(x), math.sin(x)), _ulps(L.oracle_nv_cos(x), math.cos(x)))
    for x in (0.0, math.pi / 2, -math.pi / 2, math.pi, -math.pi, math.pi / 4, 3.3, -3.3):
        assert abs(L.oracle_nv_sin(x) - math.sin(x)) < 4e-16
        assert abs(L.oracle_nv_cos(x) - math.cos(x)) < 4e-16
    assert L.oracle_nv_sin(0.0) == 0.0 and L.oracle_nv_cos(0.0) == 1.0
    assert worst <= 2.0, worst


def test_atan_close_to_libm():
    L = binding.lib()
    rng = random.Random(2)
    worst = 0.0
    for _ in range(200000):
        x = rng.uniform(-1, 1) * 10 ** rng.uniform(-3, 3)
        worst = max(worst, _ulps(L.oracle_nv_atan(x), math.atan(x)))
    assert worst <= 2.0, worst


def test_pyround_matches_cpython_round():
    L = binding.lib()
    rng = random.Random(3)
    cases = [0.25, 0.35, 2.675, -0.05, 0.5, 1.5, 2.5, -0.5, -1.5, 0.05, 0.15, 1.005, 359.995, -179.995, 0.0, -0.0,
             1e-9, -1e-9, 7.55, -7.65, 3.15, 0.45]
    for _ in range(300000):
        k = rng.choice([0, 1, 2])
        if rng.random() < 0.5:  # near ties, where naive scaling goes wrong
            n = rng.randint(-80000, 80000)
            x = (n + 0.5) / 10 ** k + rng.choice([0.0, 1e-17, -1e-17, 1e-15, -1e-15, 1e-13, -1e-13])
        else:
            x = rng.uniform(-400, 400)
        cases.append((x, k))
    for c in cases:
        x, ks = (c, (0, 1, 2)) if not isinstance(c, tuple) else (c[0], (c[1],))
        for k in ks:
            want = float(round(x)) if k == 0 else round(x, k)
            assert L.oracle_nv_pyround(x, k) == want, (x, k)
            assert L.ref_pyround(x, k) == want, (x, k)


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    L = binding.lib()
    out = (ctypes.c_uint32 * 4)()
    kats = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in kats:
        L.oracle_philox(*ctr, *key, out)
        assert tuple(out) == want


def test_goal_uniform_packing_matches_cpython_random():
    """nv_u53 packs two 32-bit words like random.random(): (a>>5)*2^26 + (b>>6) over 2^53."""
    a, b = 0xdeadbeef, 0x12345678
    want = ((a >> 5) * 67108864.0 + (b >> 6)) / 9007199254740992.0
    assert 0.0 <= want < 1.0
    L = binding.lib()
    L.shim_goal_uniforms.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32,
                                     ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    out = (ctypes.c_uint32 * 4)()
    L.oracle_philox(5, 0, 9, 0, 77, 0, out)
    ux, uy = ctypes.c_double(), ctypes.c_double()
    L.shim_goal_uniforms(77, 9, 5, ctypes.byref(ux), ctypes.byref(uy))
    assert ux.value == ((out[0] >> 5) * 67108864.0 + (out[1] >> 6)) / 9007199254740992.0
    assert uy.value == ((out[2] >> 5) * 67108864.0 + (out[3] >> 6)) / 9007199254740992.0


def test_ref_odometry_matches_python_formula():
    """The C restatement of Env.getOdometry against a line-by-line Python evaluation."""
    L = binding.lib()
    rng = random.Random(4)
    y_, r_, d_ = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
    for _ in range(20000):
        th = rng.uniform(-math.pi, math.pi)
        px, py = rng.uniform(-4, 4), rng.uniform(-4, 4)
        gx, gy = rng.choice([rng.uniform(-3.6, 3.6), round(px, 1)]), rng.choice([rng.uniform(-3.6, 3.6), round(py, 1)])
        qz, qw = math.sin(th / 2), math.cos(th / 2)
        yaw = round(math.degrees(math.atan2(2 * (qw * qz), 1 - 2 * (qz * qz))))
        yaw = yaw if yaw >= 0 else yaw + 360
        rx, ry = round(gx - px, 1), round(gy - py, 1)
        if rx > 0 and ry > 0: theta = math.atan(ry / rx)
        elif rx > 0 and ry < 0: theta = 2 * math.pi + math.atan(ry / rx)
        elif rx < 0 and ry < 0: theta = math.pi + math.atan(ry / rx)
        elif rx < 0 and ry > 0: theta = math.pi + math.atan(ry / rx)
        elif rx == 0 and ry > 0: theta = 1 / 2 * math.pi
        elif rx == 0 and ry < 0: theta = 3 / 2 * math.pi
        elif ry == 0 and rx > 0: theta = 0
        else: theta = math.pi
        rel = round(math.degrees(theta), 2)
        diff = yaw - rel
        if 0 <= diff <= 180 or -180 <= diff < 0: diff = round(diff, 2)
        elif diff < -180: diff = round(360 + diff, 2)
        else: diff = round(-360 + diff, 2)
        L.ref_odometry(px, py, qz, qw, gx, gy, ctypes.byref(y_), ctypes.byref(r_), ctypes.byref(d_))
        assert (y_.value, r_.value, d_.value) == (yaw, rel, diff)


def test_library_exports_every_declared_symbol():
    """Every function include/*.h declares is exported by libnavbot_b200.so and bound."""
    L = _capi.lib()
    declared = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        text = open(os.path.join(ROOT, "include", fn)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        declared |= set(re.findall(r"\b(nav[a-z_0-9]+)\s*\(", text))
    assert declared, "no declarations found"
    bound = set(_capi.NAVSIM_SYMBOLS) | set(_capi.NAVPPO_SYMBOLS)
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/ but not exported"
        assert name in bound, f"{name} has no ctypes prototype in _capi"
    assert L.navsim_abi_version() >= 1


def test_default_cfg_holds_the_reference_constants():
    cfg = _capi.default_cfg(7)
    assert cfg.num_agents == 7 and cfg.num_beams == 10
    assert cfg.diag_norm == math.sqrt(2) * (3.8 + 3.8)       # environment_new.py:21
    assert (cfg.arrive_threshold, cfg.collision_range) == (0.2, 0.2)
    assert (cfg.reward_scale, cfg.reward_collide, cfg.reward_arrive) == (500.0, -100.0, 120.0)
    assert list(cfg.reset_rects[:4]) == [1.7, 2.3, -1.2, 1.2]
    assert list(cfg.respawn_rects[:4]) == [1.6, 2.4, -1.4, 1.4]


def test_create_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        return
    h = ctypes.c_void_p()
    cfg = _capi.default_cfg(4)
    rc = _capi.lib().navsim_create(ctypes.byref(h), ctypes.byref(cfg))
    assert rc == -19 and b"no CPU fallback" in _capi.lib().nav_last_error()


def _naive_scan(x, y, th, seg, nb, off=-0.032, rmin=0.12, rmax=3.5, fov=1.5707975):
    """Brute-force fp64 ray cast with libm trig, two-sided walls, no culling."""
    ox, oy = x + off * math.cos(th), y + off * math.sin(th)
    out = []
    for i in range(nb):
        a = th + (-fov + i * (2 * fov) / (nb - 1))
        dx, dy = math.cos(a), math.sin(a)
        best, sin_inc = math.inf, 1.0
        for x0, y0, x1, y1 in seg:
            ex, ey = x1 - x0, y1 - y0
            den = dx * ey - dy * ex
            if den == 0:
                continue
            t = ((x0 - ox) * ey - (y0 - oy) * ex) / den
            u = ((x0 - ox) * dy - (y0 - oy) * dx) / den
            if t >= 0 and 0 <= u <= 1 and t < best:
                best, sin_inc = t, abs(den) / math.hypot(ex, ey)
        out.append((best, sin_inc))
    return np.array(out)


def test_shared_header_raycast_against_bruteforce():
    """Row R: the fp32 culled/inverse-distance ray cast of navsim_math.h agrees with a naive
    fp64 two-sided ray cast, and the +-inf gates fire identically (except within a hair of the
    gate) — i.e. range and back-face culling never change a scan.  The fp32 beam direction
    is good to ~1e-7 rad, so the range error bound is 1e-6 * t / sin(incidence) + 1e-6 m
    (micrometres head-on, tens of micrometres for a beam grazing a wall; the sensor's own
    resolution is 15 mm).  Maps: stage_1, stage_2, a dense house-like box map; 10/36 beams."""
    from navbot_ppo_b200 import maps
    rng = np.random.RandomState(0)
    for name, nb in (("stage_1", 10), ("stage_2", 10), ("dense", 36)):
        if name == "dense":
            seg = maps.synthetic_map(40, seed=5)
            boxes = None
        else:
            seg = maps.get_map(name)
        cfg = _capi.default_cfg(1)
        cfg.num_beams = nb
        sim = binding.OracleSim(cfg, seg)
        checked = 0
        for _ in range(1500):
            lim = 3.8 if name != "dense" else 7.0
            x, y, th = rng.uniform(-lim, lim), rng.uniform(-lim, lim), rng.uniform(-math.pi, math.pi)
            both = _naive_scan(x, y, th, seg, nb)
            want, sin_inc = both[:, 0], both[:, 1]
            # skip poses inside an obstacle (odd number of crossings along some beam is hard to
            # tell cheaply: use the nearest-wall distance instead) or hugging a wall
            if np.min(want) < 0.05:
                continue
            if name == "stage_2" and (abs(abs(x) - 2) < 0.12 and abs(y) < 1.02 or abs(abs(y) - 2) < 0.12 and abs(x) < 1.02):
                continue
            if name == "dense":
                inside = False
                for k in range(0, len(seg), 4):
                    e = seg[k:k + 4]
                    cr = [(e[j][2] - e[j][0]) * (y - e[j][1]) - (e[j][3] - e[j][1]) * (x - 0.032 * math.cos(th) - e[j][0])
                          for j in range(4)]
                    inside |= all(c > -1e-3 for c in cr)
                if inside:
                    continue
            sim.arr["x"][0], sim.arr["y"][0], sim.arr["th"][0] = x, y, th
            got = sim.scan()[0]
            for g_, w_, si in zip(got, want, sin_inc):
                if abs(w_ - 3.5) < 1e-3 or abs(w_ - 0.12) < 1e-3 or si < 1e-3:
                    continue
                if w_ > 3.5:
                    assert g_ == math.inf
                elif w_ < 0.12:
                    assert g_ == -math.inf
                else:
                    assert abs(g_ - w_) < 1e-6 * w_ / si + 1e-6, (name, x, y, th, g_, w_, si)
                    assert g_ == float(np.float32(g_))  # travels as float32
            checked += 1
        assert checked > 800

def test_markstein_division_equals_division_for_every_small_integer():
    """nv_div_scale (three fp64 ops instead of a division) is what nv_pyround1/2 finish with."""
    L = binding.lib()
    L.oracle_nv_div_scale_mismatches.argtypes = [ctypes.c_long]
    L.oracle_nv_div_scale_mismatches.restype = ctypes.c_long
    assert L.oracle_nv_div_scale_mismatches(1 << 26) == 0


def test_tabulated_bearing_features_equal_reference_formula():
    """nv_rel_theta_centideg / nv_diff_angle_centideg (the integer form the step kernel looks up)
    against the restated Env.getOdometry on the whole +-20 m offset grid and every whole-degree yaw."""
    L = binding.lib()
    L.oracle_nv_bearing_mismatches.argtypes = [ctypes.c_int]
    L.oracle_nv_bearing_mismatches.restype = ctypes.c_long
    assert L.oracle_nv_bearing_mismatches(200) == 0
    L.oracle_nv_rel_theta_centideg.argtypes = [ctypes.c_int, ctypes.c_int]
    L.oracle_nv_rel_theta_centideg.restype = ctypes.c_int
    for nx, ny, want in ((10, 10, 4500), (0, 5, 9000), (-3, 0, 18000), (0, 0, 18000), (4, 0, 0), (0, -1, 27000),
                         (10, -10, 31500), (-10, -10, 22500), (-10, 10, 13500)):
        assert L.oracle_nv_rel_theta_centideg(nx, ny) == want
        rx, ry = round(nx / 10, 1), round(ny / 10, 1)
        if rx > 0 and ry > 0:
            assert want == round(round(math.degrees(math.atan(ry / rx)), 2) * 100)
