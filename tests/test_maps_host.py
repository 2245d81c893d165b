"""Host-side map compiler (SDF collision shapes -> wall segments): stored maps, cylinders as
polygons, and - where the reference assets exist - a re-compilation of the stored box lists."""
import math
import os

import numpy as np
import pytest

from navbot_ppo_b200 import _capi, maps
from oracle import binding

REF = "/root/reference/turtlebot3_simulations/turtlebot3_gazebo"


def test_stored_maps_have_the_surveyed_geometry():
    s1 = maps.get_map("stage_1")
    assert s1.shape == (16, 4)
    # inner faces of the four 8.1 x 0.1 walls at +-4 m: free space |x|, |y| < 3.95
    coords = np.unique(np.round(np.abs(s1), 3))
    assert np.isclose(coords, 3.95, atol=1e-3).any() and np.isclose(coords, 4.05, atol=1e-3).any()
    assert maps.get_map("stage_2").shape == (32, 4) and maps.get_map("house").shape == (208, 4)
    sx, sy, _ = maps.SPAWN["house"]
    d = min(_pt_seg(np.array([sx, sy]), s) for s in maps.get_map("house"))
    assert d > 1.0                                   # the spawn pose is in open space
    with pytest.raises(KeyError):
        maps.get_map("nope")


def _pt_seg(p, s):
    a, b = np.array(s[:2]), np.array(s[2:])
    ab = b - a
    t = np.clip(np.dot(p - a, ab) / np.dot(ab, ab), 0, 1)
    return float(np.linalg.norm(p - (a + t * ab)))


def test_cylinder_polygon_is_closed_ccw_and_ranges_match_the_circle():
    poly = maps.polygon_segments(1.5, 0.0, 0.4, sides=64)
    assert np.allclose(poly[:, 2:], np.roll(poly[:, :2], -1, axis=0))          # closed chain
    area = 0.5 * sum(x0 * y1 - x1 * y0 for x0, y0, x1, y1 in poly - np.array([1.5, 0, 1.5, 0]))
    assert area > 0 and abs(area - math.pi * 0.4 ** 2) < 0.01 * math.pi * 0.4 ** 2
    # LaserScan of the polygon (host build of the physics header) vs the exact ray-circle distance
    cfg = _capi.default_cfg(1)
    cfg.num_beams = 36
    sim = binding.OracleSim(cfg, poly)
    sim.reset()
    r = sim.scan()[0]
    ang = np.linspace(cfg.fov_min, cfg.fov_max, 36)
    ox = cfg.lidar_offset_x
    hit = 0
    for a, got in zip(ang, r):
        dx, dy = math.cos(a), math.sin(a)
        # |o + t d - c|^2 = R^2 with o = (ox, 0), c = (1.5, 0), R = 0.4
        bq = dx * (ox - 1.5)
        disc = bq * bq - ((ox - 1.5) ** 2 - 0.4 ** 2)
        if disc > 1e-4:
            t = -bq - math.sqrt(disc)
            assert abs(got - t) < 2e-3, (a, got, t)
            hit += 1
        elif disc < -1e-4:
            assert np.isinf(got)
    assert hit >= 3


@pytest.mark.skipif(not os.path.exists(REF), reason="reference assets only exist in the build container")
def test_stored_box_lists_recompile_from_the_reference_assets():
    assert maps.compile_sdf(os.path.join(REF, "worlds/train_world1.world")) == maps.map_boxes("stage_1")
    assert maps.compile_sdf(os.path.join(REF, "worlds/train_world_new.world")) == maps.map_boxes("stage_2")
    house = maps.compile_sdf(os.path.join(REF, "models/turtlebot3_house/model.sdf"))
    assert len(house) == 52 and np.allclose(np.array(house), np.array(maps.map_boxes("house")), atol=1e-6)
    cyl = maps.compile_sdf(os.path.join(REF, "models/turtlebot3_world/model.sdf"), want="cylinder")
    assert len(cyl) == 9 and all(abs(c[2] - 0.15) < 1e-9 for c in cyl)
    assert maps.compile_sdf_segments(os.path.join(REF, "models/turtlebot3_world/model.sdf")).shape == (9 * 16, 4)
