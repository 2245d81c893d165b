"""Static obstacle maps as wall-segment arrays.

The reference hands Gazebo an SDF world (turtlebot3_stage_1.launch:8 ->
worlds/train_world1.world); the ray sensor sees the collision boxes of that world.  Here
each collision box that crosses the LiDAR plane (z = 0.182 m, urdf.xacro:8-12,134-138)
becomes four wall segments {x0, y0, x1, y1}.  The named maps below were compiled with
`compile_sdf()` from the reference's assets and are stored as plain box lists
(centre x, centre y, size x, size y, yaw) so the GPU box needs no access to the SDF files:

  stage_1  worlds/train_world1.world:85-252      four 8.1 x 0.1 walls at +-4 m
  stage_2  worlds/train_world_new.world:85-420   8 x 0.2 outer walls + four 2 x 0.2 inner walls
           (turtlebot3_stage_2.launch:8 names a world file that is absent from the
           reference; train_world_new.world is the in-repo "added boxes" map that the goal
           rejection rectangles of environment_new.py:340-343 were written for)
"""
from __future__ import annotations

import math
import re
import xml.etree.ElementTree as ET

import numpy as np

LIDAR_Z = 0.182

# (cx, cy, sx, sy, yaw) per collision box, yaw exactly as printed in the world file.
_BOXES = {
    "stage_1": [
        (4.0, 0.0, 8.1, 0.1, -1.5708),
        (0.0, -4.0, 8.10002, 0.1, 3.14159),
        (-4.0, 0.0, 8.1, 0.1, 1.5708),
        (0.0, 4.0, 8.1, 0.1, 0.0),
    ],
    "stage_2": [
        (4.0, 0.0, 8.0, 0.2, -1.5708),
        (0.0, -4.0, 8.0002, 0.2, 3.14159),
        (-4.0, 0.0, 8.0, 0.2, 1.5708),
        (0.0, 4.0, 8.0, 0.2, 0.0),
        (2.0, 0.0, 2.0, 0.2, -1.5708),
        (0.0, -2.0, 2.0, 0.2, 3.14159),
        (-2.0, 0.0, 2.0, 0.2, 1.5708),
        (0.0, 2.0, 2.0, 0.2, 0.0),
    ],
}


def boxes_to_segments(boxes) -> np.ndarray:
    """Four edges per box, counter-clockwise, as float64 [4*len(boxes), 4]."""
    segs = []
    for cx, cy, sx, sy, yaw in boxes:
        c, s = math.cos(yaw), math.sin(yaw)
        hx, hy = sx / 2.0, sy / 2.0
        corners = [(-hx, -hy), (hx, -hy), (hx, hy), (-hx, hy)]
        pts = [(cx + c * px - s * py, cy + s * px + c * py) for px, py in corners]
        for k in range(4):
            x0, y0 = pts[k]
            x1, y1 = pts[(k + 1) % 4]
            segs.append((x0, y0, x1, y1))
    return np.asarray(segs, dtype=np.float64).reshape(-1, 4)


def get_map(name: str) -> np.ndarray:
    if name not in _BOXES:
        raise KeyError(f"unknown map {name!r}; known: {sorted(_BOXES)}")
    return boxes_to_segments(_BOXES[name])


def map_boxes(name: str):
    return list(_BOXES[name])


def synthetic_map(num_boxes: int, seed: int = 0, extent: float = 7.5) -> np.ndarray:
    """Outer room of `extent` half-width plus `num_boxes - 4` random interior boxes kept
    away from the origin — a stand-in for dense maps (house-like segment counts) in tests
    and beam/segment sweeps."""
    rng = np.random.RandomState(seed)
    e = extent
    boxes = [(e, 0.0, 2 * e + 0.1, 0.1, math.pi / 2), (-e, 0.0, 2 * e + 0.1, 0.1, math.pi / 2),
             (0.0, e, 2 * e + 0.1, 0.1, 0.0), (0.0, -e, 2 * e + 0.1, 0.1, 0.0)]
    while len(boxes) < num_boxes:
        cx, cy = rng.uniform(-e + 1, e - 1, size=2)
        if math.hypot(cx, cy) < 1.5:
            continue
        boxes.append((float(cx), float(cy), float(rng.uniform(0.2, 1.5)), float(rng.uniform(0.1, 0.6)),
                      float(rng.uniform(-math.pi, math.pi))))
    return boxes_to_segments(boxes)


def _floats(text):
    return [float(t) for t in re.split(r"\s+", text.strip()) if t]


def compile_sdf(path: str, model: str | None = None, z_plane: float = LIDAR_Z):
    """SDF world/model -> list of (cx, cy, sx, sy, yaw) for every box *collision* whose
    z-extent contains `z_plane`.  Handles model pose + link pose + collision pose with
    yaw-only rotations (all walls in the reference's worlds are upright boxes)."""
    root = ET.parse(path).getroot()
    out = []
    for mdl in root.iter("model"):
        if model is not None and mdl.get("name") != model:
            continue
        mpose = mdl.find("pose")
        mp = _floats(mpose.text) if mpose is not None else [0.0] * 6
        for link in mdl.findall("link"):
            lpose = link.find("pose")
            lp = _floats(lpose.text) if lpose is not None else [0.0] * 6
            for col in link.findall("collision"):
                box = col.find("geometry/box/size")
                if box is None:
                    continue
                sx, sy, sz = _floats(box.text)
                cpose = col.find("pose")
                cp = _floats(cpose.text) if cpose is not None else [0.0] * 6
                # compose planar transforms model * link * collision
                x, y, z, yaw = 0.0, 0.0, 0.0, 0.0
                for p in (mp, lp, cp):
                    c, s = math.cos(yaw), math.sin(yaw)
                    x, y = x + c * p[0] - s * p[1], y + s * p[0] + c * p[1]
                    z += p[2]
                    yaw += p[5]
                if not (z - sz / 2.0 <= z_plane <= z + sz / 2.0):
                    continue
                if sx >= 50 or sy >= 50:  # ground plane
                    continue
                out.append((x, y, sx, sy, yaw))
        if model is not None:
            break
    return out
