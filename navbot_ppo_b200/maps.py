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
  house    models/turtlebot3_house/model.sdf     52 boxes = 208 wall segments, 15 x 10.5 m: the
           vendored stand-in for the un-vendored aws_robomaker_small_house_world that
           project_ppo/launch/navbot_small_house.launch:11 names (BASELINE configs[4]).
           Robot spawn (-3, 1): the launch file's (0, 0) (navbot_small_house.launch:5-6) lies
           0.1 m from a wall of THIS model; (-3, 1) is in the open-space family of
           spawn_goal_sampler.py:5 and has 1.1 m clearance
"""

from __future__ import annotations

import math
import re
import xml.etree.ElementTree as ET

import numpy as np

LIDAR_Z = 0.182

# robot spawn pose per map (x, y, yaw): turtlebot3_stage_1.launch:3-5; house: see the docstring
SPAWN = {"stage_1": (0.0, 0.0, 0.0), "stage_2": (0.0, 0.0, 0.0), "house": (-3.0, 1.0, 0.0),
         "turtlebot3_world": (-2.0, -0.5, 0.0)}     # turtlebot3_world.launch:3-5
# goal sampling square (environment_new.py:337) and whether the stage rejection rectangles
# (:340-343) apply; the house has no such rectangles (goals are only used for the features)
GOAL_RANGE = {"stage_1": (-3.6, 3.6, True), "stage_2": (-3.6, 3.6, True), "house": (-4.5, 4.5, False),
              "turtlebot3_world": (-1.9, 1.9, False)}

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
    # models/turtlebot3_house/model.sdf: the 52 box collisions (of 100) that cross the LiDAR plane
    "house": [
        (-0.05, 3.1, 4.5, 0.15, -1.5708),
        (2.299992, 5.21625, 0.267504, 0.15, 1.5708),
        (2.300002, 2.516248, 3.3325, 0.15, 1.5708),
        (7.122235, -0.175, 0.905529, 0.15, 0.0),
        (5.297235, -0.175, 0.944471, 0.15, 0.0),
        (1.96856, -0.17488, 0.812647, 0.15, 0.0),
        (-2.231444, -0.17488, 5.78735, 0.15, 0.0),
        (5.07752, 5.2688, 5.0, 0.15, 3.14159),
        (1.25, 5.26876, 2.75, 0.15, 3.14159),
        (-0.05599, 5.274994, 0.161986, 0.15, 3.14159),
        (-2.555993, 5.275, 4.83801, 0.15, 3.14159),
        (-6.2, 5.275, 2.75, 0.15, 3.14159),
        (-7.500004, 1.99694, 2.29388, 0.15, -1.5708),
        (-7.499996, 4.24694, 2.20612, 0.15, -1.5708),
        (-7.500001, -1.78462, 4.43076, 0.15, -1.5708),
        (-7.499992, 0.71538, 0.569241, 0.15, -1.5708),
        (-6.325, -3.925, 2.5, 0.15, 0.0),
        (-5.15151, -1.49625, 5.0, 0.15, 1.5708),
        (1.125, 0.925, 2.5, 0.15, 0.0),
        (2.300003, 0.926561, 0.146998, 0.15, -1.57091),
        (2.29988, -0.148439, 0.203002, 0.15, -1.57091),
        (3.59994, -0.17494, 2.75, 0.15, -4.6e-05),
        (4.9, -2.725, 5.25, 0.15, -1.5708),
        (6.2, -5.275, 2.75, 0.15, 0.0),
        (7.500009, -5.05429, 0.591431, 0.15, 1.5708),
        (7.499999, -2.429285, 4.65857, 0.15, 1.5708),
        (7.500007, 0.50174, 1.50348, 0.15, 1.5708),
        (7.499997, 3.25174, 3.99652, 0.15, 1.5708),
        (-7.167111, 0.925, 0.815777, 0.15, 0.0),
        (-5.467111, 0.925, 0.784223, 0.15, 0.0),
        (-5.149994, 1.47884, 1.25768, 0.15, 1.5708),
        (-5.150007, 4.93695, 0.826108, 0.15, 1.5708),
        (-6.54397, 5.19986, 0.9, 0.01, 0.0),
        (-6.09397, 4.99986, 0.02, 0.4, 0.0),
        (-6.99397, 4.99986, 0.02, 0.4, 0.0),
        (4.72359, 5.18421, 0.9, 0.01, 0.0),
        (5.17359, 4.98421, 0.02, 0.4, 0.0),
        (4.27359, 4.98421, 0.02, 0.4, 0.0),
        (5.64449, 5.18412, 0.9, 0.01, 0.0),
        (6.09449, 4.98412, 0.02, 0.4, 0.0),
        (5.19449, 4.98412, 0.02, 0.4, 0.0),
        (-5.237, -1.57624, 0.02, 0.45, 0.0),
        (-5.472, -1.34124, 0.45, 0.02, 0.0),
        (-5.472, -1.81124, 0.45, 0.02, 0.0),
        (-5.23789, -2.06545, 0.02, 0.45, 0.0),
        (-5.47289, -1.83045, 0.45, 0.02, 0.0),
        (-5.47289, -2.30045, 0.45, 0.02, 0.0),
        (-7.183631, 1.01299, 0.02, 0.45, -1.5708),
        (-6.94863, 1.247989, 0.45, 0.02, -1.5708),
        (-7.41863, 1.247991, 0.45, 0.02, -1.5708),
        (6.35919, -3.19202, 0.042, 0.042, 0.0),
        (6.35903, -2.27759, 0.042, 0.042, 0.0),
    ],
}


# Start poses (x, y, yaw) and goal points (x, y) of the reference's GoalSpawnSampler, per world_type
# (spawn_goal_sampler.py:5-35; data tables, reproduced as data).  `--use_external_sampler` (arguments.py:41) is
# parsed and never read by the reference; VecEnv(use_external_sampler=...) honours it.
SAMPLER_TABLES = {
    "small_house": (
        [(-3.5, 1.0, 0.0), (-3.0, 0.5, 1.57), (-2.5, 1.5, -1.57), (-3.0, 2.0, 0.0),
         (-1.0, 0.0, 0.0), (-0.5, 0.5, 1.57), (0.0, 0.0, -1.57),
         (2.0, 1.5, 3.14), (2.5, 0.5, -1.57), (3.0, 1.0, 0.0),
         (1.0, -2.0, 1.57), (0.5, -2.5, 0.0), (1.5, -2.0, -1.57),
         (-1.5, 2.5, 0.0), (-2.0, 3.0, 1.57),
         (0.0, 1.0, 0.0), (-1.0, 1.5, 1.57), (1.0, 1.0, -1.57)],
        [(-3.5, 0.5), (-3.0, 1.5), (-2.5, 2.0), (-3.5, 2.5), (-4.0, 1.0), (-2.0, 1.0), (-3.0, 0.0),
         (-1.0, 0.5), (-0.5, 0.0), (0.0, 0.5), (-1.5, 0.0), (0.5, 0.0), (-1.0, -0.5),
         (2.0, 0.5), (2.5, 1.0), (3.0, 1.5), (2.0, 2.0), (3.5, 1.0), (2.5, 0.0), (3.0, 0.5),
         (1.0, -2.5), (0.5, -2.0), (1.5, -2.5), (1.0, -3.0), (0.0, -2.5), (1.5, -1.5),
         (-1.5, 2.0), (-2.0, 2.5), (-1.0, 3.0), (-2.5, 2.5), (-1.5, 3.5),
         (0.0, 1.5), (-1.0, 1.0), (1.0, 0.5), (0.5, 1.5), (-0.5, 1.0), (0.0, 2.0), (1.0, 1.5),
         (-4.0, 3.0), (3.5, 2.0), (2.0, -3.0), (-2.0, -1.0)]),
    "stage1": (
        [(0.0, 0.0, 0.0), (0.5, 0.5, 0.785), (-0.5, 0.5, 2.356), (0.5, -0.5, -0.785),
         (-0.5, -0.5, -2.356), (1.0, 0.0, 0.0), (0.0, 1.0, 1.57), (-1.0, 0.0, 3.14),
         (0.0, -1.0, -1.57), (1.0, 1.0, 0.785)],
        [(3.0, 3.0), (3.5, 2.5), (2.5, 3.5), (4.0, 3.0), (-3.0, 3.0), (-3.5, 2.5), (-2.5, 3.5), (-4.0, 3.0),
         (3.0, -3.0), (3.5, -2.5), (2.5, -3.5), (4.0, -3.0), (-3.0, -3.0), (-3.5, -2.5), (-2.5, -3.5), (-4.0, -3.0),
         (4.0, 0.0), (-4.0, 0.0), (0.0, 4.0), (0.0, -4.0), (3.0, 0.0), (-3.0, 0.0), (0.0, 3.0), (0.0, -3.0),
         (2.0, 2.0), (-2.0, 2.0), (2.0, -2.0), (-2.0, -2.0)]),
}
# which table set a named map uses when the caller only says use_external_sampler=True
SAMPLER_FOR_MAP = {"stage_1": "stage1", "stage_2": "stage1", "house": "small_house"}

# upright cylinders (cx, cy, radius) crossing the LiDAR plane: models/turtlebot3_world/model.sdf (the nine pillars
# of the standard TurtleBot3 world; its hexagonal outer wall is a mesh and is not compiled)
_CYLINDERS = {
    "turtlebot3_world": [(-1.1, -1.1, 0.15), (-1.1, 0.0, 0.15), (-1.1, 1.1, 0.15), (0.0, -1.1, 0.15), (0.0, 0.0, 0.15),
                         (0.0, 1.1, 0.15), (1.1, -1.1, 0.15), (1.1, 0.0, 0.15), (1.1, 1.1, 0.15)],
}


def sampler_tables(world_type: str):
    """(starts [n, 3], goals [n, 2]) float64 arrays of GoalSpawnSampler(world_type)."""
    if world_type not in SAMPLER_TABLES:
        raise ValueError(f"Unknown world_type: {world_type}")          # spawn_goal_sampler.py:49
    st, go = SAMPLER_TABLES[world_type]
    return np.asarray(st, dtype=np.float64).reshape(-1, 3), np.asarray(go, dtype=np.float64).reshape(-1, 2)


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
    if name in _CYLINDERS:
        return np.concatenate([polygon_segments(cx, cy, r) for cx, cy, r in _CYLINDERS[name]])
    if name not in _BOXES:
        raise KeyError(f"unknown map {name!r}; known: {sorted(_BOXES) + sorted(_CYLINDERS)}")
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


def polygon_segments(cx: float, cy: float, radius: float, sides: int = 16) -> np.ndarray:
    """Counter-clockwise regular polygon circumscribing a circle (an upright cylinder cut by the
    LiDAR plane): `sides` wall segments, closed, so it can share a map with box obstacles."""
    r = radius / math.cos(math.pi / sides)          # the circle touches every edge from inside
    pts = [(cx + r * math.cos(2 * math.pi * k / sides), cy + r * math.sin(2 * math.pi * k / sides)) for k in range(sides)]
    return np.asarray([(*pts[k], *pts[(k + 1) % sides]) for k in range(sides)], dtype=np.float64)


def compile_sdf_segments(path: str, model: str | None = None, z_plane: float = LIDAR_Z, cylinder_sides: int = 16) -> np.ndarray:
    """SDF world/model -> wall segments of every box AND cylinder collision crossing `z_plane`
    (cylinders as circumscribed `cylinder_sides`-gons).  Mesh collisions are not handled."""
    segs = [boxes_to_segments(compile_sdf(path, model, z_plane))]
    for cx, cy, radius in compile_sdf(path, model, z_plane, want="cylinder"):
        segs.append(polygon_segments(cx, cy, radius, cylinder_sides))
    segs = [s for s in segs if len(s)]
    return np.concatenate(segs) if segs else np.zeros((0, 4))


def compile_sdf(path: str, model: str | None = None, z_plane: float = LIDAR_Z, want: str = "box"):
    """SDF world/model -> list of (cx, cy, sx, sy, yaw) for every box *collision* whose
    z-extent contains `z_plane` (want="cylinder": list of (cx, cy, radius) for upright
    cylinders).  Handles model pose + link pose + collision pose with yaw-only rotations (all
    walls in the reference's worlds are upright boxes)."""
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
                cyl = col.find("geometry/cylinder")
                if want == "cylinder":
                    if cyl is None:
                        continue
                    rad = float(cyl.find("radius").text)
                    sx = sy = 2.0 * rad
                    sz = float(cyl.find("length").text)
                else:
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
                out.append((x, y, sx / 2.0) if want == "cylinder" else (x, y, sx, sy, yaw))
        if model is not None:
            break
    return out
