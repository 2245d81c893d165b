"""fake_ros.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Runs the reference's *unmodified* `environment_new.Env` (and `ppo.PPO`, `NetActor`,
`NetCritic`) in a container that has no ROS and no Gazebo, by

  * registering import-only stand-ins for the seven ROS modules environment_new.py imports
    (:4-17) and for the three unused imports of ppo.py (:7,13,17), and
  * implementing the process underneath them — gzserver with the diff-drive and ray-sensor
    plugins — as a tiny kinematic + ray-cast world whose arithmetic is oracle/liboracle.so's
    `shim_*` functions, i.e. the SAME `navsim_math.h` the CUDA kernels compile.

`rospy.wait_for_message('scan')` is the reference's clock (environment_new.py:281-286,
352-357): each call integrates the last `cmd_vel` for one LiDAR period (0.2 s), delivers
an Odometry message to the subscriber callback (`Env.getOdometry`) and returns the
LaserScan of the new pose.  `/gazebo/reset_world` puts the robot back at the spawn pose with
zero commanded velocity.  `random.uniform`, which the reference uses unseeded for goal
spawning (:337-345), is redirected per Env instance to the Philox stream the simulator
uses, so goals agree draw for draw.

Only usable where /root/reference exists (this container); the GPU box consumes the
committed tests/golden/*.npz instead.
"""
from __future__ import annotations

import ctypes
import os
import random
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("NAVBOT_REFERENCE_SRC", "/root/reference/project_ppo/src")

_lib = None


def oracle_lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/liboracle.so missing: run `make -C oracle`")
        lib = ctypes.CDLL(path)
        d, pd = ctypes.c_double, ctypes.POINTER(ctypes.c_double)
        lib.shim_drive.argtypes = [pd, pd, pd, d, d, d]
        lib.shim_drive.restype = None
        lib.shim_quat.argtypes = [d, pd, pd]
        lib.shim_quat.restype = None
        vp = ctypes.c_void_p
        lib.shim_beam_table.argtypes = [ctypes.c_int32, d, d, vp, vp]
        lib.shim_beam_table.restype = None
        lib.shim_pack_segments.argtypes = [vp, ctypes.c_int32, d, vp]
        lib.shim_pack_segments.restype = None
        lib.shim_scan.argtypes = [d, d, d, vp, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, vp, vp, d, d, d, pd]
        lib.shim_scan.restype = None
        lib.shim_goal_uniforms.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32, pd, pd]
        lib.shim_goal_uniforms.restype = None
        _lib = lib
    return _lib


def _pd(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


class PhiloxUniform:
    """Drop-in for random.uniform: call 2k is the x draw and call 2k+1 the y draw of Philox
    block k of (seed, agent) — the same stream navsim_math.h::nv_goal_uniforms yields."""

    def __init__(self, seed: int, agent: int):
        self.seed, self.agent, self.draws, self._pending = seed, agent, 0, None

    def __call__(self, a, b):
        if self._pending is None:
            ux, uy = ctypes.c_double(), ctypes.c_double()
            oracle_lib().shim_goal_uniforms(self.seed, self.agent, self.draws, ctypes.byref(ux), ctypes.byref(uy))
            self.draws += 1
            self._pending = uy.value
            u = ux.value
        else:
            u, self._pending = self._pending, None
        return a + (b - a) * u  # CPython Lib/random.py uniform()


class FakeGazebo:
    """gzserver + libgazebo_ros_diff_drive + libgazebo_ros_laser for one robot."""

    def __init__(self, segments, num_beams=10, dt=0.2, lidar_offset_x=-0.032, lidar_min=0.12, lidar_max=3.5,
                 fov_min=-1.5707975, fov_max=1.5707975, start=(0.0, 0.0, 0.0), closed_boxes=True):
        self.seg = np.ascontiguousarray(segments, dtype=np.float64).reshape(-1, 4)
        self.nb, self.dt, self.off = int(num_beams), float(dt), float(lidar_offset_x)
        self.rmin, self.rmax = float(lidar_min), float(lidar_max)
        self.bc = np.zeros(self.nb, np.float32)
        self.bs = np.zeros(self.nb, np.float32)
        oracle_lib().shim_beam_table(self.nb, fov_min, fov_max, self.bc.ctypes.data, self.bs.ctypes.data)
        self.segf = np.zeros((len(self.seg), 8), np.float32)
        oracle_lib().shim_pack_segments(self.seg.ctypes.data, len(self.seg), self.rmax, self.segf.ctypes.data)
        self.closed_boxes = 1 if closed_boxes else 0
        self.start = tuple(float(v) for v in start)
        self.x, self.y, self.th = (ctypes.c_double(v) for v in self.start)
        self.cmd = (0.0, 0.0)
        self.odom_callbacks = []

    # -- services -------------------------------------------------------------------
    def reset_world(self):
        self.x.value, self.y.value, self.th.value = self.start
        self.cmd = (0.0, 0.0)  # the diff-drive plugin's Reset() zeroes its command

    # -- topics ---------------------------------------------------------------------
    def publish_cmd_vel(self, twist):
        self.cmd = (float(twist.linear.x), float(twist.angular.z))

    def next_scan(self):
        lib = oracle_lib()
        lib.shim_drive(ctypes.byref(self.x), ctypes.byref(self.y), ctypes.byref(self.th), self.cmd[0], self.cmd[1], self.dt)
        qz, qw = ctypes.c_double(), ctypes.c_double()
        lib.shim_quat(self.th.value, ctypes.byref(qz), ctypes.byref(qw))
        odom = _Msg()
        odom.pose.pose.position.x = self.x.value
        odom.pose.pose.position.y = self.y.value
        odom.pose.pose.position.z = 0.0
        odom.pose.pose.orientation.x = 0.0
        odom.pose.pose.orientation.y = 0.0
        odom.pose.pose.orientation.z = qz.value
        odom.pose.pose.orientation.w = qw.value
        for cb in self.odom_callbacks:
            cb(odom)
        ranges = np.zeros(self.nb)
        lib.shim_scan(self.x.value, self.y.value, self.th.value, self.segf.ctypes.data, len(self.seg),
                      self.closed_boxes, self.nb, self.bc.ctypes.data, self.bs.ctypes.data, self.off, self.rmin, self.rmax, _pd(ranges))
        scan = _Msg()
        scan.ranges = [float(r) for r in ranges]
        return scan


class _Msg:
    """Attribute bag standing in for any ROS message: unknown fields spring into being
    (nested messages) and numeric leaves default to 0.0 when first read."""

    _LEAVES = {"x", "y", "z", "w"}

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        val = 0.0 if name in _Msg._LEAVES else _Msg()
        object.__setattr__(self, name, val)
        return val


CURRENT: FakeGazebo | None = None  # the world the stubbed rospy talks to


def _make_msg_class(name):
    return type(name, (_Msg,), {})


class _Publisher:
    def __init__(self, topic, cls, queue_size=None):
        self.topic = topic

    def publish(self, msg):
        if self.topic == "cmd_vel":
            CURRENT.publish_cmd_vel(msg)


class _Subscriber:
    def __init__(self, topic, cls, callback):
        if topic == "odom":
            CURRENT.odom_callbacks.append(callback)


class _ServiceProxy:
    def __init__(self, name, cls):
        self.name = name.lstrip("/")

    def __call__(self, *args, **kw):
        if self.name in ("gazebo/reset_world", "gazebo/reset_simulation"):
            CURRENT.reset_world()
        return _Msg()


def install_stubs():
    """Register the import-only stand-ins.  Idempotent."""
    if "rospy" in sys.modules and getattr(sys.modules["rospy"], "_navbot_fake", False):
        return
    rospy = types.ModuleType("rospy")
    rospy._navbot_fake = True
    rospy.Publisher = _Publisher
    rospy.Subscriber = _Subscriber
    rospy.ServiceProxy = _ServiceProxy
    rospy.ServiceException = type("ServiceException", (Exception,), {})
    rospy.ROSException = type("ROSException", (Exception,), {})
    rospy.wait_for_service = lambda *a, **k: None
    rospy.wait_for_message = lambda topic, cls, timeout=None: CURRENT.next_scan()
    rospy.init_node = lambda *a, **k: None
    rospy.is_shutdown = lambda: False
    rospy.sleep = lambda *a, **k: None
    sys.modules["rospy"] = rospy
    sys.modules["roslaunch"] = types.ModuleType("roslaunch")
    for mod, names in {
        "geometry_msgs.msg": ["Twist", "Point", "Pose"],
        "sensor_msgs.msg": ["LaserScan", "Image"],
        "nav_msgs.msg": ["Odometry"],
        "std_srvs.srv": ["Empty"],
        "gazebo_msgs.srv": ["SpawnModel", "DeleteModel", "GetModelState"],
    }.items():
        pkg = mod.split(".")[0]
        sys.modules.setdefault(pkg, types.ModuleType(pkg))
        m = types.ModuleType(mod)
        for n in names:
            setattr(m, n, _make_msg_class(n))
        sys.modules[mod] = m
        setattr(sys.modules[pkg], mod.split(".")[1], m)
    # ppo.py imports these three names and never uses them on the LiDAR path (:7,13,17)
    sys.modules.setdefault("gym", types.ModuleType("gym"))
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
    if "tensorboardX" not in sys.modules:
        tbx = types.ModuleType("tensorboardX")

        class SummaryWriter:  # logging sink; scalars are not part of the parity contract
            def __init__(self, *a, **k):
                pass

            def add_scalar(self, *a, **k):
                pass

            def flush(self):
                pass

            def close(self):
                pass

        tbx.SummaryWriter = SummaryWriter
        sys.modules["tensorboardX"] = tbx
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    sys.dont_write_bytecode = True


def reference_available() -> bool:
    return os.path.exists(os.path.join(REF_SRC, "environment_new.py"))


class RefEnv:
    """One instance of the reference's Env over its own FakeGazebo world and goal stream."""

    def __init__(self, segments, seed=0, agent=0, is_training=True, **world_kw):
        global CURRENT
        install_stubs()
        import environment_new  # the reference's file, unmodified

        self._mod = environment_new
        self.sim = FakeGazebo(segments, **world_kw)
        self.sampler = PhiloxUniform(seed, agent)
        CURRENT = self.sim
        self.env = environment_new.Env(is_training)

    def _enter(self):
        global CURRENT
        CURRENT = self.sim
        random.uniform = self.sampler  # environment_new calls random.uniform(...) by attribute

    def reset(self):
        self._enter()
        return self.env.reset()

    def step(self, action, past_action):
        self._enter()
        return self.env.step(action, past_action)

    def state(self):
        e = self.env
        return dict(x=self.sim.x.value, y=self.sim.y.value, th=self.sim.th.value,
                    gx=e.goal_position.position.x, gy=e.goal_position.position.y,
                    past=e.past_distance, draws=self.sampler.draws)
