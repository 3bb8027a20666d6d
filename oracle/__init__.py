"""CPU ORACLE -- test infrastructure, NOT the product.

ctypes binding of ``oracle/sag_oracle.c`` (a plain-C float64 restatement of the reference's per-step
environment loop; see the header of ``sag_oracle.h`` for the pinning status).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import
this module; the product package ``safe_adaptation_gym_b200`` never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libsag_oracle.so")

MAX_OBJ = 32
NONE, HAZARD, VASE, GREMLIN, PILLAR, GOAL, BUTTON, BOX, ROD, BALL = range(10)
POINT, CAR = 0, 1
TASK_NAMES = [
    "catch_goal", "collect", "dribble_ball", "go_to_goal", "go_to_goal_damping", "go_to_goal_motor",
    "go_to_goal_scarce", "haul_box", "press_buttons", "press_buttons_scarce", "push_box", "push_box_scarce",
    "roll_rod", "unsupervised",
]
TASK_ID = {n: i for i, n in enumerate(TASK_NAMES)}


class Config(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "placements_margin", "robot_keepout", "hazards_size", "vases_size", "pillars_size", "gremlins_size",
        "hazards_keepout", "gremlins_keepout", "vases_keepout", "pillars_keepout", "gremlins_travel",
        "robot_ctrl_range_scale", "action_noise", "max_bound")] + [("random_bound", C.c_int), ("max_layout_draws", C.c_int),
                                                                  ("num_gremlins", C.c_int)]


class Obj(C.Structure):
    _fields_ = [("type", C.c_int), ("group", C.c_int), ("x", C.c_double), ("y", C.c_double), ("yaw", C.c_double),
                ("vx", C.c_double), ("vy", C.c_double), ("w", C.c_double), ("keepout", C.c_double)]


class Contact(C.Structure):
    _fields_ = [("ba", C.c_int), ("bb", C.c_int), ("sa", C.c_int), ("sb", C.c_int), ("ga", C.c_int), ("gb", C.c_int),
                ("nx", C.c_double), ("ny", C.c_double), ("px", C.c_double), ("py", C.c_double), ("dist", C.c_double)]


def build(force=False):
    """Compile the oracle with gcc (recipe: oracle/Makefile)."""
    src = os.path.join(_HERE, "sag_oracle.c")
    hdr = os.path.join(_HERE, "sag_oracle.h")
    dm = os.path.join(_HERE, "..", "include", "sag_detmath.h")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(hdr), os.path.getmtime(dm))):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    dp = C.POINTER(C.c_double)
    L.orc_default_config.argtypes = [C.POINTER(Config)]
    L.orc_env_create.restype = C.c_void_p
    L.orc_env_create.argtypes = [C.c_int, C.c_int, C.POINTER(Config)]
    L.orc_env_destroy.argtypes = [C.c_void_p]
    L.orc_env_seed.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32]
    L.orc_env_set_replay.argtypes = [C.c_void_p, dp, C.c_int]
    L.orc_env_replay_pos.argtypes = [C.c_void_p]
    L.orc_env_reset.argtypes = [C.c_void_p, C.c_uint32]
    L.orc_env_step.argtypes = [C.c_void_p, dp, dp, dp, dp, C.POINTER(C.c_int)]
    L.orc_env_observation.argtypes = [C.c_void_p, dp]
    L.orc_env_obs_dim.argtypes = [C.c_void_p]
    L.orc_env_nobj.argtypes = [C.c_void_p]
    L.orc_env_get_robot.argtypes = [C.c_void_p, dp]
    L.orc_env_set_robot.argtypes = [C.c_void_p, dp]
    L.orc_env_get_obj.argtypes = [C.c_void_p, C.c_int, C.POINTER(Obj)]
    L.orc_env_get_robot_ext.argtypes = [C.c_void_p, dp]
    L.orc_env_set_robot_ext.argtypes = [C.c_void_p, dp]
    L.orc_env_set_obj.argtypes = [C.c_void_p, C.c_int, C.POINTER(Obj)]
    L.orc_env_get_task_state.argtypes = [C.c_void_p, dp]
    L.orc_env_set_task_state.argtypes = [C.c_void_p, dp]
    L.orc_env_set_dyn_params.argtypes = [C.c_void_p, C.c_double, C.c_double]
    L.orc_env_set_ctrlrange.argtypes = [C.c_void_p, dp, dp]
    L.orc_env_bound.restype = C.c_double
    L.orc_env_bound.argtypes = [C.c_void_p]
    L.orc_env_new_task.argtypes = [C.c_void_p]
    L.orc_env_get_ctrlrange.argtypes = [C.c_void_p, dp, dp]
    L.orc_phys_set_control.argtypes = [C.c_void_p, dp]
    L.orc_phys_step.argtypes = [C.c_void_p, C.c_int]
    L.orc_phys_forward.argtypes = [C.c_void_p]
    L.orc_phys_ncon.argtypes = [C.c_void_p]
    L.orc_phys_get_contact.argtypes = [C.c_void_p, C.c_int, C.POINTER(Contact)]
    L.orc_phys_sensors.argtypes = [C.c_void_p, dp]
    L.orc_phys_error.argtypes = [C.c_void_p]
    L.orc_phys_time.restype = C.c_double
    L.orc_phys_time.argtypes = [C.c_void_p]
    L.orc_phys_clear.argtypes = [C.c_void_p]
    L.orc_phys_set_mocap_pos.argtypes = [C.c_void_p, C.c_double, C.c_double]
    L.orc_env_get_gremlin_state.argtypes = [C.c_void_p, dp]
    L.orc_env_set_gremlin_state.argtypes = [C.c_void_p, dp]
    L.orc_env_num_gremlins.argtypes = [C.c_void_p]
    L.orc_env_slot_types.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    L.orc_phys_add_obj.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]
    L.orc_lidar.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, dp, dp, dp]
    L.orc_lidar_literal.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, dp, dp, dp]
    L.orc_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.orc_philox_uniform2.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, dp]
    L.orc_draw_placement.argtypes = [dp, C.c_double, C.c_double, C.c_double, dp]
    L.orc_collide_circle_circle.argtypes = [C.c_double] * 6 + [dp]
    L.orc_collide_circle_box.argtypes = [C.c_double] * 8 + [C.c_int, dp]
    L.orc_collide_box_box.argtypes = [C.c_double] * 10 + [dp]
    L.orc_task_nobj.argtypes = [C.c_int]
    L.orc_task_slot_types.argtypes = [C.c_int, C.POINTER(C.c_int)]
    L.orc_batch_rollout.restype = C.c_long
    L.orc_batch_rollout.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, dp, dp]
    L.orc_detmath.argtypes = [C.c_int, dp, dp, C.c_int, dp, dp]
    _lib = L
    return L


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def default_config(**over):
    cfg = Config()
    lib().orc_default_config(C.byref(cfg))
    for k, v in over.items():
        if not hasattr(cfg, k):
            continue  # dead keys such as obstacles_size_noise_scale (world.py:29)
        setattr(cfg, k, v)
    return cfg


class OracleEnv:
    """One environment of the CPU restatement (safe_adaptation_gym.py:21-257 + world.py + tasks)."""

    def __init__(self, robot="point", task="go_to_goal", config=None, seed=0, env_gid=0):
        self.L = lib()
        self.cfg = default_config(**(config or {}))
        self.robot = {"point": POINT, "car": CAR}[robot]
        self.task = TASK_ID[task] if isinstance(task, str) else int(task)
        self.h = self.L.orc_env_create(self.robot, self.task, C.byref(self.cfg))
        self.L.orc_env_seed(self.h, seed, env_gid)
        self._replay = None

    def __del__(self):
        try:
            self.L.orc_env_destroy(self.h)
        except Exception:
            pass

    # rng
    def seed(self, seed, env_gid=0):
        self.L.orc_env_seed(self.h, seed, env_gid)

    def set_replay(self, u):
        self._replay = np.ascontiguousarray(u, dtype=np.float64)
        self.L.orc_env_set_replay(self.h, _dp(self._replay), len(self._replay))

    @property
    def replay_pos(self):
        return self.L.orc_env_replay_pos(self.h)

    def new_task(self):
        """the next reset() belongs to a fresh Task instance: redraws ctrl-range scale / constraint bound (world.py:72-78)"""
        self.L.orc_env_new_task(self.h)

    @property
    def bound(self):
        return self.L.orc_env_bound(self.h)

    @property
    def ctrlrange(self):
        lo, hi = np.zeros(2), np.zeros(2)
        self.L.orc_env_get_ctrlrange(self.h, _dp(lo), _dp(hi))
        return lo, hi

    # env API
    @property
    def obs_dim(self):
        return self.L.orc_env_obs_dim(self.h)

    def reset(self, episode=0):
        return self.L.orc_env_reset(self.h, episode)

    def step(self, action):
        a = np.ascontiguousarray(action, dtype=np.float64)
        obs = np.zeros(self.obs_dim)
        rew = np.zeros(2)
        cost = C.c_double()
        done = C.c_int()
        rc = self.L.orc_env_step(self.h, _dp(a), _dp(obs), _dp(rew), C.byref(cost), C.byref(done))
        return obs, rew, cost.value, bool(done.value), rc

    def observation(self):
        obs = np.zeros(self.obs_dim)
        self.L.orc_env_observation(self.h, _dp(obs))
        return obs

    # state
    @property
    def nobj(self):
        return self.L.orc_env_nobj(self.h)

    @property
    def robot_state(self):
        o = np.zeros(6)
        self.L.orc_env_get_robot(self.h, _dp(o))
        return o

    @robot_state.setter
    def robot_state(self, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        self.L.orc_env_set_robot(self.h, _dp(v))

    @property
    def robot_ext(self):
        """car extras: wheel rates (2), castor quaternion (4)"""
        o = np.zeros(6)
        self.L.orc_env_get_robot_ext(self.h, _dp(o))
        return o

    @robot_ext.setter
    def robot_ext(self, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        self.L.orc_env_set_robot_ext(self.h, _dp(v))

    def get_obj(self, s):
        o = Obj()
        self.L.orc_env_get_obj(self.h, s, C.byref(o))
        return o

    def set_obj(self, s, **kw):
        o = self.get_obj(s)
        for k, v in kw.items():
            setattr(o, k, v)
        self.L.orc_env_set_obj(self.h, s, C.byref(o))

    def objects(self):
        """array [nobj, 8]: type, group, x, y, yaw, vx, vy, w"""
        out = np.zeros((self.nobj, 8))
        for s in range(self.nobj):
            o = self.get_obj(s)
            out[s] = [o.type, o.group, o.x, o.y, o.yaw, o.vx, o.vy, o.w]
        return out

    @property
    def task_state(self):
        o = np.zeros(16)
        self.L.orc_env_get_task_state(self.h, _dp(o))
        return o

    @task_state.setter
    def task_state(self, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        self.L.orc_env_set_task_state(self.h, _dp(v))

    # physics-level
    def set_control(self, u):
        u = np.ascontiguousarray(u, dtype=np.float64)
        self.L.orc_phys_set_control(self.h, _dp(u))

    def phys_step(self, n):
        self.L.orc_phys_step(self.h, n)

    def forward(self):
        self.L.orc_phys_forward(self.h)

    def contacts(self):
        out = []
        for i in range(self.L.orc_phys_ncon(self.h)):
            c = Contact()
            self.L.orc_phys_get_contact(self.h, i, C.byref(c))
            out.append(c)
        return out

    def sensors(self):
        o = np.zeros(24)
        self.L.orc_phys_sensors(self.h, _dp(o))
        return o[:self.obs_dim - 48]

    @property
    def error(self):
        return self.L.orc_phys_error(self.h)

    @property
    def time(self):
        return self.L.orc_phys_time(self.h)

    def set_mocap_pos(self, x, y):
        """data.mocap_pos of the gremlins' mocap bodies (mujoco_bridge.py:232-233; world.py:157-165 writes one value to all)"""
        self.L.orc_phys_set_mocap_pos(self.h, float(x), float(y))

    @property
    def num_gremlins(self):
        return self.L.orc_env_num_gremlins(self.h)

    @property
    def gremlin_state(self):
        """[mocap_pos (2), mocap position seen by the last kinematics pass (2), spawn x, y, yaw per gremlin]"""
        o = np.zeros(4 + 3 * 4)
        self.L.orc_env_get_gremlin_state(self.h, _dp(o))
        return o[:4 + 3 * self.num_gremlins]

    @gremlin_state.setter
    def gremlin_state(self, v):
        o = np.zeros(4 + 3 * 4)
        o[:len(v)] = v
        self.L.orc_env_set_gremlin_state(self.h, _dp(o))

    def slot_types(self):
        t = (C.c_int * 32)()
        self.L.orc_env_slot_types(self.h, t)
        return list(t)[:self.nobj]

    def clear_world(self):
        self.L.orc_phys_clear(self.h)

    def add_obj(self, type_, x, y, yaw=0.0, keepout=0.0, group=0):
        return self.L.orc_phys_add_obj(self.h, type_, x, y, yaw, keepout, group)


def lidar(rx, ry, ryaw, xs, ys, literal=False):
    """one 16-bin pseudo-lidar; literal=True evaluates the line-by-line form of safe_adaptation_gym.py:204-223,
    the default the folded form the environment and the GPU use (include/sag_detmath.h: sag_lidar_bin16)"""
    xs = np.ascontiguousarray(xs, dtype=np.float64)
    ys = np.ascontiguousarray(ys, dtype=np.float64)
    out = np.zeros(16)
    (lib().orc_lidar_literal if literal else lib().orc_lidar)(rx, ry, ryaw, len(xs), _dp(xs), _dp(ys), _dp(out))
    return out


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return [int(x) for x in o]


def philox_uniform2(seed, ctr, episode, gid, stream):
    o = np.zeros(2)
    lib().orc_philox_uniform2(seed, ctr, episode, gid, stream, _dp(o))
    return o


def task_slot_types(task):
    t = (C.c_int * 32)()
    tid = TASK_ID[task] if isinstance(task, str) else task
    lib().orc_task_slot_types(tid, t)
    return list(t)[:lib().orc_task_nobj(tid)]


def batch_rollout(envs, steps, nthreads):
    arr = (C.c_void_p * len(envs))(*[e.h for e in envs])
    sr = C.c_double()
    sc = C.c_double()
    n = lib().orc_batch_rollout(arr, len(envs), steps, nthreads, C.byref(sr), C.byref(sc))
    return n, sr.value, sc.value


def detmath(fn, a, b=None):
    """fn: 'sincos' | 'atan2' | 'log' evaluated by include/sag_detmath.h"""
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b if b is not None else a, dtype=np.float64)
    o1, o2 = np.zeros_like(a), np.zeros_like(a)
    lib().orc_detmath({"sincos": 0, "atan2": 1, "log": 2}[fn], _dp(a), _dp(b), len(a), _dp(o1), _dp(o2))
    return (o1, o2) if fn == "sincos" else o1
