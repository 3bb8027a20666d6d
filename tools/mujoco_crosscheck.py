#!/usr/bin/env python
"""One-command cross-check of the restated physics (oracle/sag_oracle.c == the CUDA kernels, bit for bit) against the
REAL reference running on MuJoCo.

    python tools/mujoco_crosscheck.py [--reference /path/to/safe-adaptation-gym] [--steps 1000] [--json out.json]
    python tools/mujoco_crosscheck.py --selftest      # harness check without MuJoCo (reference logic over oracle physics)

Needs `dm_control` (MuJoCo), `gym` and `xmltodict` -- none of them exists in the build image, which is why physics parity
is "unpinned" (DESIGN.md 2).  Wherever they are installed this script runs the UNMODIFIED reference
(safe_adaptation_gym.make(...), reset, step) and the oracle side by side:

  1. the reference samples its layout (its own RandomState) and builds its MuJoCo model;
  2. every body's pose is read back through the reference's own accessors (MujocoBridge.body_pos / body_mat,
     mujoco_bridge.py:211-219) and injected into an oracle environment of the same robot / task;
  3. both are stepped with the same action sequence (action_noise = 0) and compared step by step.

Three tiers, as SURVEY.md 7.3 proposes, each with a scripted policy that produces the situation:
  free     random actions; compared until the first contact in either simulator
  static   drive into the pillar; compared from the start until 100 steps after the first contact
  movable  drive into the nearest vase; robot compared as above, plus the vase's displacement

Exit code 0: every tier within TOLERANCE; 1: some tier outside (the table printed says which); 3: MuJoCo stack missing.
The tolerances are the STATED ones of the north star ("point and car trajectories within a stated position / velocity
tolerance over 1000 steps"); they are commitments to be confirmed or revised by the first run of this script.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# max abs error over the compared window: position [m], yaw [rad], linear velocity [m/s]; `disp` = relative error of
# the pushed vase's displacement
TOLERANCE = {
    ("point", "free"): {"pos": 2e-3, "yaw": 2e-2, "vel": 1e-2},      # planar model is exact up to integrator details
    ("point", "static"): {"pos": 2e-2, "yaw": 1e-1, "vel": 1e-1},    # soft-contact solver: PGS (here) vs Newton (MuJoCo)
    ("point", "movable"): {"pos": 5e-2, "yaw": 3e-1, "vel": 2e-1, "disp": 0.3},
    ("car", "free"): {"pos": 5e-2, "yaw": 1e-1, "vel": 1e-1},        # reduced planar differential drive vs 3-D free body
    ("car", "static"): {"pos": 1e-1, "yaw": 3e-1, "vel": 3e-1},
    ("car", "movable"): {"pos": 2e-1, "yaw": 5e-1, "vel": 5e-1, "disp": 0.5},
    # a user-defined task with one gremlin (Task.obstacles[2] = 1): the gremlin's position while it follows its mocap target
    # (weld as three planar rows vs MuJoCo's 6-row weld; `gpos` = max position error of the gremlin, robot compared as in "free")
    ("point", "gremlin"): {"pos": 2e-3, "yaw": 2e-2, "vel": 1e-2, "gpos": 3e-2},
    ("car", "gremlin"): {"pos": 5e-2, "yaw": 1e-1, "vel": 1e-1, "gpos": 3e-2},
}


# body-frame COM offset of the whole robot from its root body origin (uniform-density rule on point.xml / car.xml;
# same numbers as sag_core.cuh kPtMc / kPtM and car_model())
COM_OFFSET = {"point": (1e-4 / (4.0 / 3.0 * np.pi * 1e-3 + 1e-3), 0.0), "car": (0.0, 0.007403508202017577)}


def yaw_of(mat):
    m = np.asarray(mat).reshape(3, 3)
    return float(np.arctan2(m[1, 0], m[0, 0]))


def wrap(a):
    return (a + np.pi) % (2 * np.pi) - np.pi


class RefEnv:
    """The reference environment behind the accessors this script needs."""

    def __init__(self, robot, task, seed, selftest, reference_path, gremlins=0):
        self.selftest = selftest
        self.robot_name = robot
        if selftest:
            sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
            import make_golden as G
            if "dm_control" not in sys.modules:
                G.install_stubs()
            cfg = {"action_noise": 0.0}
            if gremlins:
                cfg["num_gremlins"] = gremlins
            self.env = G.make_env(task, config=cfg, seed=seed, robot=robot)
            np.random.RandomState = G._RealRS   # make_env installs a recording RandomState for the golden generator
        else:
            if reference_path:
                sys.path.insert(0, reference_path)
            import safe_adaptation_gym  # the unmodified reference
            self.env = safe_adaptation_gym.make(robot, task, seed=seed, config={"action_noise": 0.0},
                                                render_lidar_and_collision=False)
            if gremlins:  # a user-defined task: the registry task with Task.obstacles[2] = gremlins (task.py:70)
                from safe_adaptation_gym.benchmark import TASKS
                base = TASKS[task]
                counts = list(base().obstacles)
                cls = type(base.__name__ + "WithGremlins", (base,), {"obstacles": property(lambda self: [counts[0], counts[1], gremlins, counts[3]])})
                self.env.set_task(cls())
            self.env.reset()
        self.bridge = self.env.mujoco_bridge
        self.names = [n for n in self.env._world._layout.keys() if n != "robot"]   # placement order (world.py:83-90)

    def robot(self):
        """x, y, yaw and the velocity of the robot's root body origin (the point the oracle integrates)"""
        if self.selftest:  # raw state of the stand-in physics (yaw unwrapped: sin / cos of yaw + 2 pi differ by an ulp)
            s = self.bridge.env.robot_state
            return np.array([s[0], s[1], s[2], s[3], s[4]])
        p, v = self.bridge.robot_pos(), self.bridge.robot_vel()     # subtree_linvel: velocity of the robot's COM
        yaw = yaw_of(self.bridge.robot_mat())
        w = float(np.asarray(self.bridge.get_sensor('gyro'))[2])    # yaw rate (body z = world z for planar motion)
        cx, cy = COM_OFFSET[self.robot_name]
        ox, oy = cx * np.cos(yaw) - cy * np.sin(yaw), cx * np.sin(yaw) + cy * np.cos(yaw)
        return np.array([p[0], p[1], yaw, v[0] + w * oy, v[1] - w * ox])   # v_origin = v_com - w x r_com

    def body(self, name):
        if self.selftest:  # the stand-in bridge answers body_mat for the robot only; its objects carry their yaw
            o = self.bridge.env.objects()[self.names.index(name)]
            return np.array([o[2], o[3], o[4]])
        p = self.bridge.body_pos(name)
        return np.array([p[0], p[1], yaw_of(self.bridge.body_mat(name))])

    def step(self, a):
        return self.env.step(np.asarray(a, dtype=np.float64))

    def ncontacts(self):
        c = self.bridge.contacts
        c = c() if callable(c) else c
        return sum(1 for g1, g2 in c if ("robot" in str(g1)) != ("robot" in str(g2)) and "floor" not in str(g1) + str(g2))


def mirror_oracle(ref, robot, task, gremlins=0):
    import oracle as O
    o = O.OracleEnv(robot, task, config={"action_noise": 0.0, "num_gremlins": gremlins}, seed=0, env_gid=0)
    assert o.reset(0) == 0
    r = ref.robot()
    st = o.robot_state
    st[:] = [r[0], r[1], r[2], 0.0, 0.0, 0.0]
    o.robot_state = st
    objs = o.objects()
    assert len(ref.names) == len(objs), (ref.names, len(objs))
    for s, name in enumerate(ref.names):
        b = ref.body(name)
        o.set_obj(s, x=float(b[0]), y=float(b[1]), yaw=float(b[2]), vx=0.0, vy=0.0, w=0.0)
    ts = o.task_state    # distances the task remembers (go_to_goal.py:50-57) follow from the injected poses
    if gremlins:  # weld anchors = the poses just injected; mocap bodies at the world origin (fresh physics)
        anchors = []
        for s, name in enumerate(ref.names):
            if name.startswith("gremlins"):
                anchors += list(ref.body(name))
        o.gremlin_state = [0.0, 0.0, 0.0, 0.0] + anchors
    o.forward()
    return o, ts


def policy(tier, o, rng, t):
    import oracle as O
    s = o.robot_state
    objs = o.objects()
    kinds = objs[:, 0].astype(int)
    if tier in ("free", "gremlin"):
        return rng.uniform(-1, 1, 2)
    want = O.PILLAR if tier == "static" else O.VASE
    cand = objs[kinds == want][:, 2:4]
    tgt = cand[np.argmin(np.linalg.norm(cand - s[:2], axis=1))]
    d = tgt - s[:2]
    if o.robot == O.CAR:  # the car's front is body -y (car.xml:19-20)
        err = wrap(np.arctan2(d[1], d[0]) - (s[2] - np.pi / 2))
        fwd = np.clip(1.0 - abs(err), 0.0, 1.0) * 0.02
        return np.clip(np.array([fwd + 0.01 * np.clip(err, -1, 1), fwd - 0.01 * np.clip(err, -1, 1)]), -1, 1)
    err = wrap(np.arctan2(d[1], d[0]) - s[2])
    return np.array([np.clip(1.0 - abs(err), 0.02, 1.0), np.clip(2.0 * err, -1, 1)])


def run_tier(robot, tier, steps, seed, selftest, reference_path):
    import oracle as O
    ng = 1 if tier == "gremlin" else 0
    ref = RefEnv(robot, "go_to_goal", seed, selftest, reference_path, gremlins=ng)
    o, _ = mirror_oracle(ref, robot, "go_to_goal", gremlins=ng)
    gslots = [s for s, name in enumerate(ref.names) if name.startswith("gremlins")]
    rng = np.random.RandomState(100 + seed)
    objs0 = o.objects()
    err = {"pos": 0.0, "yaw": 0.0, "vel": 0.0}
    first_contact, compared = None, 0
    for t in range(steps):
        a = policy(tier, o, rng, t)
        ref.step(a)
        o.step(a)
        touching = len([c for c in o.contacts() if c.ba == 0 or c.bb == 0]) > 0 or ref.ncontacts() > 0
        if touching and first_contact is None:
            first_contact = t
        if tier in ("free", "gremlin") and first_contact is not None:
            break
        if tier != "free" and first_contact is not None and t > first_contact + 100:
            break
        r, s = ref.robot(), o.robot_state
        err["pos"] = max(err["pos"], float(np.hypot(r[0] - s[0], r[1] - s[1])))
        err["yaw"] = max(err["yaw"], float(abs(wrap(r[2] - s[2]))))
        err["vel"] = max(err["vel"], float(np.hypot(r[3] - s[3], r[4] - s[4])))
        for gs in gslots:
            b, oo = ref.body(ref.names[gs]), o.objects()[gs]
            err["gpos"] = max(err.get("gpos", 0.0), float(np.hypot(b[0] - oo[2], b[1] - oo[3])))
        compared += 1
    out = dict(err, steps_compared=compared, first_contact=first_contact)
    if tier == "movable":
        objs1 = o.objects()
        k = int(np.argmax(np.linalg.norm(objs1[:, 2:4] - objs0[:, 2:4], axis=1)))
        d_or = float(np.linalg.norm(objs1[k, 2:4] - objs0[k, 2:4]))
        b = ref.body(ref.names[k])
        d_ref = float(np.linalg.norm(b[:2] - objs0[k, 2:4]))
        out["disp"] = abs(d_or - d_ref) / max(d_ref, 1e-6) if max(d_or, d_ref) > 1e-3 else 0.0
        out["disp_oracle_m"], out["disp_reference_m"] = d_or, d_ref
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default=os.environ.get("SAG_REFERENCE", "/root/reference"))
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--seeds", type=int, default=3)
    ap.add_argument("--selftest", action="store_true", help="reference logic over the oracle's physics: errors must be 0")
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    if not args.selftest:
        missing = []
        for m in ("dm_control", "gym", "xmltodict"):
            try:
                __import__(m)
            except Exception:
                missing.append(m)
        if missing:
            print("MuJoCo cross-check NOT RUN: missing " + ", ".join(missing) + " (physics parity stays unpinned; "
                  "`--selftest` checks the harness itself)")
            return 3
    rows, ok = [], True
    for robot in ("point", "car"):
        for tier in ("free", "static", "movable", "gremlin"):
            worst = {}
            for seed in range(args.seeds):
                r = run_tier(robot, tier, args.steps, seed, args.selftest, args.reference)
                for k, v in r.items():
                    if isinstance(v, float):
                        worst[k] = max(worst.get(k, 0.0), v)
                worst["steps_compared"] = worst.get("steps_compared", 0) + r["steps_compared"]
            tol = TOLERANCE[(robot, tier)]
            within = all(worst.get(k, 0.0) <= (1e-9 if args.selftest else v) for k, v in tol.items())
            ok &= within
            rows.append({"robot": robot, "tier": tier, "worst": worst, "tolerance": tol, "within": within})
            print(f"{robot:5s} {tier:8s} " + " ".join(f"{k}={worst.get(k, 0.0):.3e}(tol {v:g})" for k, v in tol.items()) +
                  f"  steps={worst['steps_compared']}  {'OK' if within else 'OUT OF TOLERANCE'}")
    if args.json:
        json.dump({"selftest": args.selftest, "rows": rows}, open(args.json, "w"), indent=1)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
