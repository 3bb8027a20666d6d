"""Known-answer tests that pin the oracle independently of the goldens (SURVEY 8c "pins the new repo must create")."""
import numpy as np

import oracle as O


def test_philox4x32_10_random123_vectors():
    assert O.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert O.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert O.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == [
        0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_lidar_closed_form_single_object():
    """safe_adaptation_gym.py:208-222: object at ego angle a, distance d -> bins (bin, bin+1, bin-1)"""
    bs = 2 * np.pi / 16
    for k in range(16):
        for frac, d in ((0.25, 1.0), (0.5, 2.5), (0.9, 4.0)):
            a = (k + frac) * bs
            out = O.lidar(0.3, -0.2, 0.0, [0.3 + d * np.cos(a)], [-0.2 + d * np.sin(a)])
            want = np.zeros(16)
            s = (5 - d) / 5
            want[k] = s; want[(k + 1) % 16] = frac * s; want[(k - 1) % 16] = (1 - frac) * s
            np.testing.assert_allclose(out, want, atol=1e-12)
    assert O.lidar(0, 0, 0, [6.0], [0.0]).max() == 0.0            # beyond LIDAR_MAX_DIST
    np.testing.assert_allclose(O.lidar(0, 0, 0, [0.0], [0.0])[0], 1.0)  # on top of the robot
    # robot yaw rotates the ego frame
    np.testing.assert_allclose(O.lidar(0, 0, np.pi / 2, [0.0], [1.0]), O.lidar(0, 0, 0, [1.0], [0.0]), atol=1e-15)


def test_hazard_boundary_counts_and_goal_distance_is_3d():
    e = O.OracleEnv("point", "go_to_goal", config={"action_noise": 0.0}, seed=1)
    assert e.reset(0) == 0
    # put hazard 0 exactly 0.2 away along x: dist == size counts (world.py:152 `<=`)
    s = e.robot_state
    s[3:] = 0
    e.robot_state = s
    for k in range(e.nobj):
        o = e.get_obj(k)
        if o.type in (O.HAZARD, O.VASE, O.PILLAR):
            e.set_obj(k, x=s[0] + 3.0 + 0.3 * k, y=s[1] + 3.0)
    e.set_obj(0, x=s[0] + 0.2, y=s[1])
    obs, rew, cost, done, rc = e.step([0.0, 0.0])
    assert cost == 1.0 and not done
    # goal-met uses the 3-D distance with dz = 0.06: xy distance 0.2945 is NOT met, 0.2930 is (go_to_goal.py:34-40)
    for dxy, met in ((0.2945, False), (0.2930, True)):
        e2 = O.OracleEnv("point", "go_to_goal", config={"action_noise": 0.0}, seed=2)
        e2.reset(0)
        s = e2.robot_state
        g = [k for k in range(e2.nobj) if e2.get_obj(k).type == O.GOAL][0]
        e2.set_obj(g, x=s[0] + dxy, y=s[1])
        last = e2.task_state[0]
        obs, rew, cost, done, rc = e2.step([0.0, 0.0])
        d3 = np.sqrt(dxy ** 2 + 0.06 ** 2)
        assert abs(rew[0] - ((last - d3) + (1.0 if met else 0.0))) < 1e-12


def test_point_free_space_steady_states():
    """point.xml: thrust saturates at 0.3*0.05 N against damping 0.01 -> terminal speed 1.5 m/s"""
    e = O.OracleEnv("point", "go_to_goal", config={"action_noise": 0.0}, seed=3)
    e.reset(0)
    e.clear_world()
    e.robot_state = [0, 0, 0.3, 0, 0, 0]
    for _ in range(600):
        e.set_control([1.0, 0.0]); e.phys_step(5)
    s = e.robot_state
    assert abs(np.hypot(s[3], s[4]) - 1.5) < 1e-3
    # yaw servo: bounded chatter around the target rate u / 0.3
    e.robot_state = [0, 0, 0, 0, 0, 0]
    ws = []
    for _ in range(300):
        e.set_control([0.0, 0.5]); e.phys_step(1); ws.append(e.robot_state[5])
    assert 0.5 < np.mean(ws[100:]) < 2.5 and max(np.abs(ws)) < 10


def test_collision_primitives():
    L = O.lib()
    import ctypes as C
    out = (C.c_double * 10)()
    assert L.orc_collide_circle_circle(0, 0, 0.1, 0.25, 0, 0.1, out) == 0
    assert L.orc_collide_circle_circle(0, 0, 0.1, 0.15, 0, 0.1, out) == 1
    np.testing.assert_allclose(list(out)[:5], [1, 0, 0.075, 0, -0.05], atol=1e-15)
    assert L.orc_collide_circle_circle(0, 0, 0.1, 0.2, 0, 0.1, out) == 1 and out[4] == 0.0  # touching counts
    assert L.orc_collide_circle_box(0.0, 0.0, 0.1, 0.19, 0.0, 0.0, 0.1, 0.1, 1, out) == 1
    np.testing.assert_allclose(list(out)[:5], [1, 0, 0.095, 0, -0.01], atol=1e-15)
    n = L.orc_collide_box_box(0, 0, 0.0, 0.1, 0.1, 0.19, 0.02, 0.0, 0.1, 0.1, out)
    assert n == 2 and abs(out[4] + 0.01) < 1e-15 and abs(out[9] + 0.01) < 1e-15 and out[0] == 1.0
    n = L.orc_collide_box_box(0, 0, 0.0, 0.1, 0.1, 0.23, 0.0, np.pi / 4, 0.1, 0.1, out)  # corner against face
    assert n == 1 and abs(out[4] - (0.23 - 0.1 - 0.1 * np.sqrt(2))) < 1e-12
