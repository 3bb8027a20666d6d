"""Oracle restatement vs golden vectors produced by the reference's own Python code
(tests/golden/make_golden.py).  CPU only."""
import json
import os

import numpy as np
import pytest

import oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")


def _episodes():
    out = []
    for name in ("episodes.npz", "episodes_gremlins.npz"):  # the second: a user-defined task with Task.obstacles[2] = 2
        raw = np.load(os.path.join(G, name))["data"]
        out += json.loads(raw.tobytes().decode())
    return out


EPISODES = _episodes()
PAD = np.random.RandomState(0).uniform(size=4000).tolist()


def test_lidar_matches_reference_lidar():
    """safe_adaptation_gym.py:174-223 executed by the reference vs orc_lidar"""
    cases = json.load(open(os.path.join(G, "lidar_kat.json")))
    assert len(cases) > 150
    for c in cases:
        pts = np.asarray(c["pts"])
        for literal in (True, False):  # the line-by-line form and the folded form the env / GPU evaluate
            out = O.lidar(c["robot"][0], c["robot"][1], c["robot"][2], pts[:, 0], pts[:, 1], literal=literal)
            np.testing.assert_allclose(out, c["out"], rtol=1e-12, atol=1e-14)


def test_layout_sampler_matches_reference_world():
    """world.py:104-137,172-217 + utils.py:28-70 executed by the reference vs orc_env_reset (replayed draws)"""
    layouts = json.load(open(os.path.join(G, "layouts.json")))
    n = 0
    for L in layouts:
        if L["fail"]:
            continue
        e = O.OracleEnv("point", L["task"])
        # World.sample_layout stops before task.reset (world.py:104-106): pad the stream for the oracle's goal draw
        e.set_replay(L["replay"] + PAD)
        assert e.reset(0) == 0
        objs = e.objects()
        names = [k for k in L["order"] if k != "robot"]
        rs = e.robot_state
        # the reference's goal is re-sampled by task.reset (not run by World.sample_layout): compare pre-reset items
        np.testing.assert_allclose(rs[:2], L["layout"]["robot"], rtol=0, atol=1e-15)
        assert abs(rs[2] - L["robot_rot"]) < 1e-15
        for s, name in enumerate(names):
            if name == "goal":
                continue
            if name == "box" and L["task"] == "haul_box":
                continue
            np.testing.assert_allclose(objs[s, 2:4], L["layout"][name], rtol=0, atol=1e-15, err_msg=f"{L['task']} {name}")
            if name in L["yaws"]:
                d = (objs[s, 4] - L["yaws"][name] + np.pi) % (2 * np.pi) - np.pi
                assert abs(d) < 1e-12, (L["task"], name)
            assert abs(e.get_obj(s).keepout - L["keepouts"][name]) < 1e-15
        n += 1
    assert n >= 60


@pytest.mark.parametrize("ep", EPISODES, ids=[f"{e.get('robot', 'point')}-{e['task']}-{e['seed']}" + ("-gremlins" if e["config"].get("num_gremlins") else "") for e in EPISODES])
def test_episode_matches_reference_step_loop(ep):
    """safe_adaptation_gym.py:56-107 + world.py + tasks/*.py executed by the reference over oracle physics,
    vs the oracle's own restatement of that logic, on the same recorded random stream."""
    cfg = dict(ep["config"])
    e = O.OracleEnv(ep.get("robot", "point"), ep["task"], config=cfg)
    e.set_replay(ep["replay"])
    for k, seg in enumerate(ep["segments"]):
        assert e.reset(k) == 0
        np.testing.assert_allclose(e.robot_state, seg["layout"]["robot"], rtol=0, atol=1e-14)
        ref_objs = np.asarray(seg["layout"]["objects"])
        np.testing.assert_array_equal(e.objects()[:, 2:], ref_objs[:, 2:])
        np.testing.assert_array_equal(e.objects()[:, 1], ref_objs[:, 1])
        np.testing.assert_allclose(e.observation(), seg["obs0"], rtol=1e-12, atol=1e-13)
        for t, a in enumerate(seg["actions"]):
            obs, rew, cost, done, rc = e.step(a)
            assert rc == 0 and not done
            msg = f"{ep['task']} seg {k} step {t}"
            np.testing.assert_allclose(e.robot_state, seg["robot"][t], rtol=0, atol=1e-12, err_msg=msg)
            ref_r = seg["reward"][t]
            if len(ref_r) == 2:
                np.testing.assert_allclose(rew, ref_r, rtol=1e-10, atol=1e-13, err_msg=msg)
            else:
                np.testing.assert_allclose(rew[0], ref_r[0], rtol=1e-10, atol=1e-13, err_msg=msg)
            assert cost == seg["cost"][t], msg
            np.testing.assert_allclose(obs, seg["obs"][t], rtol=1e-9, atol=1e-11, err_msg=msg)  # alias = frac(angle/bin) is ill-conditioned near 0
        np.testing.assert_allclose(e.objects()[:, 2:], np.asarray(seg["final_objects"])[:, 2:], rtol=0, atol=1e-11)
    assert e.replay_pos == len(ep["replay"])


def test_geometry_constants_match_reference_xml():
    """sizes / heights the reference's primitive_objects.py emitted into the XML vs the oracle's constants"""
    want = {"hazards": ("cylinder", [0.2, 0.01], 0.02), "vases": ("box", [0.1, 0.1, 0.1], 0.1 - 4e-5),
            "pillars": ("cylinder", [0.2, 0.5], 0.5), "goal": ("cylinder", [0.3, 0.15], 0.16),
            "buttons": ("sphere", [0.1, 0.1, 0.1], 0.1), "box": ("box", [0.2, 0.2, 0.2], 0.2),
            "gremlins": ("box", [0.1, 0.1, 0.1], 0.1)}  # primitive_objects.py:63-72
    # the 'box' body of roll_rod.py:27-34 (cylinder radius 0.08, half length 0.3) and dribble_ball.py:24-30 (sphere 0.14)
    box_by_task = {"roll_rod": ("cylinder", [0.08, 0.3], 0.08), "dribble_ball": ("sphere", [0.14], 0.14)}
    seen = set()
    for ep in EPISODES:
        for name, typ, size, z in ep["sizes"]:
            key = next(k for k in want if name.startswith(k))
            if key == "box" and ep["task"] in box_by_task:
                assert (typ, z) == (box_by_task[ep["task"]][0], box_by_task[ep["task"]][2])
                np.testing.assert_allclose(size, box_by_task[ep["task"]][1], atol=1e-12)
                seen.add(ep["task"])
                continue
            assert typ == want[key][0]
            np.testing.assert_allclose(size, want[key][1], atol=1e-12)
            assert abs(z - want[key][2]) < 1e-12
            seen.add(key)
    assert seen == set(want) | set(box_by_task)
