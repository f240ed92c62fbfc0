#!/usr/bin/env python
"""Generate tests/golden/* by RUNNING the reference's own NumPy code in this container.

Needs /root/reference (absent on the GPU box) -- the produced fixtures are committed.
  * fk_<robot>.npz     : RobotMixin.forward_kinematics (gpflow_vgpmp/utils/robot_mixin.py:32-58) on seeded joint
                         configurations, with the robot tables of data/robots/<robot>/config.yaml.
  * sdf_small.sdf/.npz : a seeded random grid written in the reference's text format, read back with
                         SignedDistanceField.from_sdf (utils/sdf_utils.py:195-210) and queried with the NumPy
                         get_distance / get_distance_grad (utils/sdf_utils.py:68-98).  TensorFlow is absent, so the
                         module is executed with a stub `tensorflow` (only tf.constant/tf.reshape are touched, at
                         construction time; the NumPy methods never use them) and `np.int = int` (removed in NumPy>=1.24).
  * ur10_dh_theta0.npz : the literal UR10 DH matrices of tests/test_robot.py:14-42 (parsed from the test source).
"""
import importlib.util
import json
import re
import sys
import types
from pathlib import Path
from unittest.mock import MagicMock

import numpy as np
import yaml

REF = Path("/root/reference")
ROOT = Path(__file__).resolve().parents[1]
OUT = ROOT / "tests" / "golden"
OUT.mkdir(parents=True, exist_ok=True)


def load_by_path(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def golden_fk():
    mixin = load_by_path("ref_robot_mixin", REF / "gpflow_vgpmp" / "utils" / "robot_mixin.py")
    problems = json.loads((ROOT / "vgpmp_b200" / "data" / "problemsets.json").read_text())
    rng = np.random.default_rng(20240917)
    for name in ("franka", "kuka", "wam", "ur10"):
        cfg = yaml.safe_load(open(REF / "data" / "robots" / name / "config.yaml"))
        pos, orn = problems[name]["bookshelves"]["pos_and_orn"]
        x, y, z, w = orn
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        base = np.eye(4)
        base[:3, :3] = R
        base[:3, 3] = pos
        rob = mixin.RobotMixin(robot_name=name, dh_parameters=cfg["dh_parameters"], dof=cfg["dof"], twist=cfg["twist"],
                               fk_slice=cfg["fk_slice"], craig_dh_convention=cfg["craig_dh_convention"],
                               joint_limits=cfg["joint_limits"], velocity_limits=cfg["velocity_limits"],
                               base_pose=base, vectorize=False)
        D = cfg["dof"]
        lim = np.array(cfg["joint_limits"]).reshape(D, 2)
        thetas = [np.zeros(D), np.full(D, 0.1)]  # 0.1 is the config of tests/test_robot.py:95-107
        thetas += [lim[:, 1] + (lim[:, 0] - lim[:, 1]) * rng.uniform(size=D) for _ in range(14)]
        thetas = np.array(thetas)
        frames = np.array([rob.forward_kinematics(t.reshape(D, 1)) for t in thetas])
        np.savez(OUT / f"fk_{name}.npz", thetas=thetas, frames=frames, base_pose=base)
        print(name, frames.shape)


def golden_sdf():
    tf_stub = types.ModuleType("tensorflow")
    for attr in ("constant", "reshape", "float64", "int64", "custom_gradient", "function"):
        setattr(tf_stub, attr, MagicMock())
    tf_stub.custom_gradient = lambda f: f
    sys.modules["tensorflow"] = tf_stub
    np.int = int  # the reference predates NumPy 1.24
    sdfmod = load_by_path("ref_sdf_utils", REF / "gpflow_vgpmp" / "utils" / "sdf_utils.py")
    rng = np.random.default_rng(7)
    nx, ny, nz = 9, 7, 5
    data = np.round(rng.normal(size=(nx, ny, nz)), 6)
    data[2:4, 1:3, 1:4] = 0.25            # a flat patch so some central differences are exactly 0
    origin, delta = np.array([-0.4, -0.3, -0.2]), 0.1
    path = OUT / "sdf_small.sdf"
    with open(path, "w") as fh:
        fh.write(f"{nx} {ny} {nz}\n{origin[0]} {origin[1]} {origin[2]}\n{delta}\n")
        for k in range(nz):
            for j in range(ny):
                for i in range(nx):
                    fh.write(f"{data[i, j, k]}\n")
    ref = sdfmod.SignedDistanceField.from_sdf(str(path))
    assert np.array_equal(ref.data, data)
    pts = rng.uniform(-0.6, 0.7, size=(400, 3))     # some outside the grid on every side (clipping)
    pts[:8] = origin + delta * np.array([[0, 0, 0], [8, 6, 4], [8.999, 6.999, 4.999], [2.5, 1.5, 2.5], [3.5, 2.5, 1.5],
                                         [0.5, 0.5, 0.5], [-3, 2, 2], [4, 20, 2]])
    np.savez(OUT / "sdf_small.npz", points=pts, dist=ref.get_distance(pts), grad=ref.get_distance_grad(pts),
             idx=ref._rel_pos_to_idxes(pts), data=data, origin=origin, delta=delta)
    print("sdf", ref.data.shape, ref.get_distance(pts)[:3])


def golden_ur10_dh():
    src = (REF / "tests" / "test_robot.py").read_text()
    mats = {}
    for m in re.finditer(r"(h\d\d) = np\.array\((\[\[.*?\]\]), dtype=np\.float64\)", src, re.S):
        mats[m.group(1)] = np.array(eval(m.group(2)))
    assert sorted(mats) == ["h01", "h12", "h23", "h34", "h45", "h56"], sorted(mats)
    np.savez(OUT / "ur10_dh_theta0.npz", **mats)
    print("ur10 dh", list(mats))


if __name__ == "__main__":
    golden_fk()
    golden_sdf()
    golden_ur10_dh()
