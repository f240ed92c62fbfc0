"""Robot constants of the hot path.

The reference `Robot` (gpflow_vgpmp/utils/robot.py:64-162) is a pybullet wrapper; pybullet only supplies
constants to the ELBO path (base pose :196-203, sphere offsets :482-499, spheres per link :534-550).  This class
holds exactly those constants, read either from the tables lifted into vgpmp_b200/data/*.json
(tools/lift_reference_data.py) or straight from a robot config.yaml + URDF.  Motor control, collision
filtering and trajectory execution are simulator features and are out of scope.
"""
from __future__ import annotations

import json
import xml.etree.ElementTree as ET
from pathlib import Path
from typing import List, Optional, Sequence

import numpy as np

DATA_DIR = Path(__file__).resolve().parents[1] / "data"


def quat_to_rotmat(q: Sequence[float]) -> np.ndarray:
    x, y, z, w = (float(v) for v in q)
    n = x * x + y * y + z * z + w * w
    s = 2.0 / n
    return np.array([[1 - s * (y * y + z * z), s * (x * y - z * w), s * (x * z + y * w)],
                     [s * (x * y + z * w), 1 - s * (x * x + z * z), s * (y * z - x * w)],
                     [s * (x * z - y * w), s * (y * z + x * w), 1 - s * (x * x + y * y)]])


def get_base(rotation, translation) -> np.ndarray:
    T = np.eye(4)
    T[:3, :3] = np.asarray(rotation, dtype=np.float64).reshape(3, 3)
    T[:3, 3] = np.asarray(translation, dtype=np.float64).reshape(3)
    return T


def _load_json(name):
    return json.loads((DATA_DIR / name).read_text())


def urdf_sphere_visuals(urdf_path):
    """Sphere <visual>s in pybullet link order; offsets are visual origin minus inertial origin (what
    p.getVisualShapeData()[5] reports).  Returns (spheres_per_link, offsets[P,3])."""
    root = ET.parse(urdf_path).getroot()
    links = {l.get("name"): l for l in root.findall("link")}
    joints = root.findall("joint")
    child_names = {j.find("child").get("link") for j in joints}
    roots = [n for n in links if n not in child_names]
    if len(roots) != 1:
        raise ValueError(f"{urdf_path}: expected one root link, found {roots}")
    order, stack = [], [roots[0]]
    while stack:
        name = stack.pop()
        order.append(name)
        kids = [j.find("child").get("link") for j in joints if j.find("parent").get("link") == name]
        stack.extend(reversed(kids))
    per_link, offsets = [], []
    for name in order:
        link = links[name]
        inertial = link.find("inertial")
        io = np.zeros(3)
        if inertial is not None and inertial.find("origin") is not None and inertial.find("origin").get("xyz"):
            io = np.array([float(v) for v in inertial.find("origin").get("xyz").split()])
        count = 0
        for vis in link.findall("visual"):
            geom = vis.find("geometry")
            if geom is None or geom.find("sphere") is None:
                continue
            org = vis.find("origin")
            vo = np.array([float(v) for v in org.get("xyz").split()]) if org is not None and org.get("xyz") else np.zeros(3)
            offsets.append(vo - io)
            count += 1
        if count:
            per_link.append(count)
    return per_link, np.array(offsets, dtype=np.float64).reshape(-1, 3)


class Robot:
    """Constants-only stand-in for the reference Robot.  Attribute names follow utils/robot.py / robot_mixin.py."""

    def __init__(self, robot_name: str, dof: int, radius, joint_limits, velocity_limits, dh_parameters, twist, fk_slice,
                 craig_dh_convention: bool, num_spheres_per_link, sphere_offsets, base_pose=None, **_ignored):
        self.name = robot_name
        self.dof = int(dof)
        self.sphere_radii = [float(r) for r in radius]
        self.num_spheres = len(self.sphere_radii)
        if len(joint_limits) != 2 * self.dof:
            raise AssertionError("Cannot set joint limits for a different number than the total active joints")
        if len(velocity_limits) != 2 * self.dof:
            raise AssertionError(f"Velocity limits must be of length {self.dof}")
        self.joint_limits = [float(v) for v in joint_limits]          # [hi, lo] per joint
        self.velocity_limits = [float(v) for v in velocity_limits]
        self.DH = np.asarray(dh_parameters, dtype=np.float64).reshape(-1, 3)
        self.twist = np.asarray(twist, dtype=np.float64).reshape(-1, 1)
        self.fk_slice = [int(i) for i in fk_slice]
        self.craig_notation = bool(craig_dh_convention)
        self.num_spheres_per_link = [int(n) for n in num_spheres_per_link]
        self.num_frames_for_spheres = len(self.num_spheres_per_link)
        self.sphere_offsets = np.asarray(sphere_offsets, dtype=np.float64).reshape(-1, 3)
        if self.sphere_offsets.shape[0] != self.num_spheres or sum(self.num_spheres_per_link) != self.num_spheres:
            raise AssertionError("sphere offsets / radii / spheres-per-link disagree")
        self.base_pose = np.eye(4) if base_pose is None else np.asarray(base_pose, dtype=np.float64).reshape(4, 4)
        self.is_initialized = True

    # ---- constructors ----------------------------------------------------------------------------
    @classmethod
    def from_tables(cls, robot_name: str, environment: Optional[str] = None, pos_and_orn=None) -> "Robot":
        entry = _load_json("robots.json")[robot_name]
        if pos_and_orn is None and environment is not None:
            pos_and_orn = _load_json("problemsets.json")[robot_name][environment]["pos_and_orn"]
        rob = cls(robot_name=robot_name, dof=entry["dof"], radius=entry["radius"], joint_limits=entry["joint_limits"],
                  velocity_limits=entry["velocity_limits"], dh_parameters=entry["dh_parameters"], twist=entry["twist"],
                  fk_slice=entry["fk_slice"], craig_dh_convention=entry["craig_dh_convention"],
                  num_spheres_per_link=entry["num_spheres_per_link"], sphere_offsets=entry["sphere_offsets_urdf"])
        if pos_and_orn is not None:
            rob.reset_pos_and_orn(*pos_and_orn)
        return rob

    @classmethod
    def from_urdf(cls, robot_name: str, config: dict, urdf_path, pos_and_orn=None) -> "Robot":
        per_link, offsets = urdf_sphere_visuals(urdf_path)
        rob = cls(robot_name=robot_name, dof=config["dof"], radius=config["radius"], joint_limits=config["joint_limits"],
                  velocity_limits=config["velocity_limits"], dh_parameters=config["dh_parameters"], twist=config["twist"],
                  fk_slice=config["fk_slice"], craig_dh_convention=config["craig_dh_convention"],
                  num_spheres_per_link=per_link, sphere_offsets=offsets)
        if pos_and_orn is not None:
            rob.reset_pos_and_orn(*pos_and_orn)
        return rob

    # ---- reference-named helpers -------------------------------------------------------------------
    def reset_pos_and_orn(self, pos, orn):
        self.position, self.orientation = list(pos), list(orn)
        self.base_pose = get_base(quat_to_rotmat(orn), pos)

    def get_base_pose(self) -> np.ndarray:
        return self.base_pose

    @property
    def limits_hi(self) -> np.ndarray:
        return np.asarray(self.joint_limits, dtype=np.float64).reshape(self.dof, 2)[:, 0].copy()

    @property
    def limits_lo(self) -> np.ndarray:
        return np.asarray(self.joint_limits, dtype=np.float64).reshape(self.dof, 2)[:, 1].copy()
