#!/usr/bin/env python
"""Lift the reference's *data tables* (not code) into JSON the B200 package ships.

Run in the build container only (needs /root/reference).  Outputs:
  vgpmp_b200/data/robots.json       <- data/robots/*/config.yaml + URDF sphere visuals
  vgpmp_b200/data/problemsets.json  <- data/problemsets/{franka,kuka,wam,ur10}.py tables

Sphere offsets replace what pybullet's getVisualShapeData()[5] supplies to the reference
(gpflow_vgpmp/utils/robot.py:482-499): the visual origin expressed in the link's *inertial*
frame, i.e. visual_xyz - inertial_xyz when the inertial rpy is zero (true for all four URDFs).
Franka has all-zero inertial origins, so its constants are exact; the others are flagged
"constants re-derived" (SURVEY.md section 8c).
"""
import json
import sys
import types
import xml.etree.ElementTree as ET
from pathlib import Path

import yaml

REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
OUT = Path(__file__).resolve().parents[1] / "vgpmp_b200" / "data"


def _xyz(elem, default=(0.0, 0.0, 0.0)):
    if elem is None or elem.get("xyz") is None:
        return list(default)
    return [float(v) for v in elem.get("xyz").split()]


def urdf_spheres(path):
    """Sphere visuals in pybullet link-index order: base link first, then joints' children in file order."""
    root = ET.parse(path).getroot()
    links = {l.get("name"): l for l in root.findall("link")}
    children = {j.find("child").get("link") for j in root.findall("joint")}
    parents = [n for n in links if n not in children]
    assert len(parents) == 1, parents
    # pybullet numbers links by a depth-first walk of the joint tree from the root link
    order = []

    def walk(name):
        order.append(name)
        for j in root.findall("joint"):
            if j.find("parent").get("link") == name:
                walk(j.find("child").get("link"))

    walk(parents[0])
    per_link, offsets = [], []
    for name in order:
        link = links[name]
        inertial = link.find("inertial")
        io = _xyz(inertial.find("origin")) if inertial is not None else [0.0, 0.0, 0.0]
        n = 0
        for vis in link.findall("visual"):
            geom = vis.find("geometry")
            if geom is not None and geom.find("sphere") is not None:
                vo = _xyz(vis.find("origin"))
                offsets.append([vo[k] - io[k] for k in range(3)])
                n += 1
        if n:
            per_link.append(n)
    return per_link, offsets


def lift_robots():
    out = {}
    for name in ("franka", "kuka", "wam", "ur10"):
        cfg = yaml.safe_load(open(REF / "data" / "robots" / name / "config.yaml"))
        per_link, offsets = urdf_spheres(REF / "data" / "robots" / name / cfg["path"])
        assert len(offsets) == cfg["num_spheres"] == len(cfg["radius"]), (name, len(offsets))
        assert len(per_link) == cfg["num_frames_for_spheres"] == len(cfg["fk_slice"]), (name, per_link)
        out[name] = {
            "robot_name": name,
            "dof": cfg["dof"],
            "radius": cfg["radius"],
            "joint_limits": cfg["joint_limits"],          # [hi, lo] per joint (likelihood.py:45-52)
            "velocity_limits": cfg["velocity_limits"],
            "dh_parameters": cfg["dh_parameters"],        # (d, a, alpha) per joint
            "twist": cfg["twist"],
            "fk_slice": cfg["fk_slice"],
            "craig_dh_convention": bool(cfg["craig_dh_convention"]),
            "num_frames_for_spheres": cfg["num_frames_for_spheres"],
            "num_spheres_per_link": per_link,
            "sphere_offsets_urdf": offsets,               # before Sampler.get_mat (sampler.py:68-101)
            "constants_exact": name == "franka",
            "q_mu": cfg.get("q_mu"),
        }
    return out


def lift_problemsets():
    fake = types.ModuleType("problemset")
    fake.AbstractProblemset = type("AbstractProblemset", (), {})
    sys.modules["problemset"] = fake
    out = {}
    for name in ("franka", "kuka", "wam", "ur10"):
        ns = {}
        exec(compile((REF / "data" / "problemsets" / f"{name}.py").read_text(), name, "exec"), ns)
        P = ns["Problemset"]
        out[name] = {}
        for env in ("bookshelves", "industrial", "boxes", "lab"):
            try:
                n_states, states = P.states(env)
            except (ValueError, SystemExit):
                continue
            entry = {"states": states}
            for key, fn in (("pos_and_orn", P.pos_and_orn), ("object_positions", P.object_positions),
                            ("planner_params", P.planner_params)):
                try:
                    entry[key] = fn(env)
                except (ValueError, SystemExit):
                    entry[key] = None
            out[name][env] = entry
    return out


if __name__ == "__main__":
    OUT.mkdir(parents=True, exist_ok=True)
    (OUT / "robots.json").write_text(json.dumps(lift_robots(), indent=1))
    (OUT / "problemsets.json").write_text(json.dumps(lift_problemsets(), indent=1))
    tp = yaml.safe_load(open(REF / "parameters.yaml"))
    trainable = [d["trainable_params"] for d in tp if "trainable_params" in d][0]
    (OUT / "trainable_params.json").write_text(json.dumps(trainable, indent=1))
    # scene meshes (data assets, convex pieces): input of the GPU SDF producer (vgpmp_b200/utils/gen_sdf.py)
    import shutil
    (OUT / "scenes").mkdir(exist_ok=True)
    shutil.copy(REF / "data" / "scenes" / "bookshelves" / "bookshelves_center.obj", OUT / "scenes")
    shutil.copy(REF / "data" / "scenes" / "industrial" / "industrial-acd.obj", OUT / "scenes")
    print("wrote", sorted(p.name for p in OUT.iterdir()))
