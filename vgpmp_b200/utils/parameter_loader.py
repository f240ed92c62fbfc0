"""`ParameterLoader` (reference: gpflow_vgpmp/utils/parameter_loader.py:47-159): the reference's own
`parameters.yaml` drives this engine unchanged.

Same outputs: `robot_params`, `scene_params` (with `queries = combinations(states, 2)`, `robot_pos_and_orn`,
`sdf_path`, `environment_path`), `planner_params`, `trainable_params`, `graphics_params`.  Robot config / URDF / problem
tables are read from a checkout of the reference's `data/` directory when `data_dir` is given, otherwise from the tables
lifted into vgpmp_b200/data/*.json (tools/lift_reference_data.py).  `build_environment` replaces the pybullet wiring of
SimulationManager (utils/simulation_manager.py:45-58) for the hot path: Robot constants, Sampler, SignedDistanceField.
"""
from __future__ import annotations

import copy
import itertools
import json
from pathlib import Path
from typing import Optional

import yaml

from .robot import DATA_DIR, Robot
from .sampler import Sampler
from .sdf_utils import SignedDistanceField

__all__ = ["ParameterLoader", "load_yaml_config", "build_environment"]


def load_yaml_config(path):
    with open(path, "r") as stream:
        return yaml.safe_load(stream)


class ParameterLoader:
    def __init__(self, data_dir: Optional[Path] = None):
        self.is_initialized = False
        self._params = None
        self.trainable_params = self.planner_params = self.graphics_params = None
        self.robot_params = self.scene_params = None
        self.data_dir_path = Path(data_dir) if data_dir is not None else None

    @property
    def params(self) -> dict:
        assert self._params is not None, "Parameter Loader must be initialized before it can be accessed"
        return self._params

    def initialize(self, file_path: Path = None, params: Optional[dict] = None):
        if file_path is not None:
            self._params = self.set_params(load_yaml_config(file_path))
        else:
            assert params is not None, "Either parameter_file_path or params must be specified"
            self._params = self.set_params(params)
        return self

    def set_params(self, params):
        robot_params, scene_params, trainable_params, graphic_params = params
        self.scene_params = scene_params["scene"]
        self.robot_params = robot_params["robot"]
        self.trainable_params = trainable_params["trainable_params"]
        self.graphics_params = graphic_params["graphics"]
        self.get_robot_config(self.robot_params)
        self.get_scene_config(self.scene_params)
        self.is_initialized = True
        return {"robot_params": self.robot_params, "scene_params": self.scene_params,
                "planner_params": self.planner_params, "trainable_params": self.trainable_params,
                "graphics_params": self.graphics_params}

    # ---- robot ---------------------------------------------------------------------------------------------------
    def get_robot_config(self, robot_params: dict):
        name = robot_params["robot_name"]
        if self.data_dir_path is not None:
            robot_path = self.data_dir_path / "robots" / name
            config = load_yaml_config(robot_path / "config.yaml")
            config["urdf_path"] = robot_path / config["path"]
        else:
            tables = json.loads((DATA_DIR / "robots.json").read_text())
            if name not in tables:
                raise SystemExit("Robot not available. Check params file and try again...")
            config = dict(tables[name])
            config["urdf_path"] = None
        self.robot_params = {**config, **robot_params}   # main-file entries take precedence (parameter_loader.py:97-99)

    # ---- scene ---------------------------------------------------------------------------------------------------
    def get_scene_config(self, scene_params: dict):
        assert scene_params["benchmark"] is not None, "Benchmark attribute is not specified"
        assert type(scene_params["benchmark"]) is bool, "Benchmark attribute must be a boolean"
        if scene_params["benchmark"] is False:
            attrs = scene_params["non_benchmark_attributes"]
            states, planner_params = attrs["states"], attrs["planner_params"]
            robot_pos_and_orn = tuple(attrs["robot_pos_and_orn"])
        else:
            name = self.robot_params["robot_name"]
            problemset = scene_params["benchmark_attributes"]["problemset_name"]
            table = json.loads((DATA_DIR / "problemsets.json").read_text())[name]
            if problemset not in table:
                raise ValueError("Unknown problem set: {}".format(problemset))
            states = table[problemset]["states"]
            planner_params = table[problemset]["planner_params"]
            robot_pos_and_orn = tuple(table[problemset]["pos_and_orn"])
        scene_params = copy.deepcopy(scene_params)
        scene_params["queries"] = list(itertools.combinations(states, 2))          # parameter_loader.py:138
        scene_params["robot_pos_and_orn"] = robot_pos_and_orn
        env, env_file, sdf_file = (scene_params["environment_name"], scene_params["environment_file_name"],
                                   scene_params["sdf_file_name"])
        if self.data_dir_path is not None:
            scenes = self.data_dir_path / "scenes" / env
            scene_params["environment_path"] = scenes / (env_file + ".urdf")
            scene_params["sdf_path"] = scenes / (sdf_file + ".sdf")
        else:
            scene_params["environment_path"] = scene_params["sdf_path"] = None
        scene_params.pop("benchmark_attributes", None)
        scene_params.pop("non_benchmark_attributes", None)
        self.scene_params = scene_params
        self.planner_params = planner_params


def build_environment(loader: ParameterLoader, sdf: Optional[SignedDistanceField] = None):
    """(robot, sampler, sdf) for the loader's configuration.  The reference asserts that the .sdf file exists
    (parameter_loader.py:157); its grids are missing blobs in the snapshot, so a caller may pass `sdf` instead."""
    rp, sp = loader.robot_params, loader.scene_params
    if rp.get("urdf_path") is not None:
        robot = Robot.from_urdf(rp["robot_name"], rp, rp["urdf_path"], sp["robot_pos_and_orn"])
    else:
        robot = Robot.from_tables(rp["robot_name"], pos_and_orn=sp["robot_pos_and_orn"])
    sampler = Sampler(loader, robot)
    if sdf is None:
        path = sp.get("sdf_path")
        assert path is not None and Path(path).exists(), f"SDF file {path} does not exist"
        sdf = SignedDistanceField.from_sdf(path)
    return robot, sampler, sdf
