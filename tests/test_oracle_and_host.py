"""CPU suite: the oracle against the reference-derived golden vectors, oracle self-consistency (finite differences),
host-side logic, and the C-ABI library's symbol table.  No GPU compute."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import vgpmp_oracle as O
from tests import helpers as H


# ---------------------------------------------------------------- oracle vs reference-derived goldens
@pytest.mark.parametrize("name", ["franka", "kuka", "wam", "ur10"])
def test_oracle_fk_matches_reference_numpy_fk(name, golden_dir):
    g = np.load(golden_dir / f"fk_{name}.npz")
    rob = H.oracle_robot(name)
    assert np.array_equal(rob.base_pose, g["base_pose"])
    for th, frames in zip(g["thetas"], g["frames"]):
        assert np.allclose(O.forward_kinematics_np(rob, th), frames, rtol=0, atol=1e-15)
    ft = O.fk_torch(rob, torch.as_tensor(g["thetas"])).numpy()
    assert np.allclose(ft, g["frames"], rtol=0, atol=1e-14)


def test_oracle_dh_matches_ur10_test_matrices(golden_dir):
    g = np.load(golden_dir / "ur10_dh_theta0.npz")   # tests/test_robot.py:14-42 of the reference (8 printed digits)
    rob = H.oracle_robot("ur10")
    for i, key in enumerate(["h01", "h12", "h23", "h34", "h45", "h56"]):
        T = O.dh_transform(0.0, *rob.dh[i], craig=False)
        assert np.allclose(T, g[key], rtol=0, atol=5e-9)


def test_ur10_base_pose_pin():
    # tests/test_robot.py:69-73: quaternion [0,0,-1,0] -> diag(-1,-1,1)
    rob = H.oracle_robot("ur10", "bookshelves")
    assert np.array_equal(rob.base_pose, np.diag([-1.0, -1.0, 1.0, 1.0]))


def test_oracle_sdf_matches_reference_numpy_sdf(golden_dir):
    g = np.load(golden_dir / "sdf_small.npz")
    s = O.OracleSDF.from_sdf(golden_dir / "sdf_small.sdf")
    assert np.array_equal(s.data, g["data"])
    assert np.array_equal(s.idx(g["points"]), g["idx"])
    assert np.array_equal(s.distance(g["points"]), g["dist"])
    assert np.array_equal(s.distance_grad(g["points"], zero_rule=False), g["grad"])
    zr = s.distance_grad(g["points"])
    assert np.all(zr[g["grad"] == 0] == 0.1) and np.array_equal(zr[g["grad"] != 0], g["grad"][g["grad"] != 0])
    assert (g["grad"] == 0).sum() > 0


def test_franka_dh_frames_vs_urdf_joint_origins():
    """DH chain vs the URDF joint origins of franka_spheres.urdf (known-answer: d1=.333, d3=.316, d5=.384, a4=.0825, a7=.088)."""
    rob = H.oracle_robot("franka")
    fr = O.forward_kinematics_np(rob, np.zeros(7))
    assert np.allclose(fr[1][:3, 3], [0, 0, 0.333], atol=1e-12)
    assert np.allclose(fr[3][:3, 3], [0, 0, 0.333 + 0.316], atol=1e-5)
    assert np.allclose(fr[7][:3, 3], [0.088, 0, 0.333 + 0.316 + 0.384], atol=1e-4)


# ---------------------------------------------------------------- oracle self-consistency
def test_oracle_gradients_match_finite_differences():
    case = H.make_case(num_problems=1, S=3, N=9, M=5, B=16, seed=3)
    p, d = case["oracle"][0], case["draws"][0]
    args = [case["q_mu"][0], case["q_sqrt"][0], case["ls"][0], case["var"][0]]
    out = O.elbo_and_grads(p, *args, d)

    def val(a):
        return float(p.elbo(*[O._t(x) for x in a], d))
    # The SDF term is piecewise constant in value (nearest voxel) while its gradient is *defined* by the stencil,
    # so finite differences can only validate the smooth part: check the KL-only gradient (alpha = 0).
    p0 = O.OracleProblem(**{**p.__dict__, "alpha": 0.0})
    out0 = O.elbo_and_grads(p0, *args, d)
    for k, name in enumerate(["q_mu", "q_sqrt", "lengthscales", "variances"]):
        g = out0["d_" + name]
        idx = tuple(np.unravel_index(np.argmax(np.abs(g)), g.shape))
        eps = 1e-6
        hi, lo = [x.copy() for x in args], [x.copy() for x in args]
        hi[k][idx] += eps
        lo[k][idx] -= eps
        fd = (float(p0.elbo(*[O._t(x) for x in hi], d)) - float(p0.elbo(*[O._t(x) for x in lo], d))) / (2 * eps)
        assert abs(fd - g[idx]) <= 5e-4 * max(1.0, abs(g[idx])), (name, fd, g[idx])   # Khat is ill-conditioned (cond ~1e7)
    assert np.isfinite(out["elbo"]) and val(args) == pytest.approx(out["elbo"])


def test_oracle_kl_zero_at_prior():
    """KL(q||p) = 0 when q equals the conditioned prior: whitened mean 0, q_sqrt = I."""
    case = H.make_case(num_problems=1, S=2, N=5, M=6, B=8, perturb=False)
    p = case["oracle"][0]
    ls, var = O._t(case["ls"][0]), O._t(case["var"][0])
    K = O.kuu(O._t(p.Zy), ls, var, O.JITTER)
    qs = O._t(p.joint_sigmoid_inv(p.query_states))
    pm = (K[:, :, :2] @ torch.linalg.solve(K[:, :2, :2], qs.T[..., None]))[:, 2:, 0].T    # prior mean at Z
    kl = p.prior_kl(pm, O._t(np.broadcast_to(np.eye(6), (p.robot.dof, 6, 6)).copy()), ls, var)
    assert abs(float(kl)) < 1e-8


def test_oracle_pathwise_sample_moments():
    """Sanity of the decoupled sampler at the inducing inputs with a smooth mean: samples stay near the variational mean
    and near the conditioned endpoints.  (Khat (Khat + jitter I)^-1 is not the identity on the tiny-eigenvalue directions
    of Khat, and the Fourier prior sees all D input columns -- Appendix B quirk -- so neither is exact.)"""
    case = H.make_case(num_problems=1, S=400, N=7, M=5, B=256, seed=5, perturb=False)
    p = case["oracle"][0]
    f = p.sample_paths(p.Zy, O._t(case["q_mu"][0]), O._t(case["q_sqrt"][0]), O._t(case["ls"][0]), O._t(case["var"][0]),
                       case["draws"][0]).numpy()                     # [S,Mp,D]
    mu = p.q_mu_full(O._t(case["q_mu"][0])).numpy()
    assert np.abs(f[:, :2] - mu[None, :2]).max() < 0.1
    assert np.abs(f.mean(0)[2:] - mu[2:]).max() < 0.05


def test_adam_matches_closed_form_first_step():
    st = O.AdamState()
    params = {"a": np.array([1.0, -2.0])}
    O.adam_step(params, {"a": np.array([0.5, -0.25])}, st, lr=0.09)
    # first Keras-Adam step moves every coordinate by ~lr * sign(g)
    assert np.allclose(params["a"], [1.0 - 0.09, -2.0 + 0.09], atol=1e-6)


# ---------------------------------------------------------------- host logic
@pytest.mark.parametrize("name", ["franka", "kuka", "wam", "ur10"])
def test_host_constants_match_oracle(name):
    from vgpmp_b200.utils.robot import Robot
    from vgpmp_b200.utils.sampler import Sampler
    c = Sampler(None, Robot.from_tables(name, "bookshelves")).constants()
    o = H.oracle_robot(name)
    assert np.array_equal(c.sphere_offsets, o.sphere_offsets)
    assert np.array_equal(c.sphere_frame, o.sphere_frame)
    assert np.array_equal(c.base_pose, o.base_pose)
    assert np.array_equal(c.limits_lo, o.limits_lo) and np.array_equal(c.limits_hi, o.limits_hi)
    assert c.sphere_radii.shape[0] == sum(Robot.from_tables(name).num_spheres_per_link)


def test_sdf_text_loader_roundtrip(tmp_path, golden_dir):
    from vgpmp_b200.utils.sdf_utils import SignedDistanceField
    g = np.load(golden_dir / "sdf_small.npz")
    s = SignedDistanceField.from_sdf(golden_dir / "sdf_small.sdf")
    assert np.array_equal(s.data, g["data"]) and np.array_equal(s.origin, g["origin"]) and s.delta == float(g["delta"])
    s.to_sdf(tmp_path / "rt.sdf")
    s2 = SignedDistanceField.from_sdf(tmp_path / "rt.sdf")
    assert np.array_equal(s2.data, s.data)


def test_init_trainset_and_problemset():
    from vgpmp_b200.utils.miscellaneous import init_trainset, load_problemset
    ps = load_problemset("franka", "bookshelves")
    assert len(ps["states"]) == 11 and len(ps["queries"]) == 55
    assert ps["planner_params"]["num_inducing"] == 24 and ps["planner_params"]["num_samples"] == 7
    X, y, Xnew = init_trainset(70, 150, 7, 7, ps["queries"][0][0], ps["queries"][0][1], scale=1)
    assert X.shape == (70, 7) and Xnew.shape == (150, 7) and y.shape == (2, 7)
    assert np.all(X[:, 0] == X[:, 6]) and X[0, 0] == 0.0 and X[-1, 0] == 1.0


REFERENCE_PARAMETERS_YAML = """
- robot:
    robot_name: "wam"
- scene:
      position: [ 0.85, -0.15, 0.834]
      orientation: [ 0.0, 0.0, 0.0, 1.0 ]
      environment_name: "bookshelves"
      environment_file_name: "bookshelves_mesh"
      sdf_file_name: "bookshelves_center_vgpmp"
      objects: []
      objects_position: []
      objects_orientation: []
      benchmark: True
      non_benchmark_attributes:
        states: [[0.0, 0.0, 0.0, 0.0, 0.0, 0.0], [0.0, 0.0, 0.0, 0.0, 0.0, 0.0]]
        robot_pos_and_orn: [[ 0.0, 0.0, 0.0 ], [ 0.0, 0.0, 0.0, 1.0 ]]
        planner_params: {sigma_obs: 0.005, epsilon: 0.05, lengthscales: [5.0, 5.0, 5.0, 2.0, 5.0, 5.0], variance: 0.25,
                         alpha: 100, num_samples: 7, num_inducing: 24, learning_rate: 0.09, num_steps: 130,
                         time_spacing_X: 70, time_spacing_Xnew: 150}
      benchmark_attributes:
        problemset_name: "bookshelves"
- trainable_params: {q_mu: True, q_sqrt: True, lengthscales: True, kernel_variance: True, sigma_obs: False,
                     inducing_variable: False, alpha: False}
- graphics: {visuals: True, quality: high, GUI: True}
"""


def test_parameter_loader_reads_the_reference_yaml_layout(tmp_path):
    """Same list-of-dicts layout as the reference's parameters.yaml (shipped configured for WAM + bookshelves)."""
    from vgpmp_b200.utils.parameter_loader import ParameterLoader, build_environment
    f = tmp_path / "parameters.yaml"
    f.write_text(REFERENCE_PARAMETERS_YAML)
    loader = ParameterLoader().initialize(file_path=f)
    p = loader.params
    assert p["robot_params"]["robot_name"] == "wam" and p["robot_params"]["dof"] == 7
    assert len(p["scene_params"]["queries"]) == 55                      # combinations of 11 states
    assert p["scene_params"]["robot_pos_and_orn"][0] == [0.0, 0.0, 0.346]
    assert p["planner_params"]["num_inducing"] == 15 and p["planner_params"]["num_samples"] == 20
    assert p["trainable_params"]["q_mu"] is True and p["trainable_params"]["alpha"] is False
    assert "benchmark_attributes" not in p["scene_params"]
    from vgpmp_b200.utils.sdf_utils import SignedDistanceField
    robot, sampler, sdf = build_environment(loader, sdf=SignedDistanceField(np.zeros((2, 2, 2)), np.zeros(3), 0.1))
    o = H.oracle_robot("wam", "bookshelves")
    assert np.array_equal(sampler.constants().sphere_offsets, o.sphere_offsets)
    assert np.array_equal(robot.base_pose, o.base_pose)
    with pytest.raises(AssertionError, match="SDF file"):
        build_environment(loader)                                         # the reference asserts the .sdf exists


def test_rejects_unsupported_configurations():
    from vgpmp_b200.kernels import Matern52
    with pytest.raises(ValueError):
        Matern52(variance=-1.0)
    from vgpmp_b200.utils.robot import Robot
    with pytest.raises(AssertionError):
        Robot("x", 2, [0.1], [1, -1], [1, -1, 1, -1], [0] * 6, [0, 0], [0], False, [1], [[0, 0, 0]])


# ---------------------------------------------------------------- C-ABI library
def test_library_exports_every_declared_symbol():
    from vgpmp_b200 import _cabi
    from vgpmp_b200.build import build
    build()
    header = (H.ROOT / "include" / "vgpmp_b200.h").read_text()
    declared = set(re.findall(r"\b(vgpmp_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_cabi.SYMBOLS), declared ^ set(_cabi.SYMBOLS)
    lib = ctypes.CDLL(str(_cabi.LIB_PATH))
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in _cabi.load().vgpmp_version()
    # and the ctypes table agrees with the header on every prototype's arity (catches ABI drift on the Python side)
    protos = re.findall(r"\b(vgpmp_[a-z_0-9]+)\s*\(([^;{]*?)\)\s*;", re.sub(r"/\*.*?\*/", "", header, flags=re.S))
    assert {n for n, _ in protos} == declared
    for name, args in protos:
        args = " ".join(args.split())
        n = 0 if args in ("", "void") else args.count(",") + 1
        assert n == len(_cabi.SYMBOLS[name][1]), (name, n, len(_cabi.SYMBOLS[name][1]))


def test_product_fails_loudly_without_gpu():
    from vgpmp_b200 import _cabi
    from vgpmp_b200.engine import Engine, RobotConstants
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_cabi.VgpmpError):
        Engine(RobotConstants.dummy(3), np.zeros((1, 1, 1)), (0, 0, 0), 1.0)


def test_product_never_imports_oracle():
    for path in (H.ROOT / "vgpmp_b200").rglob("*.py"):
        src = path.read_text()
        assert "import oracle" not in src and "from oracle" not in src, path


def test_oracle_matern52_matches_scikit_learn():
    """GPflow's Matern52 is not installable here; scikit-learn's Matern(nu=2.5) is an independent implementation of the
    same kernel.  Pins `matern52`, `k_conditioned` and `kuu` (lengthscale scaling, variance, jitter)."""
    from sklearn.gaussian_process.kernels import Matern
    rng = np.random.default_rng(0)
    Zy = np.sort(rng.uniform(0, 1, size=(9, 1)), axis=0) * np.ones((1, 3))      # rank-1 inputs, 3 latents
    X = rng.uniform(-0.2, 1.2, size=(11, 1)) * np.ones((1, 3))
    ls, var = np.array([0.3, 1.7, 4.0]), np.array([0.5, 0.25, 2.0])
    K = O.k_conditioned(O._t(Zy), O._t(X), O._t(ls), O._t(var)).numpy()          # [D, 9, 11]
    Kuu = O.kuu(O._t(Zy), O._t(ls), O._t(var), 1e-6).numpy()
    for l in range(3):
        sk = Matern(length_scale=ls[l], nu=2.5)
        assert np.allclose(K[l], var[l] * sk(Zy[:, l:l + 1], X[:, l:l + 1]), rtol=1e-12, atol=1e-14)
        assert np.allclose(Kuu[l], var[l] * sk(Zy[:, l:l + 1]) + 1e-6 * np.eye(9), rtol=1e-12, atol=1e-14)


def test_oracle_whitened_kl_matches_torch_distributions():
    """gauss_kl(white=True) = KL(N(m, S S^T) || N(0, I)) per latent; torch.distributions is an independent implementation.
    The conditioned-prior mean shift of prior_kl.py is removed by placing q_mu on the prior mean plus a known whitened
    offset, so the closed form applies exactly."""
    import torch.distributions as td
    case = H.make_case(num_problems=1, S=2, N=5, M=6, B=8, perturb=True, seed=3)
    p = case["oracle"][0]
    D, M = p.robot.dof, 6
    ls, var = O._t(case["ls"][0]), O._t(case["var"][0])
    K = O.kuu(O._t(p.Zy), ls, var, O.JITTER)
    L = torch.linalg.cholesky(K)
    qs = O._t(p.joint_sigmoid_inv(p.query_states))
    prior_mean = (K[:, :, :2] @ torch.cholesky_solve(qs.T[..., None], L[:, :2, :2]))[:, :, 0]          # [D, Mp]
    rng = np.random.default_rng(4)
    white = O._t(rng.standard_normal((D, M)))                                                          # whitened offset a
    full = prior_mean + (L @ torch.cat([torch.zeros(D, 2, dtype=torch.float64), white], 1)[..., None])[:, :, 0]
    q_mu = full[:, 2:].T                                                                                # [M, D]
    q_sqrt = O._t(case["q_sqrt"][0])
    got = float(p.prior_kl(q_mu, q_sqrt, ls, var))
    want = sum(float(td.kl_divergence(td.MultivariateNormal(white[l], scale_tril=torch.tril(q_sqrt[l])),
                                      td.MultivariateNormal(torch.zeros(M, dtype=torch.float64),
                                                            torch.eye(M, dtype=torch.float64)))) for l in range(D))
    assert abs(got - want) <= 1e-9 * abs(want)


def test_oracle_fourier_features_reproduce_the_matern52_kernel():
    """The random-Fourier prior of GPflowSampling must have the Matern-5/2 covariance: with omega = N(0,I)/sqrt(Gamma(5/2,
    rate 5/2)), E[phi(x)^T phi(x')] = k(x, x').  Monte-Carlo over 200 000 bases, 1-D inputs (tolerance 4 sigma)."""
    rng = np.random.default_rng(7)
    d = O.make_draws(rng, D=1, S=1, B=200_000, Mp=3, Din=1)
    omega, tau = d["omega"][0, :, 0], d["tau"][0]
    x = np.array([0.0, 0.13, 0.4, 1.0, 2.5])
    ell, var = 0.8, 1.3
    phi = np.sqrt(2.0 * var / omega.size) * np.cos(np.outer(x, omega) / ell + tau[None, :])            # [5, B]
    est = phi @ phi.T
    r = np.abs(x[:, None] - x[None, :]) / ell
    want = var * O.matern52(O._t(r)).numpy()
    assert np.abs(est - want).max() < 4 * var * np.sqrt(0.5 / omega.size) * 2


# ---------------------------------------------------------------- URDF extraction vs the lifted tables
REFERENCE_ROBOTS = Path("/root/reference/data/robots")


@pytest.mark.skipif(not REFERENCE_ROBOTS.exists(), reason="needs the reference checkout (its URDFs are not copied into this repo)")
@pytest.mark.parametrize("name,urdf", [("franka", "franka_spheres.urdf"), ("kuka", "kuka.urdf"), ("wam", "wam.urdf"),
                                       ("ur10", "ur10.urdf")])
def test_robot_from_urdf_reproduces_the_lifted_tables(name, urdf):
    """`Robot.from_urdf` (replaces the pybullet visual-shape walk of utils/robot.py:482-550) run on the reference's own URDF +
    config.yaml must give exactly the constants lifted into vgpmp_b200/data/robots.json (what every test and bench uses)."""
    import yaml
    from vgpmp_b200.utils.robot import Robot
    cfg = yaml.safe_load((REFERENCE_ROBOTS / name / "config.yaml").read_text())
    cfg = cfg["robot"] if "robot" in cfg else cfg
    a = Robot.from_urdf(name, cfg, REFERENCE_ROBOTS / name / urdf)
    b = Robot.from_tables(name)
    assert a.dof == b.dof and a.num_spheres == b.num_spheres and a.craig_notation == b.craig_notation
    assert list(a.num_spheres_per_link) == list(b.num_spheres_per_link) and list(a.fk_slice) == list(b.fk_slice)
    assert np.array_equal(np.asarray(a.sphere_offsets), np.asarray(b.sphere_offsets))
    assert np.array_equal(np.asarray(a.sphere_radii), np.asarray(b.sphere_radii))
    assert np.array_equal(a.DH, b.DH) and np.array_equal(a.twist, b.twist)
    assert np.array_equal(np.asarray(a.joint_limits), np.asarray(b.joint_limits))
    # and, through the Sampler's get_mat remaps, the constants the kernels receive
    from vgpmp_b200.utils.sampler import Sampler
    ca, cb = Sampler(None, a).constants(), Sampler(None, b).constants()
    for f in ("dh", "twist", "base_pose", "sphere_frame", "sphere_offsets", "sphere_radii", "limits_lo", "limits_hi"):
        assert np.array_equal(getattr(ca, f), getattr(cb, f)), f


def test_mesh_sdf_vectorised_matches_the_scalar_checker():
    """bench.py's CPU arm builds its bookshelves grid with `mesh_sdf_vectorised`; it must equal the scalar brute force that
    checks the GPU producer."""
    from vgpmp_b200.utils.gen_sdf import grid_geometry, load_obj_convex_pieces, scene_mesh_path
    tri, plane, piece_end = load_obj_convex_pieces(scene_mesh_path("bookshelves"))
    origin, shape = grid_geometry(tri, 0.15, 1)
    a = O.mesh_sdf_np(tri, plane, piece_end, origin, 0.15, shape)
    b = O.mesh_sdf_vectorised(tri, plane, piece_end, origin, 0.15, shape, chunk=97)
    assert (a < 0).any() and np.abs(a - b).max() < 1e-12
