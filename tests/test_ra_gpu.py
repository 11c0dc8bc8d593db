"""GPU parity tests of rotation averaging against oracle/ra_oracle.py.
Thresholds are the reference's (rotation_averager_test.cc:167,210,261: all-pairs
relative rotation error < 1e-2 deg noise-free; :310: < 3 deg with noise and
outliers)."""
import numpy as np
import pytest

from glomap_b200 import estimators as E, geometry as G, synthetic as S
from oracle import ra_oracle as RA

pytestmark = pytest.mark.gpu


def _device(vg, R0=None, **kw):
    opts = E.RotationEstimatorOptions(pcg_rel_tolerance=1e-12, **kw)
    est = E.RotationEstimator(opts)
    ok, R = est.EstimateRotations(vg, R0)
    return ok, R, est.summary


def _pairwise_err(R, R_gt):
    return RA.max_pairwise_rotation_error_deg(G.so3_log(R), R_gt)


def test_ring_noise_free_with_mst_initialisation():
    vg = S.make_ring_view_graph(100, 5, seed=1)
    ok, R, st = _device(vg)
    assert ok and _pairwise_err(R, vg.R_gt) < 1e-2          # rotation_averager_test.cc:167


@pytest.mark.parametrize("weight_type", [0, 1])
def test_noisy_graph_matches_oracle_step_by_step(weight_type):
    """Same initial rotations as the oracle, tight PCG: identical L1/IRLS
    iteration counts and the same rotations."""
    vg = S.make_random_view_graph(150, 10, seed=3, noise_deg=1.0, outlier_ratio=0.05)
    theta0 = np.zeros((vg.n_images, 3))
    ok, R, st = _device(vg, skip_initialization=True, weight_type=weight_type)
    th, info = RA.estimate_rotations(vg.n_images, vg.ei, vg.ej, vg.R_rel, theta0,
                                     opts=RA.RAOptions(weight_type="HALF_NORM" if weight_type else "GEMAN_MCCLURE"))
    assert ok
    assert (st.l1_iterations, st.irls_iterations) == (info["l1_iterations"], info["irls_iterations"])
    assert st.admm_iterations == info["admm_iterations"]
    err = np.abs(R - RA.aa_to_R(th)).max()          # (arccos-based angles bottom out at ~1e-6 deg)
    assert err < 1e-8, err
    assert _pairwise_err(R, vg.R_gt) < 3.0                  # rotation_averager_test.cc:310


def test_outliers_from_mst_init_within_reference_threshold():
    vg = S.make_random_view_graph(300, 16, seed=5, noise_deg=2.0, outlier_ratio=0.10)
    ok, R, st = _device(vg)
    R0 = E.initialize_from_maximum_spanning_tree(vg)
    th, info = RA.estimate_rotations(vg.n_images, vg.ei, vg.ej, vg.R_rel, G.so3_log(R0))
    assert ok and _pairwise_err(R, vg.R_gt) < 3.0
    assert np.abs(R - RA.aa_to_R(th)).max() < 1e-7


def test_edge_weights():
    vg = S.make_random_view_graph(80, 8, seed=7, noise_deg=1.0, outlier_ratio=0.05)
    rng = np.random.default_rng(0)
    vg.weight = rng.uniform(0.2, 2.0, size=vg.E)
    ok, R, st = _device(vg, skip_initialization=True, use_weight=True)
    th, info = RA.estimate_rotations(vg.n_images, vg.ei, vg.ej, vg.R_rel, np.zeros((vg.n_images, 3)), vg.weight,
                                     RA.RAOptions(use_weight=True))
    assert ok and np.abs(R - RA.aa_to_R(th)).max() < 1e-8


def test_idempotent_on_converged_solution():
    vg = S.make_random_view_graph(100, 10, seed=9, noise_deg=0.5)
    ok, R, st = _device(vg)
    ok2, R2, st2 = _device(vg, R, skip_initialization=True)
    assert ok and ok2 and st2.irls_iterations <= 2
    assert G.rotation_angle_deg(R, R2).max() < 0.1


def test_text_formats_roundtrip(tmp_path):
    """config 1 plumbing: relpose file (docs/rotation_averager.md:43-46) -> view graph -> rotations file."""
    vg = S.make_ring_view_graph(100, 5, seed=1, noise_deg=0.5)
    p = tmp_path / "relpose.txt"
    S.write_relpose_file(str(p), vg)
    vg2, names = S.read_relpose_file(str(p))
    assert vg2.E == vg.E and len(names) == 100
    ok, R, st = _device(vg2)
    assert ok
    out = tmp_path / "rot.txt"
    S.write_global_rotation_file(str(out), names, R)
    assert len(out.read_text().strip().splitlines()) == 100
    # names are numbered in order of first appearance (pose_io.cc:46-59): map back before comparing
    idx = np.array([int(nm[3:]) for nm in names])
    assert RA.max_pairwise_rotation_error_deg(G.so3_log(R), vg.R_gt[idx]) < 3.0


def _gravity_setup(vg, frac, seed):
    rng = np.random.default_rng(seed)
    hg = rng.uniform(size=vg.n_images) < frac
    g = np.full((vg.n_images, 3), np.nan)
    g[hg] = vg.R_gt[hg][:, :, 1]          # gravity = second column of R (docs/rotation_averager.md:51-60)
    return hg, g


@pytest.mark.parametrize("frac", [1.0, 0.6])
def test_gravity_aligned_frames_match_oracle(frac):
    """use_gravity: 1-DoF frames, mixed 1-DoF/3-DoF pairs (global_rotation_averaging.cc:207-217,386-421)."""
    vg = S.make_random_view_graph(80, 8, seed=13, noise_deg=1.0, outlier_ratio=0.05)
    hg, g = _gravity_setup(vg, frac, 2)
    est = E.RotationEstimator(E.RotationEstimatorOptions(use_gravity=True, pcg_rel_tolerance=1e-12))
    ok, R = est.EstimateRotations(vg, gravity=g)
    st = est.summary
    R_align = np.tile(np.eye(3), (vg.n_images, 1, 1))
    for i in np.nonzero(hg)[0]:
        R_align[i] = E.get_align_rot(g[i])
    Ro, info = RA.estimate_rotations_gravity(vg.n_images, vg.ei, vg.ej, vg.R_rel, np.tile(np.eye(3), (vg.n_images, 1, 1)), hg, R_align)
    assert ok
    assert (st.l1_iterations, st.irls_iterations, st.admm_iterations) == (info["l1_iterations"], info["irls_iterations"], info["admm_iterations"])
    assert np.abs(R - Ro).max() < 1e-8
    assert _pairwise_err(R, vg.R_gt) < 2.0                  # rotation_averager_test.cc:359-362
    # gravity is honoured exactly: second column of every gravity frame's rotation
    assert np.abs(R[hg][:, :, 1] - g[hg]).max() < 1e-12


def test_two_level_preconditioner_and_gather_laplacian_do_not_change_the_result(monkeypatch):
    """The CSR-by-node gather Laplacian and the two-level (aggregation) preconditioner are solver internals: with a tight
    PCG tolerance the L1 / IRLS iteration counts and the rotations are those of the edge-parallel Jacobi-PCG path."""
    vg = S.make_lattice_view_graph(4096, 24, seed=3, noise_deg=1.0, outlier_ratio=0.05)
    R0 = E.initialize_from_maximum_spanning_tree(vg)
    out = {}
    for name, csr, lvl, fused in (("reference", "0", "0", "0"), ("csr", "1", "0", "0"), ("csr+2lvl", "1", "1", "0"),
                                  ("csr+2lvl fused", "1", "1", "1")):
        monkeypatch.setenv("B200SFM_RA_CSR", csr)
        monkeypatch.setenv("B200SFM_RA_2LVL", lvl)
        monkeypatch.setenv("B200SFM_RA_FUSED", fused)   # the four-kernel iteration (opt-in: measured slower at config 5)
        est = E.RotationEstimator(E.RotationEstimatorOptions(skip_initialization=True, pcg_rel_tolerance=1e-11))
        ok, R = est.EstimateRotations(vg, R0)
        assert ok
        out[name] = (R, est.summary.l1_iterations, est.summary.irls_iterations, est.summary.pcg_iterations)
    for name in ("csr", "csr+2lvl", "csr+2lvl fused"):
        assert out[name][1:3] == out["reference"][1:3], (name, out[name][1:], out["reference"][1:])
        assert np.abs(out[name][0] - out["reference"][0]).max() < 1e-8, name
    assert out["csr+2lvl"][3] < 0.5 * out["csr"][3], (out["csr+2lvl"][3], out["csr"][3])     # and it pays: far fewer PCG iterations
