"""CPU tests pinning the RA oracle to synthetic ground truth with the
reference's thresholds (rotation_averager_test.cc:167,310)."""
import numpy as np

from glomap_b200 import estimators as E, geometry as G, synthetic as S
from oracle import ra_oracle as RA


def test_exp_log_roundtrip_and_small_angle_fallback():
    rng = np.random.default_rng(0)
    v = rng.normal(size=(200, 3))
    v *= (rng.uniform(0, np.pi - 1e-3, size=(200, 1)) / np.linalg.norm(v, axis=1, keepdims=True))
    assert np.abs(RA.R_to_aa(RA.aa_to_R(v)) - v).max() < 1e-9
    tiny = np.array([[1e-13, -2e-13, 3e-13]])
    R = RA.aa_to_R(tiny)[0]
    assert R[0, 1] == -tiny[0, 2] and R[2, 2] == 1.0          # math/rigid3d.cc:50-61


def test_noise_free_graph_exact():
    vg = S.make_random_view_graph(60, 8, seed=1)
    R0 = E.initialize_from_maximum_spanning_tree(vg)
    th, info = RA.estimate_rotations(vg.n_images, vg.ei, vg.ej, vg.R_rel, G.so3_log(R0))
    assert RA.max_pairwise_rotation_error_deg(th, vg.R_gt) < 1e-2


def test_noisy_graph_with_outliers():
    vg = S.make_random_view_graph(200, 12, seed=3, noise_deg=1.0, outlier_ratio=0.05)
    th, info = RA.estimate_rotations(vg.n_images, vg.ei, vg.ej, vg.R_rel, np.zeros((200, 3)))
    assert RA.max_pairwise_rotation_error_deg(th, vg.R_gt) < 3.0


def test_mst_initialisation_is_exact_without_noise():
    vg = S.make_ring_view_graph(100, 5, seed=1)
    R0 = E.initialize_from_maximum_spanning_tree(vg)
    # relative rotations along every edge are reproduced exactly
    assert G.rotation_angle_deg(R0[vg.ej] @ np.swapaxes(R0[vg.ei], -1, -2), vg.R_rel).max() < 1e-5   # arccos resolution near 0 is ~1e-6 deg
