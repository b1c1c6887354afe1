"""Pin the oracle (and the host camera helper) against fixtures produced by the REFERENCE's
own Python (tests/golden/make_golden.py -> ref_pins.npz).  CPU only."""
import os

import numpy as np
import torch

from oracle import splat_ref as O
from harness import cameras

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_pins.npz"))


def test_sh_basis_matches_reference_eval_sh():
    d = torch.from_numpy(G["sh_dirs"])
    sh = torch.from_numpy(G["sh_coeffs"])
    for deg in range(4):
        got = O.eval_sh_basis_dot(deg, sh, d).numpy()
        np.testing.assert_allclose(got, G[f"sh_eval_deg{deg}"], rtol=1e-5, atol=2e-6)


def test_rgb2sh_constant():
    np.testing.assert_allclose((G["rgb2sh_in"] - 0.5) / O.SH_C0, G["rgb2sh_out"], rtol=1e-6)


def test_cov3d_matches_reference_build_covariance():
    s, q = torch.from_numpy(G["cov_scales"]), torch.from_numpy(G["cov_rots"])
    for mod in (1.0, 0.7):
        got = O.cov3d_from_scale_rot(s, q, mod, torch.float32).numpy()
        ref = G[f"cov3d_mod{mod}"]
        np.testing.assert_allclose(got, ref, rtol=2e-4, atol=1e-9)


def test_camera_helper_matches_reference_rcamera():
    for k in range(int(G["n_cams"][0])):
        radius, theta, phi, fov, h, w = G[f"cam{k}_args"]
        pose = cameras.orbit_pose(radius, theta, phi)
        np.testing.assert_allclose(pose, G[f"cam{k}_pose"], atol=2e-6)
        cam = cameras.camera_from_pose(G[f"cam{k}_pose"], float(fov), int(h), int(w))
        np.testing.assert_allclose(cam.world_view_transform.numpy(), G[f"cam{k}_view"], atol=1e-6)
        np.testing.assert_allclose(cam.full_proj_transform.numpy(), G[f"cam{k}_fullproj"], atol=1e-6)
        np.testing.assert_allclose(cam.camera_center.numpy(), G[f"cam{k}_center"], atol=1e-6)
        assert abs(cam.FoVy - float(G[f"cam{k}_fovy"][0])) < 1e-12


def test_projection_convention_w_clip_is_view_depth():
    # graphics_utils.py:61-81 with z_sign=+1: full_proj's 4th column equals the view z column
    for k in range(int(G["n_cams"][0])):
        V, F = G[f"cam{k}_view"], G[f"cam{k}_fullproj"]
        np.testing.assert_allclose(F[:, 3], V[:, 2], atol=1e-6)
