"""Generate golden fixtures FROM THE REFERENCE'S OWN PYTHON (run in the build container only;
/root/reference does not exist on the GPU box).  Output: tests/golden/ref_pins.npz.

What the reference pins in-tree for the rasterizer path:
  (i)   SH basis:  utils/sh_utils.py eval_sh / RGB2SH
  (ii)  covariance: gs_renderer.py build_rotation / build_scaling_rotation / strip_symmetric
        via GaussianModel.setup_functions' build_covariance_from_scaling_rotation
  (iii) cameras: utils/cam_utils.py circle_poses + RCamera, utils/graphics_utils.py
The reference hard-codes device="cuda"; this script redirects those allocations to the CPU
(no reference file is modified or copied).

    python tests/golden/make_golden.py
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
sys.path.insert(0, REF)

# --- make the reference importable on a CPU-only box with missing optional deps ------------
for name in ["simple_knn", "simple_knn._C", "plyfile", "open3d", "point_e", "point_e.diffusion",
             "point_e.diffusion.configs", "point_e.diffusion.sampler", "point_e.models",
             "point_e.models.configs", "point_e.models.download", "point_e.util",
             "point_e.util.plotting", "loguru", "omegaconf", "omegaconf.dictconfig"]:
    if name not in sys.modules:
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m


class _Any:
    def __init__(self, *a, **k): pass
    def __call__(self, *a, **k): return _Any()
    def __getattr__(self, k): return _Any()


for name, attrs in {"simple_knn._C": ["distCUDA2"], "plyfile": ["PlyData", "PlyElement"],
                    "point_e.diffusion.configs": ["DIFFUSION_CONFIGS", "diffusion_from_config"],
                    "point_e.diffusion.sampler": ["PointCloudSampler"],
                    "point_e.models.configs": ["MODEL_CONFIGS", "model_from_config"],
                    "point_e.models.download": ["load_checkpoint"],
                    "point_e.util.plotting": ["plot_point_cloud"],
                    "loguru": ["logger"], "omegaconf": ["OmegaConf"],
                    "omegaconf.dictconfig": ["DictConfig"]}.items():
    for a in attrs:
        setattr(sys.modules[name], a, _Any())

_zeros = torch.zeros


def _cpu_zeros(*a, **k):
    k.pop("device", None)
    return _zeros(*a, **k)


torch.zeros = _cpu_zeros
torch.Tensor.cuda = lambda self, *a, **k: self

from utils.sh_utils import eval_sh, RGB2SH                    # noqa: E402
from utils.graphics_utils import fov2focal, focal2fov          # noqa: E402
from utils.cam_utils import circle_poses, RCamera              # noqa: E402
import gs_renderer                                              # noqa: E402

rng = np.random.RandomState(1234)
out = {}

# (i) SH: eval_sh(deg, sh[...,C,K], dirs) for deg 0..3
N = 257
dirs = rng.normal(size=(N, 3)).astype(np.float32)
dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
sh = rng.normal(0, 0.5, size=(N, 16, 3)).astype(np.float32)     # rasterizer layout [P,M,3]
out["sh_dirs"], out["sh_coeffs"] = dirs, sh
for deg in range(4):
    r = eval_sh(deg, torch.from_numpy(sh).transpose(1, 2), torch.from_numpy(dirs))
    out[f"sh_eval_deg{deg}"] = r.numpy()
out["rgb2sh_in"] = rng.random_sample((8, 3)).astype(np.float32)
out["rgb2sh_out"] = RGB2SH(torch.from_numpy(out["rgb2sh_in"])).numpy()

# (ii) covariance
gm = gs_renderer.GaussianModel.__new__(gs_renderer.GaussianModel)
gm.setup_functions()
scal = np.exp(rng.normal(-3, 1, size=(N, 3))).astype(np.float32)
rot = rng.normal(size=(N, 4)).astype(np.float32)
rot /= np.linalg.norm(rot, axis=1, keepdims=True)
out["cov_scales"], out["cov_rots"] = scal, rot
for mod in (1.0, 0.7):
    cov = gm.covariance_activation(torch.from_numpy(scal), mod, torch.from_numpy(rot))
    out[f"cov3d_mod{mod}"] = cov.numpy()

# (iii) cameras: circle_poses -> (R,T) as cam_utils.py:1383-1386 -> RCamera
class Opt:
    image_w = 640
    image_h = 480
    SSAA = 1

cams = []
for k, (radius, theta, phi, fov) in enumerate([(3.5, 60.0, 0.0, 0.55), (3.5, 60.0, 45.0, 0.55),
                                               (2.0, 80.0, 200.0, 0.96), (5.0, 30.0, 300.0, 0.4)]):
    poses = circle_poses(radius=torch.tensor([radius]), theta=torch.tensor([theta]),
                         phi=torch.tensor([phi]))
    matrix = np.linalg.inv(poses[0])
    R = -np.transpose(matrix[:3, :3])
    R[:, 0] = -R[:, 0]
    T = -matrix[:3, 3]
    fovy = focal2fov(fov2focal(fov, Opt.image_h), Opt.image_w)
    cam = RCamera(R=R, T=T, FoVx=fov, FoVy=fovy, delta_polar=0, delta_azimuth=0, delta_radius=0,
                  opt=Opt)
    out[f"cam{k}_args"] = np.array([radius, theta, phi, fov, Opt.image_h, Opt.image_w], np.float64)
    out[f"cam{k}_pose"] = poses[0]
    out[f"cam{k}_view"] = cam.world_view_transform.numpy()
    out[f"cam{k}_fullproj"] = cam.full_proj_transform.numpy()
    out[f"cam{k}_center"] = cam.camera_center.numpy()
    out[f"cam{k}_fovy"] = np.array([cam.FoVy], np.float64)
out["n_cams"] = np.array([4])

dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_pins.npz")
np.savez_compressed(dst, **out)
print("wrote", dst, {k: v.shape for k, v in out.items()})
