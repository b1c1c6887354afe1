"""Debug helper: find Gaussians whose CUDA radius differs from the oracle's and dump their inputs
and the oracle's intermediates to gpurun_out/radii_mismatch.npz (run on the GPU box)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import splat_ref as O
from tests import util_scene as U
from dreamscene_b200 import rasterizer as R

found = {}
for seed in range(100, 160):
    for phi in (float(seed % 7) * 50.0,):
        P, H, W = 1_000_000, 1024, 1024
        sc, cam, deg = U.make_inputs(P, H, W, seed=seed, phi=phi, exact_knn=False)
        dev = torch.device("cuda", 0)
        t = {k: v.to(dev) for k, v in sc.items()}
        S = U.cuda_settings(cam, deg, device=dev)
        with torch.no_grad():
            color, radii, da, _, st = R._forward_impl(S, t["means3D"], t["shs"], None, t["opacities"], t["scales"],
                                                      t["rotations"], None)
            pre = O.preprocess(U.oracle_settings(cam, deg), sc["means3D"], sc["opacities"], shs=sc["shs"],
                               scales=sc["scales"], rotations=sc["rotations"])
        torch.cuda.synchronize()
        bad = torch.nonzero(radii.cpu() != pre["radii"]).flatten()
        print("seed", seed, "phi", phi, "mismatches", bad.numel(), flush=True)
        if bad.numel():
            i = bad[:8]
            found = dict(idx=i.numpy(), means3D=sc["means3D"][i].numpy(), scales=sc["scales"][i].numpy(),
                         rotations=sc["rotations"][i].numpy(), cuda_radii=radii.cpu()[i].numpy(),
                         oracle_radii=pre["radii"][i].numpy(), view=cam.world_view_transform.numpy(),
                         proj=cam.full_proj_transform.numpy(), tanfovx=np.array([cam.tanfovx, cam.tanfovy]),
                         cov2d=np.stack([c[i].numpy() for c in pre["cov2d"]]), depth=pre["depth"][i].numpy(),
                         geom=U.decode_saved(st.saved, P, H, W, st.capacity)["geom_f32"][i.numpy()])
            break
    if found:
        break
os.makedirs("gpurun_out", exist_ok=True)
np.savez("gpurun_out/radii_mismatch.npz", **found)
print("saved", list(found))
