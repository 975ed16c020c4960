"""TEST INFRASTRUCTURE ONLY — golden vectors for the create_data `rand` host logic, produced by the REAL
reference (run in the build container, where /root/reference exists): the pose / focal random sequence of
utils/create_data.py:815-818 with dataset/load_blender.py:359-368 `get_rand_pose` (AST-extracted, because
load_blender.py imports imageio at module level), under np.random.seed(0).

    python oracle/make_golden_create_data.py   ->  tests/golden/create_data.npz
"""
import ast
import os
import sys

import numpy as np
import torch

REF = os.environ.get("R2L_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    sys.path.insert(0, REF)
    import utils.run_nerf_raybased_helpers as Hh   # to_tensor
    torch.autograd.set_detect_anomaly(False)
    src = open(os.path.join(REF, "dataset", "load_blender.py")).read()
    tree = ast.parse(src)
    keep = []
    for node in tree.body:
        if isinstance(node, ast.Assign) and any(getattr(t, "id", "") in ("trans_t", "rot_phi", "rot_theta") for t in node.targets):
            keep.append(node)
        if isinstance(node, ast.FunctionDef) and node.name in ("pose_spherical", "get_rand_pose"):
            keep.append(node)
    ns = {"torch": torch, "np": np, "to_tensor": Hh.to_tensor}
    exec(compile(ast.Module(keep, []), "load_blender.py", "exec"), ns)
    np.random.seed(0)
    poses, fdraws = [], []
    for _ in range(8):
        poses.append(ns["get_rand_pose"]().cpu().numpy())      # 2 draws
        fdraws.append(np.random.rand())                        # focal_ = focal * (rand + 1)
    perm1 = np.random.permutation(1000)
    perm2 = np.random.permutation(1000)
    np.savez(os.path.join(OUT, "create_data.npz"), poses=np.stack(poses), fdraws=np.array(fdraws), perm1=perm1,
             perm2=perm2, fixed_pose=ns["pose_spherical"](37.5, -42.0, 4.0).numpy())
    print("wrote", os.path.join(OUT, "create_data.npz"))


if __name__ == "__main__":
    main()
