"""Generate tests/golden/raygen.npz from the UNMODIFIED reference ray generator.

Run in the build container only (needs /root/reference):
    python tests/golden/make_raygen_golden.py
For every case of nrh_testlib.RAYGEN_CASES this builds the reference RayGenerator (/root/reference/camera/ray_generator.py),
sets its parameters / noise buffers to the seeded values, runs forward on the seeded RawPixelBundle, back-propagates a seeded
linear functional of the five RayBundle fields, and stores inputs, outputs and parameter gradients.  The oracle
(oracle/raygen_oracle.py) is run on the same inputs and must agree, which is what pins it to the reference.
"""
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parent.parent))
import nrh_testlib as T  # noqa: E402
from oracle import raygen_oracle as rgo  # noqa: E402

REF = Path("/root/reference")


def main():
    sys.path.insert(0, str(REF))
    from camera.camera_model import CameraModel
    from camera.ray_generator import RayGenerator, RayGeneratorConfig
    from data.data_loader import RawPixelBundle
    store = {}
    for name, case in T.RAYGEN_CASES.items():
        inp = T.raygen_inputs(case)
        cam = inp["camera"]
        cfg = RayGeneratorConfig(override_near_far_from_sphere=case["override_near_far"], cam_opt_mode=case["cam_opt_mode"],
                                 pl_opt=case["pl_opt"],
                                 cam_position_noise_std=0.01 if case["noise"] else 0.0,
                                 cam_orientation_noise_std=0.01 if case["noise"] else 0.0,
                                 pl_position_noise_std=0.01 if case["noise"] else 0.0)
        gen = RayGenerator(CameraModel(H=cam["H"], W=cam["W"], cx=cam["cx"], cy=cam["cy"], fx=cam["fx"], fy=cam["fy"],
                                       zn=cam["zn"], zf=cam["zf"]), inp["n_cameras"], cfg)
        with torch.no_grad():
            if case["cam_opt_mode"] != "off":
                gen.cam_pose_adjustment.copy_(inp["cam_pose_adjustment"])
            if case["pl_opt"]:
                gen.pl_adjustment.copy_(inp["pl_adjustment"])
            if case["noise"]:
                gen.cam_pose_noise.copy_(inp["cam_pose_noise"])
                gen.pl_noise.copy_(inp["pl_noise"])
        img = inp["img_indices"]
        bundle = RawPixelBundle(img_indices=img[:, None] if img is not None else None, h_indices=inp["h_indices"][:, None],
                                w_indices=inp["w_indices"][:, None], poses=inp["poses"], pls=inp["pls"], rgb_gt=None)
        out = gen(bundle)
        fields = {"origins": out.origins, "directions": out.directions, "pl_positions": out.pl_positions,
                  "nears": out.nears, "fars": out.fars}
        cot = T.raygen_cotangents(case)
        params = [p for p in gen.parameters()]
        grads = {}
        if params:
            loss = sum((fields[k] * cot[k]).sum() for k in fields)
            if loss.requires_grad:                       # video views touch no parameter
                loss.backward()
            for n, p in gen.named_parameters():
                grads[n] = p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p)
        want = rgo.raygen_forward(cam, case["cam_opt_mode"], case["override_near_far"], inp["w_indices"], inp["h_indices"], img,
                                  inp["poses"], inp["pls"], inp.get("cam_pose_noise") if case["noise"] else None,
                                  inp.get("pl_noise") if case["noise"] else None,
                                  inp.get("cam_pose_adjustment") if case["cam_opt_mode"] != "off" else None,
                                  inp.get("pl_adjustment") if case["pl_opt"] else None)
        for k, v in fields.items():
            err = float((v.detach() - want[k]).abs().max())
            assert err < 2e-6, (name, k, err)
            store[f"{name}.out_{k}"] = v.detach().numpy()
        for n, g in grads.items():
            store[f"{name}.grad_{n}"] = g.numpy()
        print(name, "ok;", {n: float(g.abs().max()) for n, g in grads.items()})
    # the constructor's RNG draws (ray_generator.py:62-73): noise buffers for a fixed seed, and the state_dict layout
    torch.manual_seed(123)
    gen = RayGenerator(CameraModel(H=8, W=8, cx=4.0, cy=4.0, fx=10.0, fy=10.0, zn=0.1, zf=10.0), 5,
                       RayGeneratorConfig(cam_opt_mode="SE3", pl_opt=True, cam_position_noise_std=0.02, cam_orientation_noise_std=0.01,
                                          pl_position_noise_std=0.03))
    for k, v in gen.state_dict().items():
        store[f"ctor_seed123.{k}"] = v.detach().numpy()
    store["ctor_seed123.keys"] = np.array(list(gen.state_dict().keys()))
    np.savez_compressed(HERE / "raygen.npz", **store)
    print("written", HERE / "raygen.npz")


if __name__ == "__main__":
    main()
