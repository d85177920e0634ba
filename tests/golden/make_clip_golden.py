"""Generate tests/golden/clip_*.npz from the UNMODIFIED reference: the non-default `n_shadow_importance_clip > 0` option
(models/neus_hint_model.py:554-576: one shadow march per group of samples).  Run in the build container only:
    python tests/golden/make_clip_golden.py
The oracle does not restate this option (it is served by nrhints_b200/hint_fallback.py, checked against the live reference on the
CPU in tests/test_hint_fallback.py), so these fixtures hold the reference's inputs and RenderOutput only."""
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE))
import nrh_testlib as T  # noqa: E402
from make_golden import import_reference  # noqa: E402
import nrhints_b200 as nb  # noqa: E402

CASES = {
    # name: (clip, weights, R, renderer kwargs shared by the reference and the product config)
    "clip4_sharp_40x64": (4, "sharp", 40, dict(n_samples=32, n_importance_samples=32, n_shadow_samples=32, n_shadow_importance_samples=32)),
    "clip16_init_24x128": (16, "init", 24, dict()),
}


def main():
    M, RayBundle = import_reference()
    torch.set_num_threads(8)
    from oracle import nrh_oracle as orc
    for name, (clip, weights, R, kw) in CASES.items():
        cfg = nb.NeuSModelConfig(renderer=nb.NeuSRendererConfig(**kw))
        sd = T.make_state(weights, cfg)
        ref = M.NeuSHintRenderer(M.NeuSModelConfig(renderer=M.NeuSRendererConfig(n_shadow_importance_clip=clip, **kw)))
        ref.load_state_dict(sd, strict=True)
        rays = orc.synthetic_rays(R, seed=100 + clip, crop=300)
        with torch.no_grad():
            out = ref(RayBundle(**rays), background_rgb=torch.ones(1, 3))
        data = {"in_" + k: v.numpy() for k, v in rays.items()}
        for k in T.OUT_FIELDS:
            v = getattr(out, k, None)
            if v is not None:
                data["out_" + k] = v.detach().numpy()
        data["digest"] = np.array(T.state_digest(sd))
        data["clip"] = np.array(clip)
        np.savez_compressed(HERE / f"{name}.npz", **data)
        print(name, {k: v.shape for k, v in data.items() if k.startswith("out_")}, "mean visibility", float(out.visibilities.mean()))


if __name__ == "__main__":
    main()
