"""TEST INFRASTRUCTURE ONLY.  tests/golden/msda_module_init.npz: the initial parameters of the reference's own MSDeformAttn3D
(ops/modules/ms_deform_attn.py:117-165) for the shipped geometry (d_model 256, 4 levels, 8 heads, 8 points) under
torch.manual_seed(0): the deterministic offset bias in full, and the two Xavier-initialised projections as checksums + a corner.
Run in the build container only:   python -m oracle.make_golden_msda_init"""
import importlib
import os
import numpy as np
import torch

from .ref_import import import_reference
from .make_golden import GOLD


def main():
    import_reference()
    mod = importlib.import_module("ops.modules.ms_deform_attn")
    torch.manual_seed(0)
    m = mod.MSDeformAttn3D(d_model=256, n_levels=4, n_heads=8, n_points=8)
    sd = {k: v.detach().numpy() for k, v in m.state_dict().items()}
    np.savez_compressed(os.path.join(GOLD, "msda_module_init.npz"),
                        sampling_offsets_bias=sd["sampling_offsets.bias"],
                        value_proj_corner=sd["value_proj.weight"][:4, :8].copy(), output_proj_corner=sd["output_proj.weight"][:4, :8].copy(),
                        value_proj_sum=np.float64(sd["value_proj.weight"].astype(np.float64).sum()),
                        output_proj_sum=np.float64(sd["output_proj.weight"].astype(np.float64).sum()),
                        zero_max=np.float32(max(np.abs(sd[k]).max() for k in ("sampling_offsets.weight", "attention_weights.weight",
                                                                               "attention_weights.bias", "value_proj.bias", "output_proj.bias"))))
    print({k: v.shape for k, v in sd.items()})


if __name__ == "__main__":
    main()
