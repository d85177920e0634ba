"""Golden fixture of the hash encoding from the UNMODIFIED reference (fields/encodings.py:237-371, torch backend).
Run in the build container only:  python tests/golden/make_hash_golden.py
Stores seeded points, a seeded (small, log2 T = 10) table, and the reference output + table gradient."""
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, "/root/reference")
sys.path.insert(0, str(HERE.parent.parent))


def main():
    from fields.encodings import HashEncoding
    from oracle import hash_oracle as ho
    torch.manual_seed(1234)
    enc = HashEncoding(num_levels=16, min_res=16, max_res=1024, log2_hashmap_size=10, features_per_level=2,
                       hash_init_scale=0.001, implementation="torch")
    g = torch.Generator().manual_seed(5)
    pts = torch.rand(300, 3, generator=g)
    pts[:8] = torch.tensor([[0, 0, 0], [1, 1, 1], [0.5, 0.5, 0.5], [1, 0, 0.25], [0.0625, 0.125, 0.9375],
                            [0.999999, 1e-7, 0.3], [0.25, 0.75, 1.0], [1.0, 0.0, 0.0]])       # grid-aligned / boundary points
    out = enc(pts)
    d_out = torch.randn(out.shape, generator=g)
    (out * d_out).sum().backward()
    table = enc.hash_table.detach().numpy()
    want = ho.hash_encode(pts.numpy(), table, enc.scalings.numpy(), 10)
    assert np.array_equal(want, out.detach().numpy()), "oracle is not bit-identical to the reference"
    np.testing.assert_array_equal(ho.scalings(), enc.scalings.numpy())
    gt = ho.hash_encode_table_grad(pts.numpy(), d_out.numpy(), enc.scalings.numpy(), 10, table.shape[0], 2)
    np.testing.assert_allclose(gt, enc.hash_table.grad.numpy(), rtol=2e-5, atol=1e-7)
    np.savez_compressed(HERE / "hash_16x2_T10.npz", pts=pts.numpy(), table=table, scalings=enc.scalings.numpy(),
                        out=out.detach().numpy(), d_out=d_out.numpy(), table_grad=enc.hash_table.grad.numpy())
    print("hash fixture written; oracle bit-identical to the reference on", pts.shape[0], "points")


if __name__ == "__main__":
    main()
