"""Hash-grid encoding: oracle vs the committed reference fixture (CPU), CUDA operator vs
oracle and fixture (GPU, bit-exact for the forward), ragged / empty inputs, table gradient."""
import numpy as np
import pytest
import torch

import nrh_testlib as T
from oracle import hash_oracle as ho

FX = T.GOLDEN_DIR / "hash_16x2_T10.npz"


def test_oracle_bit_exact_against_reference_fixture():
    fx = np.load(FX)
    np.testing.assert_array_equal(ho.scalings(), fx["scalings"])
    got = ho.hash_encode(fx["pts"], fx["table"], fx["scalings"], 10)
    np.testing.assert_array_equal(got, fx["out"])
    g = ho.hash_encode_table_grad(fx["pts"], fx["d_out"], fx["scalings"], 10, fx["table"].shape[0], 2)
    np.testing.assert_allclose(g, fx["table_grad"], rtol=2e-5, atol=1e-7)


def test_hash_has_no_32bit_wrap():
    """(y * 2654435761) is evaluated in int64 (fields/encodings.py:317-319): differs from the uint32 Instant-NGP hash."""
    ix, iy, iz = np.array([3]), np.array([1000]), np.array([7])
    h = ho.hash_fn(ix, iy, iz, np.array([0]), 19)
    want = ((3 * 1) ^ (1000 * 2654435761) ^ (7 * 805459861)) % (1 << 19)
    wrapped = ((3 * 1) ^ ((1000 * 2654435761) & 0xFFFFFFFF) ^ ((7 * 805459861) & 0xFFFFFFFF)) % (1 << 19)
    assert int(h[0]) == want
    assert want == wrapped          # low 19 bits agree: the wrap only matters above bit 31 -> both forms are equivalent mod 2^19


def test_module_mirrors_reference_constructor():
    import nrhints_b200.encodings as E
    torch.manual_seed(1234)
    enc = E.HashEncoding(log2_hashmap_size=10)
    fx = np.load(FX)
    np.testing.assert_array_equal(enc.hash_table.detach().numpy(), fx["table"])       # same seeded init as the reference
    np.testing.assert_array_equal(enc.scalings.numpy(), fx["scalings"])
    assert enc.get_out_dim() == 32 and list(enc.state_dict()) == ["hash_table"]
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        enc(torch.rand(4, 3))


@pytest.mark.gpu
def test_cuda_forward_bit_exact_and_table_gradient():
    import nrhints_b200.encodings as E
    fx = np.load(FX)
    torch.manual_seed(1234)
    enc = E.HashEncoding(log2_hashmap_size=10).cuda()
    pts = torch.tensor(fx["pts"]).cuda()
    out = enc(pts)
    np.testing.assert_array_equal(out.detach().cpu().numpy(), fx["out"])               # vs the reference itself
    (out * torch.tensor(fx["d_out"]).cuda()).sum().backward()
    np.testing.assert_allclose(enc.hash_table.grad.cpu().numpy(), fx["table_grad"], rtol=1e-4, atol=1e-7)
    # batched shape, ragged N, empty input
    assert enc(pts.reshape(3, 100, 3)).shape == (3, 100, 32)
    assert enc(pts[:0]).shape == (0, 32)
    for n in (1, 63, 65, 129):
        np.testing.assert_array_equal(enc(pts[:n]).detach().cpu().numpy(), fx["out"][:n])


@pytest.mark.gpu
def test_cuda_full_size_table_matches_oracle():
    """Reference-default table (16 levels x 2^19 x 2 = 64 MiB) on 100k points: bit-exact vs the numpy oracle, including
    points outside [0,1] (negative cells: floor-mod semantics of torch's %)."""
    import nrhints_b200.encodings as E
    torch.manual_seed(7)
    enc = E.HashEncoding().cuda()
    g = torch.Generator().manual_seed(3)
    pts = torch.rand(100000, 3, generator=g)
    pts[:1000] = pts[:1000] * 3.0 - 1.0
    out = enc(pts.cuda()).detach().cpu().numpy()
    want = ho.hash_encode(pts.numpy(), enc.hash_table.detach().cpu().numpy(), enc.scalings.numpy(), 19)
    np.testing.assert_array_equal(out, want)
