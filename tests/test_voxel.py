"""update_voxel (SURVEY.md §8a R9, h3dgsv3.py:227-316): integer work, BIT-EXACT against the reference's op sequence restated
on the CPU (oracle/voxel_ref.py).  CPU test pins the oracle's properties; GPU tests compare the hash-table kernels with it."""
import pytest
import torch

from oracle import voxel_ref


def _cloud(n, seed, extent=3.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(n, 3, generator=g) * 2 - 1) * extent


def test_oracle_voxel_properties():
    xyz = torch.tensor([[0.01, 0.01, 0.01], [0.02, 0.03, 0.04], [0.05, 0.05, 0.05], [0.51, 0.0, 0.0], [0.52, 0.0, 0.0]])
    cls = torch.tensor([[7], [3], [3], [5], [2]])
    new = torch.tensor([[0.03, 0.03, 0.03], [0.55, 0.01, 0.01], [2.0, 2.0, 2.0], [2.01, 2.01, 2.01], [-1.0, 0.0, 0.0]])
    o, n, cnt = voxel_ref.update_voxel(new, xyz, cls, 0.1)
    assert o.squeeze(-1).tolist() == [3, 3, 3, 2, 2]          # majority (3 beats 7); tie 5 vs 2 -> the smaller id
    assert n.squeeze(-1).tolist()[:2] == [3, 2] and cnt == 2  # hits take the voxel's class
    assert n.squeeze(-1).tolist()[2:] == [9, 9, 8]            # new voxels: max_cls+1 + rank in sorted hash order (x-major)
    ids, c = voxel_ref.update_voxel(new, xyz[:0], cls[:0], 0.1)
    assert c == 4 and ids.squeeze(-1).tolist() == [1, 2, 3, 3, 0]


@pytest.mark.gpu
@pytest.mark.parametrize("N,M,ncls", [(0, 5000, 1), (20000, 8000, 50), (200000, 60000, 4000), (3000, 0, 10)])
def test_update_voxel_bit_exact(cuda, N, M, ncls):
    from artdeco_b200.voxel import update_voxel
    xyz, new = _cloud(N, 1), _cloud(M, 2, extent=3.5)
    g = torch.Generator().manual_seed(3)
    cls = torch.randint(0, ncls, (N, 1), generator=g)
    ref = voxel_ref.update_voxel(new, xyz, cls, 0.1)
    out = update_voxel(new.to(cuda), xyz.to(cuda), cls.to(cuda), 0.1)
    assert len(out) == len(ref)
    if N == 0:
        assert out[1] == ref[1] and torch.equal(out[0].cpu(), ref[0])
    else:
        assert out[2] == ref[2], "number of new voxels"
        assert torch.equal(out[0].cpu(), ref[0]), "majority class of every original point's voxel"
        assert torch.equal(out[1].cpu(), ref[1]), "class ids of the new points"
        assert out[0].dtype == torch.long and out[0].shape == (N, 1) and out[1].shape == (M, 1)
