"""peer.PeerExchange (csrc/peer_exchange.cu): the gradient exchange of the multi-view step by this library's own kernels over
peer memory.  On a one-GPU box the protocol is exercised by several "ranks" living in one process — each with its own region,
flags and stream, and a pointer table naming the others' regions — which runs exactly the kernels and the signal / wait
protocol of the multi-process case (the only thing left out is the CUDA IPC mapping, covered by the 2-GPU test below and by
bench.py --gpus N)."""
import ctypes as C
import os
import subprocess
import sys

import pytest
import torch

from artdeco_b200 import synthetic
from helpers import rel_err

KEYS = ("means", "quats", "scales", "opacities", "sh")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _virtual_ranks(N, Cl, world, dev, timeout_s=5.0):
    from artdeco_b200 import _lib
    from artdeco_b200.peer import PeerExchange, PeerRegion
    words = PeerRegion(N, Cl, world).words
    bases = []
    with torch.cuda.device(dev):
        for _ in range(world):
            p = C.c_void_p()
            _lib.call("adb_peer_alloc", words * 4, C.byref(p))
            bases.append(p.value)
    exs = [PeerExchange(N, Cl, dev, rank=r, world=world, base=bases[r], peers=bases, timeout_s=timeout_s) for r in range(world)]
    return exs, bases


def _free(bases, dev):
    from artdeco_b200 import _lib
    torch.cuda.synchronize(dev)
    for b in bases:
        _lib.call("adb_peer_free", C.c_void_p(b))


@pytest.mark.gpu
@pytest.mark.parametrize("world,Cl,N", [(2, 2, 10000), (4, 1, 4096), (8, 1, 20004), (3, 2, 1000)])
def test_gather_and_reduce_kernels_match_torch(cuda, world, Cl, N):
    """Two steps (both parities of the colour table): every rank ends with every rank's masked colour rows, the camera centres and
    the rank-ordered sum of the geometry blocks; all ranks hold bit-identical sums."""
    from artdeco_b200 import raster as R
    exs, bases = _virtual_ranks(N, Cl, world, cuda)
    g = torch.Generator().manual_seed(world * 100 + Cl)
    for step in range(3):
        splats = [torch.randn(Cl, N, 12, generator=g).to(cuda) for _ in range(world)]
        v_splats = [torch.randn(Cl, N, 12, generator=g).to(cuda) for _ in range(world)]
        campos = [torch.randn(Cl, 3, generator=g).to(cuda) for _ in range(world)]
        parts = [torch.randn(N * 11, generator=g).to(cuda) for _ in range(world)]
        for r, ex in enumerate(exs):
            for c in range(Cl):
                if step == 1:       # the ready-made-row entry point (autograd path)
                    row = R.mask_rgb_grad(splats[r][c], v_splats[r][c], torch.empty(N, 3, device=cuda))
                    ex.start_gather_view(c, row, campos[r] if c == 0 else None)
                else:
                    ex.push_view(c, splats[r][c], v_splats[r][c], campos[r] if c == 0 else None)
            ex.geom_in[:N * 11].copy_(parts[r])
            ex.start_reduce()
        outs = []
        for ex in exs:
            g_all, p_all = ex.wait_gather()
            ex.wait_reduce()
            outs.append((g_all.clone(), p_all.clone(), torch.cat([ex.views[k].reshape(-1) for k in
                                                                  ("v_means", "v_quats", "v_scales", "v_opac")]).clone()))
        torch.cuda.synchronize()
        for ex in exs:
            ex.check()
        want_g = torch.stack([R.mask_rgb_grad(splats[r][c], v_splats[r][c], torch.empty(N, 3, device=cuda))
                              for c in range(Cl) for r in range(world)])
        want_p = torch.stack([campos[r][c] for c in range(Cl) for r in range(world)])
        want_sum = parts[0].clone()
        for r in range(1, world):
            want_sum += parts[r]                       # same (rank) order as the kernel: bit-exact
        for g_all, p_all, red in outs:
            assert torch.equal(g_all, want_g), f"step {step}: gathered colour rows"
            assert torch.equal(p_all, want_p), f"step {step}: camera centres"
            assert torch.equal(red, want_sum), f"step {step}: reduced geometry block"
    _free(bases, cuda)


@pytest.mark.gpu
def test_wait_times_out_instead_of_hanging(cuda):
    """A rank that never signals must not hang the GPU: the wait kernel gives up after its timeout and sets the error word."""
    from artdeco_b200 import _lib
    exs, bases = _virtual_ranks(1024, 1, 2, cuda, timeout_s=0.05)
    exs[0].push_view(0, torch.randn(1, 1024, 12, device=cuda)[0], torch.randn(1, 1024, 12, device=cuda)[0], torch.zeros(1, 3, device=cuda))
    exs[0].wait_gather()                    # rank 1 never pushes
    torch.cuda.synchronize()
    with pytest.raises(_lib.ArtdecoB200Error, match="timed out"):
        exs[0].check()
    _free(bases, cuda)


@pytest.mark.gpu
def test_multi_view_backward_through_peer_exchange_matches_single_process(cuda):
    """4 views split over two virtual ranks, gradients exchanged by PeerExchange: every rank must end with the single-process
    4-view gradients (geometry: sum of two partial sums; SH: expanded from the gathered colour rows)."""
    from artdeco_b200 import raster as R
    N, W, H = 20000, 480, 272
    sc = synthetic.raster_scene(N, seed=8)
    cams = [synthetic.camera(W, H, view=v) for v in (1.0, 3.0, 5.0, 7.0)]
    Vs = torch.stack([c[0] for c in cams]).to(cuda)
    Ks = torch.stack([c[1] for c in cams]).to(cuda)
    P = torch.inverse(Vs)[:, :3, 3].contiguous()
    t = {k: sc[k].to(cuda) for k in KEYS}
    g = torch.Generator().manual_seed(6)
    vc, va = torch.randn(4, H, W, 4, generator=g).to(cuda), torch.randn(4, H, W, generator=g).to(cuda)

    def local(idx):
        Cn = len(idx)
        radii = torch.empty(Cn, N, 2, dtype=torch.int32, device=cuda)
        splats = torch.empty(Cn, N, 12, device=cuda)
        tpg = torch.empty(Cn, N, dtype=torch.int32, device=cuda)
        v_splats = torch.zeros(Cn, N, 12, device=cuda)
        for j, c in enumerate(idx):
            R.project(t["means"], t["quats"], t["scales"], t["opacities"], t["sh"], 3, Vs[c], Ks[c], P[c], W, H, 0.01, 0.01,
                      1e10, 0.0, out=(radii[j], splats[j], tpg[j]))
            keys, vals, offs, _ = R.intersect(radii[j], splats[j], tpg[j], W, H)
            col, alp, last = R.blend_forward(W, H, N, splats[j], vals, offs)
            R.blend_backward(W, H, N, splats[j], vals, offs, alp, last, vc[c].contiguous(), va[c].contiguous(), out=v_splats[j])
        return radii, splats, v_splats

    ra, sa, va_ = local([0, 1, 2, 3])
    full = R.multi_view_backward(t["means"], t["quats"], t["scales"], t["sh"], 3, Vs, Ks, P, W, H, ra, sa, va_)
    world = 2
    exs, bases = _virtual_ranks(N, 2, world, cuda)
    # Both "ranks" are driven by ONE host thread here: rank 0's wait kernels spin until this thread has enqueued rank 1's work, so
    # nothing launched in between may be a first launch (lazy module loading waits for the device to drain).  Load the two
    # kernels of the split SH backward now (the single-process call above used the fused one).
    from artdeco_b200 import _lib
    z = torch.zeros(4, 48, device=cuda)
    _lib.call("adb_raster_sh_dir_bwd_multi", 4, 1, _lib.ptr(z), _lib.ptr(z), 3, _lib.ptr(z), _lib.ptr(z), _lib.ptr(z),
              _lib.ptr(torch.zeros(4, 3, device=cuda)), None, _lib.stream())
    _lib.call("adb_raster_sh_expand_multi", 4, 1, _lib.ptr(z), 3, _lib.ptr(z), _lib.ptr(z), _lib.ptr(torch.zeros(4, 48, device=cuda)),
              _lib.stream())
    torch.cuda.synchronize()
    views = [[0, 2], [1, 3]]                               # parallel.views_for_rank
    loc = [local(v) for v in views]
    streams = [torch.cuda.Stream(device=cuda) for _ in range(world)]
    res = [None] * world
    torch.cuda.synchronize()
    for r in range(world):                                  # one stream per "rank": a rank's waits must not block the other's kernels
        with torch.cuda.stream(streams[r]):
            ra_, sa_, va2 = loc[r]
            for c in range(2):
                exs[r].push_view(c, sa_[c], va2[c], P[views[r]].contiguous() if c == 0 else None)
    for r in range(world):
        with torch.cuda.stream(streams[r]):
            ra_, sa_, va2 = loc[r]
            out = {"v_sh": torch.empty(N, 16, 3, device=cuda)}
            out.update(exs[r].views)
            res[r] = R.multi_view_backward(t["means"], t["quats"], t["scales"], t["sh"], 3, Vs[views[r]], Ks[views[r]],
                                           P[views[r]].contiguous(), W, H, ra_, sa_, va2, out=out, exchange=exs[r])
    torch.cuda.synchronize()
    for r in range(world):
        exs[r].check()
        for k, name in enumerate(("v_means", "v_quats", "v_scales", "v_opac", "v_sh")):
            assert rel_err(res[r][k], full[k]) < 2e-5, f"rank {r}: {name}"
    for k in range(4):
        assert torch.equal(res[0][k], res[1][k]), "every rank holds bit-identical reduced gradients"
    _free(bases, cuda)


@pytest.mark.gpu
def test_two_processes_over_cuda_ipc(cuda):
    """Real peer mapping (CUDA IPC) between two processes on two GPUs: tools/check_peer_exchange.py compares the peer-memory
    exchange with the NCCL exchange on the same step."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "check_peer_exchange.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0 and "PEER-EXCHANGE-OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
