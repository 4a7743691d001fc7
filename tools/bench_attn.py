"""Fused attention kernel alone at the two MASt3R shapes of the B=4 bench (encoder: 8 images x 16 heads, decoder side: 4 x 12;
1024 queries x 1024 keys, head_dim 64), against its tensor-pipe floor.  With ADB_CUPROF=1 one call per shape is bracketed by
cudaProfilerStart/Stop for `ncu --profile-from-start off`."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from artdeco_b200.mast3r import ops  # noqa: E402


def run(B, h, N, reps=20, Nk=None):
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    mk = lambda *s: ops.split(torch.randn(*s, generator=g).to(dev))
    Nk = N if Nk is None else Nk
    q, k = mk(B * h * N, 64), mk(B * h * Nk, 64)
    vt = mk(B * h * 64, Nk)
    f = lambda: ops.attention(q, k, vt, B, h, N, Nk, Nk, 0.125, x3=True)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        f()
    b.record()
    torch.cuda.synchronize()
    if os.environ.get("ADB_CUPROF"):
        torch.cuda.cudart().cudaProfilerStart()
        f()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
    us = a.elapsed_time(b) / reps * 1e3
    # tensor work actually issued per (query tile, key block): QK hi*hi (4 MMAs N=128) + QK x3 (12) + PV x3 (24 MMAs N=64)
    flop = B * h * (N // 128) * (Nk // 128) * (16 * 2 * 128 * 128 * 16 + 24 * 2 * 128 * 64 * 16)
    algo = 4.0 * B * h * N * Nk * 64
    return {"B": B, "heads": h, "N": N, "Nk": Nk, "us": us, "issued_tensor_TFLOPs": flop / us / 1e6, "algorithmic_TFLOPs": algo / us / 1e6}


if __name__ == "__main__":
    if os.environ.get("ADB_SWEEP"):
        print(json.dumps([run(8, 16, 1024, Nk=nk) for nk in (128, 256, 512, 1024, 2048)]))
    else:
        print(json.dumps([run(8, 16, 1024), run(4, 12, 1024), run(4, 12, 768)]))
