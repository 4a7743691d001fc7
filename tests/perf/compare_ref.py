"""'Reference CUDA build on the same box': times the reference's OWN extensions (built from /root/reference sources into
oracle/_ref by oracle/build_ref.py) and the reference's torch-eager MASt3R path against ours, and checks results."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from artdeco_b200 import synthetic  # noqa: E402
from artdeco_b200.knn import distCUDA2, distIndex2  # noqa: E402
from artdeco_b200.ssim import fused_ssim  # noqa: E402
from oracle import build_ref  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


# ---------------- fused-ssim ----------------
try:
    ref = build_ref.load("fused_ssim_ref")
    for shape in ((1, 3, 1080, 1920), (5, 5, 1080, 1920)):
        g = torch.Generator().manual_seed(0)
        i1 = torch.rand(*shape, generator=g).to(dev)
        i2 = torch.rand(*shape, generator=g).to(dev)
        C1, C2 = 0.01 ** 2, 0.03 ** 2

        def ref_fb():
            m, d1, d2, d3 = ref.fusedssim(C1, C2, i1, i2, True)
            s = m.mean()
            dl = torch.full_like(i1, 1.0 / i1.numel())
            return s, ref.fusedssim_backward(C1, C2, i1, i2, dl, d1, d2, d3)

        def our_fb():
            x = i1.detach().requires_grad_(True)
            s = fused_ssim(x, i2)
            s.backward()
            return s, x.grad
        sr, gr = ref_fb()
        so, go = our_fb()
        t_ref, t_our = timeit(ref_fb), timeit(our_fb)
        n = i1.numel()
        print(f"fused-ssim {shape}: reference {t_ref:.3f} ms, ours {t_our:.3f} ms  ({t_ref / t_our:.2f}x); "
              f"|mean diff| {abs(float(sr) - float(so)):.2e}, grad rel diff {float((gr - go).abs().max() / gr.abs().max()):.2e}; "
              f"ours {52 * n / (t_our * 1e-3) / 1e9:.0f} GB/s algorithmic")
except Exception as e:  # noqa: BLE001
    print("fused-ssim reference unavailable:", e)

# ---------------- simple-knn ----------------
try:
    ref = build_ref.load("simple_knn_ref")
    for P in (100_000, 1_000_000):
        pts = synthetic.raster_scene(P, seed=0)["means"].to(dev)
        r = ref.distCUDA2(pts)
        o = distCUDA2(pts)
        same = bool(torch.equal(r, o))
        t_ref, t_our = timeit(lambda: ref.distCUDA2(pts), 5, 1), timeit(lambda: distCUDA2(pts), 5, 1)
        print(f"distCUDA2 P={P}: reference {t_ref:.2f} ms, ours {t_our:.2f} ms ({t_ref / t_our:.1f}x); bit-identical to the reference: {same}"
              + ("" if same else f" (max rel diff {float(((r - o).abs() / r.abs().clamp_min(1e-30)).max()):.2e})"))
        K = 8
        dr, ir = ref.distIndex2(pts, K)
        do, io = distIndex2(pts, K)
        dr_s = dr.view(P, K).sort(dim=1).values
        print(f"distIndex2 K={K} P={P}: sorted distances identical to the reference: {bool(torch.equal(dr_s, do.view(P, K)))}; "
              f"reference {timeit(lambda: ref.distIndex2(pts, K), 3, 1):.2f} ms, ours {timeit(lambda: distIndex2(pts, K), 3, 1):.2f} ms")
except Exception as e:  # noqa: BLE001
    print("simple-knn reference unavailable:", e)

# ---------------- MASt3R: the reference's GPU path is torch eager fp32 with TF32 enabled (croco.py:13, run_system.py:73) ----------------
from artdeco_b200.mast3r import FULL_CFG, AsymmetricMASt3R, forward_pair  # noqa: E402
from artdeco_b200.mast3r.shapes import random_state_dict  # noqa: E402
from oracle import mast3r_torch as mt  # noqa: E402

sd = random_state_dict(FULL_CFG, dev)
img1 = torch.rand(1, 3, 512, 512, device=dev) * 2 - 1
img2 = torch.rand(1, 3, 512, 512, device=dev) * 2 - 1
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
with torch.inference_mode():
    t_tf32 = timeit(lambda: mt.forward_pair(sd, FULL_CFG, img1, img2), 5, 2)
    ref_tf32 = mt.forward_pair(sd, FULL_CFG, img1, img2)
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
with torch.inference_mode():
    t_fp32 = timeit(lambda: mt.forward_pair(sd, FULL_CFG, img1, img2), 3, 1)
    ref_fp32 = mt.forward_pair(sd, FULL_CFG, img1, img2)
m = AsymmetricMASt3R(**FULL_CFG).load_state_dict(sd).to(dev)
t_our = timeit(lambda: forward_pair(m, img1, img2), 10, 3)
ours = forward_pair(m, img1, img2)
rel = lambda a, b: float((a.double() - b.double()).abs().max() / b.double().abs().max())
print(f"MASt3R 512^2 pair (batch 1): torch-eager TF32 (the reference's GPU mode) {t_tf32:.2f} ms, torch-eager fp32 {t_fp32:.2f} ms, "
      f"ours bf16x3 {t_our:.2f} ms ({t_tf32 / t_our:.2f}x vs TF32 eager)")
print(f"  pts3d error vs fp32 eager: ours {rel(ours[0]['pts3d'], ref_fp32[0]['pts3d']):.2e}, TF32 eager {rel(ref_tf32[0]['pts3d'], ref_fp32[0]['pts3d']):.2e}")
