"""Times MASt3R pair inference (2x encode, decoder, 2x head) with CUDA events; ADB_PREC=bf16x3|bf16, ADB_B=batch."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from artdeco_b200 import _lib  # noqa: E402
from artdeco_b200.mast3r import FULL_CFG, AsymmetricMASt3R  # noqa: E402
from artdeco_b200.mast3r.model import forward_pair  # noqa: E402


def random_state(cfg, dev, seed=0):
    """Random-init weights generated ON THE DEVICE (fast); same scaling rules as synthetic.det_weights."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from artdeco_b200.mast3r.shapes import param_shapes
    g = torch.Generator(device=dev).manual_seed(seed)
    sd = {}
    for name, shape in param_shapes(cfg).items():
        t = torch.randn(*shape, generator=g, device=dev)
        if len(shape) == 1:
            t = 1.0 + 0.1 * t if name.endswith(".weight") else 0.02 * t
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            t = t * (1.0 / fan_in) ** 0.5
            if name.endswith(".dpt.head.4.weight") or name.endswith(".head_local_features.fc2.weight"):
                t = t * 0.2
        sd[name] = t
    return sd


if __name__ == "__main__":
    dev = torch.device("cuda:0")
    prec = os.environ.get("ADB_PREC", "bf16x3")
    B = int(os.environ.get("ADB_B", "1"))
    reps = int(os.environ.get("ADB_REPS", "3"))
    t0 = time.time()
    m = AsymmetricMASt3R(precision=prec, **FULL_CFG).load_state_dict(random_state(FULL_CFG, dev)).to(dev)
    print(f"load {time.time() - t0:.1f}s")
    img1 = torch.rand(B, 3, 512, 512, device=dev) * 2 - 1
    img2 = torch.rand(B, 3, 512, 512, device=dev) * 2 - 1
    ev = lambda: torch.cuda.Event(enable_timing=True)
    for it in range(reps):
        if os.environ.get("ADB_CUPROF") and it == reps - 1:
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaProfilerStart()     # ncu --profile-from-start off captures only this forward
        e = [ev() for _ in range(5)]
        e[0].record()
        f, pos, _ = m._encode_image(torch.cat((img1, img2), 0), None)
        e[1].record()
        (f1, f2), (p1, p2) = f.chunk(2, 0), pos.chunk(2, 0)
        d1, d2 = m._decoder(f1.contiguous(), p1.contiguous(), f2.contiguous(), p2.contiguous())
        e[2].record()
        r1 = m._downstream_head(1, d1, (512, 512))
        r2 = m._downstream_head(2, d2, (512, 512))
        e[3].record()
        torch.cuda.synchronize()
        if os.environ.get("ADB_CUPROF") and it == reps - 1:
            torch.cuda.cudart().cudaProfilerStop()
        print(f"[{prec} B={B}] enc {e[0].elapsed_time(e[1]):.2f} ms  dec {e[1].elapsed_time(e[2]):.2f} ms  heads {e[2].elapsed_time(e[3]):.2f} ms  "
              f"total {e[0].elapsed_time(e[3]):.2f} ms -> {B / e[0].elapsed_time(e[3]) * 1e3:.2f} pairs/s")
    a, b = ev(), ev()
    forward_pair(m, img1, img2); torch.cuda.synchronize()
    a.record()
    for _ in range(5):
        forward_pair(m, img1, img2)
    b.record(); torch.cuda.synchronize()
    print(f"[{prec} B={B}] forward_pair (two-stream heads): {a.elapsed_time(b) / 5:.2f} ms -> {B / (a.elapsed_time(b) / 5) * 1e3:.2f} pairs/s")
    if os.environ.get("ADB_GRAPH", "1") == "1":
        from artdeco_b200.mast3r import GraphedForwardPair
        ref1, ref2 = forward_pair(m, img1, img2)
        ref_p = ref1["pts3d"].clone(); ref_q = ref2["desc"].clone()
        g = GraphedForwardPair(m, B, 512, 512)
        o1, o2 = g(img1, img2)
        torch.cuda.synchronize()
        print("graph == eager:", bool(torch.equal(o1["pts3d"], ref_p)), bool(torch.equal(o2["desc"], ref_q)))
        a.record()
        for _ in range(10):
            g(img1, img2)
        b.record(); torch.cuda.synchronize()
        print(f"[{prec} B={B}] CUDA-graph replay: {a.elapsed_time(b) / 10:.2f} ms -> {B / (a.elapsed_time(b) / 10) * 1e3:.2f} pairs/s")
    if os.environ.get("ADB_STAGE"):
        _lib.TIMER = _lib.StageTimer()
        _lib.LAUNCHES.update({k: 1 for k in ("adb_gemm_bf16", "adb_gemm_bf16_rope", "adb_layernorm", "adb_split_bf16", "adb_rope_heads",
                                             "adb_softmax_rows", "adb_im2col_patch16")})
        forward_pair(m, img1, img2)
        torch.cuda.synchronize()
        for k, (ms, n) in sorted(_lib.TIMER.totals_ms().items(), key=lambda kv: -kv[1][0]):
            print(f"  {k:22s} {ms:8.2f} ms over {n} calls")
