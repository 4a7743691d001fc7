"""Dense matching at the BASELINE image size: our four kernels against the reference flow on the same box
(reference kernels from oracle/_ref/mast3r_matching_ref.so + the PyTorch glue of VSLAM/utils_matching.py as restated in
oracle/matching_ref.py).  Prints one JSON line; run under gpurun (never under a profiler)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from artdeco_b200 import matching as M  # noqa: E402
from oracle import build_ref, matching_ref as mr  # noqa: E402
from test_matching import CFG, _scene  # noqa: E402


def timed(fn, reps=20, warm=3):
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


def main():
    dev = torch.device("cuda:0")
    B = int(os.environ.get("ADB_B", "8"))       # mast3r_match_symmetric on 4 pairs matches 2*4 image pairs at once
    X11, X21, D11, D21 = [t.to(dev).contiguous() for t in _scene(b=B, h=512, w=512, seed=5, rot_deg=0.8)]
    c = CFG["matching"]
    out = {"workload": f"utils_matching.match, b={B}, 512x512, 24-dim descriptors, radius 4, dilation_max 5"}
    out["ours_ms"] = timed(lambda: M.match(CFG, X11, X21, D11, D21))
    # per kernel
    rays, pts, p_init = M.prep_for_iter_proj(X11, X21, None)
    out["prep_ms"] = timed(lambda: M.prep_for_iter_proj(X11, X21, None))
    out["iter_proj_ms"] = timed(lambda: M.iter_proj(rays, pts, p_init, c["max_iter"], c["lambda_init"], c["convergence_thresh"]))
    p, conv = M.iter_proj(rays, pts, p_init, c["max_iter"], c["lambda_init"], c["convergence_thresh"])
    p1 = p.long()
    a, q = D11.half().contiguous(), D21.reshape(B, -1, 24).half().contiguous()
    out["refine_ms"] = timed(lambda: M.refine_matches(a, q, p1, c["radius"], c["dilation_max"]))
    try:
        ref = build_ref.load("mast3r_matching_ref")
    except Exception as e:  # noqa: BLE001
        out["reference"] = f"unavailable: {e}"
        print(json.dumps(out))
        return

    def ref_flow():
        r, pt, pi = mr.prep_for_iter_proj(X11, X21, None)
        pp, cv = ref.iter_proj(r, pt, pi, c["max_iter"], c["lambda_init"], c["convergence_thresh"])
        pl = pp.long()
        bi = torch.arange(B, device=dev)[:, None].repeat(1, 512 * 512)
        d = torch.linalg.norm(X11[bi, pl[..., 1], pl[..., 0], :].reshape(B, 512, 512, 3) - X21, dim=-1)
        v = cv & (d < c["dist_thresh"]).view(B, -1)
        (pl,) = ref.refine_matches(D11.half(), D21.view(B, 512 * 512, -1).half(), pl, c["radius"], c["dilation_max"])
        return pl[..., 0] + 512 * pl[..., 1], v

    out["reference_flow_ms"] = timed(ref_flow)
    out["reference_iter_proj_ms"] = timed(lambda: ref.iter_proj(rays, pts, p_init, c["max_iter"], c["lambda_init"], c["convergence_thresh"]))
    out["reference_refine_ms"] = timed(lambda: ref.refine_matches(a, q, p1, c["radius"], c["dilation_max"]))
    out["speedup_flow"] = out["reference_flow_ms"] / out["ours_ms"]
    idx, valid = M.match(CFG, X11, X21, D11, D21)
    idx_r, valid_r = ref_flow()
    out["idx_agreement"] = float((idx == idx_r).float().mean())
    out["valid_agreement"] = float((valid[..., 0] == valid_r).float().mean())
    out["matches_per_s"] = B * 512 * 512 / (out["ours_ms"] * 1e-3)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
