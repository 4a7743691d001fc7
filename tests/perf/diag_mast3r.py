"""Stage-by-stage error of the CUDA MASt3R model against the plain-torch oracle run on the GPU in fp32 (TF32 off)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from artdeco_b200.mast3r import FULL_CFG, AsymmetricMASt3R  # noqa: E402
from oracle import mast3r_torch as mt  # noqa: E402
from tools.prof_mast3r import random_state  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")
cfg = FULL_CFG
H = W = int(os.environ.get("ADB_HW", "512"))
sd = random_state(cfg, dev)
m = AsymmetricMASt3R(**cfg).load_state_dict(sd).to(dev)
img1 = torch.rand(1, 3, H, W, device=dev) * 2 - 1
img2 = torch.rand(1, 3, H, W, device=dev) * 2 - 1


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-20))


with torch.inference_mode():
    f1, p1 = mt.encode_image(sd, cfg, img1)
    f2, p2 = mt.encode_image(sd, cfg, img2)
    d1, d2 = mt.decoder(sd, cfg, f1, p1, f2, p2)
    raw_ref = mt.downstream_head(sd, cfg, 1, d1, H, W, raw=True)
    ref = mt.postprocess(raw_ref)
g1, q1, _ = m._encode_image(img1, None)
g2, q2, _ = m._encode_image(img2, None)
e1, e2 = m._decoder(g1, q1, g2, q2)
print("enc", rel(g1, f1), "dec last", rel(e1[-1], d1[-1]))
# head on the ORACLE's decoder outputs: isolates the head's own error
raw_iso = m._downstream_head(1, [t for t in d1], (H, W), raw=True)
print("head (oracle inputs) raw pts4", rel(raw_iso[:, :4], raw_ref[:, :4]), "raw lf", rel(raw_iso[:, 4:], raw_ref[:, 4:]))
raw_all = m._downstream_head(1, e1, (H, W), raw=True)
print("head (own inputs)    raw pts4", rel(raw_all[:, :4], raw_ref[:, :4]), "raw lf", rel(raw_all[:, 4:], raw_ref[:, 4:]))
out = m._downstream_head(1, e1, (H, W))
for k in ("pts3d", "conf", "desc", "desc_conf"):
    print(k, rel(out[k], ref[k]), "absmax", float(ref[k].abs().max()))
d = raw_ref[:, :3].norm(dim=1)
print("log-depth d: max", float(d.max()), "mean", float(d.mean()))
