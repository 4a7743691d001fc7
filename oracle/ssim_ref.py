"""TEST INFRASTRUCTURE ONLY.  SSIM oracle = the conv2d formulation the reference's own test asserts against
(/root/reference Reconstruct/submodules/fused-ssim/tests/test.py:14-54: ``gaussian``, ``create_window``,
``_ssim``; window 11, sigma 1.5, zero 'same' padding, C1=0.01^2, C2=0.03^2).  PINNED: the reference test
(tests/test.py:82-91) requires fused_ssim to be ``isclose`` (rtol 1e-5) to exactly this computation, and the
gradient w.r.t. img1 to match its autograd.  Restated here (not imported) because the reference test file needs
a GPU and ``pytorch_msssim`` at import time."""
from math import exp

import torch
import torch.nn.functional as F


def gaussian(window_size: int, sigma: float) -> torch.Tensor:
    g = torch.tensor([exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)])
    return g / g.sum()


def create_window(window_size: int, channel: int, dtype=torch.float32) -> torch.Tensor:
    w1 = gaussian(window_size, 1.5).unsqueeze(1)
    w2 = w1.mm(w1.t()).float().unsqueeze(0).unsqueeze(0)
    return w2.expand(channel, 1, window_size, window_size).contiguous().to(dtype)


def ssim_map(img1: torch.Tensor, img2: torch.Tensor, padding: str = "same") -> torch.Tensor:
    ch = img1.size(-3)
    win = create_window(11, ch, img1.dtype).to(img1.device)
    pad = 5 if padding == "same" else 0
    mu1 = F.conv2d(img1, win, padding=pad, groups=ch)
    mu2 = F.conv2d(img2, win, padding=pad, groups=ch)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    s1 = F.conv2d(img1 * img1, win, padding=pad, groups=ch) - mu1_sq
    s2 = F.conv2d(img2 * img2, win, padding=pad, groups=ch) - mu2_sq
    s12 = F.conv2d(img1 * img2, win, padding=pad, groups=ch) - mu1_mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    return ((2 * mu1_mu2 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))


def ssim(img1, img2, padding="same"):
    return ssim_map(img1, img2, padding).mean()
