"""TEST INFRASTRUCTURE ONLY.  Plain-PyTorch (fp32, CPU or GPU) restatement of the reference's MASt3R inference path,
written functionally over a state dict that uses the REFERENCE'S parameter names, so the real checkpoint or any
reference-initialised model's ``state_dict()`` loads unchanged.

Restates (all under /root/reference/VSLAM/thirdparty/mast3r/):
  _encode_image        dust3r/dust3r/model.py:127-140, patch_embed.py:19-29, croco/models/blocks.py:94-130
  RoPE2D               dust3r/croco/models/pos_embed.py:112-159
  _decoder             dust3r/dust3r/model.py:172-191, blocks.py:140-191
  _downstream_head     mast3r/catmlp_dpt_head.py:71-96, dust3r/heads/dpt_head.py:34-65, croco/models/dpt_block.py:79-218,356-410
  postprocess          mast3r/catmlp_dpt_head.py:17-39, dust3r/heads/postprocess.py:22-58
PINNED BY RUNNING THE REFERENCE: tests/golden/make_mast3r_golden.py imports the real module in this container, loads the
same deterministic weights into it and checks this file agrees (max rel err < 1e-5) before writing the golden outputs
that travel to the GPU box (the reference has no tests or vectors of its own for this path)."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _ln(x, sd, pre, eps=1e-6):
    return F.layer_norm(x, (x.shape[-1],), sd[pre + ".weight"], sd[pre + ".bias"], eps)


def _lin(x, sd, pre):
    return F.linear(x, sd[pre + ".weight"], sd.get(pre + ".bias"))


def rope2d(tokens, positions, base=100.0):
    """tokens [B,h,N,D], positions [B,N,2] (y,x) int64."""
    D = tokens.shape[-1] // 2
    inv = 1.0 / (base ** (torch.arange(0, D, 2, device=tokens.device).float() / D))

    def rope1d(t, p):
        fr = p[..., None].to(inv.dtype) * inv
        fr = torch.cat([fr, fr], -1)[:, None].to(t.dtype)
        rot = torch.cat([-t[..., D // 2:], t[..., :D // 2]], -1)
        return t * fr.cos() + rot * fr.sin()

    y, x = tokens.chunk(2, dim=-1)
    return torch.cat([rope1d(y, positions[..., 0]), rope1d(x, positions[..., 1])], -1)


def _attention(x, pos, sd, pre, heads):
    B, N, C = x.shape
    qkv = _lin(x, sd, pre + ".qkv").reshape(B, N, 3, heads, C // heads).transpose(1, 3)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    q, k = rope2d(q, pos), rope2d(k, pos)
    a = ((q @ k.transpose(-2, -1)) * (C // heads) ** -0.5).softmax(-1)
    return _lin((a @ v).transpose(1, 2).reshape(B, N, C), sd, pre + ".proj")


def _cross_attention(xq, y, qpos, kpos, sd, pre, heads):
    B, Nq, C = xq.shape
    Nk = y.shape[1]
    hd = C // heads
    q = _lin(xq, sd, pre + ".projq").reshape(B, Nq, heads, hd).permute(0, 2, 1, 3)
    k = _lin(y, sd, pre + ".projk").reshape(B, Nk, heads, hd).permute(0, 2, 1, 3)
    v = _lin(y, sd, pre + ".projv").reshape(B, Nk, heads, hd).permute(0, 2, 1, 3)
    q, k = rope2d(q, qpos), rope2d(k, kpos)
    a = ((q @ k.transpose(-2, -1)) * hd ** -0.5).softmax(-1)
    return _lin((a @ v).transpose(1, 2).reshape(B, Nq, C), sd, pre + ".proj")


def _mlp(x, sd, pre):
    return _lin(F.gelu(_lin(x, sd, pre + ".fc1")), sd, pre + ".fc2")


def positions(B, H, W, device):
    y, x = torch.arange(H // 16, device=device), torch.arange(W // 16, device=device)
    return torch.cartesian_prod(y, x).view(1, -1, 2).expand(B, -1, 2).clone()


def encode_image(sd, cfg, img):
    B, _, H, W = img.shape
    x = F.conv2d(img, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=16).flatten(2).transpose(1, 2)
    pos = positions(B, H, W, img.device)
    for i in range(cfg["enc_depth"]):
        p = f"enc_blocks.{i}"
        x = x + _attention(_ln(x, sd, p + ".norm1"), pos, sd, p + ".attn", cfg["enc_num_heads"])
        x = x + _mlp(_ln(x, sd, p + ".norm2"), sd, p + ".mlp")
    return _ln(x, sd, "enc_norm"), pos


def _dec_block(x, y, xpos, ypos, sd, p, heads):
    x = x + _attention(_ln(x, sd, p + ".norm1"), xpos, sd, p + ".attn", heads)
    y_ = _ln(y, sd, p + ".norm_y")
    x = x + _cross_attention(_ln(x, sd, p + ".norm2"), y_, xpos, ypos, sd, p + ".cross_attn", heads)
    return x + _mlp(_ln(x, sd, p + ".norm3"), sd, p + ".mlp")


def decoder(sd, cfg, f1, pos1, f2, pos2):
    out = [(f1, f2)]
    f1, f2 = _lin(f1, sd, "decoder_embed"), _lin(f2, sd, "decoder_embed")
    cur = (f1, f2)
    for i in range(cfg["dec_depth"]):
        n1 = _dec_block(cur[0], cur[1], pos1, pos2, sd, f"dec_blocks.{i}", cfg["dec_num_heads"])
        n2 = _dec_block(cur[1], cur[0], pos2, pos1, sd, f"dec_blocks2.{i}", cfg["dec_num_heads"])
        cur = (n1, n2)
        out.append(cur)
    out[-1] = (_ln(out[-1][0], sd, "dec_norm"), _ln(out[-1][1], sd, "dec_norm"))
    return [o[0] for o in out], [o[1] for o in out]


def _conv(x, sd, pre, **kw):
    return F.conv2d(x, sd[pre + ".weight"], sd.get(pre + ".bias"), **kw)


def _rcu(x, sd, pre):
    out = _conv(F.relu(x), sd, pre + ".conv1", padding=1)
    out = _conv(F.relu(out), sd, pre + ".conv2", padding=1)
    return out + x


def _fusion(sd, pre, x0, x1=None):
    out = x0
    if x1 is not None:
        out = out + _rcu(x1, sd, pre + ".resConfUnit1")
    out = _rcu(out, sd, pre + ".resConfUnit2")
    out = F.interpolate(out, scale_factor=2, mode="bilinear", align_corners=True)
    return _conv(out, sd, pre + ".out_conv")


def dpt(sd, pre, decout, H, W, hooks):
    nh, nw = H // 16, W // 16
    layers = [decout[h].transpose(1, 2).reshape(decout[h].shape[0], -1, nh, nw) for h in hooks]
    ap = pre + ".act_postprocess"
    l0 = F.conv_transpose2d(_conv(layers[0], sd, ap + ".0.0"), sd[ap + ".0.1.weight"], sd[ap + ".0.1.bias"], stride=4)
    l1 = F.conv_transpose2d(_conv(layers[1], sd, ap + ".1.0"), sd[ap + ".1.1.weight"], sd[ap + ".1.1.bias"], stride=2)
    l2 = _conv(layers[2], sd, ap + ".2.0")
    l3 = _conv(_conv(layers[3], sd, ap + ".3.0"), sd, ap + ".3.1", stride=2, padding=1)
    ls = [F.conv2d(l, sd[f"{pre}.scratch.layer{i + 1}_rn.weight"], None, padding=1) for i, l in enumerate((l0, l1, l2, l3))]
    p4 = _fusion(sd, pre + ".scratch.refinenet4", ls[3])[:, :, :ls[2].shape[2], :ls[2].shape[3]]
    p3 = _fusion(sd, pre + ".scratch.refinenet3", p4, ls[2])
    p2 = _fusion(sd, pre + ".scratch.refinenet2", p3, ls[1])
    p1 = _fusion(sd, pre + ".scratch.refinenet1", p2, ls[0])
    out = _conv(p1, sd, pre + ".head.0", padding=1)
    out = F.interpolate(out, scale_factor=2, mode="bilinear", align_corners=True)
    out = F.relu(_conv(out, sd, pre + ".head.2", padding=1))
    return _conv(out, sd, pre + ".head.4")


def postprocess(out, desc_dim=24):
    """depth_mode=('exp',-inf,inf), conf_mode=('exp',1,inf), desc 'norm', two_confs, desc_conf_mode=('exp',0,inf)."""
    fmap = out.permute(0, 2, 3, 1)
    xyz = fmap[..., 0:3]
    d = xyz.norm(dim=-1, keepdim=True)
    res = dict(pts3d=xyz / d.clip(min=1e-8) * torch.expm1(d))
    res["conf"] = 1 + fmap[..., 3].exp()
    desc = fmap[..., 4:4 + desc_dim]
    res["desc"] = desc / desc.norm(dim=-1, keepdim=True)
    res["desc_conf"] = 0 + fmap[..., 4 + desc_dim].exp()
    return res


def downstream_head(sd, cfg, head_num, decout, H, W, raw=False):
    pre = f"downstream_head{head_num}"
    l2 = cfg["dec_depth"]
    pts = dpt(sd, pre + ".dpt", decout, H, W, [0, l2 * 2 // 4, l2 * 3 // 4, l2])
    cat = torch.cat([decout[0], decout[-1]], -1)
    B = cat.shape[0]
    lf = _mlp(cat, sd, pre + ".head_local_features")
    lf = F.pixel_shuffle(lf.transpose(-1, -2).reshape(B, -1, H // 16, W // 16), 16)
    out = torch.cat([pts, lf], 1)
    return out if raw else postprocess(out)


def forward_pair(sd, cfg, img1, img2):
    """2x _encode_image + _decoder + 2x _downstream_head — the unit of work BASELINE.json counts as one 'pair'."""
    H, W = img1.shape[-2:]
    f1, p1 = encode_image(sd, cfg, img1)
    f2, p2 = encode_image(sd, cfg, img2)
    d1, d2 = decoder(sd, cfg, f1, p1, f2, p2)
    return downstream_head(sd, cfg, 1, d1, H, W), downstream_head(sd, cfg, 2, d2, H, W)


FULL_CFG = dict(enc_embed_dim=1024, enc_depth=24, enc_num_heads=16, dec_embed_dim=768, dec_depth=12, dec_num_heads=12)
SMALL_CFG = dict(enc_embed_dim=128, enc_depth=2, enc_num_heads=2, dec_embed_dim=128, dec_depth=12, dec_num_heads=2)


def param_shapes(cfg):
    """Every tensor of the reference state dict that inference reads: name -> shape."""
    E, Dd = cfg["enc_embed_dim"], cfg["dec_embed_dim"]
    s = {"patch_embed.proj.weight": (E, 3, 16, 16), "patch_embed.proj.bias": (E,)}

    def ln(p, d):
        s[p + ".weight"] = (d,); s[p + ".bias"] = (d,)

    def lin(p, o, i):
        s[p + ".weight"] = (o, i); s[p + ".bias"] = (o,)
    for i in range(cfg["enc_depth"]):
        p = f"enc_blocks.{i}"
        ln(p + ".norm1", E); lin(p + ".attn.qkv", 3 * E, E); lin(p + ".attn.proj", E, E)
        ln(p + ".norm2", E); lin(p + ".mlp.fc1", 4 * E, E); lin(p + ".mlp.fc2", E, 4 * E)
    ln("enc_norm", E); lin("decoder_embed", Dd, E)
    for blk in ("dec_blocks", "dec_blocks2"):
        for i in range(cfg["dec_depth"]):
            p = f"{blk}.{i}"
            ln(p + ".norm1", Dd); lin(p + ".attn.qkv", 3 * Dd, Dd); lin(p + ".attn.proj", Dd, Dd)
            ln(p + ".norm2", Dd); ln(p + ".norm3", Dd); ln(p + ".norm_y", Dd)
            for n in ("projq", "projk", "projv", "proj"):
                lin(p + ".cross_attn." + n, Dd, Dd)
            lin(p + ".mlp.fc1", 4 * Dd, Dd); lin(p + ".mlp.fc2", Dd, 4 * Dd)
    ln("dec_norm", Dd)
    ld = [96, 192, 384, 768]
    dims = [E, Dd, Dd, Dd]
    for hn in (1, 2):
        p = f"downstream_head{hn}"
        lin(p + ".head_local_features.fc1", 4 * (E + Dd), E + Dd)
        lin(p + ".head_local_features.fc2", 25 * 256, 4 * (E + Dd))
        d = p + ".dpt"
        for k in range(4):
            s[f"{d}.act_postprocess.{k}.0.weight"] = (ld[k], dims[k], 1, 1); s[f"{d}.act_postprocess.{k}.0.bias"] = (ld[k],)
            s[f"{d}.scratch.layer{k + 1}_rn.weight"] = (256, ld[k], 3, 3)
        s[f"{d}.act_postprocess.0.1.weight"] = (96, 96, 4, 4); s[f"{d}.act_postprocess.0.1.bias"] = (96,)
        s[f"{d}.act_postprocess.1.1.weight"] = (192, 192, 2, 2); s[f"{d}.act_postprocess.1.1.bias"] = (192,)
        s[f"{d}.act_postprocess.3.1.weight"] = (768, 768, 3, 3); s[f"{d}.act_postprocess.3.1.bias"] = (768,)
        for r in (1, 2, 3, 4):
            rp = f"{d}.scratch.refinenet{r}"
            s[rp + ".out_conv.weight"] = (256, 256, 1, 1); s[rp + ".out_conv.bias"] = (256,)
            for u in (1, 2):
                for c in (1, 2):
                    s[f"{rp}.resConfUnit{u}.conv{c}.weight"] = (256, 256, 3, 3); s[f"{rp}.resConfUnit{u}.conv{c}.bias"] = (256,)
        s[d + ".head.0.weight"] = (128, 256, 3, 3); s[d + ".head.0.bias"] = (128,)
        s[d + ".head.2.weight"] = (128, 128, 3, 3); s[d + ".head.2.bias"] = (128,)
        s[d + ".head.4.weight"] = (4, 128, 1, 1); s[d + ".head.4.bias"] = (4,)
    return s
