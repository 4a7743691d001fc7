"""Names and shapes of every reference-state-dict tensor that MASt3R inference reads (used to build random-init
weights when no checkpoint is available; the reference checkpoint's own keys are a superset)."""
from __future__ import annotations


def param_shapes(cfg):
    E, Dd = cfg["enc_embed_dim"], cfg["dec_embed_dim"]
    s = {"patch_embed.proj.weight": (E, 3, 16, 16), "patch_embed.proj.bias": (E,)}

    def ln(p, d):
        s[p + ".weight"] = (d,)
        s[p + ".bias"] = (d,)

    def lin(p, o, i):
        s[p + ".weight"] = (o, i)
        s[p + ".bias"] = (o,)

    for i in range(cfg["enc_depth"]):
        p = f"enc_blocks.{i}"
        ln(p + ".norm1", E); lin(p + ".attn.qkv", 3 * E, E); lin(p + ".attn.proj", E, E)
        ln(p + ".norm2", E); lin(p + ".mlp.fc1", 4 * E, E); lin(p + ".mlp.fc2", E, 4 * E)
    ln("enc_norm", E)
    lin("decoder_embed", Dd, E)
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
            s[f"{d}.act_postprocess.{k}.0.weight"] = (ld[k], dims[k], 1, 1)
            s[f"{d}.act_postprocess.{k}.0.bias"] = (ld[k],)
            s[f"{d}.scratch.layer{k + 1}_rn.weight"] = (256, ld[k], 3, 3)
        s[f"{d}.act_postprocess.0.1.weight"] = (96, 96, 4, 4); s[f"{d}.act_postprocess.0.1.bias"] = (96,)
        s[f"{d}.act_postprocess.1.1.weight"] = (192, 192, 2, 2); s[f"{d}.act_postprocess.1.1.bias"] = (192,)
        s[f"{d}.act_postprocess.3.1.weight"] = (768, 768, 3, 3); s[f"{d}.act_postprocess.3.1.bias"] = (768,)
        for r in (1, 2, 3, 4):
            rp = f"{d}.scratch.refinenet{r}"
            s[rp + ".out_conv.weight"] = (256, 256, 1, 1); s[rp + ".out_conv.bias"] = (256,)
            for u in (1, 2):
                for c in (1, 2):
                    s[f"{rp}.resConfUnit{u}.conv{c}.weight"] = (256, 256, 3, 3)
                    s[f"{rp}.resConfUnit{u}.conv{c}.bias"] = (256,)
        s[d + ".head.0.weight"] = (128, 256, 3, 3); s[d + ".head.0.bias"] = (128,)
        s[d + ".head.2.weight"] = (128, 128, 3, 3); s[d + ".head.2.bias"] = (128,)
        s[d + ".head.4.weight"] = (4, 128, 1, 1); s[d + ".head.4.bias"] = (4,)
    return s


def random_state_dict(cfg, device, seed: int = 0):
    """Random-init weights generated on `device` (fast; used by the benchmark: there is no checkpoint offline).
    Same scaling rules as synthetic.det_weights (unit-variance activations, log-depth <~ 3.5)."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    sd = {}
    for name, shape in param_shapes(cfg).items():
        t = torch.randn(*shape, generator=g, device=device)
        if len(shape) == 1:
            t = 1.0 + 0.1 * t if name.endswith(".weight") else 0.02 * t
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            if ".act_postprocess." in name and name.endswith(".1.weight") and shape[0] == shape[1] and shape[2] in (2, 4):
                fan_in = shape[0]
            t = t * (1.0 / fan_in) ** 0.5
            if name.endswith(".dpt.head.4.weight") or name.endswith(".head_local_features.fc2.weight"):
                t = t * 0.2
        sd[name] = t
    return sd
