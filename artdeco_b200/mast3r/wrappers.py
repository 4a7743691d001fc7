"""Host-side mirror of the reference's MASt3R call wrappers (VSLAM/utils_mast3r.py:30-36, 42-71, 116-141, 176-191;
SURVEY.md §8a M10, §8f rank 2): same names, argument meaning and returned stacking, over ``artdeco_b200.mast3r``.

``mast3r_decode_symmetric_batch`` is where this differs structurally from the reference: the reference loops over the
batch in Python and runs the decoder twice and the heads four times PER PAIR (utils_mast3r.py:45-59); here the whole batch
goes through ONE decoder call on the concatenation [i|j] x [j|i] and ONE call per head, which is the same computation
(pairs are independent; both directions share the weights) on a 2B batch.

``mast3r_match_symmetric / mast3r_match_asymmetric`` chain into ``artdeco_b200.matching`` (the dense matching kernels).
"""
from __future__ import annotations

import torch

from .. import matching
from .model import AsymmetricMASt3R


def load_mast3r(path=None, device="cuda:0"):
    """utils_mast3r.py:10-17."""
    weights_path = "models/MASt3R_ViTLarge_BaseDecoder_512_catmlpdpt_metric.pth" if path is None else path
    return AsymmetricMASt3R.from_pretrained(weights_path).to(device)


def _hw(shape):
    if isinstance(shape, torch.Tensor):
        hw = shape.reshape(-1, 2)
        return int(hw[0, 0]), int(hw[0, 1])
    return int(shape[0]), int(shape[1])


def _heads(model, dec1, dec2, shape1, shape2):
    """Both heads, on two streams when the model allows it (they are independent)."""
    if not model.concurrent:
        return model._downstream_head(1, dec1, shape1), model._downstream_head(2, dec2, shape2)
    s1 = torch.cuda.current_stream()
    s2 = model._side_stream(dec1[0].device)
    s2.wait_stream(s1)
    with torch.cuda.stream(s2):
        res2 = model._downstream_head(2, dec2, shape2)
    res1 = model._downstream_head(1, dec1, shape1)
    s1.wait_stream(s2)
    for v in res2.values():
        v.record_stream(s1)
    return res1, res2


@torch.inference_mode()
def decoder(model, feat1, feat2, pos1, pos2, shape1, shape2):
    """utils_mast3r.py:30-36."""
    dec1, dec2 = model._decoder(feat1.contiguous(), pos1.contiguous(), feat2.contiguous(), pos2.contiguous())
    return _heads(model, dec1, dec2, _hw(shape1), _hw(shape2))


@torch.inference_mode()
def mast3r_decode_symmetric_batch(model, feat_i, pos_i, feat_j, pos_j, shape_i, shape_j):
    """utils_mast3r.py:42-71.  Returns X[4,b,h,w,3], C[4,b,h,w], D[4,b,h,w,24], Q[4,b,h,w] ordered (ii, ji, jj, ij).
    NOTE (as the reference): assumes all images share one shape."""
    B = feat_i.shape[0]
    si = shape_i[0] if not isinstance(shape_i, torch.Tensor) or shape_i.dim() > 1 else shape_i
    sj = shape_j[0] if not isinstance(shape_j, torch.Tensor) or shape_j.dim() > 1 else shape_j
    if _hw(si) != _hw(sj):
        raise ValueError("mast3r_decode_symmetric_batch: image shapes must match")
    f1, p1 = torch.cat((feat_i, feat_j), 0), torch.cat((pos_i, pos_j), 0)
    f2, p2 = torch.cat((feat_j, feat_i), 0), torch.cat((pos_j, pos_i), 0)
    r1, r2 = decoder(model, f1, f2, p1, p2, si, sj)      # r1[:B]=ii r1[B:]=jj ; r2[:B]=ji r2[B:]=ij
    out = []
    for k in ("pts3d", "conf", "desc", "desc_conf"):
        out.append(torch.stack((r1[k][:B], r2[k][:B], r1[k][B:], r2[k][B:]), 0))
    return tuple(out)


@torch.inference_mode()
def mast3r_asymmetric_inference(model, frame_i, frame_j, embeddings_i=None, embeddings_j=None):
    """utils_mast3r.py:116-141.  ``frame.img`` is [3,H,W] in [-1,1]."""
    img_i, img_j = frame_i.img[None], frame_j.img[None]
    shape_i, shape_j = tuple(img_i.shape[2:]), tuple(img_j.shape[2:])
    if embeddings_i is not None:
        feat1, pos1 = embeddings_i[0], embeddings_i[1]
    else:
        feat1, pos1, _ = model._encode_image(img_i, None)
    if embeddings_j is not None:
        feat2, pos2 = embeddings_j[0], embeddings_j[1]
    else:
        feat2, pos2, _ = model._encode_image(img_j, None)
    res11, res21 = decoder(model, feat1, feat2, pos1, pos2, shape_i, shape_j)
    X, C, D, Q = (torch.stack((res11[k][0], res21[k][0])) for k in ("pts3d", "conf", "desc", "desc_conf"))
    return X, C, D, Q, feat1, pos1


@torch.inference_mode()
def mast3r_inference_mono(model, frame):
    """utils_mast3r.py:176-191: the frame against itself; returns Xii[hw,3], Cii[hw,1], feat, pos."""
    img = frame.img[None]
    shape = tuple(img.shape[2:])
    feat, pos, _ = model._encode_image(img, None)
    res11, _ = decoder(model, feat, feat, pos, pos, shape, shape)
    Xii = res11["pts3d"][0].reshape(-1, 3)
    Cii = res11["conf"][0].reshape(-1, 1)
    return Xii, Cii, feat, pos


def mast3r_match_symmetric(config, model, feat_i, pos_i, feat_j, pos_j, shape_i, shape_j):
    """utils_mast3r.py:74-112: decode both directions, match j->i and i->j in one batched call."""
    X, C, D, Q = mast3r_decode_symmetric_batch(model, feat_i, pos_i, feat_j, pos_j, shape_i, shape_j)
    b = X.shape[1]
    Xii, Xji, Xjj, Xij = X[0], X[1], X[2], X[3]
    Dii, Dji, Djj, Dij = D[0], D[1], D[2], D[3]
    Qii, Qji, Qjj, Qij = Q[0], Q[1], Q[2], Q[3]
    X11, X21 = torch.cat((Xii, Xjj), 0), torch.cat((Xji, Xij), 0)
    D11, D21 = torch.cat((Dii, Djj), 0), torch.cat((Dji, Dij), 0)
    idx_1_to_2, valid_match_2 = matching.match(config, X11, X21, D11, D21)
    match_b = X11.shape[0] // 2
    return (idx_1_to_2[:match_b], idx_1_to_2[match_b:], valid_match_2[:match_b], valid_match_2[match_b:],
            Qii.reshape(b, -1, 1), Qjj.reshape(b, -1, 1), Qji.reshape(b, -1, 1), Qij.reshape(b, -1, 1))


def mast3r_match_asymmetric(config, model, frame_i, frame_j, idx_i2j_init=None, embeddings_i=None, embeddings_j=None):
    """utils_mast3r.py:144-171."""
    X, C, D, Q, feat1, pos1 = mast3r_asymmetric_inference(model, frame_i, frame_j, embeddings_i=embeddings_i,
                                                          embeddings_j=embeddings_j)
    b = X.shape[0] // 2
    Xii, Xji = X[:b], X[b:]
    Dii, Dji = D[:b], D[b:]
    idx_i2j, valid_match_j = matching.match(config, Xii.contiguous(), Xji.contiguous(), Dii.contiguous(),
                                            Dji.contiguous(), idx_1_to_2_init=idx_i2j_init)
    hw = X.shape[1] * X.shape[2]
    Xii, Xji = X.reshape(2 * b, hw, 3)
    Cii, Cji = C.reshape(2 * b, hw, 1)
    Qii, Qji = Q.reshape(2 * b, hw, 1)
    return idx_i2j, valid_match_j, Xii, Cii, Qii, Xji, Cji, Qji, feat1, pos1
