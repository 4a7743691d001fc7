"""TEST INFRASTRUCTURE ONLY.  ``update_voxel`` oracle = the reference function restated line for line
(/root/reference Reconstruct/scene/scene_models/h3dgsv3.py:227-316) in plain PyTorch on the CPU, with two substitutions:
  * ``torch_scatter.scatter_max`` (not installed here) -> an explicit first-maximum scan over the ascending pair list, which
    is what torch_scatter's CPU kernel does (strict '>' update): on equal counts the SMALLEST class id wins;
  * the voxel division follows the reference's CUDA arithmetic: PyTorch evaluates ``tensor / python_float`` on the GPU as
    ``tensor * (1 / python_float)`` in fp32 (BinaryDivTrueKernel.cu, CPU-scalar fast path), so the oracle multiplies by the
    fp32 reciprocal.  PARITY: pinned by construction (it is the reference's own op sequence), not by a reference test."""
import numpy as np
import torch


def _vidx(p, mn, voxel_size):
    inv = torch.tensor(np.float32(1.0) / np.float32(voxel_size))
    return torch.floor((p - mn) * inv).long()


def update_voxel(new_xyz, xyz, cls_id, voxel_size=0.1):
    num_new, num_orig = new_xyz.shape[0], xyz.shape[0]
    if num_orig == 0:
        v_min = new_xyz.min(dim=0).values
        v_idx = _vidx(new_xyz, v_min, voxel_size)
        v_max = v_idx.max(dim=0).values + 1
        stride = torch.tensor([v_max[1] * v_max[2], v_max[2], 1])
        h_new = (v_idx * stride).sum(dim=1)
        u_hashes, u_inv = torch.unique(h_new, return_inverse=True)
        return u_inv.unsqueeze(-1), u_hashes.shape[0]
    cls_id_1d = cls_id.squeeze(-1)
    max_cls = cls_id_1d.max().item()
    all_p = torch.cat([xyz, new_xyz], dim=0)
    min_c = all_p.min(dim=0).values
    v_idx_all = _vidx(all_p, min_c, voxel_size)
    v_max = v_idx_all.max(dim=0).values + 1
    stride = torch.tensor([v_max[1] * v_max[2], v_max[2], 1])
    h_all = (v_idx_all * stride).sum(dim=1)
    h_orig, h_new = h_all[:num_orig], h_all[num_orig:]
    unique_voxels, inv_idx = torch.unique(h_orig, return_inverse=True)
    offset = max_cls + 1
    pair_id = inv_idx * offset + cls_id_1d
    pair_unique_ids, pair_counts = torch.unique(pair_id, return_counts=True)
    v_indices_in_pair = pair_unique_ids // offset
    c_labels_in_pair = pair_unique_ids % offset
    # scatter_max(pair_counts, v_indices_in_pair): arg of the FIRST maximum per voxel
    nv = unique_voxels.shape[0]
    best_c = torch.full((nv,), -1, dtype=torch.long)
    max_indices = torch.zeros(nv, dtype=torch.long)
    pc, vi = pair_counts.numpy(), v_indices_in_pair.numpy()
    bc, mi = best_c.numpy(), max_indices.numpy()
    for k in range(len(pc)):
        if pc[k] > bc[vi[k]]:
            bc[vi[k]] = pc[k]
            mi[vi[k]] = k
    voxel_mode_labels = c_labels_in_pair[torch.from_numpy(mi)]
    updated_orig_cls_id = voxel_mode_labels[inv_idx].unsqueeze(-1)
    pos = torch.searchsorted(unique_voxels, h_new)
    pos_clamped = pos.clamp(max=unique_voxels.shape[0] - 1)
    mask = unique_voxels[pos_clamped] == h_new
    updated_new_cls_id = torch.zeros(num_new, dtype=torch.long)
    if mask.any():
        updated_new_cls_id[mask] = voxel_mode_labels[pos_clamped[mask]]
    new_voxel_count = 0
    if (~mask).any():
        unmatched_h = h_new[~mask]
        u_new_h, u_new_inv = torch.unique(unmatched_h, return_inverse=True)
        new_voxel_count = u_new_h.shape[0]
        updated_new_cls_id[~mask] = u_new_inv + max_cls + 1
    return updated_orig_cls_id, updated_new_cls_id.unsqueeze(-1), new_voxel_count
