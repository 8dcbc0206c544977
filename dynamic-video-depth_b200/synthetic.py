"""Deterministic synthetic frame-pair batches of the reference's pair-file format.

Format follows what `datasets/davis_sequence.py:86-113` hands to the Model (after the DataLoader's
leading batch dim of 1) and the packing of `scripts/preprocess/davis/generate_sequence_midas.py:61-76,
146-147,179,187`: pose / intrinsic tensors hold TRANSPOSES so that `row @ M` works
(R_i = R_c2w_i^T, R_i_T = R_c2w_i, K = K^T, K_inv = (K^-1)^T). Distributions: SURVEY.md §8(d).
"""
import math

import torch

GAPS = (1, 2, 4, 6, 8)


def pair_list(n_frames=80, gaps=GAPS):
    """All (f, f+g) with f < n_frames-1-g — the reference's rule
    (scripts/preprocess/davis/generate_sequence_midas.py:187): 374 pairs for 80 frames."""
    out = []
    for g in gaps:
        for f in range(0, n_frames - 1 - g):
            out.append((f, f + g))
    return out


def _rot_y(a):
    c, s = math.cos(a), math.sin(a)
    return torch.tensor([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]], dtype=torch.float64)


def camera(frame_id, H, W):
    """c2w rotation, translation and intrinsics of one synthetic frame (float64)."""
    R = _rot_y(0.01 + 0.005 * frame_id)
    t = torch.tensor([0.05 * frame_id, 0.0, 0.0], dtype=torch.float64)
    f = 300.0 * W / 384.0
    K = torch.tensor([[f, 0.0, (W - 1) / 2.0], [0.0, f, (H - 1) / 2.0], [0.0, 0.0, 1.0]],
                     dtype=torch.float64)
    return R, t, K


def make_batch(pairs, H=224, W=384, n_frames=80, seed=0, smooth_flow=False, dtype=torch.float32,
               leading_dim=True, mask_p=0.9, flow_sigma=3.0):
    """Build one batch dict for `pairs` = [(frame_id_1, frame_id_2), ...] (B pairs).

    Returned tensors live on the CPU; `leading_dim` adds the DataLoader's batch dim of 1 that
    `Model._train_on_batch` squeezes (models/scene_flow_motion_field.py:177-179)."""
    B = len(pairs)
    g = torch.Generator().manual_seed(seed)

    def rnd(*shape):
        return torch.rand(*shape, generator=g, dtype=torch.float64)

    def nrm(*shape):
        return torch.randn(*shape, generator=g, dtype=torch.float64)

    def flow():
        if smooth_flow:
            lo = nrm(B, 2, (H + 7) // 8 + 1, (W + 7) // 8 + 1) * flow_sigma
            f = torch.nn.functional.interpolate(lo, size=(H, W), mode='bilinear', align_corners=True)
            return f.permute(0, 2, 3, 1).contiguous()
        return nrm(B, H, W, 2) * flow_sigma

    b = {}
    b['img_1'] = rnd(B, 3, H, W)
    b['img_2'] = rnd(B, 3, H, W)
    b['flow_1_2'] = flow()
    b['flow_2_1'] = flow()
    b['mask_1'] = (rnd(B, H, W, 1, 1) < mask_p).double()
    b['mask_2'] = (rnd(B, H, W, 1, 1) < mask_p).double()
    b['motion_seg_1'] = torch.ones(B, H, W, 1, 1, dtype=torch.float64)
    R1, R2, t1, t2, Ks, Kinvs = [], [], [], [], [], []
    for (f1, f2) in pairs:
        Ra, ta, K = camera(f1, H, W)
        Rb, tb, _ = camera(f2, H, W)
        R1.append(Ra), R2.append(Rb), t1.append(ta), t2.append(tb)
        Ks.append(K), Kinvs.append(torch.linalg.inv(K))
    R1, R2 = torch.stack(R1), torch.stack(R2)
    b['R_1'] = R1.transpose(1, 2).reshape(B, 1, 1, 3, 3).contiguous()
    b['R_1_T'] = R1.reshape(B, 1, 1, 3, 3).contiguous()
    b['R_2'] = R2.transpose(1, 2).reshape(B, 1, 1, 3, 3).contiguous()
    b['R_2_T'] = R2.reshape(B, 1, 1, 3, 3).contiguous()
    b['t_1'] = torch.stack(t1).reshape(B, 1, 1, 1, 3)
    b['t_2'] = torch.stack(t2).reshape(B, 1, 1, 1, 3)
    b['K'] = torch.stack(Ks).transpose(1, 2).reshape(B, 1, 1, 3, 3).contiguous()
    b['K_inv'] = torch.stack(Kinvs).transpose(1, 2).reshape(B, 1, 1, 3, 3).contiguous()
    fid1 = torch.tensor([p[0] for p in pairs], dtype=torch.float64)
    fid2 = torch.tensor([p[1] for p in pairs], dtype=torch.float64)
    b['time_stamp_1'] = (fid1 / n_frames).reshape(B, 1, 1, 1).expand(B, 1, H, W).contiguous()
    b['time_stamp_2'] = (fid2 / n_frames).reshape(B, 1, 1, 1).expand(B, 1, H, W).contiguous()
    b['frame_id_1'] = fid1.clone()
    b['frame_id_2'] = fid2.clone()
    b['depth_pred_1'] = torch.ones(B, 1, H, W, dtype=torch.float64)
    out = {}
    for k, v in b.items():
        v = v.to(dtype)
        out[k] = v.unsqueeze(0) if leading_dim else v
    out['time_step'] = torch.tensor([1.0 / n_frames], dtype=dtype)
    out['pair_path'] = ['synthetic_%03d_%03d' % p for p in pairs]
    return out


def make_depths(B, H, W, seed=1, lo=2.0, hi=8.0, dtype=torch.float32):
    """Smooth positive depth maps for op-level tests (the depth nets are tested separately)."""
    g = torch.Generator().manual_seed(seed)
    lo_res = torch.rand(B, 1, (H + 15) // 16 + 1, (W + 15) // 16 + 1, generator=g, dtype=torch.float64)
    d = torch.nn.functional.interpolate(lo_res, size=(H, W), mode='bilinear', align_corners=True)
    return (lo + (hi - lo) * d).to(dtype).contiguous()
