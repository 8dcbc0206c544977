"""Host-side mirror of losses/scene_flow_projection.py: same class names, same `forward` signatures, same
result-dict keys / shapes / dtypes, dispatching to the fused CUDA kernels through the C ABI.

The training fast path does NOT go through these modules (Model uses ops.reproject_loss, which never
materialises per-pixel tensors); they exist for the operator-level drop-in and for visualised batches.
Outputs are returned in the reference's layouts ([B,H,W,1,3] / [B,H,W,2] / [B,1,H,W]) as views of the
channel-planar kernel outputs and are differentiable w.r.t. depth_1, depth_2 and the scene flow for any
downstream loss (ops.ReprojectMaterialize -> dvd_reproject_materialize / dvd_reproject_materialize_bwd).
"""
from torch import nn

from .. import ops


def _bhw13(x):      # [B,3,H,W] -> [B,H,W,1,3]
    return x.permute(0, 2, 3, 1).unsqueeze(3)


def _bhw2(x):       # [B,2,H,W] -> [B,H,W,2]
    return x.permute(0, 2, 3, 1)


class unproject_ptcld(nn.Module):
    """losses/scene_flow_projection.py:48-67."""

    def __init__(self, is_one_way=True):
        super().__init__()

    def forward(self, depth_1, R_1, t_1, K_inv):
        # the C ABI wants column-vector matrices: R_1 holds R_c2w^T, K_inv holds (K^-1)^T
        B = depth_1.shape[0]
        poses = ops.pack_poses(K_inv.reshape(B, 3, 3).transpose(1, 2), K_inv, R_1.reshape(B, 3, 3).transpose(1, 2),
                               R_1.reshape(B, 3, 3).transpose(1, 2), t_1, t_1)
        return _bhw13(ops.unproject(depth_1, poses, 1))


class flow_by_depth(nn.Module):
    """losses/scene_flow_projection.py:95-153."""

    def __init__(self, is_one_way=True):
        super().__init__()
        self.one_way = is_one_way

    def forward(self, depth_1, depth_2, flow_1_2, R_1, R_2, R_1_T, R_2_T, t_1, t_2, K, K_inv):
        poses = ops.pack_poses(K, K_inv, R_1_T, R_2_T, t_1, t_2)
        o = ops.reproject_tensors(depth_1, depth_2, None, flow_1_2.contiguous(), poses)
        # with no scene flow the projected flow is the static one
        return {'dflow_1_2': _bhw2(o['staticflow_1_2']), 'sf_by_depth': _bhw13(o['sf_by_depth']),
                'warped_global_p2': _bhw13(o['warped_global_p2']), 'global_p1': _bhw13(o['global_p1'])}


class scene_flow_projection_slack(nn.Module):
    """losses/scene_flow_projection.py:204-278."""

    def __init__(self, is_one_way=False):
        super().__init__()
        self.is_one_way = is_one_way

    def forward(self, depth_1, depth_2, flow_1_2, flow_2_1, R_1, R_2, R_1_T, R_2_T, t_1, t_2, K, K_inv,
                sflow_1_2, sflow_2_1):
        poses = ops.pack_poses(K, K_inv, R_1_T, R_2_T, t_1, t_2)
        sf = sflow_1_2.squeeze(3).permute(0, 3, 1, 2).contiguous()   # [B,H,W,1,3] -> [B,3,H,W]
        o = ops.reproject_tensors(depth_1, depth_2, sf, flow_1_2.contiguous(), poses)
        return {'dflow_1_2': _bhw2(o['dflow_1_2']), 'depth_image_1_2': o['depth_image_1_2'],
                'depth_warp_1_2': o['depth_warp_1_2'], 'depth_1': depth_1, 'depth_2': depth_2,
                'scenef_1_2': sflow_1_2, 'global_p1': _bhw13(o['global_p1']),
                'staticflow_1_2': _bhw2(o['staticflow_1_2']), 'p1_camera_2': _bhw13(o['p1_camera_2']),
                'warped_p2_camera_2': _bhw13(o['warped_p2_camera_2'])}
