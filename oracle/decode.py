"""torch-CPU restatement of the post-sampling decode (oracle; test infrastructure only).

* inv_transform          reference data_loaders/humanml/data/dataset.py:478-479
* recover_root_rot_pos   reference data_loaders/humanml/scripts/motion_process.py:389-411
* recover_from_ric       reference data_loaders/humanml/scripts/motion_process.py:444-461
* qrot                   reference data_loaders/humanml/common/quaternion.py:88-99

Pinned by tests/golden/make_golden_decode.py (runs the reference's own recover_from_ric)."""
import torch


def qrot(q, vec):
    s, u = q[..., :1], q[..., 1:]
    u, vec = torch.broadcast_tensors(u, vec)
    uv = torch.cross(u, vec, dim=-1)
    uuv = torch.cross(u, uv, dim=-1)
    return vec + 2 * (s * uv + uuv)


def recover_root_rot_pos(data):
    rot_vel = data[..., 0]
    ang = torch.zeros_like(rot_vel)
    ang[..., 1:] = rot_vel[..., :-1]
    ang = torch.cumsum(ang, dim=-1)
    quat = torch.zeros(data.shape[:-1] + (4,))
    quat[..., 0] = torch.cos(ang)
    quat[..., 2] = torch.sin(ang)
    pos = torch.zeros(data.shape[:-1] + (3,))
    pos[..., 1:, [0, 2]] = data[..., :-1, 1:3]
    pos = qrot(quat, pos)
    pos = torch.cumsum(pos, dim=-2)
    pos[..., 1] = data[..., 3]
    return quat, pos


def recover_from_ric(data, joints_num):
    """data [..., T, F] -> [..., T, J, 3]"""
    quat, pos = recover_root_rot_pos(data)
    p = data[..., 4:(joints_num - 1) * 3 + 4]
    p = p.reshape(p.shape[:-1] + (-1, 3))
    p = qrot(quat[..., None, :].expand(p.shape[:-1] + (4,)), p)
    p = p.clone()
    p[..., 0] += pos[..., 0:1]
    p[..., 2] += pos[..., 2:3]
    return torch.cat([pos.unsqueeze(-2), p], dim=-2)


def decode_motion(sample, mean, std, joints_num):
    """sample [B,F,1,T] -> [B,1,T,J,3]"""
    return recover_from_ric(sample.permute(0, 2, 3, 1) * std + mean, joints_num)
