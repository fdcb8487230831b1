"""Post-sampling decode of the hml_vec representation (reference
``data_loaders/humanml/scripts/motion_process.py:389-461``) - the step every caller runs right after the sampling
loop (``sample/demo_style_transfer.py:265-268``, ``train/finetune_style_diffusion.py:331-333``).

Only the decode on the sampling path is mirrored here (``recover_root_rot_pos``, ``recover_from_ric``); the dataset
pre-processing functions of the reference module stay in the reference.  The work is one CUDA kernel
(``mst_recover_from_ric``); ``decode_motion`` additionally fuses ``inv_transform`` and the ``permute(0, 2, 3, 1)``.
"""
import torch

from .... import engine as K


def _require_cuda(t):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RuntimeError("recover_from_ric: the mst decode runs on CUDA tensors only (no CPU fallback)")


def recover_from_ric(data, joints_num):
    """data [..., T, F] de-normalised hml_vec (the reference's layout, e.g. [B, 1, T, F]) -> [..., T, J, 3]."""
    _require_cuda(data)
    lead, T, F = data.shape[:-2], data.shape[-2], data.shape[-1]
    x = data.reshape(-1, T, F).permute(0, 2, 1).float().contiguous().unsqueeze(2)   # [N, F, 1, T]
    out = K.recover_from_ric(x, joints_num)                                         # [N, 1, T, J, 3]
    return out.reshape(*lead, T, joints_num, 3)


def decode_motion(sample, mean, std, joints_num):
    """sample [B, F, 1, T] straight from p_sample_loop; mean / std [F] -> joints [B, 1, T, J, 3]
    (= recover_from_ric(inv_transform(sample.permute(0, 2, 3, 1)), joints_num) in one kernel)."""
    _require_cuda(sample)
    dev = sample.device
    mean = torch.as_tensor(mean, dtype=torch.float32, device=dev).contiguous()
    std = torch.as_tensor(std, dtype=torch.float32, device=dev).contiguous()
    return K.recover_from_ric(sample.float().contiguous(), joints_num, mean, std)
