"""Feature-row inpainting masks, one table-driven implementation for every skeleton.

The reference keeps four near-identical copies of this logic
(``data_loaders/{stylexia_posrot,bandai_posrot,humanml_posrot,humanml}_utils.py``).
They differ only in the joint list, the lower-body joint set and the feature
layout of a frame:

* ``posrot``  : [root rot-vel, root lin-vel x2, root y | (J-1)*3 ric | J*6 rot]
* ``hml_vec`` : [4 root | (J-1)*3 ric | (J-1)*6 rot | J*3 local vel | 4 foot contacts]

``get_inpainting_mask(name, shape)`` returns the float64 ndarray ``[B, F, 1, T]`` in {0,1} the
reference returns (callers ``.float()`` it, ``sample/demo_style_transfer.py:229``); 1 = keep the
given motion, 0 = let the sampler generate.  Masks must match the reference bit-exactly.
"""
import numpy as np


class MaskLayout:
    def __init__(self, joint_names, lower_body, layout, extra_names=(), right_hand=()):
        assert layout in ("posrot", "hml_vec")
        self.joint_names = list(joint_names)
        self.layout = layout
        self.J = len(self.joint_names)
        self.extra_names = set(extra_names)
        J = self.J
        is_root = np.array([True] + [False] * (J - 1))
        lower = np.array([n in lower_body for n in self.joint_names])
        self.lower_body_joints = [i for i in range(J) if lower[i]]
        zeros_j = np.zeros(J, dtype=bool)

        def rows(head4, per_joint, contacts=False):
            """Assemble a feature-row vector from the 4 root channels and a per-joint boolean."""
            per_joint = np.asarray(per_joint, dtype=bool)
            if layout == "posrot":
                parts = [np.array(head4, dtype=bool), per_joint[1:].repeat(3), per_joint.repeat(6)]
            else:
                parts = [np.array(head4, dtype=bool), per_joint[1:].repeat(3), per_joint[1:].repeat(6),
                         per_joint.repeat(3), np.array([contacts] * 4, dtype=bool)]
            return np.concatenate(parts)

        self._rows = rows
        self.root = rows([True] * 4, is_root)
        self.root_horizontal = rows([True, True, True, False], zeros_j)
        self.none = rows([False] * 4, zeros_j)
        self.y_rotation = rows([True, False, False, False], zeros_j)
        self.linear_vel = rows([False, True, True, False], zeros_j)
        self.xz_plane = rows([False, True, True, False], zeros_j)
        self.lower_body = rows([True] * 4, lower, contacts=True)
        self.upper_body = ~self.lower_body
        self.right_hand = rows([False] * 4, np.array([n in right_hand for n in self.joint_names]))
        self.traj = np.zeros_like(self.root)
        self.traj[1:3] = True
        self.num_feats = self.root.shape[0]

    # -- helpers with the reference's names ------------------------------------------------
    @staticmethod
    def expand_mask(mask, shape):
        """(num_feat[, seq_len]) -> shape (usually [B, num_feat, 1, T])."""
        _, num_feat, _, _ = shape
        return np.ones(shape) * mask.reshape((1, num_feat, 1, -1))

    def get_joints_mask(self, join_names):
        """positions (ric rows) of the named joints only"""
        hit = np.array([n in join_names for n in self.joint_names])
        if self.layout == "posrot":
            return np.concatenate(([False] * 4, hit[1:].repeat(3), np.zeros_like(hit.repeat(6))))
        return np.concatenate(([False] * 4, hit[1:].repeat(3), np.zeros_like(hit[1:].repeat(6)),
                               np.zeros_like(hit.repeat(3)), [False] * 4))

    def get_batch_joint_mask(self, shape, joint_names):
        return self.expand_mask(self.get_joints_mask(joint_names), shape)

    @staticmethod
    def get_in_between_mask(shape, lengths, prefix_end, suffix_end):
        mask = np.ones(shape)
        for i, length in enumerate(lengths):
            lo, hi = int(prefix_end * length), int(suffix_end * length)
            mask[i, :, :, lo:hi] = 0
        return mask

    def get_prefix_mask(self, shape, prefix_length=20):
        _, num_feat, _, seq_len = shape
        m = np.concatenate((np.ones((num_feat, prefix_length)), np.zeros((num_feat, seq_len - prefix_length))), axis=-1)
        return self.expand_mask(m, shape)

    def get_inpainting_mask(self, mask_name, shape, **kwargs):
        names = mask_name.split(',')
        mask = np.zeros(shape)

        def add(m):
            return np.maximum(mask, m)

        if 'in_between' in names:
            mask = add(self.get_in_between_mask(shape, **kwargs))
        if 'none' in names and 'none' in self.extra_names:
            mask = add(self.expand_mask(self.none, shape))
        if 'root' in names:
            mask = add(self.expand_mask(self.root, shape))
        if 'root_horizontal' in names:
            mask = add(self.expand_mask(self.root_horizontal, shape))
        if 'linear_vel' in names and 'linear_vel' in self.extra_names:
            mask = add(self.expand_mask(self.linear_vel, shape))
        if 'y_rotation' in names:
            mask = add(self.expand_mask(self.y_rotation, shape))
        if 'xz_plane' in self.extra_names and 'xz_plane' in mask_name:  # substring test, as in the reference
            mask = add(self.expand_mask(self.xz_plane, shape))
        if 'prefix' in names:
            mask = add(self.get_prefix_mask(shape, **kwargs))
        if 'upper_body' in names:
            mask = add(self.expand_mask(self.upper_body, shape))
        if 'lower_body' in names:
            mask = add(self.expand_mask(self.lower_body, shape))
        if 'right_hand' in names and 'right_hand' in self.extra_names:
            mask = add(self.expand_mask(self.right_hand, shape))
        return np.maximum(mask, self.get_batch_joint_mask(shape, names))


def export(layout: MaskLayout, namespace: dict):
    """Publish the reference's module-level names for one skeleton into ``namespace``."""
    namespace.update(
        HML_JOINT_NAMES=layout.joint_names, BVH_JOINT_NAMES=list(layout.joint_names), NUM_HML_JOINTS=layout.J,
        HML_LOWER_BODY_JOINTS=layout.lower_body_joints,
        SMPL_UPPER_BODY_JOINTS=[i for i in range(layout.J) if i not in layout.lower_body_joints],
        HML_ROOT_MASK=layout.root, HML_ROOT_HORIZONTAL_MASK=layout.root_horizontal,
        HML_YROTATION_MASK=layout.y_rotation, HML_LOWER_BODY_MASK=layout.lower_body,
        HML_UPPER_BODY_MASK=layout.upper_body, HML_TRAJ_MASK=layout.traj, NUM_HML_FEATS=layout.num_feats,
        expand_mask=layout.expand_mask, get_joints_mask=layout.get_joints_mask,
        get_batch_joint_mask=layout.get_batch_joint_mask, get_in_between_mask=layout.get_in_between_mask,
        get_prefix_mask=layout.get_prefix_mask, get_inpainting_mask=layout.get_inpainting_mask,
    )
