"""HumanML3D skeleton (22 joints, 263 features, hml_vec layout) - reference data_loaders/humanml_utils.py."""
from .inpainting_masks import MaskLayout, export

LAYOUT = MaskLayout(
    joint_names=['pelvis', 'left_hip', 'right_hip', 'spine1', 'left_knee', 'right_knee', 'spine2', 'left_ankle',
                 'right_ankle', 'spine3', 'left_foot', 'right_foot', 'neck', 'left_collar', 'right_collar', 'head',
                 'left_shoulder', 'right_shoulder', 'left_elbow', 'right_elbow', 'left_wrist', 'right_wrist'],
    lower_body=['pelvis', 'left_hip', 'right_hip', 'left_knee', 'right_knee', 'left_ankle', 'right_ankle',
                'left_foot', 'right_foot'],
    layout="hml_vec", extra_names=('linear_vel', 'xz_plane', 'right_hand'), right_hand=['right_wrist', 'right_elbow'])
export(LAYOUT, globals())
HML_LINEAR_VEL_MASK, HML_XZPLANE_MASK, HML_RIGHT_HAND_MASK = LAYOUT.linear_vel, LAYOUT.xz_plane, LAYOUT.right_hand
HML_RIGHT_HAND_JOINTS = [LAYOUT.joint_names.index(n) for n in ('right_wrist', 'right_elbow')]
