"""bandai posrot skeleton (21 joints, 190 features) - reference data_loaders/bandai_posrot_utils.py."""
from .inpainting_masks import MaskLayout, export

LAYOUT = MaskLayout(
    joint_names=['Hips', 'Spine', 'Chest', 'Neck', 'Head', 'Shoulder_L', 'UpperArm_L', 'LowerArm_L', 'Hand_L',
                 'Shoulder_R', 'UpperArm_R', 'LowerArm_R', 'Hand_R', 'UpperLeg_L', 'LowerLeg_L', 'Foot_L', 'Toes_L',
                 'UpperLeg_R', 'LowerLeg_R', 'Foot_R', 'Toes_R'],
    lower_body=['Hips', 'UpperLeg_L', 'LowerLeg_L', 'Foot_L', 'Toes_L', 'UpperLeg_R', 'LowerLeg_R', 'Foot_R', 'Toes_R'],
    layout="posrot", extra_names=('none',))
export(LAYOUT, globals())
HML_NONE_MASK = LAYOUT.none
