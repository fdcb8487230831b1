"""stylexia_posrot skeleton (20 joints, 181 features) - reference data_loaders/stylexia_posrot_utils.py."""
from .inpainting_masks import MaskLayout, export

LAYOUT = MaskLayout(
    joint_names=['root', 'pelvis', 'thorax', 'rclavicle', 'rhumerus', 'rradius', 'rhand', 'lclavicle', 'lhumerus',
                 'lradius', 'lhand', 'head', 'rfemur', 'rtibia', 'rfoot', 'rtoes', 'lfemur', 'ltibia', 'lfoot', 'ltoes'],
    lower_body=['root', 'pelvis', 'rfemur', 'rtibia', 'rfoot', 'rtoes', 'lfemur', 'ltibia', 'lfoot', 'ltoes'],
    layout="posrot")
export(LAYOUT, globals())
