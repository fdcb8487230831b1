"""Factories and checkpoint helpers (reference ``utils/model_util.py``).

Same names (including the reference's spelling ``creat_serval_diffusion`` /
``creat_ddpm_ddim_diffusion``), same arguments, same return tuples; they build
the B200-native classes of this package instead of the torch-eager ones.
"""
from ..diffusion import gaussian_diffusion as gd
from ..diffusion.respace import SpacedDiffusion, space_timesteps
from ..diffusion.inpainting_gaussian_diffusion import InpaintingGaussianDiffusion
from ..model.mdm_forstyledataset import DiffuseTrasnfer, MDM, StyleDiffusion  # noqa: F401


def get_cond_mode(args):
    """reference utils/parser_util.py: unconstrained -> 'no_cond'; text datasets -> 'text'; else 'action'."""
    if getattr(args, 'unconstrained', False):
        return 'no_cond'
    if args.dataset in ['kit', 'humanml', "bandai-1_posrot", "bandai-2_posrot", 'stylexia_posrot']:
        return 'text'
    return 'action'


def load_model_wo_moenc(model, state_dict):
    """strict=False load where only motion_enc.* (and zero-conv) keys may be missing (reference :9-16)."""
    missing_keys, unexpected_keys = model.load_state_dict(state_dict, strict=False)
    assert len(unexpected_keys) == 0
    assert all([k.startswith('motion_enc.') or k.startswith('input_zero.') or k.startswith('output_zero.')
                for k in missing_keys])


def load_model_wo_controlmdm(model, state_dict):
    missing_keys, unexpected_keys = model.load_state_dict(state_dict, strict=False)
    assert len(unexpected_keys) == 0
    assert all([k.startswith('controlmdm.') for k in missing_keys])


def creat_serval_diffusion(args, ModelClass=DiffuseTrasnfer, timestep_respacing=''):
    """(model, respaced InpaintingGaussianDiffusion, full SpacedDiffusion) - reference :26-30."""
    model = ModelClass(**get_transfer_args(args))
    diffusion1 = create_gaussian_diffusion(args, InpaintingGaussianDiffusion, timestep_respacing=timestep_respacing)
    diffusion2 = create_gaussian_diffusion(args)
    return model, diffusion1, diffusion2


def creat_ddpm_ddim_diffusion(args, ModelClass=DiffuseTrasnfer, timestep_respacing=''):
    """(model, respaced InpaintingGaussianDiffusion, full InpaintingGaussianDiffusion) - reference :33-37."""
    model = ModelClass(**get_transfer_args(args))
    diffusion1 = create_gaussian_diffusion(args, InpaintingGaussianDiffusion, timestep_respacing=timestep_respacing)
    diffusion2 = create_gaussian_diffusion(args, InpaintingGaussianDiffusion)
    return model, diffusion1, diffusion2


_DATASET_FEATS = {
    # dataset -> (data_rep, njoints, nfeats)   (reference :57-73, :121-136)
    'humanml': ('hml_vec', 263, 1),
    'kit': ('hml_vec', 251, 1),
    'bandai-1_posrot': ('hml_vec', 190, 1),
    'bandai-2_posrot': ('hml_vec', 190, 1),
    'stylexia_posrot': ('hml_vec', 181, 1),
}


def _model_kwargs(args, num_actions, allow_kit):
    data_rep, njoints, nfeats = 'rot6d', 25, 6  # SMPL defaults
    if args.dataset in _DATASET_FEATS and (allow_kit or args.dataset != 'kit'):
        data_rep, njoints, nfeats = _DATASET_FEATS[args.dataset]
    zero_conv = True if (hasattr(args, 'zero_conv') and args.zero_conv) else None
    return {'modeltype': '', 'njoints': njoints, 'nfeats': nfeats, 'num_actions': num_actions,
            'translation': True, 'pose_rep': 'rot6d', 'glob': True, 'glob_rot': True,
            'latent_dim': args.latent_dim, 'ff_size': 1024, 'num_layers': args.layers, 'num_heads': 4,
            'dropout': 0.1, 'activation': "gelu", 'data_rep': data_rep, 'cond_mode': get_cond_mode(args),
            'cond_mask_prob': args.cond_mask_prob, 'action_emb': 'tensor', 'arch': args.arch,
            'emb_trans_dec': args.emb_trans_dec, 'clip_version': 'ViT-B/32', 'dataset': args.dataset,
            'mdm_path': getattr(args, 'mdm_path', ""),
            'semantic_discriminator_path': getattr(args, 'semantic_discriminator_path', ""),
            'zero_conv': zero_conv,
            "inpainting_model_path": getattr(args, 'inpainting_model_path', "")}


def get_model_args(args, data):
    """kwargs of MDM(...) for a dataset loader (reference :40-105)."""
    num_actions = data.dataset.num_actions if hasattr(data.dataset, 'num_actions') else 1
    return _model_kwargs(args, num_actions, allow_kit=True)


def get_transfer_args(args):
    """kwargs of StyleDiffusion(...) (reference :108-167; no 'kit' branch there)."""
    return _model_kwargs(args, 1, allow_kit=False)


def create_gaussian_diffusion(args, DiffusionClass=SpacedDiffusion, timestep_respacing=''):
    """x0-predicting, fixed-variance diffusion on the (optionally respaced) named schedule (reference :170-212)."""
    steps = args.diffusion_steps
    print(f"number of diffusion-steps: {steps}")
    betas = gd.get_named_beta_schedule(args.noise_schedule, steps, 1.)
    if not timestep_respacing:
        timestep_respacing = [steps]
    return DiffusionClass(
        use_timesteps=space_timesteps(steps, timestep_respacing),
        betas=betas,
        model_mean_type=gd.ModelMeanType.START_X,
        model_var_type=gd.ModelVarType.FIXED_SMALL if args.sigma_small else gd.ModelVarType.FIXED_LARGE,
        loss_type=gd.LossType.MSE,
        rescale_timesteps=False,
        lambda_vel=args.lambda_vel,
        lambda_rcxyz=args.lambda_rcxyz,
        lambda_fc=args.lambda_fc,
        lambda_sty_cons=getattr(args, "lambda_sty_cons", 0),
        lambda_sty_trans=getattr(args, "lambda_sty_trans", 0),
        lambda_cont_pers=getattr(args, "lambda_cont_pers", 0),
        lambda_cont_vel=getattr(args, "lambda_cont_vel", 0),
        lambda_diff_sty=getattr(args, "lambda_diff_sty", 0),
    )


def create_model_and_diffusion(args, data):
    """Upstream-MDM spelling named by the task statement (the reference itself has no such function,
    SURVEY section 0 row 1): MDM on the dataset's feature layout + the plain spaced diffusion."""
    model = MDM(**get_model_args(args, data))
    diffusion = create_gaussian_diffusion(args)
    return model, diffusion
