"""Denoiser modules of the B200-native sampler.  Module names follow the reference's ``model`` package."""
