"""Diffusion process (host side) of the B200-native sampler: schedules, respacing,
inpainting override.  Module names follow the reference's ``diffusion`` package."""
