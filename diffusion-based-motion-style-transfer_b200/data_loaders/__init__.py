"""Inpainting masks and batch collation of the B200-native sampler (the only parts of the reference's
``data_loaders`` package that sit on the sampling hot path)."""
