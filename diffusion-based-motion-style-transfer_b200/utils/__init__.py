"""Factories of the B200-native sampler (reference ``utils`` package, model_util only)."""
