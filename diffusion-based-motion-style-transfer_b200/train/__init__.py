"""Host side of the few-shot style finetune step (reference train/training_loop.py, train/finetune_style_diffusion.py)."""
