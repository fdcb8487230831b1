"""Seeded inputs of the N4 (CompMDMGeneratedDataset) golden, shared by make_golden_n4.py and the tests."""
import numpy as np
import torch

from oracle.weights import text_features

B, F, T, N_BATCHES, SPEC, SCALE, MM_SAMPLES, MM_REPEATS = 2, 181, 20, 3, "4", 2.5, 2, 3


class FakeT2MDataset(torch.utils.data.Dataset):
    """What the class reads from dataloader.dataset: __len__, w_vectorizer, mode (only in __getitem__)."""
    mode = "train"
    w_vectorizer = {"a/DET": (np.ones(4, np.float32), np.eye(3, dtype=np.float32)[0])}

    def __init__(self):
        g = torch.Generator().manual_seed(17)
        self.items = []
        for i in range(B * N_BATCHES):
            words = ["a/DET"] * (2 + i % 3)
            self.items.append((torch.randn(F, 1, T, generator=g), f"caption number {i}", "_".join(words), T - (i % 4)))

    def __len__(self):
        return len(self.items)

    def __getitem__(self, i):
        return self.items[i]


def collate(batch):
    motion = torch.stack([b[0] for b in batch])
    y = {"text": [b[1] for b in batch], "tokens": [b[2] for b in batch], "lengths": torch.tensor([b[3] for b in batch]),
         "mask": torch.ones(len(batch), 1, 1, T, dtype=torch.bool)}
    y["text_feat"] = text_features(y["text"])  # the native side reads features; the reference side goes through stub CLIP
    return motion, {"y": y}


def loader():
    return torch.utils.data.DataLoader(FakeT2MDataset(), batch_size=B, shuffle=False, collate_fn=collate)
