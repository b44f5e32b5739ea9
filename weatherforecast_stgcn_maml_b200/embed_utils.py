"""Drop-in for the reference's ``embed_utils.KoppenEmbedding`` (embed_utils.py:30-38).

Same constructor, attributes and ``state_dict`` (``embedding.weight [31, 8]``).  On the
reference path the embedding row is concatenated into the input features once per task at
dataset-build time and detached (featurePreprocessor.py:170-177, dataset.py:50-53), so it never
receives a gradient; the lookup itself is an ordinary ``nn.Embedding``.
``time_features`` restates the arithmetic of ``add_time_embeddings`` (embed_utils.py:9-27) on plain arrays
(the xarray bookkeeping around it is ingest and out of scope).
"""
import numpy as np
import torch.nn as nn


def time_features(day_of_year, time_of_day):
    """[time, 4] f64: sin / cos of ``2 pi day_of_year / 365.25`` and of ``2 pi time_of_day / 24`` (hours, fractional), in
    the order featurePreprocessor.TIME_VARS reads them (embed_utils.py:12-26)."""
    year = 2 * np.pi * np.asarray(day_of_year) / 365.25
    day = 2 * np.pi * np.asarray(time_of_day, dtype=np.float64) / 24.0
    return np.stack([np.sin(year), np.cos(year), np.sin(day), np.cos(day)], axis=-1)


class KoppenEmbedding(nn.Module):
    def __init__(self, embedding_dim=8):
        super().__init__()
        self.num_classes = 31  # indices 0-30, 0 unused/padding
        self.embedding_dim = embedding_dim
        self.embedding = nn.Embedding(self.num_classes, embedding_dim)

    def forward(self, koppen_codes):
        return self.embedding(koppen_codes)
