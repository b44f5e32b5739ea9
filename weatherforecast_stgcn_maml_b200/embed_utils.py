"""Drop-in for the reference's ``embed_utils.KoppenEmbedding`` (embed_utils.py:30-38).

Same constructor, attributes and ``state_dict`` (``embedding.weight [31, 8]``).  On the
reference path the embedding row is concatenated into the input features once per task at
dataset-build time and detached (featurePreprocessor.py:170-177, dataset.py:50-53), so it never
receives a gradient; the lookup itself is an ordinary ``nn.Embedding``.
``add_time_embeddings`` (embed_utils.py:10-27) is host-side xarray preprocessing and out of
scope; synth.synth_features restates its four phase features for synthetic inputs.
"""
import torch.nn as nn


class KoppenEmbedding(nn.Module):
    def __init__(self, embedding_dim=8):
        super().__init__()
        self.num_classes = 31  # indices 0-30, 0 unused/padding
        self.embedding_dim = embedding_dim
        self.embedding = nn.Embedding(self.num_classes, embedding_dim)

    def forward(self, koppen_codes):
        return self.embedding(koppen_codes)
