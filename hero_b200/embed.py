"""Input embedding modules with the reference's parameter names (model/embed.py).

On the HERO encoder path these are parameter containers: `CrossModalTrm` / `TemporalTrm` feed
their tables straight into the fused gather+add+LayerNorm kernels (functional.cross_modal_embed,
functional.frame_embed). The plain `forward`s below exist for API compatibility with code that
calls the embedding modules directly (model/videoQA.py:70-79, model/tvc.py:251) and use the same
kernels through the padded<->packed adapters where that is cheap, torch ops otherwise.
"""
import torch
from torch import nn

from .layers import BertLayerNorm


class SubEmbeddings(nn.Module):
    """model/embed.py:12-86."""

    def __init__(self, config):
        super().__init__()
        self.padding_idx = 1
        self.word_embeddings = nn.Embedding(config.vocab_size, config.hidden_size,
                                            padding_idx=self.padding_idx)
        self.position_embeddings = nn.Embedding(config.max_position_embeddings,
                                                config.hidden_size)
        self.token_type_embeddings = nn.Embedding(config.type_vocab_size, config.hidden_size)
        self.LayerNorm = BertLayerNorm(config.hidden_size, eps=1e-5)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)

    def create_position_ids_from_input_ids(self, x):
        mask = x.ne(self.padding_idx).long()
        return torch.cumsum(mask, dim=1) * mask + self.padding_idx

    def forward(self, input_ids=None, position_ids=None, token_type_ids=None, inputs_embeds=None):
        # Padded-layout torch restatement (not on the hot path; the encoders call the fused kernel).
        if position_ids is None:
            if input_ids is not None:
                position_ids = self.create_position_ids_from_input_ids(input_ids)
            else:
                L = inputs_embeds.shape[1]
                position_ids = torch.arange(self.padding_idx + 1, L + self.padding_idx + 1,
                                            device=inputs_embeds.device).unsqueeze(0)
        if inputs_embeds is None:
            inputs_embeds = self.word_embeddings(input_ids)
        if token_type_ids is None:
            type_emb = self.token_type_embeddings.weight[1]
        else:
            type_emb = self.token_type_embeddings(token_type_ids)
        e = inputs_embeds + self.position_embeddings(position_ids) + type_emb
        return self.dropout(self.LayerNorm(e))


class ImageEmbeddings(nn.Module):
    """model/embed.py:89-133."""

    def __init__(self, config, img_dim, max_img_seq_len):
        super().__init__()
        self.img_linear = nn.Linear(img_dim, config.hidden_size)
        self.img_LayerNorm = BertLayerNorm(img_dim, eps=1e-5)
        self.position_embeddings = nn.Embedding(max_img_seq_len, config.hidden_size)
        self.mask_embedding = nn.Embedding(2, img_dim, padding_idx=0)
        self.LayerNorm = BertLayerNorm(config.hidden_size, eps=1e-5)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)

    def forward(self, img_feat, type_embeddings, img_pos_ids=None, img_masks=None):
        if img_pos_ids is None:
            img_pos_ids = torch.arange(img_feat.shape[1], device=img_feat.device).unsqueeze(0)
        if img_masks is not None:
            img_feat = img_feat + self.mask_embedding(img_masks.long())
        x = self.img_linear(self.img_LayerNorm(img_feat))
        x = x + self.position_embeddings(img_pos_ids) + type_embeddings
        return self.dropout(self.LayerNorm(x))


class FrameEmbeddings(nn.Module):
    """model/embed.py:136-161."""

    def __init__(self, config):
        super().__init__()
        self.position_embeddings = nn.Embedding(config.max_position_embeddings,
                                                config.hidden_size)
        self.LayerNorm = BertLayerNorm(config.hidden_size, eps=1e-5)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)

    def forward(self, frame_feat, position_ids=None):
        if position_ids is None:
            position_ids = torch.arange(frame_feat.shape[1], device=frame_feat.device).unsqueeze(0)
        e = frame_feat + self.position_embeddings(position_ids)
        return self.dropout(self.LayerNorm(e))


class QueryFeatEmbeddings(nn.Module):
    """model/embed.py:164-188 (used by QueryFeatEncoder, a 'next' row)."""

    def __init__(self, config):
        super().__init__()
        self.position_embeddings = nn.Embedding(config.max_position_embeddings,
                                                config.hidden_size)
        self.LayerNorm = BertLayerNorm(config.hidden_size, eps=1e-5)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)

    def forward(self, input_feat, position_ids=None):
        if position_ids is None:
            position_ids = torch.arange(input_feat.shape[1], device=input_feat.device).unsqueeze(0)
        e = self.LayerNorm(input_feat + self.position_embeddings(position_ids))
        return self.dropout(e)
