"""Seeded synthetic inputs with the shapes the reference data path produces.

Layout follows /root/reference/utils/image_features_reader.py:120-141 (36 regions + a global
mean-pooled row 0, 5-d box encoding), /root/reference/utils/data_utils.py:34-71 (``[CLS] cap [SEP]``
padded to max_seq_len, segments, attention mask) and
/root/reference/dataloader/dataloader_cc12m_gen.py:75-101 (decoder start token).

Every image draws from its own generator ``seed + global_image_index`` so a shard's content does
not depend on how many GPUs the job is split over.
"""
from __future__ import annotations

import torch

CLS, SEP, PAD, UNK, MASK = 101, 102, 0, 100, 103
SPECIAL_IDS = (PAD, UNK, CLS, SEP, MASK)


def _rand_tokens(n, vocab_size, g):
    lo = 1000 if vocab_size > 2000 else 104
    return torch.randint(lo, vocab_size, (n,), generator=g, dtype=torch.int64)


def synthetic_image(index: int, v_feature_size: int = 2048, num_regions: int = 36, seed: int = 1234):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed + index)
    feat = torch.zeros(num_regions + 1, v_feature_size)
    feat[1:] = torch.relu(torch.randn(num_regions, v_feature_size, generator=g))
    feat[0] = feat[1:].mean(0)
    loc = torch.zeros(num_regions + 1, 5)
    xy = torch.rand(num_regions, 2, generator=g) * 0.7
    wh = torch.rand(num_regions, 2, generator=g) * 0.2 + 0.1
    loc[1:, 0:2] = xy
    loc[1:, 2:4] = xy + wh
    loc[1:, 4] = wh[:, 0] * wh[:, 1]
    loc[0] = torch.tensor([0.0, 0.0, 1.0, 1.0, 1.0])
    return feat, loc, g


def synthetic_batch(start: int, count: int, vocab_size: int = 30522, v_feature_size: int = 2048,
                    max_seq_len: int = 256, num_regions: int = 36, seed: int = 1234,
                    cap_len=(8, 38)):
    """Batch dict with the tensor names generate.py reads (generate.py:95-111)."""
    feats, locs, ids, segs = [], [], [], []
    for i in range(start, start + count):
        feat, loc, g = synthetic_image(i, v_feature_size, num_regions, seed)
        n = int(torch.randint(cap_len[0], cap_len[1] + 1, (1,), generator=g))
        cap = _rand_tokens(n, vocab_size, g)
        row = torch.zeros(max_seq_len, dtype=torch.int64)
        row[0] = CLS
        row[1:1 + n] = cap
        row[1 + n] = SEP
        seg = torch.zeros(max_seq_len, dtype=torch.int64)
        seg[: n + 2] = 1
        feats.append(feat); locs.append(loc); ids.append(row); segs.append(seg)
    enc_input_ids = torch.stack(ids)
    return {
        "enc_image_feat": torch.stack(feats),
        "enc_image_loc": torch.stack(locs),
        "enc_image_mask": torch.ones(count, num_regions + 1),
        "enc_input_ids": enc_input_ids,
        "enc_segments": torch.stack(segs),
        "enc_att_mask": (enc_input_ids != 0).float(),
        "dec_input_ids": torch.full((count, 1), CLS, dtype=torch.int64),
        "dec_att_mask": torch.ones(count, 1),
        "image_id": torch.arange(start, start + count),
    }


def synthetic_utterance(index: int, rnd: int, vocab_size: int = 30522, max_len: int = 18, seed: int = 4321,
                        length=(6, 12)):
    """A random 6-12 token utterance ending in [SEP], zero-padded to ``max_len`` (config 2's stand-in question)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed + 1009 * index + rnd)
    n = int(torch.randint(length[0], length[1] + 1, (1,), generator=g))
    out = torch.zeros(max_len, dtype=torch.int64)
    out[: n - 1] = _rand_tokens(n - 1, vocab_size, g)
    out[n - 1] = SEP
    return out
