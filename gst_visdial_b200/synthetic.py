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


def _append(row_ids, row_seg, pos, utt, seg_value):
    """Appends the non-zero ids of ``utt`` at ``pos`` (generate.py:148-160 without the overflow case: callers keep rows short)."""
    n = int((utt != 0).sum())
    row_ids[pos:pos + n] = utt[:n]
    if seg_value is not None:
        row_seg[pos:pos + n] = seg_value
    return pos + n


def synthetic_history_batch(start: int, count: int, vocab_size: int = 30522, v_feature_size: int = 2048, max_seq_len: int = 256,
                            rounds=None, seed: int = 1234):
    """``synthetic_batch`` whose histories already hold ``rounds[i]`` question / answer rounds (questions keep segment 0, answers get
    segment 1 and no [SEP], like generate.py:148-160,214-228 leaves them): the contexts of BASELINE config 4.  Adds
    ``hist_len_bound`` (int64 [1]): the longest history of the batch, known on the host without a device read."""
    b = synthetic_batch(start, count, vocab_size, v_feature_size, max_seq_len, seed=seed)
    ids, seg = b["enc_input_ids"], b["enc_segments"]
    longest = 0
    for i in range(count):
        pos = int((ids[i] != 0).sum())
        for r in range(int(rounds[i]) if rounds is not None else 0):
            q = synthetic_utterance(start + i, 2 * r, vocab_size)
            a = synthetic_utterance(start + i, 2 * r + 1, vocab_size)
            a = a.masked_fill(a == SEP, 0)
            if pos + int((q != 0).sum()) + int((a != 0).sum()) > max_seq_len:
                break
            pos = _append(ids[i], seg[i], pos, q, None)
            pos = _append(ids[i], seg[i], pos, a, 1)
        longest = max(longest, pos)
    b["enc_att_mask"] = (ids != 0).float()
    b["hist_len_bound"] = torch.tensor([longest], dtype=torch.int64)
    return b


def synthetic_candidate_batch(start: int, items: int, candidates: int = 100, vocab_size: int = 30522, v_feature_size: int = 2048,
                              max_seq_len: int = 256, seed: int = 1234):
    """BASELINE config 5 in the shape of evaluate_disc.py:66-83 / dataloader_visdial_disc.py:320-323: every item is one
    (image, caption, r ~ U{1..10} rounds of history, question) with ``candidates`` answer options; each option is appended to the
    text stream, so every candidate is its own ``max_seq_len``-token encoder input.  Segments alternate per utterance
    (utils/data_utils.py:57), the attention mask covers positions up to the last [SEP] (train_disc.py:97-99).
    Returns image tensors per ITEM ([items, 37, ...]) and token tensors per candidate ([items, candidates, max_seq_len])."""
    feats, locs = [], []
    tokens = torch.zeros(items, candidates, max_seq_len, dtype=torch.int64)
    segments = torch.zeros(items, candidates, max_seq_len, dtype=torch.int64)
    longest = 0
    for i in range(items):
        feat, loc, g = synthetic_image(start + i, v_feature_size, 36, seed)
        feats.append(feat); locs.append(loc)
        n = int(torch.randint(8, 39, (1,), generator=g))
        row = torch.zeros(max_seq_len, dtype=torch.int64)
        sg = torch.zeros(max_seq_len, dtype=torch.int64)
        row[0] = CLS
        row[1:1 + n] = _rand_tokens(n, vocab_size, g)
        row[1 + n] = SEP
        pos, cur_seg = n + 2, 1                               # caption utterance: segment 0 ... alternate from here on
        r = int(torch.randint(1, 11, (1,), generator=g))
        for u in range(2 * (r - 1) + 1):                      # (r - 1) complete rounds, then the current question
            utt = synthetic_utterance(start + i, u, vocab_size)
            if pos + int((utt != 0).sum()) + 14 > max_seq_len:
                break
            pos = _append(row, sg, pos, utt, cur_seg)
            cur_seg ^= 1
        for c in range(candidates):
            cand = synthetic_utterance((start + i) * 131 + c, 500 + c, vocab_size)
            tokens[i, c] = row
            segments[i, c] = sg
            end = _append(tokens[i, c], segments[i, c], pos, cand, cur_seg)
            longest = max(longest, end)
    mask = (tokens != 0).float()
    return {"image_feat": torch.stack(feats), "image_loc": torch.stack(locs), "image_mask": torch.ones(items, 37),
            "tokens": tokens, "segments": segments, "mask": mask, "hist_len_bound": torch.tensor([longest], dtype=torch.int64)}
