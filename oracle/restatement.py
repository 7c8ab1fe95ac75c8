"""TEST INFRASTRUCTURE ONLY - CPU restatement (plain PyTorch, fp32) of the reference's generation hot path.

Nothing in the product path (``gst_visdial_b200/``) may import this module: it is the checker for the
``-m gpu`` parity tests, for ``__graft_entry__.smoke()``, and the thing timed by ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` leg.  It exists because the reference itself (a Python repo
under /root/reference) cannot travel to the GPU box.

Pinning: the reference has no tests, golden vectors or fixtures for this path, so the
restatement is pinned against outputs of the reference's own classes, imported under ``oracle/ref_shim.py``
in the development container on seeded weights/inputs: ``oracle/gen_golden.py`` asserts equality there and
commits the resulting vectors to ``tests/golden/``; ``tests/test_oracle_golden.py`` re-checks the restatement
against those files everywhere.  The decoder layer arithmetic is third-party (``transformers==4.16.2``
``BertLayer``, absent from /root/reference; this image has 5.5.0 whose eager ``BertLayer`` is the same post-LN
block) - its published algorithm is restated in ``decoder_layers`` and anchored on the reference call site
models/visual_dialog_decoder.py:300-311.

All functions take ``sd``: the flat ``EncoderDecoderModel.state_dict()`` (see gst_visdial_b200/weights.py).
File:line citations are into /root/reference.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

EOS, PAD, CLS = 102, 0, 101
SPECIAL_IDS = (0, 100, 101, 102, 103)
ENC = "encoder.bert_pretrained.bert."
CLS_HEAD = "encoder.bert_pretrained.cls."
DEC = "decoder.decoder.bert."
LMH = "decoder.decoder.lm_head."


def lin(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"])


def layer_norm(sd, name, x, eps=1e-12):
    """models/vilbert_dialog.py:283-296 - TF-style LayerNorm: biased variance, eps inside the sqrt."""
    u = x.mean(-1, keepdim=True)
    s = (x - u).pow(2).mean(-1, keepdim=True)
    x = (x - u) / torch.sqrt(s + eps)
    return sd[name + ".weight"] * x + sd[name + ".bias"]


def gelu(x):
    """models/vilbert_dialog.py:115-121 (erf form)."""
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def _heads(x, n):
    b, L, H = x.shape
    return x.view(b, L, n, H // n).permute(0, 2, 1, 3)


def _attend(q, k, v, add_mask, n_heads):
    """softmax(q k^T / sqrt(d) + mask) v  -  models/vilbert_dialog.py:385-407 (division by sqrt(d), additive mask)."""
    qh, kh, vh = _heads(q, n_heads), _heads(k, n_heads), _heads(v, n_heads)
    d = qh.shape[-1]
    s = torch.matmul(qh, kh.transpose(-1, -2)) / math.sqrt(d)
    s = s + add_mask
    p = torch.softmax(s, dim=-1)
    c = torch.matmul(p, vh).permute(0, 2, 1, 3).contiguous()
    return c.view(c.shape[0], c.shape[1], -1)


def text_embeddings(sd, prefix, ids, seg=None, type_vocab_size=2):
    """models/vilbert_dialog.py:324-352.  Positions are 0..L-1 (no past offset)."""
    L = ids.shape[1]
    pos = torch.arange(L, dtype=torch.long).unsqueeze(0).expand_as(ids)
    if seg is None:
        seg = torch.zeros_like(ids)
    w = sd[prefix + "word_embeddings.weight"][ids]
    p = sd[prefix + "position_embeddings.weight"][pos]
    ext = seg - type_vocab_size
    ext_mask = (ext >= 0).float()
    ext = (ext.float() * ext_mask).long()
    base_mask = (seg < type_vocab_size).float()
    base = (seg.float() * base_mask).long()
    t = sd[prefix + "token_type_embeddings.weight"][base] * base_mask.unsqueeze(-1) + \
        sd[prefix + "token_type_embeddings_extension.weight"][ext] * ext_mask.unsqueeze(-1)
    return layer_norm(sd, prefix + "LayerNorm", w + p + t)


def image_embeddings(sd, feat, loc):
    """models/vilbert_dialog.py:1420-1427."""
    p = ENC + "v_embeddings."
    return layer_norm(sd, p + "LayerNorm", lin(sd, p + "image_embeddings", feat) + lin(sd, p + "image_location_embeddings", loc))


def self_layer(sd, p, x, add_mask, n_heads):
    """Post-LN block: models/vilbert_dialog.py:380-476 (text) / :507-603 (image)."""
    c = _attend(lin(sd, p + "attention.self.query", x), lin(sd, p + "attention.self.key", x),
                lin(sd, p + "attention.self.value", x), add_mask, n_heads)
    y = layer_norm(sd, p + "attention.output.LayerNorm", lin(sd, p + "attention.output.dense", c) + x)
    z = layer_norm(sd, p + "output.LayerNorm", lin(sd, p + "output.dense", gelu(lin(sd, p + "intermediate.dense", y))) + y)
    return z


def connection_layer(sd, p, v, m_v, t, m_t, n_heads):
    """models/vilbert_dialog.py:646-773.  Contexts are swapped into the opposite stream (:765)."""
    b = p + "biattention."
    q1, k1, v1 = lin(sd, b + "query1", v), lin(sd, b + "key1", v), lin(sd, b + "value1", v)
    q2, k2, v2 = lin(sd, b + "query2", t), lin(sd, b + "key2", t), lin(sd, b + "value2", t)
    ctx_t = _attend(q2, k1, v1, m_v, n_heads)     # text queries over image keys  [B, Lt, Hb]
    ctx_v = _attend(q1, k2, v2, m_t, n_heads)     # image queries over text keys  [B, Lv, Hb]
    o = p + "biOutput."
    v_ = layer_norm(sd, o + "LayerNorm1", lin(sd, o + "dense1", ctx_v) + v)
    t_ = layer_norm(sd, o + "LayerNorm2", lin(sd, o + "dense2", ctx_t) + t)
    v2_ = layer_norm(sd, p + "v_output.LayerNorm", lin(sd, p + "v_output.dense", gelu(lin(sd, p + "v_intermediate.dense", v_))) + v_)
    t2_ = layer_norm(sd, p + "t_output.LayerNorm", lin(sd, p + "t_output.dense", gelu(lin(sd, p + "t_intermediate.dense", t_))) + t_)
    return v2_, t2_


def encoder(sd, cfg, input_ids, image_feat, image_loc, token_type_ids=None, attention_mask=None, image_attention_mask=None):
    """BertModel.forward + BertEncoder.forward: models/vilbert_dialog.py:1325-1407, :806-912.

    Returns (sequence_output_t [B,Lt,H], sequence_output_v [B,Lv,Hv])."""
    if attention_mask is None:
        attention_mask = torch.ones_like(input_ids)
    if image_attention_mask is None:
        image_attention_mask = torch.ones(image_feat.shape[0], image_feat.shape[1])
    m_t = (1.0 - attention_mask[:, None, None, :].float()) * -10000.0
    m_v = (1.0 - image_attention_mask[:, None, None, :].float()) * -10000.0
    t = text_embeddings(sd, ENC + "embeddings.", input_ids, token_type_ids, cfg.type_vocab_size)
    v = image_embeddings(sd, image_feat.float(), image_loc.float())
    v_start = t_start = 0
    for count, (v_end, t_end) in enumerate(zip(cfg.v_biattention_id, cfg.t_biattention_id)):
        for i in range(v_start, v_end):
            v = self_layer(sd, f"{ENC}encoder.v_layer.{i}.", v, m_v, cfg.v_num_attention_heads)
        for i in range(t_start, t_end):
            t = self_layer(sd, f"{ENC}encoder.layer.{i}.", t, m_t, cfg.num_attention_heads)
        v, t = connection_layer(sd, f"{ENC}encoder.c_layer.{count}.", v, m_v, t, m_t, cfg.bi_num_attention_heads)
        v_start, t_start = v_end, t_end
    for i in range(v_start, cfg.v_num_hidden_layers):
        v = self_layer(sd, f"{ENC}encoder.v_layer.{i}.", v, m_v, cfg.v_num_attention_heads)
    for i in range(t_start, cfg.num_hidden_layers):
        t = self_layer(sd, f"{ENC}encoder.layer.{i}.", t, m_t, cfg.num_attention_heads)
    return t, v


def nsp_scores(sd, seq_t, seq_v):
    """Poolers + bi_seq_relationship, fusion 'mul': models/vilbert_dialog.py:915-941, :1026-1041."""
    pt = torch.relu(lin(sd, ENC + "t_pooler.dense", seq_t[:, 0]))
    pv = torch.relu(lin(sd, ENC + "v_pooler.dense", seq_v[:, 0]))
    return lin(sd, CLS_HEAD + "bi_seq_relationship", pt * pv)


def wasted_heads(sd, seq_t, seq_v):
    """The pre-training heads the reference always evaluates (models/vilbert_dialog.py:1482) and, for enc_dec,
    throws away.  Only used so the timed CPU baseline does the same work as the reference."""
    p = CLS_HEAD + "predictions."
    h = layer_norm(sd, p + "transform.LayerNorm", gelu(lin(sd, p + "transform.dense", seq_t)))
    scores_t = F.linear(h, sd[p + "decoder.weight"]) + sd[p + "bias"]
    q = CLS_HEAD + "imagePredictions."
    hv = layer_norm(sd, q + "transform.LayerNorm", gelu(lin(sd, q + "transform.dense", seq_v)))
    scores_v = lin(sd, q + "decoder", hv)
    return scores_t, scores_v


def vlfusion(sd, seq_t, seq_v, attention_mask, image_mask):
    """models/visual_dialog_model.py:131-135 - image rows first."""
    h = torch.cat((lin(sd, "vlfusion.fc_v", seq_v), lin(sd, "vlfusion.fc_l", seq_t)), dim=1)
    m = torch.cat((image_mask.float(), attention_mask.float()), dim=1)
    return h, m


def decoder_hidden(sd, cfg, dec_ids, attention_mask, enc_hidden, enc_mask):
    """BertGenerationEncoder.forward (models/visual_dialog_decoder.py:219-323) with the HF 4.16.2 BertLayer:
    causal&pad mask (1-m)*-10000, cross mask (1-m)*-1e9, post-LN self-attn -> cross-attn -> FFN."""
    b, L = dec_ids.shape
    if attention_mask is None:
        attention_mask = torch.ones(b, L)
    i = torch.arange(L)
    causal = (i[None, None, :].repeat(b, L, 1) <= i[None, :, None]).float()
    m_self = (1.0 - causal[:, None] * attention_mask[:, None, None, :].float()) * -10000.0
    m_cross = (1.0 - enc_mask[:, None, None, :].float()) * -1e9
    h = text_embeddings(sd, DEC + "embeddings.", dec_ids, None, cfg.type_vocab_size)
    n = cfg.num_attention_heads
    for l in range(cfg.num_hidden_layers):
        p = f"{DEC}encoder.layer.{l}."
        c = _attend(lin(sd, p + "attention.self.query", h), lin(sd, p + "attention.self.key", h),
                    lin(sd, p + "attention.self.value", h), m_self, n)
        a = F.layer_norm(lin(sd, p + "attention.output.dense", c) + h, (h.shape[-1],),
                         sd[p + "attention.output.LayerNorm.weight"], sd[p + "attention.output.LayerNorm.bias"], 1e-12)
        c2 = _attend(lin(sd, p + "crossattention.self.query", a), lin(sd, p + "crossattention.self.key", enc_hidden),
                     lin(sd, p + "crossattention.self.value", enc_hidden), m_cross, n)
        b_ = F.layer_norm(lin(sd, p + "crossattention.output.dense", c2) + a, (h.shape[-1],),
                          sd[p + "crossattention.output.LayerNorm.weight"], sd[p + "crossattention.output.LayerNorm.bias"], 1e-12)
        f = lin(sd, p + "output.dense", F.gelu(lin(sd, p + "intermediate.dense", b_)))
        h = F.layer_norm(f + b_, (h.shape[-1],), sd[p + "output.LayerNorm.weight"], sd[p + "output.LayerNorm.bias"], 1e-12)
    return h


def lm_logits(sd, h):
    """models/visual_dialog_decoder.py:337-339."""
    return F.linear(h, sd[LMH + "decoder.weight"], sd[LMH + "bias"])


def decoder_forward(sd, cfg, dec_ids, attention_mask, enc_hidden, enc_mask, with_loss=False, loss_reduction=True, labels=None):
    """VisualDialogDecoder.forward (models/visual_dialog_decoder.py:33-86): label shift, **in-place** EOS->PAD on
    ``dec_ids``, logits over every position, optional CE(ignore_index=0)."""
    if labels is None:
        labels = dec_ids.new_zeros(dec_ids.shape)
        labels[:, :-1] = dec_ids[:, 1:].clone()
        dec_ids.masked_fill_(dec_ids == EOS, PAD)
    logits = lm_logits(sd, decoder_hidden(sd, cfg, dec_ids, attention_mask, enc_hidden, enc_mask))
    loss = None
    if with_loss:
        loss = F.cross_entropy(logits.reshape(-1, logits.shape[-1]), labels.view(-1), ignore_index=PAD,
                               reduction="mean" if loss_reduction else "none")
    return loss, logits


# ---------------------------------------------------------------------------------------------------------------
# utils/decoding_utils.py
# ---------------------------------------------------------------------------------------------------------------
def top_k_top_p_filter(logits, top_k=0, top_p=0.0, filter_value=-float("inf")):
    """utils/decoding_utils.py:4-35 (ties with the k-th value survive)."""
    top_k = min(top_k, logits.size(-1))
    if top_k > 0:
        remove = logits < torch.topk(logits, top_k)[0][..., -1, None]
        logits = logits.masked_fill(remove, filter_value)
    if top_p > 0.0:
        sl, si = torch.sort(logits, descending=True)
        cp = torch.cumsum(F.softmax(sl, dim=-1), dim=-1)
        rm = cp > top_p
        rm[..., 1:] = rm[..., :-1].clone()
        rm[..., 0] = 0
        logits = logits.masked_fill(rm.gather(-1, si.argsort(-1)), filter_value)
    return logits


def ngram_banned_tokens(hist, dec_prefix, n):
    """utils/decoding_utils.py:38-78 for one row: tokens w such that (last n-1 decoded tokens)+w is an n-gram of
    ``hist`` containing no special id."""
    if n <= 0:
        return []
    hist = list(hist)
    table = {}
    for i in range(len(hist) - n + 1):
        ng = tuple(hist[i:i + n])
        if set(ng) & set(SPECIAL_IDS):
            continue
        table.setdefault(ng[:-1], []).append(ng[-1])
    cur_len = len(dec_prefix)
    start = cur_len + 1 - n
    key = tuple(dec_prefix[start:cur_len]) if start >= 0 else tuple(dec_prefix[start:cur_len])
    return table.get(key, [])


def generate_greedy_or_sample(sd, enc_cfg, dec_cfg, batch, temperature=1.0, top_k=1, top_p=0.0, ngram_blocking_size=0,
                              max_new=18, generator=None, faithful_waste=False, return_logits=False):
    """EncoderDecoderModel.forward decode branch (models/visual_dialog_model.py:74-120): encoder -> fusion ->
    18 steps of full-prefix decoding with ``use_cache=False`` -> sample; then PAD after the first EOS.

    ``top_k=1`` is deterministic greedy through the reference's own sampler."""
    ids, seg, att = batch["enc_input_ids"], batch["enc_segments"], batch["enc_att_mask"]
    seq_t, seq_v = encoder(sd, enc_cfg, ids, batch["enc_image_feat"], batch["enc_image_loc"], seg, att, batch["enc_image_mask"])
    if faithful_waste:
        wasted_heads(sd, seq_t, seq_v)
    enc_h, enc_m = vlfusion(sd, seq_t, seq_v, att, batch["enc_image_mask"])
    dec_ids = batch["dec_input_ids"].clone()
    hist = (ids * (seg == 0).long()).tolist()
    out, all_logits = [], []
    for _ in range(max_new):
        _, logits = decoder_forward(sd, dec_cfg, dec_ids, None, enc_h, enc_m)
        step = logits[:, -1, :] / temperature
        if return_logits:
            all_logits.append(logits[:, -1, :].clone())
        if ngram_blocking_size > 0:
            for r in range(step.shape[0]):
                banned = ngram_banned_tokens(hist[r], dec_ids[r].tolist(), ngram_blocking_size)
                step[r, banned] = -float("inf")
        step = top_k_top_p_filter(step, top_k, top_p)
        prob = F.softmax(step, dim=-1)
        if top_k == 1:
            nxt = prob.argmax(-1, keepdim=True)
        else:
            nxt = torch.multinomial(prob, 1, generator=generator)
        dec_ids = torch.cat((dec_ids, nxt), dim=-1)
        out.append(nxt)
    seq = torch.cat(out, 1)
    mask = torch.zeros_like(seq)
    for e in (seq == EOS).nonzero(as_tuple=False):
        mask[e[0], e[1] + 1:] = 1
    seq = seq.masked_fill(mask.bool(), PAD)
    if return_logits:
        return seq, torch.stack(all_logits, 1)
    return seq


def score_answers(sd, enc_cfg, dec_cfg, batch, ans_ids, faithful_waste=False):
    """generate.py:183-209 - teacher-forced pass over the generated answer (no leading [CLS]); returns
    (loss [B, L] with reduction 'none', logits [B, L, V], ppl [B])."""
    ids, seg, att = batch["enc_input_ids"], batch["enc_segments"], batch["enc_att_mask"]
    seq_t, seq_v = encoder(sd, enc_cfg, ids, batch["enc_image_feat"], batch["enc_image_loc"], seg, att, batch["enc_image_mask"])
    if faithful_waste:
        wasted_heads(sd, seq_t, seq_v)
    enc_h, enc_m = vlfusion(sd, seq_t, seq_v, att, batch["enc_image_mask"])
    ans = ans_ids.clone()
    ans_mask = (ans != 0).float()
    loss, logits = decoder_forward(sd, dec_cfg, ans, ans_mask, enc_h, enc_m, with_loss=True, loss_reduction=False)
    ans_len = (ans != 0).sum(-1)            # counted after the in-place EOS->PAD (single-device behaviour)
    loss = loss.reshape(ans.shape[0], -1)
    ppl = torch.exp(loss.sum(-1) / ans_len)
    return loss, logits, ppl


def splice(enc_input_ids, enc_segments, enc_len, utt, segment_value, abnormal, max_seq_len=None):
    """generate.py:145-160 / :214-228 - append ``utt`` (zero-padded rows) to the history in place.
    Overflow past the row -> write a lone [SEP], length 1, mark abnormal.  Returns the per-row lengths added."""
    B, Lmax = enc_input_ids.shape
    n = (utt != 0).sum(-1)
    for i in range(B):
        s = int(enc_len[i]); e = s + int(n[i])
        if e <= Lmax:
            enc_input_ids[i, s:e] = utt[i, : int(n[i])]
        else:
            enc_input_ids[i, s:s + 1] = EOS
            n[i] = 1
            e = s + 1
            abnormal.add(i)
        if segment_value is not None:
            enc_segments[i, s:e] = segment_value
    return n
