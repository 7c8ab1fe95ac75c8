"""Checkpoint key layout and seeded synthetic weights for the generation hot path.

The key names and shapes reproduce what the reference modules register, so that the released
checkpoints (``ckpt['model_state_dict']``) load unchanged:

  * encoder  -> /root/reference/models/vilbert_dialog.py:298-352 (embeddings), :354-476 (text layer),
                :479-603 (image layer), :606-773 (connection layer), :915-941 (poolers),
                :1026-1063 (pre-training heads), :1409-1427 (image embeddings)
  * decoder  -> /root/reference/models/visual_dialog_decoder.py:116-131,184-205,326-343 plus the
                HuggingFace ``BertLayer`` sub-module names (attention / crossattention / intermediate / output)
  * fusion   -> /root/reference/models/visual_dialog_model.py:123-129

``oracle/gen_golden.py`` checks this spec key-for-key against a state_dict built by the reference
constructors (861 keys for the 6-layer/6-connect config).
"""
from __future__ import annotations

import hashlib
import json
import math
import os
import re
from collections import OrderedDict
from types import SimpleNamespace

import torch

CONFIG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "config")
DEFAULT_ENC_CONFIG = os.path.join(CONFIG_DIR, "bert_base_6layer_6conect_enc.json")
DEFAULT_DEC_CONFIG = os.path.join(CONFIG_DIR, "bert_base_6layer_6conect_dec.json")
TINY_ENC_CONFIG = os.path.join(CONFIG_DIR, "tiny_enc.json")
TINY_DEC_CONFIG = os.path.join(CONFIG_DIR, "tiny_dec.json")


def load_json_config(path: str) -> SimpleNamespace:
    with open(path, "r", encoding="utf-8") as f:
        d = json.load(f)
    return SimpleNamespace(**d)


def _linear(spec, prefix, out_f, in_f):
    spec[prefix + ".weight"] = (out_f, in_f)
    spec[prefix + ".bias"] = (out_f,)


def _ln(spec, prefix, n):
    spec[prefix + ".weight"] = (n,)
    spec[prefix + ".bias"] = (n,)


def _embeddings(spec, p, c):
    H = c.hidden_size
    spec[p + ".word_embeddings.weight"] = (c.vocab_size, H)
    spec[p + ".position_embeddings.weight"] = (c.max_position_embeddings, H)
    spec[p + ".token_type_embeddings.weight"] = (c.type_vocab_size, H)
    spec[p + ".token_type_embeddings_extension.weight"] = (10, H)
    spec[p + ".sep_embeddings.weight"] = (50, H)
    _ln(spec, p + ".LayerNorm", H)


def _self_layer(spec, p, H, F):
    """attention.self.{query,key,value}, attention.output, intermediate, output."""
    for n in ("query", "key", "value"):
        _linear(spec, f"{p}.attention.self.{n}", H, H)
    _linear(spec, f"{p}.attention.output.dense", H, H)
    _ln(spec, f"{p}.attention.output.LayerNorm", H)
    _linear(spec, f"{p}.intermediate.dense", F, H)
    _linear(spec, f"{p}.output.dense", H, F)
    _ln(spec, f"{p}.output.LayerNorm", H)


def encoder_spec(c, prefix="encoder.") -> "OrderedDict[str, tuple]":
    spec = OrderedDict()
    b = prefix + "bert_pretrained.bert"
    H, Hv, Hb = c.hidden_size, c.v_hidden_size, c.bi_hidden_size
    _embeddings(spec, b + ".embeddings", c)
    _linear(spec, b + ".v_embeddings.image_embeddings", Hv, c.v_feature_size)
    _linear(spec, b + ".v_embeddings.image_location_embeddings", Hv, 5)
    _ln(spec, b + ".v_embeddings.LayerNorm", Hv)
    for i in range(c.num_hidden_layers):
        _self_layer(spec, f"{b}.encoder.layer.{i}", H, c.intermediate_size)
    for i in range(c.v_num_hidden_layers):
        _self_layer(spec, f"{b}.encoder.v_layer.{i}", Hv, c.v_intermediate_size)
    for i in range(len(c.v_biattention_id)):
        p = f"{b}.encoder.c_layer.{i}"
        for n in ("query1", "key1", "value1"):
            _linear(spec, f"{p}.biattention.{n}", Hb, Hv)
        for n in ("query2", "key2", "value2"):
            _linear(spec, f"{p}.biattention.{n}", Hb, H)
        _linear(spec, f"{p}.biOutput.dense1", Hv, Hb)
        _ln(spec, f"{p}.biOutput.LayerNorm1", Hv)
        _linear(spec, f"{p}.biOutput.q_dense1", Hv, Hb)
        _linear(spec, f"{p}.biOutput.dense2", H, Hb)
        _ln(spec, f"{p}.biOutput.LayerNorm2", H)
        _linear(spec, f"{p}.biOutput.q_dense2", H, Hb)
        _linear(spec, f"{p}.v_intermediate.dense", c.v_intermediate_size, Hv)
        _linear(spec, f"{p}.v_output.dense", Hv, c.v_intermediate_size)
        _ln(spec, f"{p}.v_output.LayerNorm", Hv)
        _linear(spec, f"{p}.t_intermediate.dense", c.intermediate_size, H)
        _linear(spec, f"{p}.t_output.dense", H, c.intermediate_size)
        _ln(spec, f"{p}.t_output.LayerNorm", H)
    _linear(spec, b + ".t_pooler.dense", Hb, H)
    _linear(spec, b + ".v_pooler.dense", Hb, Hv)
    cls = prefix + "bert_pretrained.cls"
    spec[cls + ".predictions.bias"] = (c.vocab_size,)
    _linear(spec, cls + ".predictions.transform.dense", H, H)
    _ln(spec, cls + ".predictions.transform.LayerNorm", H)
    spec[cls + ".predictions.decoder.weight"] = (c.vocab_size, H)
    _linear(spec, cls + ".bi_seq_relationship", 2, Hb)
    _linear(spec, cls + ".imagePredictions.transform.dense", Hv, Hv)
    _ln(spec, cls + ".imagePredictions.transform.LayerNorm", Hv)
    _linear(spec, cls + ".imagePredictions.decoder", c.v_target_size, Hv)
    return spec


def decoder_spec(c, prefix="decoder.") -> "OrderedDict[str, tuple]":
    spec = OrderedDict()
    b = prefix + "decoder.bert"
    H = c.hidden_size
    _embeddings(spec, b + ".embeddings", c)
    for i in range(c.num_hidden_layers):
        p = f"{b}.encoder.layer.{i}"
        for n in ("query", "key", "value"):
            _linear(spec, f"{p}.attention.self.{n}", H, H)
        _linear(spec, f"{p}.attention.output.dense", H, H)
        _ln(spec, f"{p}.attention.output.LayerNorm", H)
        for n in ("query", "key", "value"):
            _linear(spec, f"{p}.crossattention.self.{n}", H, H)
        _linear(spec, f"{p}.crossattention.output.dense", H, H)
        _ln(spec, f"{p}.crossattention.output.LayerNorm", H)
        _linear(spec, f"{p}.intermediate.dense", c.intermediate_size, H)
        _linear(spec, f"{p}.output.dense", H, c.intermediate_size)
        _ln(spec, f"{p}.output.LayerNorm", H)
    lm = prefix + "decoder.lm_head"
    spec[lm + ".bias"] = (c.vocab_size,)
    spec[lm + ".decoder.weight"] = (c.vocab_size, H)
    spec[lm + ".decoder.bias"] = (c.vocab_size,)
    return spec


def fusion_spec(c, prefix="vlfusion.") -> "OrderedDict[str, tuple]":
    spec = OrderedDict()
    _linear(spec, prefix + "fc_l", c.hidden_size, c.hidden_size)
    _linear(spec, prefix + "fc_v", c.hidden_size, c.v_hidden_size)
    return spec


def model_spec(enc_cfg, dec_cfg) -> "OrderedDict[str, tuple]":
    """Flat ``EncoderDecoderModel.state_dict()`` layout (encoder.*, decoder.*, vlfusion.*)."""
    spec = OrderedDict()
    spec.update(encoder_spec(enc_cfg))
    spec.update(decoder_spec(dec_cfg))
    spec.update(fusion_spec(enc_cfg))
    return spec


# Tensors that alias one another in the reference at generation time:
#  * decoder embeddings are the encoder's module (generate.py:65)
#  * lm_head.bias is lm_head.decoder.bias (visual_dialog_decoder.py:333-335)
#  * the encoder's MLM output matrix is tied to the word embedding (vilbert_dialog.py:1436-1438, :997)
def _alias_of(name: str) -> str:
    if name.startswith("decoder.decoder.bert.embeddings."):
        return "encoder.bert_pretrained.bert.embeddings." + name[len("decoder.decoder.bert.embeddings."):]
    if name == "encoder.bert_pretrained.cls.predictions.decoder.weight":
        return "encoder.bert_pretrained.bert.embeddings.word_embeddings.weight"
    if name == "decoder.decoder.lm_head.decoder.bias":
        return "decoder.decoder.lm_head.bias"
    return name


def _seed_for(name: str, seed: int) -> int:
    h = hashlib.sha256(f"{seed}:{name}".encode()).digest()
    return int.from_bytes(h[:7], "little")


def synthetic_state_dict(enc_cfg, dec_cfg, seed: int = 0, dtype=torch.float32) -> "OrderedDict[str, torch.Tensor]":
    """Seeded random weights, one independent CPU generator per tensor (so any subset is reproducible).

    The reference init (vilbert_dialog.py:1076-1087: matrices N(0, 0.02), biases 0, LayerNorm identity) makes
    every sub-layer a small perturbation of the residual and ties the LM head to the input embedding, so greedy
    output degenerates and a missing bias / affine term is invisible.
    Here: matrices ~ N(0, 1/fan_in) (query/key 1.5x that, so attention is peaked rather than uniform),
    embedding tables ~ N(0, 0.25), every 1-D parameter ~ N(0, 0.01) (+1 for LayerNorm gains), and
    ``lm_head.decoder.weight`` independent of the word embedding.
    """
    spec = model_spec(enc_cfg, dec_cfg)
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, shape in spec.items():
        src = _alias_of(name)
        if src != name and src in out:
            out[name] = out[src]
            continue
        g = torch.Generator(device="cpu")
        g.manual_seed(_seed_for(src, seed))
        if len(shape) == 1:
            t = torch.randn(shape, generator=g, dtype=torch.float32) * 0.1
            if re.search(r"LayerNorm[12]?\.weight$", src):
                t = t + 1.0
        else:
            std = 1.0 / math.sqrt(shape[1])
            if re.search(r"\.(query|key|query1|key1|query2|key2)\.weight$", src):
                std *= 1.5  # peaked softmax: keeps the parity tests sensitive to mask / head-layout bugs
            elif src.endswith("_embeddings.weight"):
                std = 0.5
            t = torch.randn(shape, generator=g, dtype=torch.float32) * std
        out[name] = t.to(dtype)
    return out
