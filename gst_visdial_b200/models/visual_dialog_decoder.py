"""Drop-in ``VisualDialogDecoder`` (reference: models/visual_dialog_decoder.py:18-86) backed by the CUDA engine."""
from __future__ import annotations

import json
from dataclasses import dataclass
from typing import Optional, Tuple

import torch
from torch import nn

from .. import weights as W
from ._tree import build_tree


class BertGenerationConfig(object):
    """Attribute bag with the fields of config/bert_base_6layer_6conect_dec.json (stands in for
    transformers.BertGenerationConfig, models/visual_dialog_decoder.py:22)."""

    def __init__(self, **kw):
        self.__dict__.update(dict(bos_token_id=101, eos_token_id=102, pad_token_id=0, is_decoder=True, add_cross_attention=True,
                                  use_cache=False, layer_norm_eps=1e-12))
        self.__dict__.update(kw)

    @classmethod
    def from_json_file(cls, path):
        with open(path, "r", encoding="utf-8") as f:
            return cls(**json.load(f))

    def to_dict(self):
        return dict(self.__dict__)


@dataclass
class Seq2SeqLMOutput:
    loss: Optional[torch.Tensor] = None
    logits: Optional[torch.Tensor] = None
    past_key_values: Optional[Tuple] = None
    decoder_hidden_states: Optional[Tuple] = None
    decoder_attentions: Optional[Tuple] = None
    cross_attentions: Optional[Tuple] = None


class _DecoderBody(nn.Module):
    """``decoder`` attribute of the reference (BertForSequenceGeneration): holds ``bert`` and ``lm_head`` sub-trees."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        tree = build_tree(W.decoder_spec(config, prefix="decoder."), "decoder.decoder.")
        self.bert = tree._modules["bert"]
        self.lm_head = tree._modules["lm_head"]
        # lm_head.bias and lm_head.decoder.bias are one tensor (models/visual_dialog_decoder.py:333-335)
        self.lm_head._modules["decoder"]._parameters["bias"] = self.lm_head._parameters["bias"]

    def _reorder_cache(self, past, beam_idx):
        """models/visual_dialog_decoder.py:177-181."""
        return tuple(tuple(p.index_select(0, beam_idx) for p in layer) for layer in past)


class VisualDialogDecoder(nn.Module):
    def __init__(self, params):
        super().__init__()
        self.params = params
        self.config = BertGenerationConfig.from_json_file(params['model_dec_config'])
        self.config.__dict__['cur_device'] = params["gpu_ids"][0]
        self.decoder = _DecoderBody(self.config)
        self._owner = None

    def _reorder_cache(self, past, beam_idx):
        return self.decoder._reorder_cache(past, beam_idx)

    def forward(
        self,
        decoder_input_ids=None,
        attention_mask=None,
        encoder_hidden_states=None,
        encoder_attention_mask=None,
        labels=None,
        use_cache=False,
        output_attentions=False,
        output_hidden_states=False,
        return_dict=True,
        loss_reduction=True
    ):
        """Teacher-forced pass over ``decoder_input_ids`` [B, L] against caller-supplied encoder states (the reference
        signature, models/visual_dialog_decoder.py:33-86).  As in the reference, when ``labels`` is None the ids are
        shifted into labels and ``decoder_input_ids`` is modified IN PLACE ([SEP] -> [PAD])."""
        if self._owner is None:
            raise RuntimeError("VisualDialogDecoder must be wrapped in EncoderDecoderModel (it shares the engine and the embeddings)")
        eng = self._owner._engine(decoder_input_ids.device)
        B, Le = encoder_hidden_states.shape[0], encoder_hidden_states.shape[1]
        eng.prefill_cross(B, Le, encoder_hidden_states, encoder_attention_mask)
        return self._owner._score(eng, decoder_input_ids, attention_mask, labels, loss_reduction, want_logits=True)
