"""Drop-in ``EncoderDecoderModel`` / ``VLFusion`` (reference: models/visual_dialog_model.py:8-135) on the CUDA engine."""
from __future__ import annotations

import torch
from torch import nn

from .. import weights as W
from ._tree import build_tree
from .visual_dialog_decoder import Seq2SeqLMOutput
from .visual_dialog_encoder import _EngineOwner


class VLFusion(nn.Module):
    """Parameters of models/visual_dialog_model.py:123-129; the projection itself runs inside the engine."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        tree = build_tree(W.fusion_spec(config, prefix="vlfusion."), "vlfusion.")
        self.fc_l = tree._modules["fc_l"]
        self.fc_v = tree._modules["fc_v"]


class EncoderDecoderModel(_EngineOwner, nn.Module):
    """Convenience wrapper module, wrapping Encoder and Decoder modules (same contract as the reference).

    Extra keyword arguments (all default to the reference behaviour):
      num_beams=1        >1 switches the token selection to beam search (contract: oracle/beam.py)
      seed=None          sampling seed (a per-call counter when None)
      enc_valid_len=None   upper bound (python int) on the number of leading history positions that hold a token in ANY row.
                         The encoder, the cross-K/V prefill and the cross-attention then run on ceil32(bound) text positions
                         instead of max_seq_len: positions past the last token are padding whose keys carry zero weight and
                         whose rows nobody reads, so every valid position is bit-identical.
      reuse_encoder=False  reuse the encoder/cross-KV state left by the previous call (the caller guarantees the
                         encoder inputs are unchanged, e.g. generate.py's perplexity pass right after the answer pass)
    """

    def __init__(self, params, encoder, decoder):
        nn.Module.__init__(self)
        self.params = params
        self.encoder = encoder
        self.decoder = decoder
        self.vlfusion = VLFusion(encoder.config)
        self.config = {"encoder": encoder.config.to_dict(), "decoder": decoder.config.to_dict()}
        self._init_engine_state(params)
        self._calls = 0
        encoder._version = self._version      # one change counter for the whole model
        # plain attributes, not sub-modules: a registered back-reference would make the module tree cyclic
        object.__setattr__(encoder, "_owner", self)
        object.__setattr__(decoder, "_owner", self)

    # ---- engine plumbing --------------------------------------------------------------------------------------
    def _engine(self, device):
        return self._engine_for(device, self.encoder.config, self.decoder.config, "")

    def _score(self, eng, dec_input_ids, dec_attention_mask, dec_labels, loss_reduction, want_logits=True):
        if not dec_input_ids.is_contiguous() or dec_input_ids.dtype != torch.int64:
            raise ValueError("dec_input_ids must be a contiguous int64 tensor (it is modified in place like the reference does)")
        loss, logits = eng.score(dec_input_ids, dec_attention_mask, dec_labels, want_logits=want_logits)
        mode = self.params['mode']
        if 'train' in mode or 'eval' in mode:
            if loss_reduction:
                # CrossEntropyLoss(ignore_index=0): mean over the non-ignored targets
                if dec_labels is None:
                    n = (loss != 0).sum().clamp(min=1)   # ignored positions carry an exact 0
                else:
                    n = (dec_labels != 0).sum().clamp(min=1)
                loss = loss.sum() / n
            else:
                loss = loss.reshape(-1)
        else:
            loss = None
        return Seq2SeqLMOutput(loss=loss, logits=logits)

    def forward(
        self,
        enc_image_features=None,
        enc_image_spatials=None,
        enc_image_mask=None,
        enc_image_target=None,
        enc_image_label=None,
        enc_next_sentence_labels=None,
        enc_input_ids=None,
        enc_segments=None,
        enc_sep_indices=None,
        enc_mlm_labels=None,
        enc_attention_mask=None,
        dec_input_ids=None,
        dec_attention_mask=None,
        dec_labels=None,
        loss_reduction=True,
        **decoding_kwargs
    ):
        eng = self._engine(enc_input_ids.device)
        B = enc_input_ids.shape[0]
        if not decoding_kwargs.get("reuse_encoder", False):
            hint = decoding_kwargs.get("enc_valid_len")
            ids_e, seg_e, att_e = enc_input_ids, enc_segments, enc_attention_mask
            if hint is not None:
                Lt = enc_input_ids.shape[1]
                Lt_eff = min(Lt, max(32, (int(hint) + 31) // 32 * 32))
                if Lt_eff < Lt:
                    ids_e, seg_e, att_e = ids_e[:, :Lt_eff], seg_e[:, :Lt_eff], (att_e[:, :Lt_eff] if att_e is not None else None)
            enc = eng.encode(ids_e, enc_image_features, enc_image_spatials, seg_e, att_e, enc_image_mask)
            eng.prefill_cross(B, enc["Le"])

        if 'train' in self.params['mode'] or 'eval' in self.params['mode']:
            out = self._score(eng, dec_input_ids, dec_attention_mask, dec_labels, loss_reduction,
                              want_logits=decoding_kwargs.get("want_logits", True))
            return out.loss, out.logits

        # decode mode (models/visual_dialog_model.py:74-120): 18 new tokens, PAD after the first [SEP]
        self._calls += 1
        seed = decoding_kwargs.get("seed")
        return eng.generate(
            B,
            num_beams=int(decoding_kwargs.get("num_beams", 1)),
            temperature=decoding_kwargs['temperature'],
            top_k=decoding_kwargs['top_k'],
            top_p=decoding_kwargs['top_p'],
            ngram_blocking_size=decoding_kwargs['ngram_blocking_size'],
            seed=self._calls if seed is None else seed,
            hist_ids=enc_input_ids,
            hist_segments=enc_segments,
        )
