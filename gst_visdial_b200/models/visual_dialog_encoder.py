"""Drop-in ``VisualDialogEncoder`` (reference: models/visual_dialog_encoder.py:7-76) backed by the CUDA engine."""
from __future__ import annotations

import copy
import json

import torch
from torch import nn

from .. import weights as W
from ..engine import Engine
from ._tree import WeightVersion, build_tree


class BertConfig(object):
    """The slice of the reference's BertConfig (models/vilbert_dialog.py:131-270) callers rely on: attribute access,
    ``from_json_file`` and ``to_dict``."""

    def __init__(self, **kw):
        self.__dict__.update(dict(fast_mode=False, fixed_v_layer=0, fixed_t_layer=0, in_batch_pairs=False,
                                  fusion_method="mul", intra_gate=False, with_coattention=True, predict_feature=False))
        self.__dict__.update(kw)

    @classmethod
    def from_dict(cls, d):
        return cls(**d)

    @classmethod
    def from_json_file(cls, path):
        with open(path, "r", encoding="utf-8") as f:
            return cls.from_dict(json.load(f))

    def to_dict(self):
        return copy.deepcopy(self.__dict__)

    def to_json_string(self):
        return json.dumps(self.to_dict(), indent=2, sort_keys=True) + "\n"

    def __repr__(self):
        return self.to_json_string()


class _EngineOwner:
    """Shared machinery: one Engine per device, weights re-synchronised when the parameter version changes."""

    def _init_engine_state(self, params):
        self._engines = {}                    # device index -> [Engine, synced weight version]
        self._version = WeightVersion()
        self._engine_opts = dict(dtype=params.get("compute_dtype", "bf16"), max_batch=int(params.get("engine_max_batch", 64)),
                                 max_beams=int(params.get("engine_max_beams", 5)), max_text_len=int(params.get("max_seq_len", 256)),
                                 max_dec_len=int(params.get("max_utt_len", 25)), flags=int(params.get("engine_flags", 0)))

    def _apply(self, fn, *a, **k):            # .to() / .cuda() / .float() may move or change the parameters
        out = super()._apply(fn, *a, **k)
        if hasattr(self, "_version"):
            self._version.bump()
        return out

    def _load_from_state_dict(self, *a, **k):
        super()._load_from_state_dict(*a, **k)
        if hasattr(self, "_version"):
            self._version.bump()

    def _engine_for(self, device: torch.device, enc_cfg, dec_cfg, state_prefix: str) -> Engine:
        if device.type != "cuda":
            raise RuntimeError("gst_visdial_b200 modules run on CUDA tensors only (no CPU fallback); move the model and inputs to a B200")
        idx = device.index if device.index is not None else torch.cuda.current_device()
        slot = self._engines.get(idx)
        if slot is None:
            slot = [Engine(enc_cfg, dec_cfg, device=idx, **self._engine_opts), 0]
            self._engines[idx] = slot
        if slot[1] != self._version.value:
            sd = {k: v for k, v in self.state_dict().items()}
            slot[0].load_state_dict(sd, prefix=state_prefix, strict=True)
            slot[1] = self._version.value
        return slot[0]

    def mark_weights_changed(self):
        """Call after modifying parameters in place outside load_state_dict()/.to()."""
        self._version.bump()


class VisualDialogEncoder(_EngineOwner, nn.Module):

    def __init__(self, params):
        nn.Module.__init__(self)
        self.params = params
        self.config = BertConfig.from_json_file(params['model_enc_config'])
        self.config.__dict__['cur_device'] = params["gpu_ids"][0]
        self.config.__dict__['model_arch'] = params['model']
        self.config.__dict__['mode'] = params['mode']
        self.model_arch = params['model']
        self.bert_pretrained = build_tree(W.encoder_spec(self.config, prefix="encoder."), "encoder.bert_pretrained.")
        self._init_engine_state(params)
        self._owner = None          # set by EncoderDecoderModel: the engine then lives there

    def forward(
        self,
        input_ids,
        image_feat,
        image_loc,
        sep_indices=None,
        token_type_ids=None,
        attention_mask=None,
        masked_lm_labels=None,
        next_sentence_label=None,
        image_attention_mask=None,
        image_label=None,
        image_target=None
    ):
        """Returns the reference's 7-tuple (models/visual_dialog_encoder.py:33-76).  Inference branches only:
        enc_dec -> (.., enc_hidden_t, enc_hidden_v); enc_only -> (.., seq_relationship_score, ..).  The MLM / image heads
        the reference evaluates and discards (models/vilbert_dialog.py:1482) are not computed: prediction_scores_t is None."""
        if 'train' in self.params['mode'] and 'enc_dec' not in self.model_arch:
            raise NotImplementedError("training losses of the enc_only model are outside the generation hot path")
        if self._owner is not None:
            eng = self._owner._engine(input_ids.device)
        else:
            eng = self._engine_for(input_ids.device, self.config, None, "encoder.")
        enc_dec = 'enc_dec' in self.model_arch
        out = eng.encode(input_ids, image_feat, image_loc, token_type_ids, attention_mask, image_attention_mask,
                         want_t=enc_dec, want_v=enc_dec, want_nsp=not enc_dec)
        return (None, None, None, out["nsp"], None, out["seq_t"], out["seq_v"])
