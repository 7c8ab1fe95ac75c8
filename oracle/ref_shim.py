"""TEST INFRASTRUCTURE ONLY - imports the reference's own modules from /root/reference on CPU.

Only usable in the development container (the GPU box has no /root/reference). It is used by
``oracle/gen_golden.py`` to (1) check ``gst_visdial_b200.weights.model_spec`` against the reference's
``state_dict`` and (2) run the reference's ``EncoderDecoderModel`` on seeded weights/inputs to
produce the golden vectors under ``tests/golden/`` and to pin ``oracle/restatement.py``.

The compatibility shim follows SURVEY.md section 8c / appendix D:
  1. stub ``pytorch_transformers.modeling_bert`` / ``pytorch_pretrained_bert.file_utils`` (import-only in
     models/vilbert_dialog.py:34,37),
  2. ignore ``.to(cuda)`` when CUDA is absent (models/vilbert_dialog.py:312 moves an unused buffer),
  3. construct ``BertForMultiModalPreTraining`` / ``BertForSequenceGeneration`` directly instead of through
     ``from_pretrained('bert-base-uncased')`` (models/visual_dialog_encoder.py:17, visual_dialog_decoder.py:24-27),
  4. restore the transformers-4.16.2 helper semantics the decoder relies on
     (models/visual_dialog_decoder.py:274,285,294): causal&pad mask ``(1-m)*-10000``, cross mask ``(1-m)*-1e9``,
     ``get_head_mask`` -> ``[None]*n``; eager attention.
"""
from __future__ import annotations

import sys
import types

import torch
from torch import nn

REFERENCE_ROOT = "/root/reference"
_installed = False


def install():
    global _installed
    if _installed:
        return
    pt = types.ModuleType("pytorch_transformers")
    ptm = types.ModuleType("pytorch_transformers.modeling_bert")
    ptm.BertEmbeddings = object
    pt.modeling_bert = ptm
    pp = types.ModuleType("pytorch_pretrained_bert")
    ppf = types.ModuleType("pytorch_pretrained_bert.file_utils")
    ppf.cached_path = lambda *a, **k: None
    pp.file_utils = ppf
    sys.modules.setdefault("pytorch_transformers", pt)
    sys.modules.setdefault("pytorch_transformers.modeling_bert", ptm)
    sys.modules.setdefault("pytorch_pretrained_bert", pp)
    sys.modules.setdefault("pytorch_pretrained_bert.file_utils", ppf)

    if not torch.cuda.is_available():
        _to = torch.Tensor.to

        def to(self, *a, **k):
            if a and isinstance(a[0], torch.device) and a[0].type == "cuda":
                return self
            return _to(self, *a, **k)

        torch.Tensor.to = to

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def build_reference_model(enc_config_path: str, dec_config_path: str, model: str = "enc_dec_a", mode: str = "cc12m_gen"):
    """Returns (EncoderDecoderModel, params) built from the reference classes, eval mode, CPU fp32."""
    install()
    # our package also has a top-level-looking "models"/"utils"; make sure the reference's win here
    for m in [k for k in sys.modules if k == "models" or k.startswith("models.") or k == "utils" or k.startswith("utils.")]:
        if not getattr(sys.modules[m], "__file__", "").startswith(REFERENCE_ROOT):
            del sys.modules[m]
    import models.vilbert_dialog as vd
    import models.visual_dialog_decoder as dd
    import models.visual_dialog_encoder as ve
    import models.visual_dialog_model as vm
    from transformers import BertGenerationConfig

    def ext(self, mask, shape, device=None):
        b, L = shape
        i = torch.arange(L)
        causal = (i[None, None, :].repeat(b, L, 1) <= i[None, :, None]).to(mask.dtype)
        return (1.0 - (causal[:, None] * mask[:, None, None, :]).float()) * -10000.0

    dd.BertGenerationEncoder.get_extended_attention_mask = ext
    dd.BertGenerationEncoder.invert_attention_mask = lambda self, m: (1.0 - m[:, None, None, :].float()) * -1e9
    dd.BertGenerationEncoder.get_head_mask = lambda self, hm, n, *a: [None] * n

    params = {"model_enc_config": enc_config_path, "model_dec_config": dec_config_path,
              "gpu_ids": [0], "model": model, "mode": mode}
    enc = ve.VisualDialogEncoder.__new__(ve.VisualDialogEncoder)
    nn.Module.__init__(enc)
    enc.params = params
    enc.model_arch = params["model"]
    enc.config = vd.BertConfig.from_json_file(enc_config_path)
    enc.config.__dict__.update(cur_device=0, model_arch=params["model"], mode=params["mode"])
    enc.bert_pretrained = vd.BertForMultiModalPreTraining(enc.config)
    if "enc_only" in model:
        return enc.eval(), params
    dec = dd.VisualDialogDecoder.__new__(dd.VisualDialogDecoder)
    nn.Module.__init__(dec)
    dec.params = params
    dec.config = BertGenerationConfig.from_json_file(dec_config_path)
    dec.config.__dict__["cur_device"] = 0
    dec.config._attn_implementation = "eager"
    dec.decoder = dd.BertForSequenceGeneration(dec.config)
    dec.decoder.bert.embeddings = enc.bert_pretrained.bert.embeddings  # generate.py:65
    model_ = vm.EncoderDecoderModel(params, enc, dec).eval()
    return model_, params
