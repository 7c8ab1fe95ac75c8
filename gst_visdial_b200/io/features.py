"""Region-feature input path.

The reference keeps Faster R-CNN features in an LMDB whose values are pickled dicts with base64 strings
(utils/image_features_reader.py:55-146) and finds a record with ``list.index`` - O(N) per image - before decoding ~300 KB
of base64 per item on the dataloader workers.  At hundreds of dialogs per second per GPU that decode is the bottleneck,
so this module separates the two concerns:

* ``decode_reference_record`` / ``pad_regions`` restate the reference's per-record arithmetic (global <IMG> row = mean
  of the region features, box normalisation, area column; zero padding to ``max_regions`` with a validity mask,
  utils/data_utils.py:73-117 with mask_prob = 0) so existing LMDB dumps can be converted ONCE, offline;
* ``write_shard`` / ``FeatureShards`` store the decoded, padded tensors as flat binary files (bf16 or fp32 features)
  that are memory-mapped, looked up through a dict (O(1)), gathered into PINNED batches and handed to the GPU as they
  are - 152 KB per image in bf16;
* ``Prefetcher`` overlaps building batch i+1 with the GPU working on batch i.
"""
from __future__ import annotations

import base64
import json
import os
import queue
import threading
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

FEATURE_DIM = 2048


def decode_reference_record(item: dict) -> Tuple[np.ndarray, int, np.ndarray]:
    """One unpickled LMDB value of the reference -> (features [n+1, 2048] f32, n+1, image_location [n+1, 5] f32).
    Follows utils/image_features_reader.py:110-141: row 0 is the <IMG> token (mean feature, box [0,0,1,1,1]); the other
    rows carry x1/w, y1/h, x2/w, y2/h and the box area as a fraction of the image."""
    image_h, image_w, n = int(item["image_h"]), int(item["image_w"]), int(item["num_boxes"])
    features = np.frombuffer(base64.b64decode(item["features"]), dtype=np.float32).reshape(n, FEATURE_DIM)
    boxes = np.frombuffer(base64.b64decode(item["boxes"]), dtype=np.float32).reshape(n, 4)
    g_feat = np.sum(features, axis=0) / n
    features = np.concatenate([g_feat[None, :], features], axis=0)
    loc = np.zeros((n, 5), dtype=np.float32)
    loc[:, :4] = boxes
    loc[:, 4] = (loc[:, 3] - loc[:, 1]) * (loc[:, 2] - loc[:, 0]) / (float(image_w) * float(image_h))
    loc[:, 0] /= float(image_w)
    loc[:, 1] /= float(image_h)
    loc[:, 2] /= float(image_w)
    loc[:, 3] /= float(image_h)
    # the reference concatenates an integer row [0,0,1,1,1]; numpy promotes the result to float64 (:131-132)
    loc = np.concatenate([np.array([[0, 0, 1, 1, 1]]), loc], axis=0)
    return features, n + 1, loc


def pad_regions(features: np.ndarray, num_boxes: int, boxes: np.ndarray, max_regions: int = 37):
    """utils/data_utils.py:73-117 with mask_prob = 0: zero-pad / truncate to ``max_regions`` rows, mask = 1 on real rows.
    Returns float32 (features [R, 2048], spatials [R, 5], image_mask [R])."""
    n = min(int(num_boxes), max_regions)
    f = np.zeros((max_regions, features.shape[-1]), dtype=np.float32)
    s = np.zeros((max_regions, boxes.shape[-1]), dtype=np.float32)
    f[:n] = features[:n]
    s[:n] = boxes[:n]
    m = np.zeros(max_regions, dtype=np.float32)
    m[:n] = 1.0
    return f, s, m


# ---- shards ---------------------------------------------------------------------------------------------------------------
def _to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """fp32 -> bf16 bit patterns (uint16), round to nearest even - the same rounding the device cast applies."""
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)


def write_shard(path: str, image_ids: Sequence[int], feats: np.ndarray, locs: np.ndarray, masks: np.ndarray, dtype: str = "bf16") -> None:
    """``feats`` [N, R, 2048], ``locs`` [N, R, 5], ``masks`` [N, R] (already padded, see pad_regions) -> directory ``path``
    with index.json + feat.bin (bf16 bit patterns or fp32) + loc.bin (fp32) + mask.bin (uint8)."""
    if dtype not in ("bf16", "fp32"):
        raise ValueError("dtype must be 'bf16' or 'fp32'")
    N, R, D = feats.shape
    if len(image_ids) != N or locs.shape != (N, R, 5) or masks.shape != (N, R):
        raise ValueError("inconsistent shard arrays")
    os.makedirs(path, exist_ok=True)
    (_to_bf16_bits(feats) if dtype == "bf16" else np.ascontiguousarray(feats, dtype=np.float32)).tofile(os.path.join(path, "feat.bin"))
    np.ascontiguousarray(locs, dtype=np.float32).tofile(os.path.join(path, "loc.bin"))
    np.ascontiguousarray(masks != 0, dtype=np.uint8).tofile(os.path.join(path, "mask.bin"))
    with open(os.path.join(path, "index.json"), "w") as f:
        json.dump({"version": 1, "dtype": dtype, "count": N, "regions": R, "dim": D, "image_ids": [int(i) for i in image_ids]}, f)


class FeatureShards:
    """Memory-mapped shards with O(1) lookup by image id (the reference: ``self._image_ids.index(image_id)``,
    utils/image_features_reader.py:57)."""

    def __init__(self, paths: Iterable[str]):
        self._shards = []
        self._where: Dict[int, Tuple[int, int]] = {}
        for si, p in enumerate(paths):
            with open(os.path.join(p, "index.json")) as f:
                meta = json.load(f)
            N, R, D = meta["count"], meta["regions"], meta["dim"]
            fdt = np.uint16 if meta["dtype"] == "bf16" else np.float32
            feat = np.memmap(os.path.join(p, "feat.bin"), dtype=fdt, mode="r", shape=(N, R, D))
            loc = np.memmap(os.path.join(p, "loc.bin"), dtype=np.float32, mode="r", shape=(N, R, 5))
            mask = np.memmap(os.path.join(p, "mask.bin"), dtype=np.uint8, mode="r", shape=(N, R))
            self._shards.append((meta, feat, loc, mask))
            for row, iid in enumerate(meta["image_ids"]):
                self._where[int(iid)] = (si, row)
        if not self._shards:
            raise ValueError("no shards given")
        metas = [s[0] for s in self._shards]
        if len({(m["dtype"], m["regions"], m["dim"]) for m in metas}) != 1:
            raise ValueError("shards disagree on dtype / regions / dim")
        self.dtype, self.regions, self.dim = metas[0]["dtype"], metas[0]["regions"], metas[0]["dim"]

    def __len__(self) -> int:
        return len(self._where)

    def __contains__(self, image_id: int) -> bool:
        return int(image_id) in self._where

    def image_ids(self) -> List[int]:
        return list(self._where.keys())

    def batch(self, image_ids: Sequence[int], pin: bool = True) -> Dict[str, torch.Tensor]:
        """Gathers the rows of ``image_ids`` into (optionally pinned) host tensors with the reference dataloader's keys.
        ``enc_image_feat`` keeps the shard dtype (torch.bfloat16 or float32): ship it to the GPU as it is and widen there."""
        B = len(image_ids)
        tdt = torch.bfloat16 if self.dtype == "bf16" else torch.float32
        feat = torch.empty((B, self.regions, self.dim), dtype=tdt, pin_memory=pin)
        loc = torch.empty((B, self.regions, 5), dtype=torch.float32, pin_memory=pin)
        mask = torch.empty((B, self.regions), dtype=torch.float32, pin_memory=pin)
        fview = feat.view(torch.int16).numpy().view(np.uint16) if self.dtype == "bf16" else feat.numpy()
        lview, mview = loc.numpy(), mask.numpy()
        for i, iid in enumerate(image_ids):
            try:
                si, row = self._where[int(iid)]
            except KeyError:
                raise KeyError(f"image id {iid} is not in the shards") from None
            _, f, l, m = self._shards[si]
            fview[i] = f[row]
            lview[i] = l[row]
            mview[i] = m[row]
        return {"enc_image_feat": feat, "enc_image_loc": loc, "enc_image_mask": mask, "image_id": torch.tensor(list(map(int, image_ids)))}


class Prefetcher:
    """Builds the batches of ``id_batches`` on a background thread, ``depth`` ahead of the consumer."""

    def __init__(self, shards: FeatureShards, id_batches: Iterable[Sequence[int]], depth: int = 2, pin: bool = True):
        self._q: "queue.Queue" = queue.Queue(maxsize=max(1, depth))
        self._err: Optional[BaseException] = None

        def run():
            try:
                for ids in id_batches:
                    self._q.put(shards.batch(ids, pin=pin))
            except BaseException as e:      # surfaced to the consumer
                self._err = e
            finally:
                self._q.put(None)

        self._t = threading.Thread(target=run, daemon=True)
        self._t.start()

    def __iter__(self):
        while True:
            item = self._q.get()
            if item is None:
                if self._err is not None:
                    raise self._err
                return
            yield item


def caption_batch(captions: Sequence[Sequence[int]], max_seq_len: int = 256, max_cap_len: int = 38, cls: int = 101, sep: int = 102):
    """Text side of a generation batch from PRE-TOKENIZED captions: [CLS] caption[:38] [SEP], segment 1 on those
    positions (dataloader/dataloader_cc12m_gen.py:75-101 + utils/data_utils.py:34-71 with start_segment = 1)."""
    B = len(captions)
    ids = torch.zeros(B, max_seq_len, dtype=torch.int64)
    seg = torch.zeros(B, max_seq_len, dtype=torch.int64)
    for i, cap in enumerate(captions):
        cap = list(cap)[:max_cap_len]
        n = len(cap)
        ids[i, 0] = cls
        if n:
            ids[i, 1:1 + n] = torch.tensor(cap, dtype=torch.int64)
        ids[i, 1 + n] = sep
        seg[i, : n + 2] = 1
    return {"enc_input_ids": ids, "enc_segments": seg, "enc_att_mask": (ids != 0).float(),
            "dec_input_ids": torch.full((B, 1), cls, dtype=torch.int64), "dec_att_mask": torch.ones(B, 1)}
