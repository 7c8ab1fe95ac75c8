"""TEST INFRASTRUCTURE ONLY - pins the feature-record reader on the REFERENCE'S OWN reader.

Run in the development container only (needs /root/reference):  ``python -m oracle.gen_golden_io``

The reference's ``utils/image_features_reader.py`` imports ``lmdb`` and ``h5py`` (absent here) and opens an LMDB environment.
This script registers import-only stand-ins: ``h5py`` (never used by the class) and an ``lmdb`` whose ``open()`` returns an
in-memory environment over a dict with the same ``begin() -> txn.get(key)`` protocol, fills it with seeded records in the
reference's record format (pickled dict, base64 float32 arrays, key list under b'keys'), and runs the UNMODIFIED
``ImageFeaturesH5Reader`` (both the in-memory and the read-every-time branch, utils/image_features_reader.py:58-141) and
``utils/data_utils.encode_image_input`` (mask_prob = 0, :73-117) on them.  It asserts that oracle/io_reader.py reproduces those
outputs bit for bit and writes them as tests/golden/io_records.npz (global feature row, padded boxes and masks, SHA-256 of the
full padded feature block; the inputs are regenerated from the seed by tests/test_io_formats.py::make_reference_record).
"""
from __future__ import annotations

import base64
import hashlib
import os
import pickle
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"

CASES = [(36, 640, 480), (10, 500, 375), (50, 1024, 683), (1, 32, 32)]      # (num_boxes, image_w, image_h); 50 > 36 regions: truncated


def make_reference_record(case_index: int):
    """One record in the reference's LMDB value format (the dict that convert_to_lmdb.py pickles), seeded per case."""
    n, w, h = CASES[case_index]
    rng = np.random.default_rng(1000 + case_index)
    feats = np.maximum(rng.standard_normal((n, 2048)), 0).astype(np.float32)
    x1 = rng.uniform(0, 0.7 * w, n); y1 = rng.uniform(0, 0.7 * h, n)
    boxes = np.stack([x1, y1, x1 + rng.uniform(0.1 * w, 0.3 * w, n), y1 + rng.uniform(0.1 * h, 0.3 * h, n)], 1).astype(np.float32)
    cls = rng.uniform(0, 1, (n, 1601)).astype(np.float32)
    return {"image_id": 9000 + case_index, "image_h": h, "image_w": w, "num_boxes": n, "features": base64.b64encode(feats.tobytes()),
            "boxes": base64.b64encode(boxes.tobytes()), "cls_prob": base64.b64encode(cls.tobytes())}


def _install_stubs(store):
    class Txn:
        def __enter__(self): return self
        def __exit__(self, *a): return False
        def get(self, key): return store[key]

    class Env:
        def begin(self, write=False): return Txn()

    lm = types.ModuleType("lmdb")
    lm.open = lambda *a, **k: Env()
    sys.modules["lmdb"] = lm
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))


def main():
    from oracle import io_reader as RO
    store = {}
    keys = []
    for i in range(len(CASES)):
        rec = make_reference_record(i)
        k = str(rec["image_id"]).encode()
        store[k] = pickle.dumps(rec)
        keys.append(k)
    store[b"keys"] = pickle.dumps(keys)
    _install_stubs(store)
    sys.path.insert(0, REF)
    from utils.image_features_reader import ImageFeaturesH5Reader      # the reference's reader, unmodified
    from utils import data_utils as ref_du
    out = {}
    for in_memory in (False, True):
        reader = ImageFeaturesH5Reader("unused", in_memory=in_memory)
        assert len(reader) == len(CASES)
        for i in range(len(CASES)):
            rec = make_reference_record(i)
            feats, nb, loc, loc_ori, cls = reader[rec["image_id"]]
            if in_memory:                                                # second access comes from the cache
                f2, nb2, loc2, _, _ = reader[rec["image_id"]]
                assert nb2 == nb and np.array_equal(f2, feats) and np.array_equal(loc2, loc)
            rf, rnb, rloc = RO.read_record(rec)
            assert rnb == nb and np.array_equal(rf, feats) and rf.dtype == feats.dtype, f"case {i}: restatement != reference reader"
            assert np.array_equal(rloc, loc) and rloc.dtype == loc.dtype
            # utils/data_utils.py:73-117 with mask_prob = 0 (generation / evaluation never mask regions)
            target = np.zeros((nb, 1601), np.float32)
            tf, ts, tm, _, _ = ref_du.encode_image_input(feats, nb, loc, target, max_regions=37, mask_prob=0)
            qf, qs, qm = RO.encode_image_input(rf, rnb, rloc)
            assert np.array_equal(tf.numpy(), qf) and np.array_equal(ts.numpy(), qs) and np.array_equal(tm.numpy(), qm)
            out[f"c{i}_num_boxes"] = np.array(nb)
            out[f"c{i}_global_row"] = feats[0].copy()
            out[f"c{i}_loc"] = np.asarray(loc)
            out[f"c{i}_pad_loc"] = ts.numpy()
            out[f"c{i}_pad_mask"] = tm.numpy()
            out[f"c{i}_pad_feat_sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(tf.numpy()).tobytes()).digest(), dtype=np.uint8)
    path = os.path.join(ROOT, "tests", "golden", "io_records.npz")
    np.savez_compressed(path, **out)
    print(f"reference reader == oracle/io_reader.py on {len(CASES)} records (both reader branches); wrote {path} ({os.path.getsize(path)} bytes)")


if __name__ == "__main__":
    main()
