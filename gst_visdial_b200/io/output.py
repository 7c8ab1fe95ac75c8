"""Output path.

The reference decodes every utterance with the tokenizer inside the batch loop and keeps all dialogs in one Python list
that is written by a single ``json.dump`` at the very end (generate.py:230-258): nothing reaches the disk before the
last batch, and a crash loses the run.  Here every finished batch is appended to a JSON-lines file by a writer thread
(the GPU loop never waits for the disk), and ``jsonl_to_reference_json`` produces the reference's single-array file
(same record layout: image_id, url, caption, dialog[{question, answer, answer_ppl}]) when that is what the consumer wants.
"""
from __future__ import annotations

import json
import queue
import threading
from typing import Callable, Dict, Optional, Sequence

import torch


def strip_ids(row: Sequence[int], specials=(0, 100, 101, 102, 103)) -> list:
    """Token ids of one utterance without padding / special tokens (the id-level equivalent of
    ``tokenizer.decode(..., skip_special_tokens=True)``, generate.py:18-23)."""
    return [int(t) for t in row if int(t) not in specials]


def batch_records(image_ids, questions: torch.Tensor, answers: torch.Tensor, answer_ppl: Optional[torch.Tensor], abnormal: torch.Tensor,
                  decode: Optional[Callable[[list], str]] = None, meta: Optional[Dict[int, dict]] = None) -> list:
    """One dict per NORMAL dialog of a batch (abnormal ones - history overflow - are dropped like generate.py:236-237).
    ``decode`` maps a list of token ids to text (e.g. ``lambda ids: tokenizer.decode(ids, skip_special_tokens=True)``);
    without it the records carry the stripped token ids.  ``meta[image_id]`` may add ``url`` / ``caption``."""
    q, a = questions.cpu().tolist(), answers.cpu().tolist()
    ppl = answer_ppl.float().cpu().tolist() if answer_ppl is not None else None
    bad = abnormal.cpu().tolist()
    out = []
    for i, iid in enumerate(int(x) for x in (image_ids.tolist() if hasattr(image_ids, "tolist") else image_ids)):
        if bad[i]:
            continue
        rec = {"image_id": iid}
        if meta and iid in meta:
            rec.update({k: meta[iid][k] for k in ("url", "caption") if k in meta[iid]})
        turns = []
        for r in range(len(q[i])):
            qi, ai = strip_ids(q[i][r]), strip_ids(a[i][r])
            turn = {"question": decode(qi) if decode else qi, "answer": decode(ai) if decode else ai}
            if ppl is not None:
                turn["answer_ppl"] = ppl[i][r]
            turns.append(turn)
        rec["dialog"] = turns
        out.append(rec)
    return out


class JsonlWriter:
    """Appends records to ``path`` (one JSON object per line) from a background thread; ``close()`` drains and joins."""

    def __init__(self, path: str, queue_depth: int = 64):
        self.path = path
        self._q: "queue.Queue" = queue.Queue(maxsize=queue_depth)
        self._err: Optional[BaseException] = None
        self.count = 0
        self._f = open(path, "w")

        def run():
            try:
                while True:
                    recs = self._q.get()
                    if recs is None:
                        return
                    for r in recs:
                        self._f.write(json.dumps(r) + "\n")
                    self._f.flush()
                    self.count += len(recs)
            except BaseException as e:
                self._err = e

        self._t = threading.Thread(target=run, daemon=True)
        self._t.start()

    def _put(self, item) -> None:
        """Bounded put that cannot deadlock on a dead writer thread (full disk, ...): the error / liveness check is repeated
        while the queue is full instead of only once before a blocking put."""
        while True:
            if self._err is not None:
                raise self._err
            if not self._t.is_alive():
                raise RuntimeError(f"JsonlWriter: writer thread for {self.path} is gone")
            try:
                self._q.put(item, timeout=0.5)
                return
            except queue.Full:
                continue

    def write(self, records: list) -> None:
        self._put(list(records))

    def close(self) -> int:
        try:
            if self._t.is_alive() and self._err is None:
                self._put(None)
            self._t.join()
        finally:
            self._f.close()
        if self._err is not None:
            raise self._err
        return self.count

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def jsonl_to_reference_json(jsonl_path: str, json_path: str) -> int:
    """JSON-lines -> the single JSON array of generate.py:258, streamed (constant memory)."""
    n = 0
    with open(jsonl_path) as src, open(json_path, "w") as dst:
        dst.write("[")
        for line in src:
            line = line.strip()
            if not line:
                continue
            dst.write((", " if n else "") + line)
            n += 1
        dst.write("]")
    return n
