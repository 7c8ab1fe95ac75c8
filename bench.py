#!/usr/bin/env python
"""bench.py - synthetic dialogs/sec of the generation hot path (BASELINE.json metric) on N B200s of one node.

Workload (BASELINE.json configs[1]): teacher enc_dec_a, beam-5 answer generation, batch 64 images per GPU, 10 rounds,
bf16, random-init weights of the 12+6+6-layer ViLBERT encoder / 12-layer decoder, synthetic 37x2048 region features.
One "step" = one batch of 64 complete 10-round dialogs per GPU: per round a synthetic 6-12 token question is spliced into
the history (generate.py:148-160), the encoder + cross-KV prefill + 18 beam-search decode steps run, and the answer is
spliced back (generate.py:214-228).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this framework (N > 1: launched by torchrun)
    python bench.py --impl reference [...]                          # the reference's algorithm on the host CPU cores

Prints ONE JSON line (rank 0).  `value` = device-resident inputs; `e2e` = pinned host inputs copied in and token ids copied
out inside the timed region, through the reference-shaped module API.  `roofline` is measured live with CUDA events around
the tcgen05 GEMM launches of the timed steps.  `cpu_baseline` times oracle/ (the CPU restatement of the reference) on a
bounded sample on this host - bench.py is one of the few places allowed to execute oracle/.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "synthetic dialogs/sec (10 rounds, beam 5)"
UNIT = "dialogs/s"
TF_PER_DIALOG = 1.080            # 10 * (76.66 + 8.30 + 5 * 4.61) GFLOP


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="images per GPU")
    ap.add_argument("--rounds", type=int, default=10)
    ap.add_argument("--beams", type=int, default=5)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-trim", action="store_true",
                    help="run the encoder on all 256 history positions instead of ceil32(longest history) (A/B aid; results are identical)")
    ap.add_argument("--streams", type=int, default=3,
                    help="independent batches in flight per GPU (each on its own CUDA stream and engine context); 1 = strictly serial")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget-s", type=float, default=150.0, help="wall-clock cap of the reference arm")
    return ap.parse_args()


def workload_name(a):
    return (f"configs[1]: teacher enc_dec_a beam-{a.beams} answer generation, batch {a.batch}/GPU, {a.rounds} rounds, "
            f"{a.dtype}, 36+1 x 2048 synthetic region feats, 18 new tokens/answer")


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return d, "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                                       str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for n, v in zip(names, r[4:8]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference_round(enc_cfg, dec_cfg, sd, beams, threads):
    """One image, one round of the workload with the reference's algorithm (no KV cache: the whole prefix and the
    cross-attention K/V are recomputed every step, the discarded pre-training heads are evaluated like
    models/vilbert_dialog.py:1482 does) - oracle/beam.py over oracle/restatement.py, fp32, all host threads."""
    from gst_visdial_b200 import synthetic as S
    from oracle import beam as OB
    from oracle import restatement as R
    torch.set_num_threads(threads)
    b = S.synthetic_batch(0, 1, vocab_size=enc_cfg.vocab_size, v_feature_size=enc_cfg.v_feature_size)
    ids, seg = b["enc_input_ids"], b["enc_segments"]
    enc_len = (ids != 0).sum(-1)
    q = S.synthetic_utterance(0, 0, enc_cfg.vocab_size).unsqueeze(0)
    R.splice(ids, seg, enc_len, q, None, set())
    b["enc_att_mask"] = (ids != 0).float()
    t0 = time.perf_counter()
    with torch.no_grad():
        seq_t, seq_v = R.encoder(sd, enc_cfg, ids, b["enc_image_feat"], b["enc_image_loc"], seg, b["enc_att_mask"], b["enc_image_mask"])
        R.wasted_heads(sd, seq_t, seq_v)
        OB.beam_search(sd, enc_cfg, dec_cfg, b, num_beams=beams, precomputed_encoder=(seq_t, seq_v))
    return time.perf_counter() - t0


def run_reference(a, rank, world):
    """--impl reference: the reference's CPU path (oracle port; the Python reference cannot travel to this box)."""
    if rank != 0:
        return
    from gst_visdial_b200 import weights as W
    enc_cfg, dec_cfg = W.load_json_config(W.DEFAULT_ENC_CONFIG), W.load_json_config(W.DEFAULT_DEC_CONFIG)
    sd = W.synthetic_state_dict(enc_cfg, dec_cfg, seed=0)
    threads = os.cpu_count() or 1
    times, t_start = [], time.perf_counter()
    total = a.warmup + a.steps
    done_w = 0
    for i in range(total):
        t = cpu_reference_round(enc_cfg, dec_cfg, sd, a.beams, threads)
        if i >= a.warmup or (time.perf_counter() - t_start) > a.cpu_budget_s:
            times.append(t)
        else:
            done_w += 1
        if (time.perf_counter() - t_start) > a.cpu_budget_s and times:
            break
    t_round = statistics.mean(times)
    value = 1.0 / (a.rounds * t_round)
    sample = (f"1 image x 1 round (of {a.rounds}; per-round work is independent of the round index because the text is padded "
              f"to 256), beam {a.beams}, fp32, scaled to a {a.rounds}-round dialog; {len(times)} timed repeats")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": len(times), "warmup": done_w,
        "ms_per_step": t_round * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": workload_name(a)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


# ---------------------------------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def _quiet_stdout():
    """Everything that libraries print to fd 1 (e.g. NCCL's version banner) goes to stderr; the ONE JSON line is written to
    the original stdout by _emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


def main():
    _quiet_stdout()
    a = parse()
    from gst_visdial_b200 import dist as D
    if a.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        run_reference(a, rank, int(os.environ.get("WORLD_SIZE", "1")))
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200; there is no CPU fallback (use --impl reference for the CPU arm)")
    rank, world, local = D.init_from_env("nccl")
    if world != a.gpus and world > 1:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)

    from gst_visdial_b200 import _lib
    from gst_visdial_b200 import synthetic as S
    from gst_visdial_b200 import weights as W
    from gst_visdial_b200.dialog import generate_dialogs
    from gst_visdial_b200.models.visual_dialog_decoder import VisualDialogDecoder
    from gst_visdial_b200.models.visual_dialog_encoder import VisualDialogEncoder
    from gst_visdial_b200.models.visual_dialog_model import EncoderDecoderModel

    enc_cfg = W.load_json_config(W.DEFAULT_ENC_CONFIG)
    dec_cfg = W.load_json_config(W.DEFAULT_DEC_CONFIG)
    sd = W.synthetic_state_dict(enc_cfg, dec_cfg, seed=0)
    B = a.batch
    S_ = max(1, min(a.streams, a.steps))
    akw = dict(temperature=1.0, top_k=1, top_p=0.0, ngram_blocking_size=0, num_beams=a.beams)
    counts = [B] * world

    # One model (= one engine context: weights, workspace, KV cache, graphs) and one CUDA stream per in-flight batch.
    # Decode steps are latency-bound chains of small kernels; independent batches on different streams fill the SMs
    # those chains leave idle.  Every forward call still sees a batch of 64 images.
    slots = []
    for si in range(S_):
        params = {"model_enc_config": W.DEFAULT_ENC_CONFIG, "model_dec_config": W.DEFAULT_DEC_CONFIG, "gpu_ids": [local], "model": "enc_dec_a",
                  "mode": "cc12m_gen", "compute_dtype": a.dtype, "engine_max_batch": a.batch, "engine_max_beams": max(a.beams, 1),
                  "engine_flags": 0}
        enc, dec = VisualDialogEncoder(params), VisualDialogDecoder(params)
        dec.decoder.bert.embeddings = enc.bert_pretrained.bert.embeddings
        model = EncoderDecoderModel(params, enc, dec)
        model.load_state_dict(sd)
        model.to(dev).eval()
        eng = model._engine(dev)
        start = (rank * S_ + si) * B                     # weak scaling: every rank / stream owns its own images
        host = S.synthetic_batch(start, B, vocab_size=enc_cfg.vocab_size, v_feature_size=enc_cfg.v_feature_size)
        questions = torch.stack([torch.stack([S.synthetic_utterance(start + i, r, enc_cfg.vocab_size) for r in range(a.rounds)])
                                 for i in range(B)])
        host = {k: v.pin_memory() for k, v in host.items()}
        questions_h = questions.pin_memory()
        slots.append(dict(model=model, eng=eng, stream=torch.cuda.Stream(device=dev), host=host, questions_h=questions_h,
                          dev_batch={k: v.to(dev) for k, v in host.items()}, questions_d=questions_h.to(dev)))
    host, questions_h = slots[0]["host"], slots[0]["questions_h"]

    def step(i, from_host, to_host):
        sl = slots[i % S_]
        with torch.cuda.stream(sl["stream"]):
            batch, ques = (sl["host"], sl["questions_h"]) if from_host else (sl["dev_batch"], sl["questions_d"])
            res = generate_dialogs(sl["model"], batch, questions=ques, num_rounds=a.rounds, a_kwargs=akw, with_ppl=False, device=dev,
                                   trim_history=not a.no_trim)
            ans, abn = res.answers, res.abnormal
            if to_host:                                 # device -> pinned host, asynchronous on this slot's stream
                if "out_ans" not in sl:
                    sl["out_ans"] = torch.empty(ans.shape, dtype=ans.dtype).pin_memory()
                    sl["out_abn"] = torch.empty(abn.shape, dtype=abn.dtype).pin_memory()
                sl["out_ans"].copy_(ans, non_blocking=True)
                sl["out_abn"].copy_(abn, non_blocking=True)
                return sl["out_ans"], sl["out_abn"]
            return ans, abn

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)

    host_ms = [0.0, 0.0]

    def launches_total():
        return sum(sl["eng"].launch_count for sl in slots)

    def timed(from_host, to_host, steps, profile):
        barrier()
        if profile:
            for sl in slots:
                sl["eng"].profile_gemm(True, 1024)
        l0 = launches_total()
        cur = torch.cuda.current_stream(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(cur)
        for sl in slots:
            sl["stream"].wait_event(e0)
        t_host, c_host = time.perf_counter(), time.process_time()
        # one submitting host thread per stream: a single thread blocks on the launch queue of one stream (back-pressure of
        # ~2.7k launches per round) and starves the others; the C ABI / torch calls release the GIL
        outs = [None] * S_

        def worker(si):
            torch.cuda.set_device(dev)
            for i in range(si, steps, S_):
                outs[si] = step(i, from_host, to_host)

        if S_ == 1:
            worker(0)
        else:
            threads = [threading.Thread(target=worker, args=(si,)) for si in range(S_)]
            for th in threads:
                th.start()
            for th in threads:
                th.join()
        out = outs[0]
        if world > 1:
            # The only collective of the job: ONE final all_gather (NCCL over NVLink) of the generated token ids and flags
            # of this rank's last batch per stream, issued from the main thread after every stream has drained (collectives
            # issued concurrently from several host threads would not be ordered consistently across ranks).
            for sl in slots:
                cur.wait_stream(sl["stream"])
            mine = [o for o in outs if o is not None]
            ans = torch.cat([o[0].to(dev) for o in mine], 0)
            abn = torch.cat([o[1].to(dev) for o in mine], 0)
            g_ans, g_abn = D.gather_results([ans, abn], [ans.shape[0]] * world)
            out = (g_ans, g_abn)
        host_ms[0] = (time.perf_counter() - t_host) * 1e3 / steps      # wall time the submitting threads needed per step
        host_ms[1] = (time.process_time() - c_host) * 1e3 / steps      # CPU time (all threads of this process) per step
        for sl in slots:
            cur.wait_stream(sl["stream"])
        e1.record(cur)
        barrier()
        ms = e0.elapsed_time(e1)
        prof = None
        if profile:
            parts = [sl["eng"].profile_read() for sl in slots]
            prof = {k: sum(p_[k] for p_ in parts) for k in parts[0]}
            for sl in slots:
                sl["eng"].profile_gemm(False, 0)
        launches = launches_total() - l0
        t = torch.tensor([ms], device=dev)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item()), launches, prof, out

    n_warm = max(a.warmup, 3, S_)                      # every slot captures its decode graph during warm-up
    for i in range(n_warm):
        step(i, False, False)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    ms, launches, prof, out = timed(False, False, a.steps, S_ == 1)
    host_issue_ms, host_cpu_ms = host_ms[0], host_ms[1]
    clocks = sampler.stop() if sampler else None
    if S_ > 1:
        # With several batches in flight the event pairs around one stream's GEMMs also span other streams' kernels, so the
        # roofline of the GEMM kernel is taken from a strictly serial pass (one stream) over the same workload.
        saved = S_
        S_ = 1
        _, _, prof, _ = timed(False, False, max(2, min(a.steps, 4)), True)
        prof = dict(prof); prof["serial_pass_steps"] = max(2, min(a.steps, 4))
        S_ = saved
    for i in range(S_):
        step(i, True, True)
    ms_e2e, _, _, _ = timed(True, True, a.steps, False)

    if rank == 0:
        n_dialogs = B * world * a.steps
        value = n_dialogs / (ms / 1e3)
        e2e = n_dialogs / (ms_e2e / 1e3)
        peaks, peak_src = measured_peaks()
        h2d = sum(v.numel() * v.element_size() for k, v in host.items() if k in ("enc_image_feat", "enc_image_loc", "enc_image_mask",
                                                                                  "enc_input_ids", "enc_segments", "dec_input_ids"))
        h2d += questions_h.numel() * questions_h.element_size()
        d2h = B * world * a.rounds * 18 * 8 + B * world * 4
        peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
        ach_tf = prof["flops"] / (prof["ms"] * 1e-3) / 1e12 if prof and prof["ms"] > 0 else None
        traffic = None
        tp = os.path.join(ROOT, "profiles", "gemm_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": n_warm,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": a.dtype,
            "data": "synthetic",
            "config": {"workload": workload_name(a), "global_batch": B * world, "rounds": a.rounds, "beams": a.beams,
                       "parallelism": f"dp{world}: independent image shards, one final NCCL all_gather of token ids",
                       "history_positions": "all 256" if a.no_trim else "ceil32(longest history in the batch): padded positions are never read, results identical",
                       "streams_per_gpu": S_, "batch_per_forward": B, "host_enqueue_ms_per_step": host_issue_ms, "host_cpu_ms_per_step": host_cpu_ms,
                       "l2": "no explicit flush: each step streams >1.5 GB (0.78 GB bf16 weights, 0.69 GB cross-KV, activations), "
                             "far beyond the 126 MB L2",
                       "end_to_end_tflops": value * TF_PER_DIALOG, "tf_per_dialog": TF_PER_DIALOG},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d * world), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": ms_e2e / a.steps},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "gemm_tc_kernel (tcgen05 GEMM, launches with M >= 1024: encoder + cross-KV prefill)", "bound": "tensor",
                         "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": (ach_tf / peak_tf) if ach_tf else None,
                         "traffic": traffic, "peak_source": peak_src + ", bf16_tflops_sustained (kernel timed inside a long step)",
                         "launches": prof["launches"] if prof else 0,
                         "kernel_ms_per_step": prof["ms"] / prof.get("serial_pass_steps", a.steps) if prof else None,
                         "measured_in": "serial single-stream pass inside this run" if S_ > 1 else "the timed region",
                         "algorithmic_flops_per_launch": prof["flops"] / max(prof["launches"], 1) if prof else None},
        }
        if world == 1 and not a.no_cpu_baseline:
            threads = os.cpu_count() or 1
            cpu_reference_round(enc_cfg, dec_cfg, sd, a.beams, threads)          # warm-up
            t_round = statistics.mean(cpu_reference_round(enc_cfg, dec_cfg, sd, a.beams, threads) for _ in range(3))
            line["cpu_baseline"] = {"value": 1.0 / (a.rounds * t_round), "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"1 image x 1 round (of {a.rounds}), beam {a.beams}, fp32, reference algorithm (no KV cache, "
                                              f"discarded heads evaluated), mean of 3 repeats = {t_round:.2f} s, scaled x{a.rounds} to a dialog"}
        _emit(line)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
