#!/bin/bash
# Builder-run bench lines of every workload at N = 1 (copied to profiles/ afterwards); the driver's own are BENCH_rNN.json / SCALE_rNN.json.
set -u
mkdir -p gpurun_out
T=${1:-r2b}
python bench.py > gpurun_out/${T}_bench_gen_teacher_n1.json 2> gpurun_out/${T}_bench_gen_teacher_n1.err
python bench.py --streams 1 --no-cpu-baseline > gpurun_out/${T}_bench_gen_teacher_n1_single_stream.json 2>/dev/null
python bench.py --workload gen_qa_ppl --cpu-budget-s 60 > gpurun_out/${T}_bench_gen_qa_ppl_n1.json 2>/dev/null
python bench.py --workload select_data --cpu-budget-s 60 > gpurun_out/${T}_bench_select_data_n1.json 2>/dev/null
python bench.py --workload nsp_rank --cpu-budget-s 60 > gpurun_out/${T}_bench_nsp_rank_n1.json 2>/dev/null
for f in gpurun_out/${T}_bench_*.json; do python - "$f" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l); e = d["roofline"].get("entries", [])
        print(sys.argv[1].split("/")[-1], d["value"], d["unit"], "e2e", d["e2e"]["value"], "gemm frac", d["roofline"].get("frac"),
              "decode", [(x.get("kernel"), x.get("frac")) for x in e[1:]], "check", d["config"].get("output_check", {}).get("ids_checksum"),
              "cpu", (d.get("cpu_baseline") or {}).get("value"))
PY
done
