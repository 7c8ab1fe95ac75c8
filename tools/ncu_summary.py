"""Short text summary of an .ncu-rep (one block per profiled launch): ncu -i rep --page raw --csv | this script.

    python tools/ncu_summary.py gpurun_out/full_x.ncu-rep [...] > profiles/r1_x_ncu_summary.txt
"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"), ("launch__waves_per_multiprocessor", "waves/SM"),
    ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("dram__bytes_read.sum.per_second", "dram read rate"),
    ("dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "dram read % of ncu peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm throughput %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (mma.sync)"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor hmma subpipe active %"),
    ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor memory/MMA cycles active % (tcgen05)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe active %"),
    # L2 -> SM traffic (the budget the streams share).  `--set full` carries the lts__t_* counters.
    ("lts__t_bytes.sum", "L2 bytes (all traffic)"), ("lts__t_bytes.sum.per_second", "L2 byte rate"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "L2 sectors read by SMs (x 32 B)"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of ncu peak"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "crossbar -> SM read bytes"),
]

for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        print(f"{rep}: no data")
        continue
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print(f"== {rep}: {r[hdr.index('Kernel Name')][:150]}")
        for key, label in WANT:
            if key in hdr:
                i = hdr.index(key)
                print(f"   {label:34s} {r[i]} {units[i]}")
        print()
