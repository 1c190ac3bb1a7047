#!/usr/bin/env python
"""Key metrics of an `ncu --set full` report: python scripts/ncu_summary.py report.ncu-rep [more.ncu-rep ...]"""
import csv, subprocess, sys

WANT = [
    ("duration", "gpu__time_duration.sum"),
    ("dram read", "dram__bytes_read.sum"),
    ("dram write", "dram__bytes_write.sum"),
    ("dram throughput % of peak", "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L2 throughput % of peak", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L2 -> SM bytes", "l1tex__m_xbar2l1tex_read_bytes.sum"),
    ("tensor pipe active % (realtime, elapsed)", "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
    ("bf16->fp32 tensor ops % of peak", "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed"),
    ("tensor-memory pipe cycles active %", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("SM throughput %", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("registers/thread", "launch__registers_per_thread"),
    ("dyn smem/block", "launch__shared_mem_per_block_dynamic"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("SM clock", "sm__cycles_elapsed.avg.per_second"),
]

for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"=== {path}")
    for r in rows[2:]:
        print(f"kernel: {r[idx['Kernel Name']].strip()[:140]}")
        for label, key in WANT:
            if key in idx:
                print(f"    {label:44s} {r[idx[key]]} {units[idx[key]]}")
