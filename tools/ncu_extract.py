"""Compact per-launch table out of an `ncu --set full` report: python tools/ncu_extract.py <report.ncu-rep> [labels.txt]"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
labels = [l.strip() for l in open(sys.argv[2])] if len(sys.argv) > 2 else []
want = [("gpu__time_duration.sum", "time"), ("Grid Size", "grid"), ("launch__registers_per_thread", "regs"),
        ("launch__shared_mem_per_block_dynamic", "dyn smem"), ("sm__cycles_elapsed.max", "cycles"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "hmma subpipe %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
        ("lts__t_bytes.sum", "L2 bytes"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
        ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM read"), ("l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "L2->SM rate"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__inst_executed.sum", "warp insts")]
idx = {h: i for i, h in enumerate(hdr)}
for n, r in enumerate(data):
    print(f"--- launch {n}: {labels[n] if n < len(labels) else ''}  [{r[idx['Kernel Name']][:40]}]")
    for key, name in want:
        if key in idx:
            print(f"    {name:24s} {r[idx[key]]:>14s} {units[idx[key]]}")
