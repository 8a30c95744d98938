#!/usr/bin/env python
"""Compact summary of an Nsight Compute report (read here, no GPU): usage ncu_summary.py report.ncu-rep [out.txt]
Keeps identity, duration, DRAM bytes, tensor-pipe, crossbar, L2 and occupancy figures of every captured launch."""
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, launches = rows[0], rows[1], rows[2:]
keep = re.compile(r"^(Kernel Name|Grid Size|Block Size|gpu__time_duration\.sum$|dram__bytes_(read|write)\.sum($|\.per_second)|"
                  r"gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|launch__(grid_size|registers_per_thread$|shared_mem_per_block_dynamic|cluster)|"
                  r"sm__ops_path_tensor_op_utc\w+\.(avg|sum)\.pct_of_peak_sustained_elapsed|sm__pipe_tensor\w*cycles_active\w*\.avg\.pct_of_peak_sustained_(active|elapsed)|"
                  r"TPC\.TriageCompute\.sm__pipe_tensor_cycles_active_realtime\.avg\.pct|l1tex__m_xbar2l1tex_read_bytes\.sum($|\.per_second)|"
                  r"lts__t_sector_hit_rate\.pct|lts__t_bytes\.sum\.per_second|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"sm__warps_active\.avg\.pct_of_peak_sustained_active|smsp__cycles_active\.avg|sm__cycles_elapsed\.max|"
                  r"sm__inst_executed_pipe_(tensor|uniform)\w*\.sum$)")
out = []
for n, row in enumerate(launches):
    out.append(f"--- launch {n} of {rep.split('/')[-1]}")
    for h, u, v in zip(hdr, units, row):
        if keep.match(h) and v not in ("", "0", "n/a"):
            out.append(f"{h} [{u}] = {v}")
text = "\n".join(out) + "\n"
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text)
print(text)
