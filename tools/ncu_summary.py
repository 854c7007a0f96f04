"""Summarise an .ncu-rep (read here with `ncu -i ... --page raw --csv`) into the JSON kept under profiles/.
Usage: python tools/ncu_summary.py <file.ncu-rep> <out.json> "<command>" "<note>"."""
import csv
import io
import json
import subprocess
import sys

KEYS = ["Kernel Name", "Block Size", "Grid Size", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]

rep, out, cmd, note = sys.argv[1:5]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
launches = []
for r in rows[2:]:
    d = {}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            d[k] = (r[i] + " " + units[i]).strip()
    launches.append(d)
json.dump({"command": cmd, "note": note, "launches": launches}, open(out, "w"), indent=1)
print(json.dumps(launches, indent=1))
