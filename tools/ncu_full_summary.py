"""Key metrics of every kernel in `ncu --set full` reports, as a markdown table (for profiles/).
   python tools/ncu_full_summary.py rep1.ncu-rep [rep2 ...] > profiles/rN_ncu_full_summary.md"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "dur us"),
    ("dram__bytes_read.sum", "dram rd MB"),
    ("dram__bytes_write.sum", "dram wr MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem"),
]
print("# ncu --set full --clock-control none: per-launch summary\n")
print("| report | kernel | grid | " + " | ".join(k[1] for k in KEYS) + " |")
print("|---|---|---|" + "---:|" * len(KEYS))
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    ix = {h: i for i, h in enumerate(rows[0])}
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
        vals = [r[ix[k]] if k in ix else "" for k, _ in KEYS]
        vals = [f"{float(v):.2f}" if v.replace(".", "", 1).isdigit() and "." in v else v for v in vals]
        print(f"| {rep.split('/')[-1]} | `{name}` | {r[ix['Grid Size']]} | " + " | ".join(vals) + " |")
