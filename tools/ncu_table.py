#!/usr/bin/env python
"""One line per kernel launch from an `ncu --page raw --csv` export (tools/ncu_step.sh): duration, DRAM bytes and
throughput, tensor / XU / issue utilisation, local-memory instructions, top stall reasons.
   python tools/ncu_table.py gpurun_out/step_raw.csv [call-name ...]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
names = sys.argv[2:]


def num(d, k):
    try:
        return float(d.get(k, "nan").replace(",", ""))
    except ValueError:
        return float("nan")


STALLS = ["barrier", "long_scoreboard", "mio_throttle", "math_pipe_throttle", "short_scoreboard", "wait", "sleeping",
          "branch_resolving", "lg_throttle", "tex_throttle", "dispatch_stall", "no_instruction", "membar", "drain", "imc_miss"]
print(f"{'#':>3s} {'kernel':34s} {'grid':>8s} {'us':>8s} {'rd MB':>7s} {'wr MB':>7s} {'GB/s':>7s} {'dram%':>6s} {'tens%':>6s} {'xu%':>5s} "
      f"{'issue%':>6s} {'warps%':>6s} {'L2hit%':>6s} {'ldl':>8s} {'stl':>8s}  top stalls (per issue)")
for i, r in enumerate(rows[2:]):
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    t = num(d, "gpu__time_duration.sum")
    t_us = t / 1e3 if u.get("gpu__time_duration.sum", "") in ("ns", "nsecond") else (t if u.get("gpu__time_duration.sum") in ("us", "usecond") else t * 1e3)
    def bytes_of(k):
        v, un = num(d, k), u.get(k, "")
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(un, 1)
    rd, wr = bytes_of("dram__bytes_read.sum"), bytes_of("dram__bytes_write.sum")
    st = sorted(((num(d, f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"), s) for s in STALLS), reverse=True)
    st = ", ".join(f"{s} {v:.2f}" for v, s in st[:3] if v == v)
    k = d.get("Kernel Name", "?").replace("svol::", "").replace("void ", "")[:34]
    label = f" [{names[i]}]" if i < len(names) else ""
    print(f"{i:3d} {k:34s} {d.get('Grid Size', ''):>8s} {t_us:8.1f} {rd / 1e6:7.1f} {wr / 1e6:7.1f} {(rd + wr) / t_us / 1e3:7.0f} "
          f"{num(d, 'dram__throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} "
          f"{num(d, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'):6.1f} "
          f"{num(d, 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed'):5.1f} "
          f"{num(d, 'sm__issue_active.avg.pct_of_peak_sustained_elapsed'):6.1f} "
          f"{num(d, 'sm__warps_active.avg.pct_of_peak_sustained_active'):6.1f} {num(d, 'lts__t_sector_hit_rate.pct'):6.1f} "
          f"{num(d, 'sass__inst_executed_local_loads'):8.0f} {num(d, 'sass__inst_executed_local_stores'):8.0f}  {st}{label}")
