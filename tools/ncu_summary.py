#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into the text files kept under profiles/ and the
traffic table bench.py reads (profiles/r02_ncu_traffic.json).

    python tools/ncu_summary.py gpurun_out/<tag>/prof_predict.ncu-rep profiles/r02_<name>_ncu_summary.txt \
        [--traffic-key predict_acq_kernel] [--note "..."]

One block per profiled kernel launch: the metrics the roofline argument needs (duration, dram bytes, pipe
utilisation, L2 hit rate, occupancy, registers, stall reasons per issue)."""
import csv
import io
import json
import os
import re
import subprocess
import sys

KEEP = re.compile(
    r"^(dram__bytes_(read|write)\.sum$|gpu__dram_throughput\.avg\.pct|gpu__time_duration\.sum$|"
    r"l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum\.pct|launch__(block_size|grid_size|registers_per_thread|"
    r"shared_mem_per_block_dynamic|occupancy_limit)|lts__t_sector_hit_rate\.pct|lts__throughput\.avg\.pct|"
    r"lts__t_bytes\.sum$|sm__cycles_elapsed\.max|sm__ops_path_tensor|sm__pipe_(fp64|tensor|fma|alu)|"
    r"sm__inst_executed_pipe_(uniform|lsu)|sm__throughput\.avg\.pct|sm__warps_active\.avg\.pct|"
    r"smsp__average_warps_issue_stalled_.*_per_issue_active|smsp__issue_active\.avg\.pct|"
    r"smsp__inst_executed\.sum$|sm__mem_tensor)")


def main():
    rep, out = sys.argv[1], sys.argv[2]
    tkey = note = None
    if "--traffic-key" in sys.argv:
        tkey = sys.argv[sys.argv.index("--traffic-key") + 1]
    if "--note" in sys.argv:
        note = sys.argv[sys.argv.index("--note") + 1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu --set full --clock-control none; source report {rep} (not committed)"]
    if note:
        lines.append(f"# {note}")
    traffic = {}
    for r in rows[2:]:
        rec = dict(zip(hdr, r))
        name = rec.get("Kernel Name", "?")
        lines.append(f"# kernel: {name}   (launch id {rec.get('ID')})")
        for k, u in zip(hdr, units):
            if KEEP.search(k):
                lines.append(f"{k:<100} {u:<16} {rec[k]}")
        if tkey and tkey in name and not traffic:
            def val(key):
                v = float(rec[key].replace(",", ""))
                u = units[hdr.index(key)].lower()
                return v * {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1.0)
            traffic = {"dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
                       "kernel": name, "source": f"{os.path.basename(out)} (ncu --set full, this round)"}
        lines.append("")
    with open(out, "w") as f:
        f.write("\n".join(lines) + "\n")
    if tkey and traffic:
        tj = os.path.join(os.path.dirname(os.path.abspath(out)), "r02_ncu_traffic.json")
        cur = json.load(open(tj)) if os.path.exists(tj) else {}
        cur[tkey] = traffic
        json.dump(cur, open(tj, "w"), indent=1)
    print(f"wrote {out} ({len(rows) - 2} launches)" + (f", traffic[{tkey}] updated" if traffic else ""))


if __name__ == "__main__":
    main()
