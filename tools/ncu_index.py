#!/usr/bin/env python
"""Build profiles/ncu_index.json from `ncu --set full` reports (read here, no GPU needed).

    python tools/ncu_index.py c4=gpurun_out/r2_c4.ncu-rep c2=gpurun_out/r2_c2.ncu-rep ... [--commit HASH]

For every workload key the entry records, for the FIRST kernel of the report: kernel name, duration, DRAM bytes
read + written per launch, FP64-pipe and LSU-wavefront utilisation, registers, grid / block, plus the report's file
name and the commit the capture was taken at.  bench.py copies the entry into its JSON line as ARCHIVED evidence
(`roofline.ncu_archived`, `roofline.traffic`): numbers under a profiler are never bench values.
Also writes a text summary next to the index (profiles/<round>_ncu_<key>.txt) with tools/ncu_summary.py's layout.
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def raw_rows(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    return rows[0], rows[1], rows[2:]


def to_bytes(value, unit):
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]
    return float(value.replace(",", "")) * scale


def main(argv):
    commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, stdout=subprocess.PIPE, text=True).stdout.strip()
    pairs = []
    for a in argv:
        if a.startswith("--commit="):
            commit = a.split("=", 1)[1]
        else:
            pairs.append(a.split("=", 1))
    index_path = os.path.join(ROOT, "profiles", "ncu_index.json")
    index = json.load(open(index_path)) if os.path.exists(index_path) else {}
    for key, path in pairs:
        hdr, units, rows = raw_rows(path)
        col = {h: i for i, h in enumerate(hdr)}
        r = rows[0]

        def num(name):
            return float(r[col[name]].replace(",", ""))

        dur_unit = units[col["gpu__time_duration.sum"]]
        dur_us = num("gpu__time_duration.sum") * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(dur_unit, 1.0)
        entry = {
            "kernel": r[col["Kernel Name"]],
            "duration_us_under_ncu": dur_us,
            "dram_bytes_per_launch": to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
            + to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]]),
            "dram_read_bytes": to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]]),
            "dram_write_bytes": to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]]),
            "fp64_pipe_pct": num("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
            "lsu_wavefronts_pct": num("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
            "dram_throughput_pct": num("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "registers_per_thread": int(num("launch__registers_per_thread")),
            "grid": int(num("launch__grid_size")), "block": int(num("launch__block_size")),
            "report": os.path.basename(path), "captured_at_commit": commit,
            "how": "ncu --set full --clock-control none (cold caches, serialised): shares, not absolutes",
        }
        index[key] = entry
        print(key, json.dumps(entry))
    with open(index_path, "w") as fh:
        json.dump(index, fh, indent=1, sort_keys=True)
        fh.write("\n")


if __name__ == "__main__":
    main(sys.argv[1:])
