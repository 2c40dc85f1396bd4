#!/usr/bin/env python
"""profiles/ncu_summary.json from `ncu --set full` captures of the bench command: per kernel the DRAM bytes of one launch,
what the launch processed, its duration and the utilisation of every issue pipe.  bench.py reads it for `roofline.traffic`
and `roofline.pipe`, and ignores it (traffic null, capture_stale true) when `csrc_sha16` differs from the sources built there.
usage: python scripts/ncu_capture.py KEY=path.ncu-rep:UNITS[:label] ...   (UNITS = signatures or ranges in that launch)
       KEY in {ed25519, subchain_proofs, subchain_commit, ...}; also writes profiles/<tag>_<KEY>_ncu_full.csv via ncu_summary."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

PIPES = {
    "alu": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "fma": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "fmaheavy": "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "fp64": "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "lsu": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "fmaheavy_cycles_active_pct_elapsed": "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "alu_cycles_active_pct_elapsed": "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "fp64_cycles_active_pct_elapsed": "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "issue_active": "smsp__issue_active.avg.pct_of_peak_sustained_active",
}
UNIT_SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def main(argv):
    import bench
    tag = os.environ.get("TAG", "r02")
    summary = {"csrc_sha16": bench.csrc_sha16(), "how": "ncu --set full --clock-control none, one launch per kernel, inside the bench command",
               "kernels": {}}
    for spec in argv:
        key, rest = spec.split("=", 1)
        parts = rest.split(":")
        rep, units = parts[0], int(parts[1])
        hdr, unit, launches = raw(rep)
        col = {h: i for i, h in enumerate(hdr)}
        l = launches[-1]

        def val(name, scale_table=True):
            if name not in col or l[col[name]] in ("", "n/a"):
                return None
            v = float(l[col[name]].replace(",", ""))
            return v * UNIT_SCALE.get(unit[col[name]], 1) if scale_table else v
        dram = (val("dram__bytes_read.sum") or 0) + (val("dram__bytes_write.sum") or 0)
        out_csv = os.path.join("profiles", f"{tag}_{key}_ncu_full.csv")
        subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), rep, os.path.join(ROOT, out_csv)], check=True)
        summary["kernels"][key] = {
            "kernel_name": l[col["Kernel Name"]], "grid": l[col["Grid Size"]], "block": l[col["Block Size"]],
            "units_per_launch": units, "unit_label": parts[2] if len(parts) > 2 else "units",
            "dram_bytes_per_launch": int(dram), "duration_us": val("gpu__time_duration.sum"),
            "registers_per_thread": val("launch__registers_per_thread", False),
            "pipes": {k: val(m, False) for k, m in PIPES.items()},
            "local_load_sectors": val("l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", False),
            "l1_hit_rate_pct": val("l1tex__t_sector_hit_rate.pct", False),
            "source": out_csv,
        }
    with open(os.path.join(ROOT, "profiles", "ncu_summary.json"), "w") as f:
        json.dump(summary, f, indent=1)
    print(json.dumps(summary, indent=1))


if __name__ == "__main__":
    main(sys.argv[1:])
