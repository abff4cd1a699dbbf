#!/usr/bin/env python3
"""Turns an `ncu --set full --import-source on` capture of k_verify_fast into profiles/r2_ncu_constants.json: DRAM bytes
and executed-instruction counts per launch, tied to the hash of the sources the library was built from
(cost_model.source_sha256) so that bench.py only quotes them for the binary they were measured on.

    python tools/make_ncu_constants.py gpurun_out/e1_kvf.ncu-rep 20 "<how the capture was taken>" """
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cost_model  # noqa: E402


def main():
    rep, log2n, how = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    d = dict(zip(rows[0], rows[2]))
    units = dict(zip(rows[0], rows[1]))

    def val(name):
        v, u = float(d[name].replace(",", "")), units[name]
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12}.get(u, 1)

    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    hdr = srows[1]
    ci, si = hdr.index("Instructions Executed"), hdr.index("Source")
    n = 1 << log2n
    warps = n / 32
    total = wide = hi = 0.0
    for r in srows[2:]:
        try:
            c = float(r[ci])
        except (ValueError, IndexError):
            continue
        op = re.sub(r"^@!?U?P\d+\s+", "", r[si].strip()).split()[0] if r[si].strip() else ""
        total += c
        if op.startswith("IMAD.WIDE"):
            wide += c
        elif op.startswith("IMAD.HI"):
            hi += c
    out = {
        "capture": how,
        "source_sha256": cost_model.source_sha256(),
        "k_verify_fast": {str(log2n): {
            "dram_bytes": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
            "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
            "duration_ms_under_ncu": float(d["gpu__time_duration.sum"].replace(",", "")) * {"ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(units["gpu__time_duration.sum"], 1),
            "warp_instructions_per_verification_warp": total / warps,
            "imad_wide_per_verification": wide / warps, "imad_hi_per_verification": hi / warps,
            "registers_per_thread": int(float(d.get("launch__registers_per_thread", "0") or 0)),
            "l1_hit_pct": float(d.get("l1tex__t_sector_hit_rate.pct", "nan")), "l2_hit_pct": float(d.get("lts__t_sector_hit_rate.pct", "nan")),
            "issue_slots_busy_pct": float(d.get("smsp__issue_active.avg.pct_of_peak_sustained_active", "nan")),
            "pipe_fmaheavy_pct": float(d.get("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "nan")),
            "pipe_alu_pct": float(d.get("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "nan")),
        }},
    }
    path = os.path.join(ROOT, "profiles", "r2_ncu_constants.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
        f.write("\n")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
