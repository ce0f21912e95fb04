"""`ncu -i <rep> --page raw --csv` -> per-launch DRAM traffic / tensor-pipe table of the GEMM launches and
profiles/gemm_traffic.json, the file bench.py's `roofline.traffic` is read from.

    ncu -i gpurun_out/prof_gemm.ncu-rep --page raw --csv > gpurun_out/prof_gemm_raw.csv
    python tools/ncu_traffic.py gpurun_out/prof_gemm_raw.csv <source-name> [kernel-regex]
"""
import csv
import json
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TIME = {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3, "second": 1e6}


def main():
    path, source = sys.argv[1], sys.argv[2]
    pat = re.compile(sys.argv[3] if len(sys.argv) > 3 else "gemm_tcgen05")
    rows = list(csv.reader(l for l in open(path, newline="") if l.startswith('"')))
    head, units, body = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(head)}

    def val(r, name, table):
        v = float(r[col[name]].replace(",", "") or 0)
        return v * table.get(units[col[name]], 1.0)

    out, tot, n = [], 0.0, 0
    for r in body:
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("ctc::", "")
        if not pat.search(name):
            continue
        rd, wr = val(r, "dram__bytes_read.sum", UNIT), val(r, "dram__bytes_write.sum", UNIT)
        us = val(r, "gpu__time_duration.sum", TIME)
        tp = r[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]] \
            if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" in col else ""
        out.append({"kernel": name, "grid": r[col["Grid Size"]], "dram_read_mb": rd / 1e6, "dram_write_mb": wr / 1e6,
                    "duration_us": us, "tensor_pipe_active_pct": tp})
        tot += rd + wr
        n += 1
    print("| kernel | grid | DRAM read MB | DRAM write MB | duration us | tensor-pipe active % |\n|---|---|---|---|---|---|")
    for o in out:
        print(f"| `{o['kernel']}` | {o['grid']} | {o['dram_read_mb']:.1f} | {o['dram_write_mb']:.1f} | {o['duration_us']:.1f} | "
              f"{o['tensor_pipe_active_pct']} |")
    commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
    if n:
        (ROOT / "profiles" / "gemm_traffic.json").write_text(json.dumps(
            {"source": source, "commit": commit, "launches": n, "mean_bytes_per_launch": tot / n, "per_launch": out}, indent=1))
        print(f"\n{n} launches, mean {tot / n / 1e6:.1f} MB per launch -> profiles/gemm_traffic.json")


if __name__ == "__main__":
    main()
