#!/usr/bin/env python
"""Record the DRAM traffic of the dominant kernel from an ncu capture into profiles/traffic.json (bench.py reads it).

    ncu --set full --clock-control none -k regex:logmel_tc -c 3 -o gpurun_out/prof python bench.py --steps 2 --warmup 1 --quick --no-cpu-baseline
    python tools/ncu_traffic.py gpurun_out/prof.ncu-rep --n-mels 80 --batch 256 --out-dtype f32

Takes dram__bytes_read.sum + dram__bytes_write.sum of the LAST captured launch of the kernel (per launch, like
`roofline.achieved`), and notes the report it came from.
"""
import argparse
import csv
import io
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--kernel", default="logmel_tc_kernel")
    ap.add_argument("--key-kernel", default="tcgen05_pass")
    ap.add_argument("--n-mels", type=int, default=80)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--out-dtype", default="f32")
    ap.add_argument("--label", default=None)
    a = ap.parse_args()
    out = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    picked = [r for r in data if a.kernel in r[col["Kernel Name"]]]
    if not picked:
        raise SystemExit(f"no launch of {a.kernel} in {a.report}")
    r = picked[-1]

    def to_bytes(name):
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[col[name]]]
        return float(r[col[name]]) * scale

    rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
    path = os.path.join(ROOT, "profiles", "traffic.json")
    table = {}
    if os.path.exists(path):
        with open(path) as f:
            table = json.load(f)
    table[f"{a.key_kernel}|{a.n_mels}|{a.batch}|{a.out_dtype}"] = {
        "dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
        "gpu_time_us": float(r[col["gpu__time_duration.sum"]]) if "gpu__time_duration.sum" in col else None,
        "source": f"ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one launch ({a.label or os.path.basename(a.report)})",
    }
    with open(path, "w") as f:
        json.dump(table, f, indent=1, sort_keys=True)
    print(json.dumps(table[f"{a.key_kernel}|{a.n_mels}|{a.batch}|{a.out_dtype}"], indent=1))


if __name__ == "__main__":
    main()
