#!/usr/bin/env python
"""profiles/decode_traffic.json from an ncu `--set full` capture of the decode kernel (read by bench.py into
`roofline.traffic`).  The JSON records the sha256 of the decode kernel's sources; bench.py refuses a capture whose
hash differs from the sources it runs (a stale capture).

    # on the GPU box (one GPU; cfg2 decode launches only):
    ncu --set full --clock-control none --import-source on -k regex:paged_decode_mma -s 4 -c 2 \
        -o gpurun_out/decode python bench.py --steps 6 --warmup 4 --repeats 1 --sustain-s 0 --no-extra --no-cpu-baseline
    # here:
    ncu -i gpurun_out/decode.ncu-rep --page raw --csv > /tmp/decode_raw.csv
    python tools/decode_traffic.py /tmp/decode_raw.csv profiles/r2_decode_mma_ncu_full_raw.csv
"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mojo_opset_b200 import build  # noqa: E402


def main():
    raw, keep = sys.argv[1], sys.argv[2]
    with open(raw, newline="") as f:
        rows = list(csv.reader(f))
    header = rows[0]
    col = {name: i for i, name in enumerate(header)}
    launches = [r for r in rows[2:] if len(r) == len(header) and "paged_decode_mma" in r[col["Kernel Name"]]]
    assert launches, "no paged_decode_mma launch in the capture"
    units = rows[1]

    def metric(name):
        vals = []
        for r in launches:
            v = float(r[col[name]].replace(",", ""))
            u = units[col[name]].lower()
            v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3,
                  "msecond": 1e3, "nsecond": 1e-3}.get(u, 1)
            vals.append(v)
        return sum(vals) / len(vals)

    rd, wr = metric("dram__bytes_read.sum"), metric("dram__bytes_write.sum")
    cfg = bench.CFG2
    out = {
        "kernel": launches[0][col["Kernel Name"]],
        "source": f"{os.path.relpath(keep, ROOT)} (ncu --set full --clock-control none, cfg2: B64 32q/8kv hd128 page16 "
                  f"ctx4096, {len(launches)} launches averaged)",
        "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
        "algorithmic_bytes": bench.decode_bytes(cfg["batch"], cfg["ctx"], cfg["hq"], cfg["hkv"], cfg["d"], cfg["bs"]),
        "gpu_time_us_under_ncu": metric("gpu__time_duration.sum"),
        "source_sha256": bench.sources_sha256(),
        "kernel_sass_sha256": build.kernel_sass_sha256(os.path.join(build.OBJ_DIR, "paged_decode.o"), build.DECODE_KERNEL_SYMBOL),
        "commit": subprocess.run(["git", "-C", ROOT, "rev-parse", "HEAD"], capture_output=True, text=True).stdout.strip(),
    }
    if os.path.abspath(raw) != os.path.abspath(keep):
        shutil.copyfile(raw, keep)
    with open(os.path.join(ROOT, "profiles", "decode_traffic.json"), "w") as f:
        json.dump(out, f, indent=2)
        f.write("\n")
    print(json.dumps(out, indent=2))


if __name__ == "__main__":
    main()
