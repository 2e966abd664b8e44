"""Summarise an `ncu --set full` report (.ncu-rep) per kernel: duration, DRAM bytes, occupancy, pipe use."""
import csv
import io
import json
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma pipe %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("smsp__inst_executed.sum", "warp insts"),
]


def rows_of(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def main(path, json_out=None):
    hdr, units, rows = rows_of(path)
    ki = hdr.index("Kernel Name")
    cols = [(hdr.index(m), label, units[hdr.index(m)]) for m, label in WANT if m in hdr]
    print("| kernel | " + " | ".join(f"{l} ({u})" if u else l for _, l, u in cols) + " |")
    print("|---|" + "---:|" * len(cols))
    traffic = {}
    for r in rows:
        name = r[ki].split("(")[0].replace("void ", "").replace("sn2::", "")
        vals = [r[i] for i, _, _ in cols]
        print(f"| `{name}` | " + " | ".join(v[:12] for v in vals) + " |")
        d = {l: (r[i], u) for i, l, u in cols}

        def to_bytes(v, u):
            f = float(v.replace(",", ""))
            return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)

        if "dram read" in d:
            traffic.setdefault(name, []).append(to_bytes(*d["dram read"]) + to_bytes(*d["dram write"]))
    if json_out:
        json.dump({k: sum(v) / len(v) for k, v in traffic.items()}, open(json_out, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
