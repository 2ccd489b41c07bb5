"""Per-kernel-class numbers out of `ncu --set full` reports, stamped with the digest of the kernel sources
the library was built from: profiles/ncu_metrics.json (read by bench.py, which prints `roofline.traffic`
and `tensor_pipe_pct_ncu` only when the digest is the one of the library it runs).

    python tools/ncu_extract.py OUT.json WORKLOAD REPORT.ncu-rep:class_a,class_b,... [REPORT2:...]

The classes name the report's launches in capture order (tools/gpu/profile_r02.sh fixes the -s/-c windows);
launches of the same class are averaged."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "usecond": 1e-6,
        "ms": 1e-3, "msecond": 1e-3, "nsecond": 1e-9, "s": 1.0, "second": 1.0, "%": 1.0, "": 1.0}
KEYS = {"duration_s": "gpu__time_duration.sum", "dram_read": "dram__bytes_read.sum", "dram_write": "dram__bytes_write.sum",
        "tensor_pipe_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "xu_pipe_pct": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm_throughput_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram_throughput_pct": "dram__throughput.avg.pct_of_peak_sustained_elapsed"}


def rows_of(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    hdr, units = r[0], r[1]
    out = []
    for row in r[2:]:
        d = {"name": row[hdr.index("Kernel Name")]}
        for k, m in KEYS.items():
            if m in hdr:
                i = hdr.index(m)
                try:
                    d[k] = float(row[i].replace(",", "")) * UNIT.get(units[i], 1.0)
                except ValueError:
                    pass
        out.append(d)
    return out


def main():
    out_path, workload = sys.argv[1], sys.argv[2]
    acc = {}
    for spec in sys.argv[3:]:
        rep, classes = spec.rsplit(":", 1)
        classes = classes.split(",")
        rows = rows_of(rep)
        assert len(rows) >= len(classes), (rep, len(rows), classes)
        for cls, row in zip(classes, rows):
            acc.setdefault(cls, []).append(dict(row, report=os.path.basename(rep)))
    kernels = {}
    for cls, rows in acc.items():
        def mean(k):
            v = [r[k] for r in rows if k in r]
            return sum(v) / len(v) if v else None
        kernels[cls] = {"kernel_name": rows[0]["name"][:120], "launches_captured": len(rows),
                        "dram_bytes_per_launch": (mean("dram_read") or 0) + (mean("dram_write") or 0),
                        "duration_us_under_ncu": 1e6 * mean("duration_s") if mean("duration_s") else None,
                        "tensor_pipe_pct": mean("tensor_pipe_pct"), "xu_pipe_pct": mean("xu_pipe_pct"),
                        "issue_active_pct": mean("issue_active_pct"), "sm_throughput_pct": mean("sm_throughput_pct"),
                        "dram_throughput_pct": mean("dram_throughput_pct"), "reports": sorted({r["report"] for r in rows})}
    digest = open(os.path.join(ROOT, "sam_road_b200", "_build", "digest.txt")).read().strip()
    json.dump({"digest": digest, "workload": workload, "batch": 64,
               "source": "ncu --set full --clock-control none (tools/gpu/profile_r02.sh), per launch",
               "kernels": kernels}, open(out_path, "w"), indent=1, sort_keys=True)
    for k, v in kernels.items():
        print(k, v["kernel_name"][:50], "dram MB", round(v["dram_bytes_per_launch"] / 1e6, 1), "tensor %", v["tensor_pipe_pct"],
              "us", v["duration_us_under_ncu"])


if __name__ == "__main__":
    main()
