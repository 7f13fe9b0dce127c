"""Derive bench.py's roofline constants from an ncu capture of the head build.

On the GPU box (after the plain command has exited 0):

    ncu --metrics smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,\\
smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,\\
gpu__time_duration.sum --clock-control none -k regex:'k_symphony_fast|k_heyvaerts_fast' -c 2 --csv \\
        --log-file gpurun_out/consts.csv python tools/profile_small.py 8192 0xFF > gpurun_out/consts_run.log

Here:

    python tools/ncu_constants.py gpurun_out/consts.csv gpurun_out/consts_run.log [profiles/kernel_constants.json]

The run log carries the rule applications per point the two kernels counted (tools/profile_small.py
prints them); the CSV carries the per-launch instruction and DRAM counters.  FP64 flops per rule
application = (2 dfma + dmul + dadd thread instructions) / (points x applications per point).
"""
import csv
import json
import re
import sys


def main():
    csv_path, log_path = sys.argv[1], sys.argv[2]
    out_path = sys.argv[3] if len(sys.argv) > 3 else "profiles/kernel_constants.json"
    log = open(log_path).read()
    m = re.search(r"n (\d+) mask (\S+) kernel ms \[([^\]]*)\] apps/pt ([\d.eE+-]+) ([\d.eE+-]+)", log)
    if not m:
        raise SystemExit("run log does not hold the profile_small.py line")
    n = int(m.group(1))
    apps = {"symphony": float(m.group(4)), "heyvaerts": float(m.group(5))}

    rows = []
    with open(csv_path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        rows.append(r)
    per = {}
    for r in rows:
        name = r["Kernel Name"]
        key = "symphony" if "k_symphony_fast" in name else ("heyvaerts" if "k_heyvaerts_fast" in name else None)
        if key is None:
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "")
        metric = r["Metric Name"]
        if metric.startswith("dram__bytes"):
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
            val *= scale
        if metric == "gpu__time_duration.sum":
            scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
            val *= scale  # ms
        per.setdefault(key, {})[metric] = val
    out = {"source": f"{csv_path} + {log_path} (tools/ncu_constants.py; {n} seeded pitchy power-law points)",
           "points": n, "applications_per_point": apps, "flop_per_application": {}, "dram_bytes_per_point": {},
           "warp_instructions_per_application": {}, "ncu_duration_ms": {}}
    for key, mtr in per.items():
        flops = (2.0 * mtr["smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"] +
                 mtr["smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"] +
                 mtr["smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"])
        total_apps = apps[key] * n
        out["flop_per_application"][key] = flops / total_apps
        out["dram_bytes_per_point"][key] = (mtr["dram__bytes_read.sum"] + mtr["dram__bytes_write.sum"]) / n
        out["warp_instructions_per_application"][key] = mtr["smsp__inst_executed.sum"] / total_apps
        out["ncu_duration_ms"][key] = mtr["gpu__time_duration.sum"]
    with open(out_path, "w") as f:
        json.dump(out, f, indent=1)
        f.write("\n")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
