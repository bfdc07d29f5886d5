"""Turn .ncu-rep captures (read here, no GPU needed) into the committed summaries under profiles/.

    python tools/ncu_summary.py gpurun_out/prof_simt.ncu-rep gpurun_out/prof_tc.ncu-rep ... --tag r1

Writes profiles/<tag>_ncu_<name>.csv (one column per profiled launch) and merges per-kernel DRAM traffic
(dram__bytes_read.sum + dram__bytes_write.sum, per launch) into profiles/<tag>_traffic.json, which bench.py
reports as roofline.traffic."""
import csv
import io
import json
import os
import re
import subprocess
import sys

KEEP = [
    "Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum",
    "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct", "smsp__warps_eligible.avg.per_cycle_active",
] + [f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio" for s in
     ("long_scoreboard", "short_scoreboard", "barrier", "wait", "not_selected", "dispatch_stall", "math_pipe_throttle", "mio_throttle")]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    tag = sys.argv[sys.argv.index("--tag") + 1] if "--tag" in sys.argv else "r1"
    reps = [a for a in args if a.endswith(".ncu-rep")]
    os.makedirs("profiles", exist_ok=True)
    tpath = f"profiles/{tag}_traffic.json"
    traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}
    for rep in reps:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        ix = {h: i for i, h in enumerate(hdr)}
        name = os.path.splitext(os.path.basename(rep))[0]
        with open(f"profiles/{tag}_ncu_{name}.csv", "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(data))])
            for k in KEEP:
                if k in ix:
                    w.writerow([k, units[ix[k]]] + [r[ix[k]][:70] for r in data])
        for r in data:
            kern = re.sub(r"<.*", "", r[ix["Kernel Name"]]).replace("void ", "").replace("hept::", "").strip()
            kern = re.sub(r"\(.*", "", kern)
            tb = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]]) + \
                to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
            traffic[kern] = {"dram_bytes_per_launch": tb, "source": os.path.basename(rep)}
    json.dump(traffic, open(tpath, "w"), indent=1, sort_keys=True)
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main()
