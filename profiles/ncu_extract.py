#!/usr/bin/env python
"""profiles/ncu_extract.py <capture.ncu-rep> <out.json> [--configs N] [--note TEXT]
Reads ONE kernel launch out of an `ncu --set full` capture (ncu -i ... --page raw --csv, run here on the CPU container) and writes the
hardware-counter summary bench.py quotes in its `roofline` object: launch duration, warp instructions executed, issue-slot and pipe
utilisation, FP32 / FP64 arithmetic instruction counts (thread level), DRAM bytes.  Counts are per launch; bench.py scales them by the
number of configurations (`configs`, default: grid size x block size) because they are linear in it for a fixed scene."""
import csv
import io
import json
import subprocess
import sys


def main():
    rep, out = sys.argv[1], sys.argv[2]
    configs, note = None, ""
    if "--configs" in sys.argv:
        configs = int(sys.argv[sys.argv.index("--configs") + 1])
    if "--note" in sys.argv:
        note = sys.argv[sys.argv.index("--note") + 1]
    text = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    hdr, units, val = rows[0], rows[1], rows[2]
    col = {h: (v, u) for h, u, v in zip(hdr, units, val)}

    def num(name, scale_units=True):
        v, u = col[name]
        x = float(v.replace(",", ""))
        if scale_units:
            x *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-3, "ms": 1.0, "s": 1e3, "ns": 1e-6}.get(u, 1.0)
        return x

    cycles = num("sm__cycles_elapsed.max") if "sm__cycles_elapsed.max" in col else num("smsp__cycles_elapsed.max")
    per_cycle = lambda op: num(f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed")
    grid = [int(x) for x in col["Grid Size"][0].strip("()").split(",")] if "Grid Size" in col else None
    block = [int(x) for x in col["Block Size"][0].strip("()").split(",")] if "Block Size" in col else None
    threads = (grid[0] * grid[1] * grid[2] * block[0] * block[1] * block[2]) if grid and block else None
    fadd, fmul, ffma = per_cycle("fadd") * cycles, per_cycle("fmul") * cycles, per_cycle("ffma") * cycles
    dadd, dmul, dfma = per_cycle("dadd") * cycles, per_cycle("dmul") * cycles, per_cycle("dfma") * cycles
    res = {
        "capture": rep.split("gpurun_out/")[-1], "kernel": col["Kernel Name"][0], "grid": grid, "block": block,
        "configs": configs or threads, "duration_ms": num("gpu__time_duration.sum"), "sm_cycles_elapsed": cycles,
        "sm_clock_mhz_in_capture": round(cycles / num("gpu__time_duration.sum") / 1e3, 1),
        "warp_inst_executed": num("smsp__inst_executed.sum"),
        "issue_slots_busy_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "inst_executed_pct_of_peak_elapsed": num("sm__inst_executed.sum.pct_of_peak_sustained_elapsed"),
        "pipe_alu_pct": num("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "pipe_fma_pct": num("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        "active_threads_per_warp_inst": num("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "warps_active_pct": num("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "warps_eligible_per_scheduler": num("smsp__warps_eligible.avg.per_cycle_active"),
        "thread_inst_fp32": {"fadd": fadd, "fmul": fmul, "ffma": ffma, "flop": fadd + fmul + 2 * ffma},
        "thread_inst_fp64": {"dadd": dadd, "dmul": dmul, "dfma": dfma, "flop": dadd + dmul + 2 * dfma},
        "dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_write": num("dram__bytes_write.sum"),
        "registers_per_thread": int(float(col["Registers Per Thread"][0])) if "Registers Per Thread" in col else None,
        "note": note,
    }
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
