"""Summarise an .ncu-rep (raw page) into the handful of metrics the roofline discussion uses.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [more metrics...]
"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def main():
    """python tools/ncu_summary.py REP [--json NS]: prints the summary; with --json also rewrites
    profiles/limiter.json and profiles/traffic.json (read by bench.py) from the same capture."""
    import json
    import os

    rep = sys.argv[1]
    ns = int(sys.argv[sys.argv.index("--json") + 1]) if "--json" in sys.argv else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    limiter, traffic = {}, {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print("==", name[:90])
        val = {}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"  {w:85s} {r[i]:>16s} {units[i]}")
                try:
                    val[w] = (float(r[i].replace(",", "")), units[i])
                except ValueError:
                    pass
        short = name.replace("void ", "").replace("edgpu::", "").split("<")[0].split("(")[0]
        pipes = {"HBM (gpu__dram_throughput)": val["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"][0],
                 "L2 (lts__throughput)": val["lts__throughput.avg.pct_of_peak_sustained_elapsed"][0],
                 "L1TEX / shared-memory data pipe (l1tex__throughput)":
                     val["l1tex__throughput.avg.pct_of_peak_sustained_elapsed"][0]}
        top = max(pipes, key=pipes.get)
        stalls = {k.split("issue_stalled_")[1].split("_per_")[0]: v[0] for k, v in val.items()
                  if "issue_stalled" in k and "per_issue_active" in k}
        st = max(stalls, key=stalls.get)
        limiter[short] = (f"{top} {pipes[top]:.0f} % of peak; top stall {st} ({stalls[st]:.1f} warps per issue); "
                          f"issue slots {val['smsp__issue_active.avg.pct_of_peak_sustained_active'][0]:.0f} % busy, "
                          f"warps active {val['sm__warps_active.avg.pct_of_peak_sustained_active'][0]:.0f} % "
                          f"[{os.path.basename(rep)}]")

        def gb(key):
            v, u = val[key]
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]

        def ms(key):
            v, u = val[key]
            return v * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(u, 1.0)

        traffic[short] = {"dram_bytes_per_launch": gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum"),
                          "duration": ms("gpu__time_duration.sum"), "duration_unit": "ms"}
    if ns is not None:
        root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles")
        with open(os.path.join(root, "limiter.json"), "w") as f:
            json.dump(limiter, f, indent=1)
        with open(os.path.join(root, "traffic.json"), "w") as f:
            json.dump({"ns": ns, "source": f"{os.path.basename(rep)} (ncu --set full --clock-control none)",
                       "kernels": traffic}, f, indent=1)


if __name__ == "__main__":
    main()
