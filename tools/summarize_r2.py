#!/usr/bin/env python
"""Turn the raw captures of tools/prof_c.sh / gpu_quick.sh (gpurun_out/<tag>_*) into the committed summaries under profiles/:
  <out>_launch_list.md         per-kernel durations of one serialised cmx_map_batch_pe call (ncu gpu__time_duration)
  <out>_<kernel>_ncu.md        headline ncu --set full metrics + the top source lines by stall samples
  front_kernel_ncu.json        DRAM bytes / executed warp instructions of one seed_front_kernel launch (read by bench.py)
usage: summarize_r2.py <tag> <out-prefix> [kernel-name-substring ...]"""
import csv
import gzip
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def launch_list(tag, out):
    rows = list(csv.reader(open(os.path.join(G, tag + "_launches.csv"))))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    L = [(r[kn].split("(")[0].replace("void ", ""), float(r[mv].replace(",", ""))) for r in rows[hi + 1:] if len(r) > mv]
    starts = [i for i, (n, v) in enumerate(L) if n.startswith("seed_front_kernel")]
    call = L[starts[-1]:]
    tot = sum(v for _, v in call)
    md = ["# %s - ncu launch list of one `cmx_map_batch_pe` call" % out, "",
          "`CMX_LANES=1 ncu --metrics gpu__time_duration.sum --clock-control none python tools/profile_run.py --calls 2` (3 Gbp reference,",
          "2 M pairs, --preset chip), the second call; one lane, so the kernels run back to back.  Per-launch times under ncu are",
          "cold-cache and serialised: compare SHARES with `kernel_ms_per_step` of the bench line, not absolutes.", "",
          "| # | kernel | ms | share |", "|---|---|---|---|"]
    for i, (n, v) in enumerate(call):
        md.append("| %d | `%s` | %.3f | %.1f %% |" % (i, n[:52], v / 1e6, 100 * v / tot))
    md += ["", "total %.2f ms" % (tot / 1e6), ""]
    open(os.path.join(P, out + "_launch_list.md"), "w").write("\n".join(md))
    print("launch list: %.2f ms" % (tot / 1e6))


WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "sm__pipe_tma_cycles_active.avg.pct_of_peak_sustained_active"]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def kernels(tag, out, names):
    rows = list(csv.reader(gzip.open(os.path.join(G, tag + "_cap.raw.csv.gz"), "rt")))
    hdr, units = rows[0], rows[1]
    kcol = hdr.index("Kernel Name")
    seen = {}
    for r in rows[2:]:
        base = r[kcol].split("(")[0].replace("void ", "")
        k = seen.get(base, 0)
        seen[base] = k + 1
        if names and not any(n in base for n in names):
            continue
        label = "%s_%d" % (base.split("<")[0], k)
        md = ["# %s - `ncu --set full` of `%s` (launch %d of the captured call; 2 M pairs, 3 Gbp index, --preset chip)" % (out, base, k), "",
              "`CMX_LANES=1 ncu --set full --clock-control none --import-source on` (tools/prof_c.sh); numbers under ncu are for shares and counts,",
              "not for timing.", "", "| metric | value | unit |", "|---|---|---|"]
        got = {}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                got[w] = (r[i], units[i])
                md.append("| `%s` | %s | %s |" % (w, r[i], units[i]))
        lines = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), os.path.join(G, tag + "_cap.lines.csv.gz"), "14",
                                "%s#%d" % (base.split("<")[0].split("(")[0], k)], capture_output=True, text=True).stdout
        md += ["", "Source lines by warp-stall samples (`ncu --page source --print-source cuda,sass`): % of samples, % of executed warp instructions,",
               "active threads per instruction:", "", "```", lines.rstrip(), "```", ""]
        open(os.path.join(P, "%s_%s_ncu.md" % (out, label)), "w").write("\n".join(md))
        print("wrote", label)
        if base.startswith("seed_front_kernel"):
            rd = float(got["dram__bytes_read.sum"][0]) * SCALE[got["dram__bytes_read.sum"][1]]
            wr = float(got["dram__bytes_write.sum"][0]) * SCALE[got["dram__bytes_write.sum"][1]]
            json.dump({"kernel": "seed_front_kernel", "pairs_per_step": 2000000, "ref_bp": 3000000000, "preset": "chip", "dram_bytes_per_launch": rd + wr,
                       "dram_bytes_read": rd, "dram_bytes_write": wr, "warp_instructions_per_launch": float(got["smsp__inst_executed.sum"][0]),
                       "duration_ms_under_ncu": float(got["gpu__time_duration.sum"][0]), "source": "profiles/%s_%s_ncu.md" % (out, label)},
                      open(os.path.join(P, "front_kernel_ncu.json"), "w"), indent=1)


if __name__ == "__main__":
    tag, out = sys.argv[1], sys.argv[2]
    launch_list(tag, out)
    if os.path.exists(os.path.join(G, tag + "_cap.raw.csv.gz")):
        kernels(tag, out, sys.argv[3:])
