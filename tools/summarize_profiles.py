#!/usr/bin/env python
"""Turn the raw captures of tools/round_profiles.sh (gpurun_out/<tag>_*) into the committed summaries under profiles/."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def launch_list(tag):
    rows = list(csv.reader(open(os.path.join(ROOT, "gpurun_out", tag + "_launches.csv"))))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    L = [(r[4].split("(")[0], float(r[-1])) for r in rows[hdr + 2:]]
    starts = [i for i, (n, v) in enumerate(L) if n == "prep_kernel" and v > 100000]
    call = L[starts[-1]:]
    tot = sum(v for _, v in call)
    out = ["# Round %s - ncu launch list of one `cmx_map_batch_pe` call" % tag, "",
           "`CMX_LANES=1 ncu --metrics gpu__time_duration.sum --clock-control none python tools/profile_run.py --calls 2` (3 Gbp reference,",
           "2 M pairs, --preset chip), the second call; one lane, so the kernels run back to back.  Per-launch times under ncu are",
           "cold-cache and serialised: compare SHARES with `kernel_ms_per_step` of the bench line, not absolutes.", "",
           "| # | kernel | ms | share |", "|---|---|---|---|"]
    for i, (n, v) in enumerate(call):
        out.append("| %d | `%s` | %.3f | %.1f %% |" % (i, n[:48], v / 1e6, 100 * v / tot))
    out += ["", "total %.2f ms" % (tot / 1e6), ""]
    open(os.path.join(ROOT, "profiles", tag + "_launch_list.md"), "w").write("\n".join(out))
    return {n: v / 1e6 for n, v in call if n == "probe_kernel"}, tot / 1e6


def probe_full(tag, pairs_per_step, ref_bp, preset):
    rep = os.path.join(ROOT, "gpurun_out", tag + "_probe.ncu-rep")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, units, vals = rows[0], rows[1], rows[2]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__occupancy_limit_registers", "smsp__cycles_active.avg", "l1tex__t_sector_hit_rate.pct",
            "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]
    got = {}
    for w in want:
        if w in h:
            i = h.index(w)
            got[w] = (vals[i], units[i])
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd = float(got["dram__bytes_read.sum"][0]) * scale[got["dram__bytes_read.sum"][1]]
    wr = float(got["dram__bytes_write.sum"][0]) * scale[got["dram__bytes_write.sum"][1]]
    md = ["# Round %s - `ncu --set full` of `probe_kernel` (one launch, 2 M pairs, 3 Gbp index)" % tag, "",
          "`CMX_LANES=1 ncu --set full --clock-control none --import-source on -k regex:probe_kernel --launch-skip 1 --launch-count 1`", "",
          "| metric | value | unit |", "|---|---|---|"]
    for w in want:
        if w in got:
            md.append("| `%s` | %s | %s |" % (w, got[w][0], got[w][1]))
    md += ["", "DRAM traffic of the launch: %.3f GB read + %.3f GB written = **%.3f GB**." % (rd / 1e9, wr / 1e9, (rd + wr) / 1e9), ""]
    stalls = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, "0", "12"], capture_output=True, text=True).stdout
    md += ["Warp-stall samples by source line (top 12):", "", "```", stalls.rstrip(), "```", ""]
    open(os.path.join(ROOT, "profiles", tag + "_probe_kernel_ncu.md"), "w").write("\n".join(md))
    json.dump({"kernel": "probe_kernel", "pairs_per_step": pairs_per_step, "ref_bp": ref_bp, "preset": preset, "dram_bytes_per_launch": rd + wr,
               "dram_bytes_read": rd, "dram_bytes_write": wr, "duration_ms_under_ncu": float(got["gpu__time_duration.sum"][0]),
               "source": "profiles/%s_probe_kernel_ncu.md" % tag}, open(os.path.join(ROOT, "profiles", "probe_kernel_ncu.json"), "w"), indent=1)
    return rd + wr


def main():
    tag = sys.argv[1]
    b = json.load(open(os.path.join(ROOT, "gpurun_out", tag + "_bench_N1.json")))
    traffic = probe_full(tag, b["config"]["pairs_per_step"], b["config"]["ref_bp"], b["config"]["preset"])
    probe, tot = launch_list(tag)
    # the bench line was produced before this capture existed: complete its roofline.traffic here (same workload, same code)
    if b["roofline"].get("traffic") is None:
        b["roofline"]["traffic"] = traffic
        b["roofline"]["traffic_source"] = "profiles/%s_probe_kernel_ncu.md (ncu --set full of the same command, one launch)" % tag
    json.dump(b, open(os.path.join(ROOT, "profiles", "bench_%s_N1.json" % tag), "w"))
    r = json.load(open(os.path.join(ROOT, "gpurun_out", tag + "_bench_reference_arm.json")))
    json.dump(r, open(os.path.join(ROOT, "profiles", "bench_%s_reference_arm.json" % tag), "w"))
    print("value %.1fM e2e %.1fM reference %.3fM  probe share ncu %.3f vs events %.3f  traffic %.2f GB" % (
        b["value"] / 1e6, b["e2e"]["value"] / 1e6, r["value"] / 1e6, list(probe.values())[0] / tot, b["roofline"]["kernel_share_of_step"], traffic / 1e9))


if __name__ == "__main__":
    main()
