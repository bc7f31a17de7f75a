#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small text table: per distinct kernel the duration, DRAM /
L2 traffic and throughput, occupancy, issue utilisation, instruction count and the warp-stall breakdown.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv
import io
import subprocess
import sys

WANT = [
    ("duration_us", "gpu__time_duration.sum"),
    ("dram_read_MB", "dram__bytes_read.sum"),
    ("dram_write_MB", "dram__bytes_write.sum"),
    ("dram_throughput_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l2_sectors_read_from_sm", "lts__t_sectors_srcunit_tex_op_read.sum"),
    ("l2_throughput_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l2_hit_rate_pct", "lts__t_sector_hit_rate.pct"),
    ("l1_hit_rate_pct", "l1tex__t_sector_hit_rate.pct"),
    ("l1tex_throughput_pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("warp_instructions", "smsp__inst_executed.sum"),
    ("tensor_pipe_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("registers", "launch__registers_per_thread"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("sm_cycles", "sm__cycles_elapsed.max"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    groups = {}
    for r in rows[2:]:
        groups.setdefault(r[idx["Kernel Name"]], []).append(r)
    print(f"# {rep}: {len(rows) - 2} profiled launches, {len(groups)} distinct kernels (ncu --set full, "
          f"--clock-control none; per-launch values are means over the launches of a kernel)")
    for name, rs in groups.items():
        print(f"\n## {name[:140]}   [{len(rs)} launches]")
        for label, key in WANT:
            if key not in idx:
                continue
            vals = []
            for r in rs:
                try:
                    vals.append(float(r[idx[key]].replace(",", "")))
                except ValueError:
                    pass
            if vals:
                u = units[idx[key]]
                print(f"  {label:28s} {sum(vals) / len(vals):14.3f} {u}")
        stalls = []
        for h in hdr:
            if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
                vals = [float(r[idx[h]].replace(",", "")) for r in rs if r[idx[h]] not in ("", "n/a")]
                if vals:
                    stalls.append((sum(vals) / len(vals), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
        stalls.sort(reverse=True)
        tot = sum(v for v, _ in stalls) or 1.0
        print("  warp stalls (warps per issue-active cycle; share of warp time): "
              + ", ".join(f"{n} {v:.2f} ({100 * v / tot:.0f}%)" for v, n in stalls[:8]))


if __name__ == "__main__":
    main()
