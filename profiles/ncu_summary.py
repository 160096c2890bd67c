#!/usr/bin/env python
"""Turn an .ncu-rep (ncu --set full --import-source on ...) into the compact text summary kept under profiles/.

    python profiles/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/r01_<what>.txt

Per captured launch: duration, DRAM bytes, L2 / L1 hit rates, issue-slot utilisation, occupancy, the executed
instruction mix (per warp) and the SASS lines that collect the most warp-stall samples.  Needs `ncu` on PATH (it
reads reports without a GPU)."""
import csv
import io
import subprocess
import sys
from collections import Counter

RAW = [
    ("duration_us", "gpu__time_duration.sum"),
    ("dram_read_MB", "dram__bytes_read.sum"), ("dram_write_MB", "dram__bytes_write.sum"),
    ("dram_throughput_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l2_hit_pct", "lts__t_sector_hit_rate.pct"), ("l1_ld_hit_pct", "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct"),
    ("l2_throughput_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("issue_slots_busy_pct", "sm__inst_issued.avg.pct_of_peak_sustained_active"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("registers_per_thread", "launch__registers_per_thread"),
    ("warp_instructions", "smsp__inst_issued.sum"),
    ("threads_per_instruction", "smsp__thread_inst_executed_per_inst_executed.ratio"),
    ("stall_long_scoreboard_per_issue", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    ("stall_short_scoreboard_per_issue", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
    ("stall_wait_per_issue", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
    ("stall_barrier_per_issue", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
    ("stall_mio_throttle_per_issue", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"),
    ("stall_lg_throttle_per_issue", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"),
    ("stall_math_throttle_per_issue", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
    ("pipe_fp64_pct", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    ("pipe_alu_pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    ("pipe_fma_pct", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    ("pipe_lsu_pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    ("pipe_xu_pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
    ("local_ld_sectors", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum"), ("local_st_sectors", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum"),
]


def ncu(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep):
    raw = ncu(rep, "raw")
    hdr, units = raw[0], raw[1]
    launches = raw[2:]
    src = ncu(rep, "source")
    sections, shdr = [], None
    for r in src:
        if r and r[0] == "Kernel Name":
            sections.append([]); continue
        if r and r[0] == "Address":
            shdr = r; continue
        if sections and shdr and len(r) == len(shdr):
            sections[-1].append(r)
    sections = sections[::2] if len(sections) == 2 * len(launches) else sections     # ncu prints each kernel's listing twice
    print("# %s" % rep)
    for n, r in enumerate(launches):
        get = lambda k: r[hdr.index(k)] if k in hdr else "n/a"
        print("\n== launch %d: %s  grid %s block %s" % (n, get("Kernel Name")[:90], get("Grid Size"), get("Block Size")))
        for name, key in RAW:
            if key in hdr:
                v, u = r[hdr.index(key)], units[hdr.index(key)]
                try:
                    f = float(v.replace(",", ""))
                    if name.endswith("_MB"):
                        f = f * {"Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "byte": 1e-6}.get(u, 1.0)
                    v = "%.4g" % f
                except ValueError:
                    pass
                print("   %-34s %s" % (name, v))
        if n < len(sections) and shdr:
            data = sections[n]
            iA, iS, iT, iSrc = shdr.index("Instructions Executed"), shdr.index("# Samples"), shdr.index("Thread Instructions Executed"), shdr.index("Source")
            warps = max((int(x[iA]) for x in data), default=1) or 1
            mix, tot = Counter(), 0
            for x in data:
                e = int(x[iA])
                if not e:
                    continue
                s = x[iSrc].strip()
                if s.startswith("@"):
                    s = s.split(None, 1)[1]
                mix[s.split()[0].split(".")[0]] += e
                tot += e
            print("   executed warp-instructions per warp: %.1f  (%s)" % (tot / warps, ", ".join("%s %.1f" % (o, e / warps) for o, e in mix.most_common(10))))
            samples = sum(int(x[iS]) for x in data) or 1
            print("   top stall locations (%% of %d samples | avg active threads | SASS):" % samples)
            for i in sorted(sorted(range(len(data)), key=lambda i: -int(data[i][iS]))[:8]):
                e = int(data[i][iA])
                print("      %5.1f%%  %4.0f  [%d] %s" % (100.0 * int(data[i][iS]) / samples, int(data[i][iT]) / max(e, 1), i, data[i][iSrc].strip()[:80]))


if __name__ == "__main__":
    main(sys.argv[1])
