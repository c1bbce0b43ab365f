"""Compact text summary of an .ncu-rep (run where ncu is installed): key throughput metrics, stall
reasons per issue, and per-opcode stall-sample shares from the SASS page.

    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep [--sass] [--top N]
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
]


def page(rep, which, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", which, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    rows = page(rep, "raw")
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print("==", d.get("Kernel Name", "?")[:100])
        for k in KEYS:
            if k in d and d[k] != "":
                print(f"  {k:78s} {d[k]} {u[k]}")
        st = []
        for k in hdr:
            if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k:
                try:
                    st.append((float(d[k]), k.split("issue_stalled_")[1].split("_per_issue")[0]))
                except ValueError:
                    pass
        st.sort(reverse=True)
        print("  stalls per issue:", ", ".join(f"{n} {v:.2f}" for v, n in st[:9]))
    if "--sass" in sys.argv:
        top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 12
        rows = page(rep, "source", ["--print-source", "sass"])
        # one block per kernel: a "Kernel Name" row, a header row, then instruction rows
        i = 0
        while i < len(rows):
            if rows[i] and rows[i][0] == "Kernel Name":
                name, hdr = rows[i][1], rows[i + 1]
                j = i + 2
                data = []
                while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
                    if len(rows[j]) == len(hdr):
                        data.append(rows[j])
                    j += 1
                i = j
                cnt, ex = collections.Counter(), collections.Counter()
                for r in data:
                    op = r[1].split()
                    o = (op[1] if op[0].startswith("@") else op[0]).split(".")[0]
                    cnt[o] += int(r[2] or 0)
                    ex[o] += int(r[5] or 0)
                tot = sum(cnt.values()) or 1
                print("== SASS", name[:80], "samples", tot)
                for o, c in cnt.most_common(top):
                    print(f"  {o:10s} samples {100 * c / tot:5.1f}%  executed {ex[o]}")
                hot = sorted(range(len(data)), key=lambda k: -int(data[k][2] or 0))[:top]
                for k in sorted(hot):
                    print(f"  [{k:4d}] {data[k][1][:70]:70s} {data[k][2]}")
            else:
                i += 1


if __name__ == "__main__":
    main()
