"""Summarise an .ncu-rep (ncu --set full) into a small CSV/markdown table for profiles/: per kernel launch
duration, tensor-pipe %, DRAM bytes, DRAM %, L2 %, registers, issue-active %."""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
H, U, data = rows[0], rows[1], rows[2:]
want = [("gpu__time_duration.sum", "time"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_%"),
        ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"), ("launch__registers_per_thread", "regs"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_%"), ("sm__cycles_elapsed.avg", "sm_cycles"),
        ("l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "tma_load_bytes"), ("launch__grid_size", "grid"),
        ("launch__block_size", "block"), ("smsp__warps_eligible.avg.per_cycle_active", "eligible_warps")]
ki = H.index("Kernel Name")
print("| kernel | " + " | ".join(n for _, n in want) + " |")
print("|---|" + "---|" * len(want))
for d in data:
    cells = []
    for m, _ in want:
        if m in H:
            i = H.index(m)
            cells.append(f"{d[i]} {U[i]}".strip())
        else:
            cells.append("n/a")
    print(f"| {d[ki][:70]} | " + " | ".join(cells) + " |")
