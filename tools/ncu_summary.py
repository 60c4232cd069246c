"""Summarise an .ncu-rep (raw page) into a small markdown table for profiles/."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.max"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary of `{rep.split('/')[-1]}` (one column per captured launch)\n\n")
        name_i = hdr.index("Kernel Name")
        f.write("| metric | " + " | ".join(r[name_i].split("(")[0] for r in rows[2:]) + " |\n")
        f.write("|---|" + "---|" * len(rows[2:]) + "\n")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                f.write(f"| {w} [{units[i]}] | " + " | ".join(r[i] for r in rows[2:]) + " |\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
