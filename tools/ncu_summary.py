"""Summarise an .ncu-rep (read on the CPU box with `ncu -i`) into the few metrics the roofline
discussion needs; one block per captured launch.  Usage: python tools/ncu_summary.py rep [rep ...]"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum",
]


def main():
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        print(f"# {rep}")
        for r in rows[2:]:
            name = r[col["Kernel Name"]]
            print(f"kernel: {name[:120]}  grid={r[col['Grid Size']]} block={r[col['Block Size']]}")
            for w in WANT:
                if w in col:
                    print(f"  {w:70s} {r[col[w]]:>16s} {units[col[w]]}")
            print()


if __name__ == "__main__":
    main()
