"""Turn gpurun_out/*.ncu-rep + launch lists into the small text summaries committed here.
usage: python profiles/summarize.py <rep> <out.md> [title]"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_subunit_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else rep
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# {title}\n\nsource: `{rep}` (ncu --set full --clock-control none)\n\n")
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            f.write(f"## {name}\n\n| metric | value | unit |\n|---|---|---|\n")
            for i, h in enumerate(hdr):
                if h in WANT or "stall" in h and h.endswith("_per_warp_active.pct"):
                    f.write(f"| {h} | {r[i]} | {units[i]} |\n")
            f.write("\n")


if __name__ == "__main__":
    main()
