"""ncu report -> small text summary kept under profiles/ (the .ncu-rep files stay in gpurun_out/).
usage: python profiles/summarize.py gpurun_out/<name>.ncu-rep > profiles/<name>.txt"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# {path}: {len(rows) - 2} captured launches (ncu --set full --clock-control none)")
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")][:150])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:70s} {r[i]} {units[i]}")
        tensor = [h for h in hdr if "tensor" in h and "pct" in h]
        for k in tensor[:6]:
            if k not in KEYS:
                i = hdr.index(k)
                print(f"  {k:70s} {r[i]} {units[i]}")


if __name__ == "__main__":
    main(sys.argv[1])
