"""
Turns an .ncu-rep (brought back from the GPU box in gpurun_out/) into the small text summaries committed under
profiles/:  python scripts/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r01_name [--top 25]
Writes <out>_metrics.txt (per-kernel launch metrics) and <out>_hot_sass.txt (instructions with the most stall samples).
"""

import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum",
    "launch__grid_size",
    "launch__block_size",
    "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "smsp__mem_tensor_reads_op_ldt.sum",
]


def run(args):
    return subprocess.run(["ncu", "-i", *args], capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    rows = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    with open(out + "_metrics.txt", "w") as f:
        f.write(f"# ncu --set full --clock-control none, from {rep}\n")
        for r in rows[2:]:
            f.write("\n" + r[hdr.index("Kernel Name")][:140] + "\n")
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    f.write(f"  {m:100s} {r[i]:>18s} {units[i]}\n")
    src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv"]))))
    with open(out + "_hot_sass.txt", "w") as f:
        kernel, body = None, []

        def flush():
            if kernel is None or not body:
                return
            total = sum(int(b[2] or 0) for b in body)
            f.write(f"\n{kernel}\n  total warp-stall samples: {total}\n  samples  executed  avg-threads  instruction\n")
            for b in sorted(body, key=lambda b: -int(b[2] or 0))[:top]:
                f.write(f"  {b[2]:>7s} {b[5]:>10s} {b[8]:>4s}  {b[1].strip()[:120]}\n")

        for r in src:
            if r and r[0] == "Kernel Name":
                flush()
                kernel, body = r[1], []
            elif r and r[0].startswith("0x"):
                body.append(r)
        flush()


if __name__ == "__main__":
    main()
