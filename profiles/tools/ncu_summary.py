"""Key counters of an .ncu-rep as a markdown table (read on the CPU box with `ncu -i`).

    python profiles/tools/ncu_summary.py gpurun_out/<name>.ncu-rep [...] > profiles/<name>.md
"""
import csv
import io
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
        ("lts__t_sector_hit_rate.pct", "L2 hit rate"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % of max"),
        ("sm__icc_request_hit_rate.pct", "instruction cache (ICC) hit rate %"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction / issue"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard / issue"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
        ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall membar / issue"),
        ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch_resolving / issue"),
        ("launch__registers_per_thread", "registers / thread"), ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__cluster_size", "cluster size")]


def main():
    for rep in sys.argv[1:]:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            print(f"## `{rep.split('/')[-1]}`: `{name[:110]}`\n\n| metric | unit | value |\n|---|---|---|")
            for key, label in WANT:
                hits = [i for i, h in enumerate(hdr) if h == key]
                if hits:
                    print(f"| {label} (`{key}`) | {units[hits[0]]} | {r[hits[0]]} |")
            print()


if __name__ == "__main__":
    main()
