"""Summarise an `ncu --page raw --csv` export: one block of the listed metrics per captured launch.
python tools/ncu_summary.py raw.csv [--csv out.csv]"""
import csv
import sys

WANT = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'smsp__inst_executed.sum', 'launch__grid_size', 'launch__block_size',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum', 'dram__sectors_read.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct', 'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_barrier_per_warp_active.pct',
        'smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_wait_per_warp_active.pct',
        'smsp__warp_issue_stalled_not_selected_per_warp_active.pct', 'smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct']


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    cols = [(w, hdr.index(w)) for w in WANT if w in hdr]
    out = None
    if "--csv" in sys.argv:
        out = csv.writer(open(sys.argv[sys.argv.index("--csv") + 1], "w"))
        out.writerow([w for w, _ in cols])
        out.writerow([units[i] for _, i in cols])
    for r in rows[2:]:
        if out:
            out.writerow([r[i] for _, i in cols])
        else:
            print("---")
            for w, i in cols:
                print(f"  {w}: {r[i]} {units[i]}")


if __name__ == "__main__":
    main()
