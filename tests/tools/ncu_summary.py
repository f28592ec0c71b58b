"""Print the key per-kernel metrics of an ncu report.  Usage: python tests/tools/ncu_summary.py <report.ncu-rep>"""
import csv
import io
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_bytes.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__inst_executed.avg.per_cycle_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_bytes.sum']


def main():
    txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("---", r[idx['Kernel Name']][:60])
        for w in WANT:
            if w in idx:
                print(f"   {w:75s} {r[idx[w]]:>16s} {units[idx[w]]}")


if __name__ == "__main__":
    main()
