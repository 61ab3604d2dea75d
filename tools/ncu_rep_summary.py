"""Text summary of `ncu --set full` captures (.ncu-rep): per launch duration, tensor-pipe activity,
DRAM read / write bytes, L2->SM bytes, SM / DRAM throughput, issue activity.
    python tools/ncu_rep_summary.py a.ncu-rep b.ncu-rep ...    (needs the ncu CLI)"""
import csv
import io
import subprocess
import sys

WANT = [('gpu__time_duration.sum', 'duration'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe active %'),
        ('dram__bytes_read.sum', 'dram read'), ('dram__bytes_write.sum', 'dram write'),
        ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'dram throughput %'),
        ('l1tex__m_xbar2l1tex_read_bytes.sum', 'L2->SM (xbar2l1tex) read'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm throughput %'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue active %'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'),
        ('launch__registers_per_thread', 'registers')]


def main(paths):
    for p in paths:
        out = subprocess.run(['ncu', '-i', p, '--page', 'raw', '--csv'], capture_output=True,
                             text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            print(f'=== {p}: no launches')
            continue
        head, units = rows[0], rows[1]
        col = {n: i for i, n in enumerate(head)}
        print(f'=== {p}')
        for r in rows[2:]:
            name = r[col['Kernel Name']].split('(')[0]
            parts = [name]
            for m, label in WANT:
                if m in col:
                    parts.append(f'{label} {r[col[m]]} {units[col[m]]}'.rstrip())
            print('   ' + '; '.join(parts))


if __name__ == '__main__':
    main(sys.argv[1:])
