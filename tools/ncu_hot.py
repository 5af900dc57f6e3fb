"""Summarise one single-launch ncu report: headline metrics + the SASS lines with the most stall samples.
usage: python tools/ncu_hot.py gpurun_out/prof_x.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
det = subprocess.run(['ncu', '-i', rep, '--page', 'details'], capture_output=True, text=True).stdout
keys = ('Duration', 'SM Frequency', 'DRAM Throughput', 'L2 Cache Throughput', 'Compute (SM) Throughput', 'Executed Ipc Active',
        'Issue Slots Busy', 'Registers Per Thread', 'Theoretical Occupancy', 'Achieved Occupancy', 'Memory Throughput', 'L2 Hit Rate')
seen = set()
for line in det.splitlines():
    t = line.strip()
    for k in keys:
        if t.startswith(k) and k not in seen:
            seen.add(k)
            print(' '.join(t.split()))
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv', '--metrics',
                      'dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed.sum,gpu__time_duration.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
if len(rows) >= 3:
    for h, u, v in zip(rows[0], rows[1], rows[2]):
        if '__' in h:
            print('%s = %s %s' % (h, v, u))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ia, isamp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[isamp]) for r in data)
print('total samples %d, warp instructions %d' % (tot, sum(int(r[iex]) for r in data)))
agg = {}
for r in data:
    for i in stall_cols:
        if r[i] not in ('', '-'):
            agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i])
print('stalls:', sorted(agg.items(), key=lambda kv: -kv[1])[:8])
top = sorted(enumerate(data), key=lambda x: -int(x[1][isamp]))[:top_n]
for idx, r in sorted(top):
    st = sorted(((int(r[i]), hdr[i][6:]) for i in stall_cols if r[i] not in ('', '-')), reverse=True)[:2]
    print('%5d %6s %9s  %-64s %s' % (idx, r[isamp], r[iex], r[ia].strip()[:64], st))
