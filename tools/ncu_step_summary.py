"""Turns the raw CSV of one `ncu --set full` capture of a bench step (tools/gpu_ncu.sh) into the per-launch table and the
traffic record that bench.py reads.   usage: python tools/ncu_step_summary.py gpurun_out/prof_all_raw.csv v8 "note"
"""
import csv
import json
import sys

src, tag = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ''
rows = list(csv.reader(open(src)))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}


def g(r, name, default=0.0):
    try:
        return float(r[col[name]].replace(',', ''))
    except Exception:
        return default


def ghz(r):
    for nm in ('sm__cycles_elapsed.avg.per_second', 'smsp__cycles_elapsed.avg.per_second', 'gpc__cycles_elapsed.avg.per_second', 'gpc__cycles_elapsed.max.per_second'):
        if nm in col:
            v = g(r, nm)
            u = units[col[nm]]
            return v * {'Ghz': 1.0, 'GHz': 1.0, 'Mhz': 1e-3, 'MHz': 1e-3, 'hz': 1e-9, 'Hz': 1e-9, 'cycle/second': 1e-9, 'cycle/nsecond': 1.0, 'cycle/usecond': 1e-3}.get(u, 1e-9)
    return 0.0


units = rows[1]
out = ['ncu --set full --clock-control none; one bench step (B=16 images, T=10, 608x608), %s' % note,
       '%-26s %-5s %8s %9s %9s %8s %6s %6s %6s %6s' % ('kernel', 'layer', 'dur[us]', 'dramR[MB]', 'dramW[MB]', 'tc_pipe%', 'dram%', 'l2%', 'sm%', 'sm_GHz')]
conv_i, tot, conv_t, conv_b = 0, 0.0, 0.0, 0.0
units = rows[1]
for r in rows[2:]:
    name = r[col['Kernel Name']].split('(')[0].replace('byolo::', '').replace('void ', '')
    dur = g(r, 'gpu__time_duration.sum')
    if 'ns' in units[col['gpu__time_duration.sum']]:
        dur /= 1e3
    rd, wr = g(r, 'dram__bytes_read.sum'), g(r, 'dram__bytes_write.sum')
    for nm, val in (('dram__bytes_read.sum', 'rd'), ('dram__bytes_write.sum', 'wr')):
        u = units[col[nm]]
        f = {'Gbyte': 1e3, 'Mbyte': 1.0, 'Kbyte': 1e-3, 'byte': 1e-6}.get(u, 1.0)
        if val == 'rd':
            rd *= f
        else:
            wr *= f
    layer = ''
    if 'conv_umma' in name:
        conv_i += 1
        layer = 'L%d' % conv_i
        conv_t += dur
        conv_b += (rd + wr) * 1e6
    tot += dur
    out.append('%-26s %-5s %8.1f %9.1f %9.1f %8.1f %6.1f %6.1f %6.1f %6.2f' % (
        name[:26], layer, dur, rd, wr, g(r, 'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active') or
        g(r, 'sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active'),
        g(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'), g(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'),
        g(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'), ghz(r)))
out.append('conv launches: %d, sum duration %.1f us (%.1f%% of the step\'s kernel time), sum DRAM traffic %.1f MB' % (
    conv_i, conv_t, 100 * conv_t / tot if tot else 0, conv_b / 1e6))
open('profiles/r01/ncu_%s_step_summary.txt' % tag, 'w').write('\n'.join(out) + '\n')
json.dump({'source': 'profiles/r01/ncu_%s_step_summary.txt (ncu --set full, one step B=16 T=10 608x608)' % tag, 'conv_launches': conv_i,
           'conv_dram_bytes_per_step': conv_b, 'conv_share_of_kernel_time': conv_t / tot if tot else None},
          open('profiles/r01/ncu_%s_traffic.json' % tag, 'w'))
print('\n'.join(out[-3:]))
