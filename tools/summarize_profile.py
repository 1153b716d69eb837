"""Summarise gpurun_out ncu artefacts into profiles/ (tracked).

    python tools/summarize_profile.py <tag> <launches.csv> <report.ncu-rep> [photons_per_launch] [kernel_hash]

Writes profiles/<tag>_launches.csv (kernel share table), profiles/<tag>_trace_kernel.md (key ncu
metrics, opcode mix, hottest source lines) and updates profiles/traffic.json (dram bytes/photon,
read by bench.py for roofline.traffic)."""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    d = collections.defaultdict(list)
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(',', ''))
        except ValueError:
            continue
        scale = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'nsecond': 1e-6, 'usecond': 1e-3, 'msecond': 1.0}.get(r[ui], 1e-6)
        d[r[ki]].append(v * scale)
    tot = sum(sum(v) for v in d.values())
    out = [('kernel', 'launches', 'total_ms', 'share', 'avg_ms')]
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        out.append((k[:110], len(v), round(sum(v), 4), round(sum(v) / tot, 4), round(sum(v) / len(v), 4)))
    return out


def raw_metrics(rep):
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    return {n: (u, v) for n, u, v in zip(rows[0], rows[1], rows[2])}


def _num(x):
    try:
        return float(x or 0)
    except ValueError:
        return 0.0


def source_tables(rep):
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    cur, hdr, lines, ops = None, None, [], collections.Counter()
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Path':
            cur = r[1].split('/')[-1]
            continue
        if r[0] == 'Line No':
            hdr = r
            ii, si = hdr.index('Instructions Executed'), hdr.index('# Samples')
            continue
        if hdr is None:
            continue
        if r[0].isdigit():
            lines.append((cur, int(r[0]), r[1].strip(), _num(r[ii]), _num(r[si])))
        elif r[0] == '' and len(r) > ii and r[2].startswith('0x'):
            m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[3])
            try:
                if m:
                    ops[m.group(2).split('.')[0]] += _num(r[ii])
            except ValueError:
                pass
    return lines, ops


def main():
    tag, lpath, rep = sys.argv[1:4]
    nph = float(sys.argv[4]) if len(sys.argv) > 4 else 1e7
    os.makedirs(os.path.join(ROOT, 'profiles'), exist_ok=True)
    tab = launches(lpath)
    with open(os.path.join(ROOT, 'profiles', tag + '_launches.csv'), 'w') as f:
        csv.writer(f).writerows(tab)
    m = raw_metrics(rep)

    def val(name):
        u, v = m[name]
        v = float(v.replace(',', ''))
        return v * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1, 'ms': 1e-3, 'us': 1e-6, 'ns': 1e-9,
                    'msecond': 1e-3, 'usecond': 1e-6, 'nsecond': 1e-9}.get(u, 1)
    rd, wr, dur = val('dram__bytes_read.sum'), val('dram__bytes_write.sum'), val('gpu__time_duration.sum')
    inst = val('smsp__inst_executed.sum')
    lines, ops = source_tables(rep)
    tot_i = sum(l[3] for l in lines) or 1
    tot_s = sum(l[4] for l in lines) or 1
    tot_o = sum(ops.values()) or 1
    keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
            'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
            'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
            'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
            'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__block_size',
            'launch__grid_size', 'smsp__inst_executed.sum', 'sass__inst_executed_local_loads',
            'sass__inst_executed_local_stores', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
            'sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active',
            'sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active',
            'sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active',
            'sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active']
    with open(os.path.join(ROOT, 'profiles', tag + '_trace_kernel.md'), 'w') as f:
        f.write('# {0}: {2}, ncu --set full --clock-control none ({1:.0e} photons/launch)\n\n'.format(tag, nph, m.get('Kernel Name', ('', 'trace kernel'))[1] if 'Kernel Name' in m else 'trace kernel'))
        f.write('Source: `{0}` (scratch), launch list `{1}`.\n\n'.format(os.path.basename(rep), os.path.basename(lpath)))
        f.write('| metric | value | unit |\n|---|---|---|\n')
        for k in keys:
            if k in m:
                f.write('| {0} | {1} | {2} |\n'.format(k, m[k][1], m[k][0]))
        f.write('\nDerived: dram traffic {0:.1f} B/photon (read {1:.1f} + write {2:.1f}); {3:.0f} warp-level '
                'instructions x32 / photon; {4:.2e} photons/s under ncu (cold, serialised: compare shares only).\n'.format(
                    (rd + wr) / nph, rd / nph, wr / nph, inst * 32 / nph, nph / dur))
        f.write('\n## Share of the step (launch list)\n\n| kernel | launches | total ms | share | avg ms |\n|---|---|---|---|---|\n')
        for r in tab[1:8]:
            f.write('| `{0}` | {1} | {2} | {3} | {4} |\n'.format(r[0][:70], *r[1:]))
        f.write('\n## SASS opcode mix (executed warp instructions)\n\n| opcode | share | per photon (x32) |\n|---|---|---|\n')
        for k, v in ops.most_common(16):
            f.write('| {0} | {1:.2%} | {2:.0f} |\n'.format(k, v / tot_o, v * 32 / nph))
        f.write('\n## Hottest source lines\n\n| file:line | inst | samples | source |\n|---|---|---|---|\n')
        for l in sorted(lines, key=lambda l: -l[3])[:25]:
            f.write('| {0}:{1} | {2:.2%} | {3:.2%} | `{4}` |\n'.format(l[0], l[1], l[3] / tot_i, l[4] / tot_s,
                                                                    l[2][:90].replace('|', '\\|')))
    with open(os.path.join(ROOT, 'profiles', 'traffic.json'), 'w') as f:
        json.dump({'tag': tag, 'dram_bytes_per_photon': (rd + wr) / nph, 'read': rd / nph, 'write': wr / nph,
                   'kernel_ms_under_ncu': dur * 1e3, 'kernel_hash': sys.argv[5] if len(sys.argv) > 5 else '?',
                   'capture': 'profiles/{0}_trace_kernel.md'.format(tag)}, f)
    print('wrote profiles/{0}_*'.format(tag))


if __name__ == '__main__':
    main()
