"""Record the DRAM traffic of a kernel from an `ncu --set full` report in
profiles/ncu_traffic.json, keyed by kernel and by the hash of the kernel
sources it was captured from (bench.py quotes the figure only while the hash
matches, so it cannot silently go stale).

usage: python tools/ncu_traffic.py report.ncu-rep "kernel regex" bench_name \
           samples_per_launch [note]
"""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}


def main():
    rep, regex, name, nsamp = sys.argv[1:5]
    note = sys.argv[5] if len(sys.argv) > 5 else ''
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], dict(zip(rows[0], rows[1]))
    picked = [dict(zip(hdr, r)) for r in rows[2:]
              if re.search(regex, dict(zip(hdr, r))['Kernel Name'])]
    if not picked:
        raise SystemExit('no kernel matches ' + regex)
    d = picked[-1]

    def nbytes(key):
        return float(d[key]) * UNIT[units[key]]

    from bench import DECODE_SOURCES, _source_hash
    path = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
    table = json.load(open(path)) if os.path.exists(path) else {}
    git = subprocess.run(['git', 'rev-parse', '--short', 'HEAD'], cwd=ROOT,
                         capture_output=True, text=True).stdout.strip()
    table[name] = {
        'ncu_kernel_name': d['Kernel Name'], 'grid': d['Grid Size'],
        'report': os.path.basename(rep), 'git': git,
        'sources': list(DECODE_SOURCES),
        'csrc_sha16': _source_hash(DECODE_SOURCES),
        'dram_bytes_read': nbytes('dram__bytes_read.sum'),
        'dram_bytes_write': nbytes('dram__bytes_write.sum'),
        'gpu_time_us_under_ncu': float(d['gpu__time_duration.sum'])
        * {'ms': 1e3, 'us': 1.0, 'ns': 1e-3, 's': 1e6}[
            units['gpu__time_duration.sum']],
        'samples_per_launch': int(nsamp), 'note': note}
    with open(path, 'w') as fh:
        json.dump(table, fh, indent=1, sort_keys=True)
    e = table[name]
    print(name, (e['dram_bytes_read'] + e['dram_bytes_write'])
          / e['samples_per_launch'], 'B/sample')


if __name__ == '__main__':
    main()
