"""Instruction mix of every kernel in libbaseband_b200.so (cuobjdump -sass).
usage: python tools/sass_summary.py [regex]"""
import collections
import re
import subprocess
import sys

LIB = 'baseband_b200/libbaseband_b200.so'
want = re.compile(sys.argv[1]) if len(sys.argv) > 1 else None
sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True,
                      text=True).stdout
names = {}
name = None
for line in sass.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        name = m.group(1)
        names[name] = collections.Counter()
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m and name:
        op = m.group(1)
        base = op.split('.')[0]
        if base in ('LDG', 'STG', 'LDS', 'STS', 'SHFL'):
            base = '.'.join(p for p in op.split('.')
                            if p in (base, '64', '128', 'IDX'))
        names[name][base] += 1
demangle = subprocess.run(['c++filt'] + list(names), capture_output=True,
                          text=True).stdout.splitlines()
for mangled, pretty in sorted(zip(names, demangle), key=lambda t: t[1]):
    if want and not want.search(pretty):
        continue
    c = names[mangled]
    total = sum(c.values())
    print(re.sub(r'\(.*', '', pretty))
    print('  %d instructions: %s' % (total, ', '.join(
        '%s %d' % kv for kv in c.most_common(14))))
