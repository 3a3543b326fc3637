"""Mnemonic counts per kernel of the built library (profiles/rNN_sass_evidence.txt):  python tools/sass_evidence.py > file"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, 'oflibnumpy_b200', 'lib', 'liboflib_b200.so')
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
print("SASS evidence of liboflib_b200.so: cuobjdump -sass oflibnumpy_b200/lib/liboflib_b200.so, mnemonic counts per kernel")
print("(UTMALDG / UTMASTG = TMA tensor loads / stores, SYNCS = mbarrier operations, REDUX = warp reductions, ELECT = "
      "elect.sync, DFMA/DADD/DMUL = float64 pipe)\n")
name, cnt = None, {}


def flush():
    if name is not None:
        print("%-110s UTMALDG %3d UTMASTG %3d SYNCS %4d REDUX %3d ELECT %2d F64 %4d" % (
            name, cnt.get('UTMALDG', 0), cnt.get('UTMASTG', 0), cnt.get('SYNCS', 0), cnt.get('REDUX', 0),
            cnt.get('ELECT', 0), cnt.get('F64', 0)))


for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        flush()
        name, cnt = m.group(1), {}
        continue
    m = re.match(r'\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)', line)
    if m:
        op = m.group(1)
        for k in ('UTMALDG', 'UTMASTG', 'SYNCS', 'ELECT'):
            if op.startswith(k):
                cnt[k] = cnt.get(k, 0) + 1
        if 'REDUX' in op:
            cnt['REDUX'] = cnt.get('REDUX', 0) + 1
        if op.split('.')[0] in ('DFMA', 'DADD', 'DMUL', 'DSETP', 'DMNMX'):
            cnt['F64'] = cnt.get('F64', 0) + 1
flush()
