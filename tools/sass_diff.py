#!/usr/bin/env python
"""Which kernels differ between two builds of liblars_b200.so?  (cuobjdump -sass per function, whitespace-normalised.)
Used to show that a change to one kernel / to host code left every other kernel's machine code untouched:
    python tools/sass_diff.py old.so new.so"""
import collections
import re
import subprocess
import sys


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    table, name = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            table[name] = []
        elif name and "/*" in line:
            table[name].append(re.sub(r"\s+", " ", line).strip())
    return table


a, b = kernels(sys.argv[1]), kernels(sys.argv[2])
changed = [k for k in a if k in b and a[k] != b[k]]
print(f"{len(a)} / {len(b)} kernels; changed: {changed or 'none'}; only in the first: {[k for k in a if k not in b] or 'none'}; "
      f"only in the second: {[k for k in b if k not in a] or 'none'}")
sys.exit(1 if changed else 0)
