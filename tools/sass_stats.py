#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of a cubin / object / shared library (uses cuobjdump)."""
import collections
import re
import subprocess
import sys


def sass_stats(path):
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    cur, cnt = None, collections.OrderedDict()
    for l in txt.splitlines():
        m = re.search(r"Function : (\S+)", l)
        if m:
            cur = m.group(1)
            cnt[cur] = collections.Counter()
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
        if m and cur:
            cnt[cur][m.group(1)] += 1
    return cnt


if __name__ == "__main__":
    ops = sys.argv[2].split(",") if len(sys.argv) > 2 else None
    for f, c in sass_stats(sys.argv[1]).items():
        sel = {k: c[k] for k in ops if c[k]} if ops else dict(c.most_common(12))
        print("%-90s n=%-6d %s" % (f[:90], sum(c.values()), sel))
