"""Code size (bytes) of every kernel in an object / shared library: python scratch/ksize.py file [substring]"""
import re, subprocess, sys
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
name, last, res = None, 0, []
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        if name: res.append((last + 16, name))
        name, last = m.group(1), 0
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/", ln)
    if m: last = int(m.group(1), 16)
if name: res.append((last + 16, name))
for sz, n in sorted(res):
    if len(sys.argv) < 3 or sys.argv[2] in n: print(f"{sz:8d}  {n[:90]}")
