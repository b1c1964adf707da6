"""Where does a kernel touch local memory?  Disassembles nvit_b200/libnvit_b200.so (built with -lineinfo) and lists the source
lines carrying LDL / STL instructions of the kernels whose mangled name contains the given substring.
Usage: python scripts/spill_lines.py attn_bwd_ws3 [path/to/lib.so]"""
import collections
import os
import re
import subprocess
import sys
import tempfile

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pat = sys.argv[1]
lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(root, "nvit_b200", "libnvit_b200.so")
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, capture_output=True)
    for f in sorted(os.listdir(d)):
        txt = subprocess.run(["nvdisasm", "-g", os.path.join(d, f)], capture_output=True, text=True).stdout
        sect, cur = None, None
        hits = collections.defaultdict(collections.Counter)
        for line in txt.split("\n"):
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
            if m:
                sect = m.group(1) if pat in m.group(1) else None
                continue
            if line.lstrip().startswith(".section"):
                sect = None
            if sect is None:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', line)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
            elif re.search(r"\b(LDL|STL)\b", line):
                hits[sect][cur] += 1
        for k, c in hits.items():
            print(k)
            for (fn, ln), n in sorted(c.items()):
                src = ""
                path = os.path.join(root, "nvit_b200", "csrc", fn)
                if os.path.exists(path):
                    src = open(path).read().split("\n")[ln - 1].strip()[:110]
                print(f"  {fn}:{ln}  x{n}  {src}")
