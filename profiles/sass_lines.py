"""Static SASS instruction count per source line of one kernel: python profiles/sass_lines.py <object.o> <kernel-name-substring> [source.cu]
(cuobjdump -xelf + nvdisasm --print-line-info; the object must be compiled with -lineinfo).  Used on the CPU box to see where a kernel's
instructions go before spending GPU time."""
import os
import re
import subprocess
import sys
import tempfile

obj, pat = os.path.abspath(sys.argv[1]), sys.argv[2]
src = sys.argv[3] if len(sys.argv) > 3 else None
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=d, check=True, capture_output=True)
    cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(d, cub)], capture_output=True, text=True).stdout
secs = [m.start() for m in re.finditer(r"\.section\s+\.text\.", txt)]
for i, s in enumerate(secs):
    e = secs[i + 1] if i + 1 < len(secs) else len(txt)
    head = txt[s : txt.index("\n", s)]
    if pat not in head:
        continue
    line, cnt, ops = None, {}, {}
    for l in txt[s:e].splitlines():
        m = re.search(r'//## File "(.*?)", line (\d+)', l)
        if m:
            f = os.path.basename(m.group(1))
            line = int(m.group(2)) if (src is None or f == os.path.basename(src)) else f"{f}:{m.group(2)}"
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", l)
        if m:
            cnt[line] = cnt.get(line, 0) + 1
            ops[m.group(1)] = ops.get(m.group(1), 0) + 1
    print(head.strip()[:160])
    print("total", sum(cnt.values()))
    own = {k: v for k, v in cnt.items() if isinstance(k, int)}
    print("by line:", " ".join(f"{k}:{v}" for k, v in sorted(own.items())))
    print("other files:", sum(v for k, v in cnt.items() if not isinstance(k, int)))
    print("top ops:", " ".join(f"{k}:{v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:25]))
    break
