"""Attribute ncu per-SASS-instruction counters to CUDA source lines.

  python scripts/ncu_lines.py gpurun_out/prof.ncu-rep <kernel-substring> [top_n]
  NCU_KID='::regex:<base name>:<n>' (ncu --kernel-id) selects the n-th matching launch on the ncu side when the
  mangled substring that picks the SASS section cannot serve as ncu's kernel regex (template instances).

Joins `ncu --page source --csv` (SASS view: executed instructions, stall samples) with
`nvdisasm -g` line markers of the in-tree libccz_b200.so (compiled with -lineinfo)."""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kern = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 45
so = os.path.join(ROOT, "chinesechesszero_b200", "csrc", "libccz_b200.so")

with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=td, check=True, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=td, check=True, capture_output=True, text=True).stdout

line_of = {}
cur, inside = None, False
for ln in dis.splitlines():
    if ln.startswith("\t.section\t.text."):
        inside = kern in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
    if m:
        line_of[int(m.group(1), 16)] = cur

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", *(["--kernel-id", os.environ["NCU_KID"]] if os.environ.get("NCU_KID") else ["-k", f"regex:{kern}"])], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
base = None
inst, samp, thr = defaultdict(int), defaultdict(int), defaultdict(int)
tot_i = tot_s = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or not r[0].startswith("0x"):
        continue
    addr = int(r[0], 16)
    if base is None:
        base = addr
    key = line_of.get(addr - base)
    n = int(r[ci["Instructions Executed"]] or 0)
    s = int(r[ci["# Samples"]] or 0)
    t = int(r[ci["Thread Instructions Executed"]] or 0)
    inst[key] += n
    samp[key] += s
    thr[key] += t
    tot_i += n
    tot_s += s
src_cache = {}


def src(key):
    if key is None:
        return "?"
    f, l = key
    if f not in src_cache:
        p = os.path.join(ROOT, "chinesechesszero_b200", "csrc", f)
        src_cache[f] = open(p).read().splitlines() if os.path.exists(p) else []
    lines = src_cache[f]
    return lines[l - 1].strip()[:100] if 0 < l <= len(lines) else ""


print(f"total warp-instructions {tot_i}, samples {tot_s}")
for key, n in sorted(inst.items(), key=lambda kv: -kv[1])[:top_n]:
    f, l = key if key else ("?", 0)
    print(f"{n / tot_i * 100:5.2f}% inst {samp[key] / max(tot_s, 1) * 100:5.2f}% smp  thr/inst {thr[key] / max(n, 1):5.1f}  "
          f"{f}:{l:<4d} {src(key)}")
