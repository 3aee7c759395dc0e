#!/usr/bin/env python
"""Static evidence for the voice-parallel kernel (gas_mix_voice.cu) when no GPU is at hand: ptxas resource usage and, from
`cuobjdump -sass`, the opcode mix of every out-of-line unit function of k_mix_voice<4> plus the longest straight-line run of
float instructions in it (the 8-frame trip of the filter phase).  Writes to stdout; profiles/r02_k3_filter_tile_static.txt is its
output at the commit that introduced the filter-tile path.  Not a measurement."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "godot-audio-spatializer_b200", "csrc")
obj = os.path.join(CSRC, "gas_mix_voice.o")
sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", obj], capture_output=True, text=True, check=True).stdout
print("== cuobjdump -res-usage gas_mix_voice.o ==")
for ln in res.splitlines():
    if "Function" in ln or "REG:" in ln:
        print(re.sub(r"_ZN\d+_GLOBAL__N__[0-9a-f_]+gas_mix_voice_cu_[0-9a-f]+", "", ln.strip())[:160])
# k_mix_voice<4>: first function of the listing
funcs = re.split(r"\n\s*Function : ", sass)
body = next(f for f in funcs if "k_mix_voiceILi4E" in f.split("\n", 1)[0])
ins = []
for ln in body.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", ln)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
# out-of-line functions = targets of CALL.REL
targets = sorted({int(t, 16) for _, s in ins for t in re.findall(r"CALL\.REL\.NOINC (0x[0-9a-f]+)", s)})
ends = targets[1:] + [ins[-1][0] + 16]
print("\n== k_mix_voice<4>: %d SASS instructions; out-of-line functions at %s ==" % (len(ins), ", ".join(hex(t) for t in targets)))


def opcode(s):
    s = re.sub(r"^@!?U?P\d+\s+", "", s)
    return s.split()[0].split(".")[0]


for t0, t1 in zip(targets, ends):
    part = [(a, s) for a, s in ins if t0 <= a < t1]
    hist = collections.Counter(opcode(s) for _, s in part)
    if hist.get("BAR", 0) == 0:
        continue  # helpers (division slow path)
    keys = ["LDGSTS", "LDGDEPBAR", "DEPBAR", "BAR", "FFMA2", "FFMA", "FMUL", "FADD", "FMNMX3", "FMNMX", "LDS", "STS", "LDG", "LDL", "STL", "SHFL", "ATOMS", "ATOMG", "REDG"]
    print("\nfunction at %#x (%d instructions): " % (t0, len(part)) + ", ".join(f"{k} {hist[k]}" for k in keys if hist.get(k)))
    # longest run of float / shared-memory instructions without a branch: the unrolled 8-frame trip of the filter phase
    best, cur = [], []
    for a, s in part:
        op = opcode(s)
        if op in ("BRA", "BAR", "CALL", "RET", "BSSY", "BSYNC", "EXIT", "WARPSYNC"):
            if len(cur) > len(best):
                best = cur
            cur = []
        else:
            cur.append((a, s))
    if len(cur) > len(best):
        best = cur
    h2 = collections.Counter(opcode(s) for _, s in best)
    print("  longest straight-line run: %d instructions at %#x: " % (len(best), best[0][0]) + ", ".join(f"{k} {v}" for k, v in h2.most_common(10)))
    if "-v" in sys.argv:
        for a, s in best:
            print("    /*%05x*/ %s" % (a, s))
