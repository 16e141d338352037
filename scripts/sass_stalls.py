"""Static schedule of the loops of a kernel in an in-tree .so: decodes the stall field (bits 105-108
of the 128-bit SASS control word) and sums it per backward-branch loop.
usage: python scripts/sass_stalls.py <mangled kernel name> [min FFMA2 per loop]"""
import collections
import re
import subprocess
import sys

name = sys.argv[1]
min_ffma2 = int(sys.argv[2]) if len(sys.argv) > 2 else 100
txt = subprocess.run(["cuobjdump", "-sass", "-fun", name, "multimm_b200/libmultimm_b200.so"], capture_output=True,
                     text=True).stdout.split("\n")
ins, i = [], 0
while i < len(txt) - 1:
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/", txt[i])
    m2 = re.match(r"\s+/\* (0x[0-9a-f]{16}) \*/", txt[i + 1]) if m else None
    if m and m2:
        hi = int(m2.group(1), 16)
        ins.append((int(m.group(1), 16), m.group(2).strip(), (hi >> 41) & 0xF))
        i += 2
    else:
        i += 1
addr = {a: k for k, (a, _, _) in enumerate(ins)}
for k, (a, op, _) in enumerate(ins):
    m = re.search(r"BRA\s+(?:\w+,\s*)?0x([0-9a-f]+)", op)
    if m and int(m.group(1), 16) < a and int(m.group(1), 16) in addr:
        body = ins[addr[int(m.group(1), 16)]:k + 1]
        nf = sum("FFMA2" in b[1] for b in body)
        if nf >= min_ffma2 and len(body) < 1000:
            kinds = collections.Counter(b[1].split()[0].split(".")[0] for b in body)
            print(f"loop {int(m.group(1), 16):#x}-{a:#x}: {len(body)} instr, sum of stall fields {sum(b[2] for b in body)}, "
                  f"{kinds.most_common(8)}")
