#!/usr/bin/env python
"""Static SASS loop census: for each backward branch in a kernel, count the instructions in
the loop body by opcode class.  Usage: sass_loops.py <obj-or-so> <demangled-substring>"""
import re
import subprocess
import sys
from collections import Counter


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    sass = subprocess.run(f"cuobjdump -sass {obj} | c++filt", shell=True, capture_output=True, text=True).stdout
    blocks = re.split(r"\n\s*Function : ", sass)
    for blk in blocks[1:]:
        name = blk.split("\n", 1)[0]
        if pat not in name:
            continue
        ins = []
        for line in blk.split("\n"):
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m:
                ins.append((int(m.group(1), 16), m.group(2).strip()))
        print("==", name, "instructions:", len(ins))
        addr_index = {a: i for i, (a, _) in enumerate(ins)}
        for i, (a, t) in enumerate(ins):
            m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", t)
            if m:
                tgt = int(m.group(1), 16)
                if tgt <= a and tgt in addr_index:
                    body = ins[addr_index[tgt]: i + 1]
                    c = Counter()
                    for _, tt in body:
                        op = tt.split()[0]
                        if op.startswith("@"):
                            op = tt.split()[1]
                        c[op.split(".")[0]] += 1
                    print(f"  loop {tgt:#x}..{a:#x}: {len(body)} instr:", dict(c.most_common()))


if __name__ == "__main__":
    main()
