#!/usr/bin/env python
"""SASS opcode histogram of the hot kernels (cuobjdump -sass on the in-tree objects): which load / prefetch / async-copy
instructions the streaming loops are made of.  Usage: python scripts/sass_hist.py > profiles/r2_sass_opcodes.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJS = ["scan_literal.o", "scan_multi.o", "scan_count.o", "engine.o"]
WANT = ["k_lit_aligned4ILb0ELi4", "k_lit_window4ILb1ELb0ELi4", "k_ac_tri4ILb0ELi640", "k_ac_scanILi2ELb0E", "k_count_linesILb1ELb0ELb1", "k_count_linesILb0ELb0ELb0", "k_finish"]
NOTE = {"LDG": "global load", "LDGSTS": "cp.async (global -> shared)", "UBLKPF": "cp.async.bulk.prefetch.L2", "LDS": "shared load",
        "STS": "shared store", "VOTE": "ballot / any", "SHFL": "warp shuffle", "REDUX": "redux.sync", "ATOMG": "global atomic", "RED": "reduction atomic",
        "IMAD": "integer multiply-add (FMA pipe)", "LOP3": "3-input logic", "SHF": "funnel shift", "ISETP": "integer compare", "BAR": "barrier",
        "UTMALDG": "TMA tensor load", "STG": "global store", "LDL": "local load (spill)", "STL": "local store (spill)"}


def main():
    print("# SASS opcode histograms (sm_100a), `cuobjdump -sass` on krep_b200/build/*.o\n")
    print("No `UTMALDG` anywhere: the scans are register-streamed (LDG.E.128 straight into the filter), with `UBLKPF` bulk L2 "
          "prefetch and `LDGSTS` (cp.async) in the multi-pattern kernel — see DESIGN.md §4.\n")
    for obj in OBJS:
        path = os.path.join(ROOT, "krep_b200", "build", obj)
        if not os.path.exists(path):
            continue
        out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
        funcs = re.split(r"\n\s*Function : ", out)[1:]
        for f in funcs:
            name = f.split("\n", 1)[0].strip()
            if not any(w in name for w in WANT):
                continue
            ops = collections.Counter()
            mods = collections.Counter()
            for line in f.splitlines():
                m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
                if m:
                    ops[m.group(1)] += 1
                    if m.group(1) in ("LDG", "LDGSTS", "UBLKPF", "LDS", "ATOMG", "STG"):
                        mods[m.group(1) + m.group(2)] += 1
            total = sum(ops.values())
            demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
            print(f"## `{demangled}`\n\n{total} instructions.\n")
            print("| opcode | count | note |\n|---|---|---|")
            for op, c in ops.most_common(18):
                print(f"| {op} | {c} | {NOTE.get(op, '')} |")
            print("\nmemory instructions with modifiers: " + ", ".join(f"`{k}` × {v}" for k, v in sorted(mods.items())) + "\n")


if __name__ == "__main__":
    main()
