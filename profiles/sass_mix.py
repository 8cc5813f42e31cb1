#!/usr/bin/env python
"""Instruction mix of the hot kernels from their SASS (cuobjdump -sass of the in-tree objects; no GPU needed): opcode counts per
kernel, FP64 share, and the SASS of the two kernels BASELINE's roofline claims rest on.
usage: python profiles/sass_mix.py            -> profiles/r01_sass_mix.txt + profiles/r01_sass_<kernel>.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "finetools.jl_b200", "csrc")
WANT = {"fegpu_h8.o": ["k_h8_elasticILb1E", "k_h8_diffusionILb1ELb1E"],
        "fegpu_pattern.o": ["k_gatherILi16ELi3ELb1ELi1ELi2E", "k_gatherILi8ELi1ELb1ELi4ELi1E", "k_nbr_groupILi16ELi4ELb0E", "k_rows_sortedILi32ELi3E"],
        "fegpu_elastic.o": ["k_elastic_tiledILi20E"], "fegpu_dot.o": ["k_dot_scalarILi10ELi3ELi3E"]}
FULL = ["k_h8_elasticILb1E", "k_h8_diffusionILb1ELb1E"]   # full SASS committed for these


def functions(obj):
    txt = subprocess.run(["cuobjdump", "-sass", os.path.join(CSRC, obj)], capture_output=True, text=True).stdout
    cur, out = None, collections.OrderedDict()
    for ln in txt.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            out[cur] = []
        elif cur is not None:
            out[cur].append(ln)
    return out


def main():
    lines = ["# opcode counts from cuobjdump -sass (static instruction mix of the compiled sm_100a kernels; loops are counted once)"]
    for obj, keys in WANT.items():
        fns = functions(obj)
        for key in keys:
            name = next((f for f in fns if key in f), None)
            if name is None:
                lines.append("\n== %s: not found in %s" % (key, obj))
                continue
            ops = collections.Counter()
            for ln in fns[name]:
                m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
                if m:
                    ops[m.group(1)] += 1
            tot = sum(ops.values())
            fp64 = sum(v for k, v in ops.items() if k in ("DFMA", "DADD", "DMUL", "DSETP", "MUFU"))
            lines.append("\n== %s  (%s)\n   %d instructions, FP64 pipe (DFMA/DADD/DMUL/DSETP/MUFU.RCP64H) %d = %.1f %%"
                         % (key, obj, tot, fp64, 100.0 * fp64 / max(tot, 1)))
            lines.append("   " + ", ".join("%s %d" % kv for kv in ops.most_common(18)))
            if key in FULL:
                with open(os.path.join(ROOT, "profiles", "r01_sass_%s.txt" % key.split("IL")[0]), "w") as f:
                    f.write("# cuobjdump -sass %s, function %s\n" % (obj, name))
                    f.write("\n".join(ln for ln in fns[name] if re.search(r"/\*[0-9a-f]{4}\*/", ln)) + "\n")
    with open(os.path.join(ROOT, "profiles", "r01_sass_mix.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
