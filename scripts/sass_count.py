#!/usr/bin/env python3
"""Counts SASS opcodes of one kernel of an object file, whole function and hottest loop (largest backward-branch span).

    python scripts/sass_count.py spheral_b200/build/derivs.cu.o 'k_sph_derivsILi3ELb0ELb0ELb0ELb1ELb0E'

Offline companion to the ncu captures: the FP64 / LSU instruction count per warp-iteration of a pair loop can be read here
before GPU time is spent."""
import collections
import re
import subprocess
import sys


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    fn, cur = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            fn[cur] = []
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur:
            fn[cur].append((int(m.group(1), 16), m.group(2).strip()))
    for name, ins in fn.items():
        if pat not in name:
            continue
        print("==", name, len(ins), "instructions")
        # backward branches
        loops = []
        for addr, text in ins:
            m = re.search(r"\bBRA\S*\s+(?:\S+,\s*)*`?\(?\.?L?_?x?_?(\w+)\)?|BRA\S*\s+.*0x([0-9a-f]+)", text)
            t = re.search(r"0x([0-9a-f]+)", text) if "BRA" in text else None
            if t:
                tgt = int(t.group(1), 16)
                if tgt < addr:
                    loops.append((addr - tgt, tgt, addr))
        loops.sort(reverse=True)
        spans = [("whole", ins[0][0], ins[-1][0])] + [("loop%d" % k, l[1], l[2]) for k, l in enumerate(loops[:3])]
        for label, lo, hi in spans:
            c = collections.Counter()
            for addr, text in ins:
                if lo <= addr <= hi:
                    t = text.split()
                    op = t[1] if t[0].startswith("@") else t[0]
                    c[op.split(".")[0]] += 1
            fp64 = sum(v for k, v in c.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
            lsu = {k: v for k, v in c.items() if k in ("LDS", "LDG", "STG", "STS", "LDGSTS", "SHFL", "LDSM", "ATOMS", "RED", "ATOMG")}
            print("  %-6s [%x..%x] total %4d  FP64 %3d (DFMA %d DMUL %d DADD %d DSETP %d DMNMX %d)  MUFU %d F2I %d I2F %d  LSU %s"
                  % (label, lo, hi, sum(c.values()), fp64, c["DFMA"], c["DMUL"], c["DADD"], c["DSETP"], c["DMNMX"], c["MUFU"], c["F2I"], c["I2F"], lsu))


if __name__ == "__main__":
    main()
