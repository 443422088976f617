#!/usr/bin/env python
"""Per-kernel SASS evidence of the built library (no GPU needed): how many tcgen05 MMAs (UTCHMMA, with the A operand in tensor
memory or in shared memory), tensor-memory loads / stores (LDTM / STTM), commits (UTCBAR), bulk copies (UBLKCP) and mbarrier
operations (SYNCS) each kernel holds, next to its registers and spill bytes from the tracked ptxas logs.
    python tools/sass_evidence.py > profiles/r02_sass_evidence.txt
"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "spin-nerf_b200", "libspinnerf_b200.so")
PATTERNS = [("UTCHMMA", r"\bUTCHMMA"), ("UTCHMMA A in TMEM", r"\bUTCHMMA[.\w]* tmem"), ("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("UTCBAR", r"\bUTCBAR"),
            ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"), ("UTCATOMSWS (alloc)", r"\bUTCATOMSWS"),
            ("ELECT", r"\bELECT"), ("HMMA (legacy)", r"\bHMMA"), ("F2FP", r"\bF2FP"), ("STL/LDL (local memory: spills + indexed local arrays)", r"\b(STL|LDL)\b")]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    if not os.path.exists(LIB):
        sys.exit(f"{LIB} not built: python -c 'import __graft_entry__ as g; g.build()'")
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, cur, arch, mix = collections.OrderedDict(), None, set(), {}
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1); counts[cur] = collections.Counter(); continue
        m = re.match(r"\s*arch = (\S+)", line)
        if m:
            arch.add(m.group(1))
        if cur is None or not re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            continue
        counts[cur]["instructions"] += 1
        op = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if op:
            mix.setdefault(cur, collections.Counter())[op.group(1)] += 1
        for name, pat in PATTERNS:
            if re.search(pat, line):
                counts[cur][name] += 1
    regs = {}
    for log in glob.glob(os.path.join(ROOT, "spin-nerf_b200", "csrc", "obj", "*.log")):
        txt = open(log).read()
        for m in re.finditer(r"Function properties for (\S+)\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n"
                             r"ptxas info\s+: Used (\d+) registers", txt):
            regs[m.group(1)] = (int(m.group(5)), int(m.group(3)), int(m.group(4)))
    names = demangle(list(counts))
    print(f"# {os.path.relpath(LIB, ROOT)}: arch {sorted(arch)}; {len(counts)} kernels; counts of SASS mnemonics per kernel (cuobjdump -sass)")
    print("# tcgen05.mma = UTCHMMA, tcgen05.ld / st = LDTM / STTM, tcgen05.commit = UTCBAR, cp.async.bulk = UBLKCP, mbarrier = SYNCS")
    tc = [k for k in counts if counts[k]["UTCHMMA"]]
    rest = [k for k in counts if not counts[k]["UTCHMMA"]]
    for title, keys in (("kernels on the tcgen05 tensor cores", tc), ("CUDA-core kernels", rest)):
        print(f"\n## {title}")
        for k in keys:
            c = counts[k]
            r = regs.get(k)
            short = re.sub(r"\(.*", "", names[k]).replace("spn::", "")
            extra = f"  regs {r[0]}, spill stores/loads {r[1]}/{r[2]} B" if r else ""
            print(f"{short:44s} {c['instructions']:6d} instr{extra}")
            row = ", ".join(f"{n} {c[n]}" for n, _ in PATTERNS if c[n])
            if row:
                print(f"    {row}")
            if c["UTCHMMA"] and "mlp_" in short:        # static instruction mix of the three MLP kernels (all roles of the kernel together)
                top = mix[k].most_common(14)
                print("    mix: " + ", ".join(f"{n} {v}" for n, v in top))


if __name__ == "__main__":
    main()
