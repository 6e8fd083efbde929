#!/usr/bin/env python
"""Static evidence about the built kernels, for profiles/: `nvcc -Xptxas -v` resource usage per kernel (registers, spills,
stack, shared memory) and per-kernel counts of the SASS mnemonics that prove the hardware paths (cuobjdump -sass of the
in-tree libdronesim_b200.so): UTCHMMA / UTCBAR / LDTM / STTM (tcgen05 MMA, commit, tensor-memory load / store), UBLKCP
(cp.async.bulk = TMA 1D bulk copy), SYNCS (mbarrier), ACQBULK / PREEXIT (programmatic dependent launch), FFMA / FMUL / FADD and their packed two-per-lane forms FFMA2 / FMUL2 / FADD2.

    python tools/build_evidence.py > profiles/r02_build_evidence.txt
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "mujoco_drone_b200", "csrc")
SO = os.path.join(ROOT, "mujoco_drone_b200", "libdronesim_b200.so")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xptxas", "-v", "-c"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


print("# nvcc " + " ".join(FLAGS) + "  (nvcc " + subprocess.run(["nvcc", "--version"], capture_output=True, text=True).stdout.strip().splitlines()[-2].strip() + ")")
with tempfile.TemporaryDirectory() as td:
    for src in ("dsim_kernels.cu", "dsim_policy_mlp.cu", "dsim_policy_fp32.cu"):
        r = subprocess.run(["nvcc"] + FLAGS + [os.path.join(CSRC, src), "-o", os.path.join(td, "o.o")], capture_output=True, text=True)
        txt = r.stderr
        ents = re.findall(r"Compiling entry function '([^']+)' for 'sm_100a'\n.*?Function properties for [^\n]+\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes smem)?", txt, re.S)
        dm = demangle([e[0] for e in ents])
        print(f"\n## {src}: ptxas resource usage")
        for e in ents:
            name = re.sub(r"\(anonymous namespace\)::|dsim::", "", dm[e[0]])
            name = re.sub(r"\(.*", "", name)
            print(f"  {name:64s} regs {int(e[4]):3d}  stack {int(e[1]):4d} B  spill st/ld {e[2]}/{e[3]} B  static smem {e[6] or 0} B")

sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
cur, counts = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
    if m and cur:
        counts[cur][m.group(1)] += 1
        if m.group(1) in ("UBLKCP", "SYNCS", "LDTM", "STTM"):
            counts[cur][m.group(1) + m.group(2)] += 1
dm = demangle(list(counts))
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UBLKCP.S.G", "UBLKCP.G.S", "SYNCS", "ACQBULK", "PREEXIT", "FFMA", "FMUL", "FADD", "FFMA2", "FMUL2", "FADD2", "MUFU", "DFMA", "DADD", "DMUL"]
print("\n## SASS mnemonic counts per kernel (cuobjdump -sass mujoco_drone_b200/libdronesim_b200.so; static instruction counts)")
print("  " + "kernel".ljust(66) + " ".join(k.rjust(10) for k in KEYS) + "     total")
for fn, c in counts.items():
    name = re.sub(r"\(anonymous namespace\)::|dsim::", "", dm[fn])
    name = re.sub(r"\(.*", "", name)
    if not any(c[k] for k in KEYS[:10]) and "kernel" not in name:
        continue
    print("  " + name[:64].ljust(66) + " ".join(str(c[k]).rjust(10) for k in KEYS) + f"  {sum(v for k, v in c.items() if '.' not in k):8d}")
