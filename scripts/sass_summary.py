"""SASS evidence for K1 (no GPU needed: cuobjdump on the built object): opcode histogram of the benchmark's kernel
instantiation, the proof that no tensor-core / TMA instruction is in it, and the unrolled gather of one time step.
python scripts/sass_summary.py > profiles/r2_k1_sass.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
obj = os.path.join(ROOT, "calcium_b200", "_obj", "integrate_f64.o")
fun = "_ZN6cab20016integrate_kernelIdLi512ELi3ELi5ELb1ELi2ELi1ELi1EEEvNS_7SimArgsE"
sass = subprocess.run(["cuobjdump", "-sass", "-fun", fun, obj], capture_output=True, text=True).stdout
lines = [l for l in sass.splitlines() if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l)]
ops = collections.Counter()
for l in lines:
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", l)
    if m: ops[m.group(1)] += 1
print("kernel: integrate_kernel<double, 512 threads, 3 cells/thread, EP=5, UNIT, 2 copies, 1 CTA/SM> (the bench's K1)")
print(f"{len(lines)} SASS instructions (whole kernel, prologue included)\n\nopcode histogram:")
for op, n in ops.most_common():
    print(f"  {op:12s} {n}")
tensor = [op for op in ops if re.match(r"(HMMA|IMMA|DMMA|BMMA|QMMA|UTC|UTMA|UBLKCP|TCGEN|WGMMA|LDTM|STTM)", op)]
print("\ntensor-core / TMEM / TMA opcodes present:", tensor or "none  (a ~6-neighbour sparse stencil has no dense contraction)")
# one time step: from the first mbarrier try_wait to the next
idx = [i for i, l in enumerate(lines) if "SYNCS.PHASECHK" in l]
if len(idx) >= 2:
    step = lines[idx[0]: idx[1]]
    sop = collections.Counter()
    for l in step:
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", l)
        if m: sop[m.group(1)] += 1
    print(f"\nfirst unrolled half-step (barrier wait -> next barrier wait): {len(step)} instructions: "
          + ", ".join(f"{k} {v}" for k, v in sop.most_common(14)))
    print("\nthe gather of that step (one LDS.128 per row entry and cell, immediate = step-parity buffer; DADD chains in CSR order;"
          "\nthe DMULs inside it are the register diagonal L_ii*(c,p)):\n")
    g0 = next(i for i, l in enumerate(step) if "LDS.128" in l)
    g1 = max(i for i, l in enumerate(step) if "LDS.128" in l)
    for l in step[max(0, g0 - 4): g1 + 16]:
        print(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l))
