"""SASS evidence for libgpp_b200.so: per kernel, the count of the mnemonics that prove the hardware path (DMMA = FP64
tensor-core mma.sync, UTMALDG = TMA tensor loads, SYNCS = mbarrier ops, DFMA/MUFU for the scalar FP64 kernels, memory
ordering ops of the persistent kernels).  Runs on the CPU box: cuobjdump -sass <lib> (no GPU needed).
    python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "nonlinpdes-gpsolver_b200", "libgpp_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
MN = ["DMMA", "UTMALDG", "UTMASTG", "SYNCS", "DFMA", "DADD", "DMUL", "MUFU", "LDS", "STS", "LDG", "STG", "ATOM", "RED", "MEMBAR", "BAR", "SHFL", "LDGSTS", "CCTL", "ERRBAR"]
kern, counts, arch = None, collections.OrderedDict(), None
for line in out.splitlines():
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        arch = m.group(1)
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(anonymous namespace\)::", "", kern)
        kern = re.sub(r"\(.*", "", kern)
        counts[kern] = collections.Counter()
        counts[kern]["_arch"] = arch
        continue
    if kern is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?", line)
    if m:
        op = m.group(1)
        counts[kern]["_total"] += 1
        for k in MN:
            if op == k or op.startswith(k + "."):
                counts[kern][k] += 1
        if op == "DMMA" and m.group(2):
            counts[kern]["DMMA" + m.group(2).split(" ")[0]] += 1
        if op.startswith("LD") and m.group(2) and ("ACQUIRE" in m.group(2) or "STRONG" in m.group(2)):
            counts[kern]["ld.acquire/strong"] += 1
        if op.startswith("ST") and m.group(2) and ("RELEASE" in m.group(2) or "STRONG" in m.group(2)):
            counts[kern]["st.release/strong"] += 1
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}  (built with -gencode arch=compute_100a,code=sm_100a -lineinfo)")
print("# kernel | arch | SASS instructions | counts of the mnemonics that matter")
for k, c in counts.items():
    rest = ", ".join(f"{m}={v}" for m, v in c.items() if not m.startswith("_"))
    print(f"{k} | {c['_arch']} | {c['_total']} | {rest}")
