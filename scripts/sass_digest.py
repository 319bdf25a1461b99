"""SASS digest of the shipped library: per kernel, the count of each mnemonic that proves the Blackwell tensor path
(B200_PROFILING.md): UTCHMMA / UTCQMMA (tcgen05.mma), UTCBAR (tcgen05.commit), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG
(TMA), plus registers.  Usage: python scripts/sass_digest.py > profiles/r2_sass_digest.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "vit.rs_b200", "libvitrs.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
MNEMONICS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCATOMSWS", "SYNCS", "MUFU.EX2", "FFMA2", "REDG", "STL", "LDL"]
kernels, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    kernels[cur]["_instructions"] += 1
    for mn in MNEMONICS:
        if op == mn or op.startswith(mn + "."):
            kernels[cur][mn] += 1
    if ".2CTA" in op and op.startswith("UTCHMMA"):
        kernels[cur]["UTCHMMA.2CTA"] += 1
def demangle(name):
    out = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip()
    out = out.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "").replace("__nv_bfloat16", "bf16")
    depth = 0
    for i, ch in enumerate(out):  # cut the parameter list: the first "(" outside template brackets
        depth += ch == "<"
        depth -= ch == ">"
        if ch == "(" and depth == 0:
            out = out[:i]
            break
    return out[:90] or name[:90]
arch = re.findall(r"arch = (sm_\w+)", sass)
print(f"# {os.path.relpath(lib, ROOT)}: {len(kernels)} kernels, arch {sorted(set(arch))}")
tot = collections.Counter()
for k, c in kernels.items():
    tot.update(c)
print("# totals: " + ", ".join(f"{mn}={tot[mn]}" for mn in MNEMONICS + ["UTCHMMA.2CTA"] if tot[mn]))
cols = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "MUFU.EX2", "FFMA2", "STL", "LDL"]
print(f"{'kernel':92s} {'instr':>6s} " + " ".join(f"{c:>8s}" for c in cols))
for k, c in kernels.items():
    if c["_instructions"] < 40:
        continue
    print(f"{demangle(k):92s} {c['_instructions']:6d} " + " ".join(f"{c[x]:8d}" for x in cols))
