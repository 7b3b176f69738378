"""Regenerates profiles/wavefront_sass_r02.txt: opcode histograms and excerpts of the two production kernels (CPU only, cuobjdump)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mixed-integer-optimal-control---algorithm-tools_b200", "libbellman_b200.so")
KERNELS = [("pruned tiles, production default at config 4: wavefront_kernel<2,8,1,u8,640,false,4,128> (two-zone slices, compile-time level count)",
            "_ZN5bb20016wavefront_kernelILi2ELi8ELi1EhLi640ELb0ELi4ELi128ELi128EEEvNS_6TablesENS_7WaveCfgE"),
           ("exhaustive tiles (fallback when the bound test does not pay): wavefront_kernel<4,3,2,u8,512,false,0>",
            "_ZN5bb20016wavefront_kernelILi4ELi3ELi2EhLi512ELb0ELi0ELi0ELi0EEEvNS_6TablesENS_7WaveCfgE")]
WANT = re.compile(r"^(@!?U?P\d )?(DADD|DSETP|DMUL|DFMA|FADD|FMUL|FSETP|FMNMX|F2F|FSEL|SEL|MOV|IMAD\.MOV|LDS|STS|STG|LDG|LDL|STL|UBLKCP|SYNCS|MEMBAR|FENCE|BAR|CREDUX|REDUX|VOTE|NANOSLEEP|HMMA|UTC|WARPSYNC|CCTL|ERRBAR|CGAERRBAR)")
out = ["# SASS evidence, round 2 final build (cuobjdump -sass libbellman_b200.so, sm_100a, nvcc 12.9; tools/sass_evidence.py)", ""]
for title, sym in KERNELS:
    txt = subprocess.run(["cuobjdump", "-sass", "-fun", sym, LIB], capture_output=True, text=True).stdout
    ins = [re.sub(r"\s*/\*.*$", "", l.split("*/", 1)[1]).strip().rstrip(";").strip() for l in txt.splitlines() if re.match(r"^\s+/\*[0-9a-f]{4}\*/", l)]
    hist = collections.Counter()
    for i in ins:
        m = re.match(r"^(@!?U?P\d\s+)?(\S+)", i)
        op = ("@P " if m.group(1) else "") + m.group(2)
        if WANT.match(("@P0 " if m.group(1) else "") + m.group(2)):
            hist[op] += 1
    out += [f"## {title}", f"{len(ins)} instructions ({len(ins) * 16 / 1024:.1f} KB).  Opcode histogram (selected):", ""]
    out += [f"    {k:28s} {v}" for k, v in sorted(hist.items())]
    out += ["", "no DFMA (every add/multiply of the value path is separately rounded), no HMMA / UTC*MMA (min-plus is not a contraction), "
            f"local-memory instructions (spills): {sum(v for k, v in hist.items() if 'LDL' in k or 'STL' in k)}", ""]
    def excerpt(label, pat, before=6, after=14, which=0):
        idx = [k for k, i in enumerate(ins) if re.search(pat, i)]
        if not idx: return
        k = idx[min(which, len(idx) - 1)]
        out.append(f"### {label}")
        out.extend("    " + x for x in ins[max(0, k - before):k + after])
        out.append("")
    if "ELi4ELi128" in sym:
        excerpt("row test of the pruned scan: directed-rounding FP32 (FADD.RP / FADD.RM / F2F.*.RP), REDUX.MAX on order-preserving keys, ballot", r"CREDUX\.MAX", 14, 22)
        excerpt("level test: rd32(rd32(sf + cminf) + pmf) > UBf, four candidate blocks per trip", r"FADD\.RM", 10, 30, which=6)
        excerpt("scan of a surviving block: DADD, DSETP.GT, predicated moves (no FSEL/SEL on the update)", r"DSETP\.GT", 8, 26, which=4)
    else:
        excerpt("phase-B loop: DADD, DSETP.GT, predicated moves", r"DSETP\.GT", 8, 26, which=8)
    excerpt("bulk TMA (UBLKCP) with mbarrier transaction counts", r"UBLKCP", 6, 6)
    excerpt("publisher: fence.acq_rel.gpu + relaxed flag store", r"MEMBAR\.ALL\.GPU", 4, 6)
open(os.path.join(ROOT, "profiles", "wavefront_sass_r02.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:60]))
