"""Per-kernel counts of the Blackwell-only SASS instructions in the built objects (CPU; needs cuobjdump):
    python tools/sass_evidence.py > profiles/rNN_sass_tcgen05.txt
UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA tensor load, UTCBAR = tcgen05.commit,
SYNCS = mbarrier, VIMNMX3 = packed 3-input max of the fp16-accumulator epilogue, REDG/ATOMG.64 = the integer segment sums,
HMMA.1688.F32.TF32 = the warp-level tf32 MMAs of the fused pre_quant projection (vq_prequant.cu: an N = 32, HBM-bound GEMM)."""
import collections, glob, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAT = re.compile(r"\b(HMMA|UTCHMMA|UTCQMMA|UTCOMMA|LDTM|STTM|UTMALDG|UTMASTG|UTCBAR|UTCATOMSWS|SYNCS|VIMNMX3|VIMNMX|REDG|RED|ATOMG|SETMAXREG|USETMAXREG|ELECT|LDGSTS|UBLKCP)\b[.\w]*")
for obj in sorted(glob.glob(os.path.join(ROOT, "attention-models_b200", "lib", "vq_*.o"))):
    if "_instr" in obj or re.search(r"_(defer|inpl|bahead)\.o$", obj):
        continue
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    kernel, counts = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            kernel = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            counts[kernel] = collections.Counter()
            continue
        if kernel:
            for mm in PAT.finditer(line.split("/*")[1] if line.count("/*") >= 1 and ";" in line else ""):
                counts[kernel][mm.group(0)] += 1
    print(f"== {os.path.basename(obj)}")
    for k, c in counts.items():
        tc = {op: n for op, n in c.items() if op.split(".")[0] in ("HMMA", "LDGSTS", "UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTCBAR", "VIMNMX3", "SETMAXREG", "USETMAXREG", "SYNCS", "ELECT", "REDG", "RED", "ATOMG")}
        if tc:
            print(f"  {k}")
            print("     " + "  ".join(f"{op} x{n}" for op, n in sorted(tc.items())))
