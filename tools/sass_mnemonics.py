"""`cuobjdump -sass` of the shipped library, per kernel: instruction count and the mnemonics that prove the Blackwell paths
(profiles/r02_sass_mnemonics.txt).  Runs without a GPU."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "hipie_b200", "libhipie_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "USETMAXREG", "HMMA", "LDSM", "MUFU.EX2",
        "BAR.SYNC", "UCGABAR", "ATOMS", "REDUX", "ELECT", "LDGSTS", "F2FP"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
print("# cuobjdump -sass hipie_b200/libhipie_b200.so (sm_100a), per kernel: instruction count and the mnemonics that prove the Blackwell paths")
print("#   UTCHMMA = tcgen05.mma kind::f16, UTCQMMA = tcgen05.mma kind::f8f6f4 (the e4m3 sweep of gemm prec 6), UTMALDG / UTMASTG = TMA tensor load / store,")
print("#   LDTM / STTM = tcgen05.ld / st (tensor memory), UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, USETMAXREG = setmaxnreg,")
print("#   UCGABAR = cluster barrier (CTA pairs), HMMA / LDSM = mma.sync / ldmatrix (legacy small attention, rel-pos, fallback post-processing)")
name, cnt, n = None, collections.Counter(), 0


def flush():
    if name:
        short = re.sub(r"\(.*", "", demangle(name))
        print(f"{short}: {n} instr | " + " ".join(f"{k}={cnt[k]}" for k in KEYS if cnt[k]))


for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        flush()
        name, cnt, n = m.group(1), collections.Counter(), 0
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        n += 1
        op = m.group(1)
        for k in KEYS:
            if op == k or op.startswith(k + ".") or (k == "UCGABAR" and op.startswith(k)):
                cnt[k] += 1
flush()
