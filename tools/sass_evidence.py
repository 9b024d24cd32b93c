"""Counts, per kernel of the built objects, the SASS mnemonics that show which hardware paths the code uses (run in the build
container: cuobjdump only, no GPU): UTCIMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTMALDG / UTMASTG = TMA
load / store, SYNCS = mbarrier, IDP = dp4a, IMAD.HI = the integer requantisation.  Output: profiles/<tag>_sass_evidence.txt."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"
WANT = ["UTCIMMA", "LDTM", "UTCBAR", "UTMALDG", "UTMASTG", "SYNCS", "IDP", "IMAD.HI", "FFMA2", "DMUL", "DFMA", "LDG", "LDS", "STS", "STG"]
out = ["# static SASS mnemonic counts per kernel (cuobjdump -sass of codenet_b200/csrc/*.o, sm_100a); see tools/sass_evidence.py",
       "%-58s " % "kernel" + " ".join("%8s" % w for w in WANT)]
for obj in ("unit_fused", "unit_s2_fused", "pw_gemm", "heads_fused", "dw_tma", "deform_tile", "dw", "stem", "decode"):
    txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "codenet_b200", "csrc", obj + ".o")], capture_output=True, text=True).stdout
    fn, cnt = None, collections.OrderedDict()
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void ", "")
            cnt[fn] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m and fn:
            op = m.group(1)
            for w in WANT:
                if op == w or op.startswith(w + "."):
                    cnt[fn][w] += 1
    for fn, c in cnt.items():
        out.append("%-58s " % fn[:58] + " ".join("%8d" % c[w] for w in WANT))
open(os.path.join(ROOT, "profiles", TAG + "_sass_evidence.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
