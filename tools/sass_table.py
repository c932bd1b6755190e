"""profiles/r02_sass_mnemonics.txt: per-kernel counts of the SASS mnemonics that prove the Blackwell paths (cuobjdump -sass of the in-tree
libt2b200.so; runs without a GPU). UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA tensor load, UTMASTG = TMA tensor store,
UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, SYNCS = mbarrier, HMMA = legacy mma.sync, DFMA = fp64."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "tacotron-2_b200", "libt2b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
KEYS = [("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("UTCHMMA", r"\bUTCHMMA(?!\.2CTA)"), ("LDTM", r"\bLDTM"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"),
        ("UBLKCP", r"\bUBLKCP"), ("UTCBAR", r"\bUTCBAR"), ("SYNCS(mbarrier)", r"\bSYNCS"), ("HMMA", r"\bHMMA"), ("DFMA", r"\bDFMA"), ("MUFU.TANH", r"MUFU\.TANH")]
rows, total = [], collections.Counter()
for name, body in zip(names, re.split(r"Function : \S+", sass)[1:]):
    c = {k: len(re.findall(p, body)) for k, p in KEYS}
    if any(c.values()):
        short = re.sub(r"\(anonymous namespace\)::|\(t2::(GemmArgs|WgradArgs)\)|\(int\)", "", name)[:95]
        rows.append("%-96s %s" % (short, "  ".join("%s=%d" % (k, v) for k, v in c.items() if v)))
        total.update(c)
out = ["# SASS mnemonic counts per kernel of tacotron-2_b200/libt2b200.so (cuobjdump -sass, sm_100a), round 2 - regenerate with tools/sass_table.py",
       "# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2 CTA-pair MMA), LDTM = tcgen05.ld, UTMALDG = TMA tensor load, UTMASTG = TMA tensor STORE (epilogue tiles),",
       "# UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, HMMA = legacy mma.sync (attention filter bank only), DFMA = fp64 (STFT / Griffin-Lim)",
       "# TOTAL: " + "  ".join("%s=%d" % (k, total[k]) for k, _ in KEYS), ""] + sorted(rows)
open(os.path.join(ROOT, "profiles", "r02_sass_mnemonics.txt"), "w").write("\n".join(out) + "\n")
print(out[3])
