"""SASS opcode histogram per kernel of libiridium_b200.so (cuobjdump -sass) -> profiles/r2_sass_histogram.txt.
The mnemonics that matter: UBLKCP (cp.async.bulk / TMA 1-D), LDGSTS (cp.async), SYNCS (mbarrier), FFMA2 (packed fp32 FMA),
BAR, REDUX / VOTE / SHFL (warp collectives); no UTC*MMA / LDTM is expected (no dense contraction on this path)."""
import collections, re, subprocess, sys
so = sys.argv[1] if len(sys.argv) > 1 else "iridium-sniffer_b200/libiridium_b200.so"
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
fn, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        hist.setdefault(fn, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and fn:
        hist[fn][m.group(1)] += 1
KEY = ("UBLKCP", "LDGSTS", "SYNCS", "FFMA2", "FFMA", "DFMA", "BAR", "REDUX", "VOTE", "SHFL", "LDS", "STS", "LDG", "STG", "UTMALDG", "UTCHMMA", "LDTM")
print("%-64s %7s  %s" % ("kernel", "instrs", "  ".join(KEY)))
for fn, c in hist.items():
    if not c or not fn.startswith(("ir::", "void ir::", "k_")):
        continue
    print("%-64s %7d  %s" % (fn[:64], sum(c.values()), "  ".join("%*d" % (len(k), c.get(k, 0)) for k in KEY)))
