"""Per-kernel histogram of the Blackwell-specific SASS opcodes in libhealswin_b200.so (cuobjdump -sass):
UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA loads / stores, UTCBAR = tcgen05.commit,
SYNCS = mbarrier ops, LDGSTS = cp.async, HMMA = legacy mma.sync (none expected).
python tools/sass_hist.py [lib] > profiles/<round>_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         "heal_swin_b200", "libhealswin_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCOMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "UTCATOMSWS", "SYNCS",
        "LDGSTS", "HMMA", "LDS", "STS", "LDG", "STG", "RED", "ATOM", "MUFU", "SHFL", "F2FP", "BAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    fn, hist = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            fn = re.sub(r"\(anonymous namespace\)::", "", fn)
            fn = re.sub(r"\(.*", "", fn)[:70]
            hist[fn] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and fn:
            op = m.group(1)
            hist[fn]["_total"] += 1
            for k in KEYS:
                if op.startswith(k):
                    hist[fn][k] += 1
                    break
    print(f"# {os.path.basename(LIB)}: SASS opcode histogram per kernel (sm_100a)")
    cols = [k for k in KEYS if any(h[k] for h in hist.values())]
    print(f"{'kernel':70s} {'instr':>7s} " + " ".join(f"{c:>8s}" for c in cols))
    for fn, h in hist.items():
        print(f"{fn:70s} {h['_total']:7d} " + " ".join(f"{h[c]:8d}" for c in cols))


if __name__ == "__main__":
    main()
