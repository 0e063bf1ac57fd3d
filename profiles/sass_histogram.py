"""Opcode histogram per kernel of libdfb200.so (cuobjdump -sass): the SASS mnemonics that prove the Blackwell-native
paths (UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA tensor load, UBLKCP = bulk copy, UTCBAR =
tcgen05.commit, SYNCS = mbarrier) next to the FP32 / legacy tensor opcodes.
usage: python profiles/sass_histogram.py deepfilternet_b200/libdfb200.so > profiles/r02_sass_opcodes.md"""
import collections
import re
import subprocess
import sys

WATCH = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "UTCATOM", "SYNCS", "HMMA", "FFMA", "FFMA2",
         "MUFU", "LDG", "STG", "LDS", "STS", "SHFL", "BAR"]


def main(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    kern, hist, total = None, collections.OrderedDict(), {}
    for ln in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            kern = re.sub(r"\(.*", "", kern).replace("void ", "").replace("dfb::", "")
            hist[kern] = collections.Counter()
            total[kern] = 0
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
        if kern and m:
            op = m.group(1)
            total[kern] += 1
            for w in WATCH:
                if op == w or (w in ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "MUFU",
                                     "LDG", "STG", "LDS", "STS", "SHFL", "BAR") and op.startswith(w)):
                    hist[kern][w] += 1
                    break
    cols = [w for w in WATCH if any(h[w] for h in hist.values())]
    print("| kernel | instructions | " + " | ".join(cols) + " |")
    print("|---|---|" + "---|" * len(cols))
    for k, h in hist.items():
        print(f"| `{k}` | {total[k]} | " + " | ".join(str(h[c]) if h[c] else "" for c in cols) + " |")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "deepfilternet_b200/libdfb200.so")
