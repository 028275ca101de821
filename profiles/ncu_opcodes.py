"""Executed warp instructions per SASS opcode of one kernel in an .ncu-rep.
usage: python profiles/ncu_opcodes.py REPORT LAUNCH_INDEX [top_n]"""
import csv
import subprocess
import sys
from collections import Counter


def main():
    rep, kid = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", kid, "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    hdr, cnt = None, Counter()
    for r in csv.reader(raw.splitlines()):
        if not r:
            continue
        if r[0] == "Address":
            hdr = r
            continue
        if hdr and r[0].startswith("0x"):
            d = dict(zip(hdr, r))
            op = d["Source"].strip().split()
            op = [t for t in op if not t.startswith("@")]
            if not op:
                continue
            try:
                cnt[op[0].rstrip(";")] += int(d["Instructions Executed"])
            except ValueError:
                pass
    tot = sum(cnt.values()) or 1
    print("total warp instructions", tot)
    for op, n in cnt.most_common(top):
        print("%-28s %6.2f%%  %d" % (op, 100.0 * n / tot, n))


if __name__ == "__main__":
    main()
