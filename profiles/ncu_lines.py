"""Per-source-line hot spots of one kernel in an .ncu-rep (needs -lineinfo at compile time).
usage: python profiles/ncu_lines.py REPORT KERNEL_ID [top_n]
Prints, for the source lines with the most warp-stall samples / executed instructions:
file:line, samples, warp instructions executed, the dominant stall reasons and the source text."""
import csv
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def num(v):
    try:
        return int(v)
    except (TypeError, ValueError):
        return 0


def main():
    rep, kid = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    files = ",".join(glob.glob(os.path.join(ROOT, "hm-16.2_b200", "csrc", "*")))
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", kid, "--launch-count", "1",
                          "--resolve-source-file", files], capture_output=True, text=True).stdout
    rows, cur_file, hdr = [], None, None
    for r in csv.reader(raw.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = os.path.basename(r[1]); continue
        if r[0] == "Function Name":
            print(r[1][:150]); continue
        if r[0] == "Line No":
            hdr = r; continue
        if hdr and r[0].strip().isdigit():
            d = dict(zip(hdr[4:], r[len(r) - (len(hdr) - 4):]))   # source text may itself contain quotes / commas
            stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v)}
            rows.append((cur_file, int(r[0]), num(d.get("# Samples")), num(d.get("Instructions Executed")), stalls, r[1].strip()))
    tot_s = sum(x[2] for x in rows) or 1
    tot_i = sum(x[3] for x in rows) or 1
    print("total samples %d, total warp instructions %d" % (tot_s, tot_i))
    for key, name in ((2, "stall samples"), (3, "instructions")):
        print("---- top lines by", name)
        for f, ln, s, i, st, src in sorted(rows, key=lambda x: -x[key])[:top]:
            top_st = " ".join("%s:%d" % kv for kv in sorted(st.items(), key=lambda kv: -kv[1])[:3])
            print("%-22s %5.1f%% smp %5.1f%% inst  %-40s | %s" % ("%s:%d" % (f, ln), 100.0 * s / tot_s, 100.0 * i / tot_i, top_st, src[:90]))


if __name__ == "__main__":
    main()
