"""Summarise an `ncu --page source --csv` dump: stall reasons, executed-instruction mix, hottest instructions."""
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, data = rows[1], rows[2:]
    ismp, isrc, iex = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    tot = sum(int(r[ismp]) for r in data)
    stall = [k for k, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    agg = {hdr[k]: sum(int(r[k] or 0) for r in data) for k in stall}
    print("==", path, "SASS instructions", len(data), "samples", tot, "warp-instr executed", sum(int(r[iex]) for r in data))
    print("  stalls:", ", ".join("%s %.1f%%" % (k[6:], 100 * v / tot) for k, v in
                                 sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    ops = {}
    for r in data:
        t = r[isrc].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        ops[op] = ops.get(op, 0) + int(r[iex])
    te = sum(ops.values())
    print("  mix:", ", ".join("%s %.1f%%" % (k, 100 * v / te) for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:14]))
    top = sorted(range(len(data)), key=lambda i: -int(data[i][ismp]))[:10]
    for i in sorted(top):
        print("    #%d %.1f%% %s" % (i, 100 * int(data[i][ismp]) / tot, data[i][isrc][:70]))


if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
