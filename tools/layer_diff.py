"""Side-by-side per-layer device times of several `tools/gpu_diag.py layers` logs (first log = reference).
usage: python tools/layer_diff.py [--all] log0 log1 ...   (default: only layers that move by more than 3 %)"""
import re
import sys


def load(path):
    rows = {}
    order = []
    for l in open(path):
        m = re.match(r"(\S+)\s+(conv_tc|conv_ref|conv0|bnrelu|head)\s.*?([\d.]+) ms", l)
        if m and not l.startswith(" "):
            name = m.group(1)
            while name in rows:
                name += "'"
            rows[name] = float(m.group(3))
            order.append(name)
    return rows, order


def main():
    args = [a for a in sys.argv[1:] if a != "--all"]
    show_all = "--all" in sys.argv
    logs = [load(p) for p in args]
    ref, order = logs[0]
    print("%-46s" % "layer" + "".join("%12s" % p.split("_")[-1].replace(".log", "") for p in args))
    tot = [0.0] * len(logs)
    for n in order:
        vals = [lg[0].get(n, float("nan")) for lg in logs]
        for i, v in enumerate(vals):
            tot[i] += v if v == v else 0.0
        if show_all or any(abs(v - vals[0]) > 0.03 * vals[0] for v in vals[1:]):
            print("%-46s" % n[:46] + "".join("%12.4f" % v for v in vals))
    print("%-46s" % "TOTAL" + "".join("%12.3f" % t for t in tot))


if __name__ == "__main__":
    main()
