"""Per-class breakdown of a `tools/gpu_diag.py layers <mode> <B>` log: which layer classes take the time and what bounds
each.  usage: python tools/layer_classes.py gpurun_out/<log> [...]"""
import re
import sys


def classify(name, k, stride, cin, cout, out, bn, halo):
    if ".dense.units." in name and "conv2" in name:
        return "dense grouped kxk (128->32, g=4)"
    if k > 1:
        if cout >= 128:
            return "wide kxk (cout >= 128)"
        return "thin kxk (cout 64)"
    if ".dense.units." in name or name.endswith("convf.weight"):
        return "1x1 XF decoder (dense conv1, convf)"
    if "conv1" in name and ".units.0." not in name:
        return "1x1 XF encoder (pre-activation conv1)"
    if "conv3" in name and "shortcut" not in name:
        return "1x1 + residual (conv3, d0 at full resolution)" if name.startswith("d0.") else "1x1 + residual (conv3, d1-d3)"
    if name.startswith("d0."):
        return "1x1 plain, d0 (full resolution)"
    return "1x1 plain (unit-0 conv1, conv3+shortcut, conv_bot)"


def main(path):
    rows = {}
    other = {}
    for l in open(path):
        m = re.match(r"(\S+)\s+conv_tc\s+k(\d)x\d s(\d) cin(\d+)\s+cout(\d+)\s+out(\d+)x\d+ box=\S+?(H?) bn=(\d+)\s+([\d.]+) ms\s+([\d.]+) GFLOP", l)
        if m:
            name, k, st, cin, cout, out, halo, bn, ms, gf = m.groups()
            c = classify(name, int(k), int(st), int(cin), int(cout), int(out), int(bn), halo)
            r = rows.setdefault(c, [0, 0.0, 0.0])
            r[0] += 1; r[1] += float(ms); r[2] += float(gf)
            continue
        m = re.match(r"(\S+)\s+(conv0|bnrelu|head)\s+([\d.]+) ms", l)
        if m:
            o = other.setdefault(m.group(2), [0, 0.0])
            o[0] += 1; o[1] += float(m.group(3))
    tot = sum(r[1] for r in rows.values())
    print("## %s  (conv_tc %.2f ms, %.0f GFLOP, %.0f TFLOP/s)" % (path, tot, sum(r[2] for r in rows.values()),
                                                              sum(r[2] for r in rows.values()) / tot))
    print("| class | launches | ms | share | GFLOP | TFLOP/s |\n|---|---:|---:|---:|---:|---:|")
    for c, r in sorted(rows.items(), key=lambda x: -x[1][1]):
        print("| %s | %d | %.3f | %.1f %% | %.0f | %.0f |" % (c, r[0], r[1], 100 * r[1] / tot, r[2], r[2] / r[1]))
    for c, o in other.items():
        print("| (%s) | %d | %.3f | | | |" % (c, o[0], o[1]))


if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
