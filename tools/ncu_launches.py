"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (run on the CPU box).
usage: python tools/ncu_launches.py launches.csv "header comment" > profiles/rN_launches_*.txt"""
import csv
import re
import sys


def main(path, note=""):
    rows = list(csv.reader(l for l in open(path, errors="replace") if l.startswith('"')))
    hdr = rows[0]
    iname, imetric, ival, iunit = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, total, n = {}, 0.0, 0
    for r in rows[1:]:
        if len(r) <= ival or r[imetric] != "gpu__time_duration.sum":
            continue
        v = float(r[ival].replace(",", ""))
        v = v / 1e3 if r[iunit] in ("ns", "nsecond") else (v * 1e3 if r[iunit] in ("ms", "msecond") else v)
        name = re.sub(r"\(anonymous namespace\)", "<unnamed>", r[iname])
        name = re.sub(r"\(.*$", "", name)[:150]
        a = agg.setdefault(name, [0.0, 0])
        a[0] += v; a[1] += 1; total += v; n += 1
    if note:
        print("# " + note)
    print(f"# {n} launches, {total / 1e3:.2f} ms of kernel time in total.  Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.")
    for name, (us, k) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{us:10.1f} us {k:5d} launches {100 * us / total:5.1f}%  {name}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
