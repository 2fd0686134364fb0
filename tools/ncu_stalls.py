"""Aggregate warp-stall reasons of a .ncu-rep (source page): totals, and per reason the top SASS sites."""
import csv, io, subprocess, sys


def main(path, top=6):
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = next(r for r in rows if r and r[0] == "Address")
    body = [r for r in rows if len(r) == len(hdr) and r[0].startswith("0x")]
    cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = {hdr[i]: sum(int(r[i] or 0) for r in body) for i in cols}
    all_s = sum(tot.values()) or 1
    print(f"{len(body)} SASS lines, {all_s} stall samples")
    for name, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]:
        i = hdr.index(name)
        print(f"{100 * v / all_s:5.1f}%  {name}")
        for r in sorted(body, key=lambda r: -int(r[i] or 0))[:top]:
            if int(r[i] or 0) * 50 > v:
                print(f"        {100 * int(r[i]) / all_s:5.1f}%  {r[1].strip()[:90]}")
    # executed instruction mix
    ie = hdr.index("Instructions Executed")
    tot_i = sum(int(r[ie] or 0) for r in body) or 1
    print(f"warp instructions executed: {tot_i}")
    for r in sorted(body, key=lambda r: -int(r[ie] or 0))[:10]:
        print(f"        {100 * int(r[ie]) / tot_i:5.1f}%  {r[1].strip()[:90]}")


if __name__ == "__main__":
    main(sys.argv[1])
