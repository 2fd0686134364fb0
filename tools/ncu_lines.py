"""Per CUDA source line: stall samples and executed warp instructions of a .ncu-rep (needs -lineinfo)."""
import csv, io, subprocess, sys


def main(path, top=25):
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    fname = ""
    lines = []      # (file, line no, text, samples, instructions)
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        if len(r) > 8 and r[0].isdigit():
            lines.append((fname, int(r[0]), r[1].strip(), int(r[6]) if r[6].isdigit() else 0, int(r[7]) if r[7].isdigit() else 0))
    tot_s = sum(l[3] for l in lines) or 1
    tot_i = sum(l[4] for l in lines) or 1
    print(f"{tot_s} samples, {tot_i} warp instructions")
    print("-- by stall samples")
    for f, n, t, s, i in sorted(lines, key=lambda l: -l[3])[:top]:
        print(f"  {100 * s / tot_s:5.1f}% samples {100 * i / tot_i:5.1f}% instr  {f}:{n}  {t[:110]}")
    print("-- by instructions")
    for f, n, t, s, i in sorted(lines, key=lambda l: -l[4])[:12]:
        print(f"  {100 * s / tot_s:5.1f}% samples {100 * i / tot_i:5.1f}% instr  {f}:{n}  {t[:110]}")


if __name__ == "__main__":
    main(sys.argv[1])
