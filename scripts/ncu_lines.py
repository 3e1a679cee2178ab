"""Per-source-line totals (instructions executed, stall samples) from an .ncu-rep's source page."""
import csv, io, subprocess, sys


def main(rep, top=40):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    fpath, hdr, lines = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
            i_s, i_i = hdr.index("# Samples"), hdr.index("Instructions Executed")
        elif hdr and r[0] not in ("", "Function Name") and r[0].isdigit():
            num = lambda v: int(v) if v.strip().isdigit() else 0
            lines.append((fpath, int(r[0]), r[1].strip(), num(r[i_s]), num(r[i_i])))
    tot_s = sum(l[3] for l in lines) or 1
    tot_i = sum(l[4] for l in lines) or 1
    print("total samples %d, warp instructions %d" % (tot_s, tot_i))
    for f, n, src, s, i in sorted(lines, key=lambda l: -l[4])[:top]:
        print("%5.1f%% inst %5.1f%% smp  %s:%d  %s" % (100.0 * i / tot_i, 100.0 * s / tot_s, f, n, src[:110]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
