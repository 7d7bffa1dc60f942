"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time of the last frame + shares."""
import csv
import io
import sys


def main(path, frames):
    txt = open(path).read()
    start = txt.find('"ID"')
    rows = []
    for r in csv.DictReader(io.StringIO(txt[start:])):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            if r["Metric Unit"] in ("ns", "nsecond"):
                v /= 1000.0
            elif r["Metric Unit"] in ("ms", "msecond"):
                v *= 1000.0
            rows.append((r["Kernel Name"], v, r["Grid Size"], r["Block Size"]))
    n = len(rows) // frames
    last = rows[-n:]
    tot = sum(x[1] for x in last)
    print(f"# {path}: {len(rows)} launches, {frames} frames, last frame = {n} launches, {tot:.1f} us "
          f"(ncu per-launch times are cold-cache and serialised: compare shares)")
    for name, v, grid, block in last:
        print(f"{name[:72]:72s} {v:9.1f} us {100 * v / tot:5.1f}%  grid {grid} block {block}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 3)
