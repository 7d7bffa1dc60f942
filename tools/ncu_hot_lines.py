"""Per-source-line hot spots of one profiled kernel: `ncu --set full --import-source on` report -> top lines by
instructions executed and by stall samples.

  python tools/ncu_hot_lines.py gpurun_out/x.ncu-rep [top_n]
"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur_file, hdr, lines = "", None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Name":
            cur_file = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and r[0].isdigit():
            d = dict(zip(hdr, r))
            try:
                lines.append((int(d["Instructions Executed"]), int(d["# Samples"]), cur_file, int(r[0]), r[1].strip(),
                              {k: int(d[k]) for k in hdr if k.startswith("stall_") and "Not Issued" not in k and d[k].isdigit()}))
            except (ValueError, KeyError):
                pass
    ti = sum(x[0] for x in lines) or 1
    ts = sum(x[1] for x in lines) or 1
    print(f"# {rep}: {ti} warp instructions, {ts} stall samples")
    print("# by instructions executed")
    for n, s, f, ln, src, st in sorted(lines, key=lambda x: -x[0])[:top]:
        print(f"{100 * n / ti:5.1f}% inst {100 * s / ts:5.1f}% smpl  {f}:{ln:<4d} {src[:110]}")
    print("# by stall samples")
    for n, s, f, ln, src, st in sorted(lines, key=lambda x: -x[1])[:top]:
        why = ",".join(f"{k[6:]}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3] if v)
        print(f"{100 * s / ts:5.1f}% smpl {100 * n / ti:5.1f}% inst  {f}:{ln:<4d} {src[:80]}  [{why}]")


if __name__ == "__main__":
    main()
