"""SASS-level stall picture of one profiled kernel (`ncu --set full --import-source on` report): every instruction that
collected at least `min_pct` of the stall samples, with its dominant stall reasons -- enough to tell WHICH warp role of a
warp-specialised kernel is waiting, because inlined helpers (mbarrier waits) appear once per call site.

  python tools/ncu_sass_stalls.py gpurun_out/x.ncu-rep [min_pct]
"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.7
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None
    ins = []
    for r in rows:
        if r and r[0] == "Address":
            hdr = r
        elif hdr and r and r[0].startswith("0x"):
            ins.append(dict(zip(hdr, r)))
    tot = sum(int(d["# Samples"]) for d in ins) or 1
    print(f"# {rep}: {len(ins)} instructions, {tot} samples")
    for i, d in enumerate(ins):
        s = int(d["# Samples"])
        if 100.0 * s / tot < min_pct:
            continue
        why = sorted(((int(v), k[6:]) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v)),
                     reverse=True)[:3]
        print(f"{i:5d} {100.0 * s / tot:5.1f}%  exec={d['Instructions Executed']:>8s}  {d['Source'].strip()[:70]:70s} "
              + ",".join(f"{k}={v}" for v, k in why))


if __name__ == "__main__":
    main()
