"""Top SASS instructions by warp-stall samples of an ncu report (per kernel), with the dominant stall reasons.
python tools/ncu_hot.py report.ncu-rep [N]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    lines = out.splitlines()
    # first line: kernel name; second: header
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
    hdr, rows = rows[0], rows[1:]
    col = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[col["# Samples"]] or 0) for r in rows if len(r) == len(hdr))
    print(f"{lines[0][:160]}\ntotal samples {tot}")
    agg = {}
    for h in stall_cols:
        agg[h] = sum(int(r[col[h]] or 0) for r in rows if len(r) == len(hdr))
    print("stall totals:", ", ".join(f"{k[6:]}={v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
    idx = sorted(range(len(rows)), key=lambda i: -int(rows[i][col["# Samples"]] or 0) if len(rows[i]) == len(hdr) else 0)
    for i in idx[:top]:
        r = rows[i]
        s = int(r[col["# Samples"]] or 0)
        reasons = sorted(((int(r[col[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:3]
        ctx = rows[i - 1][col["Source"]].strip()[:50] if i > 0 else ""
        print(f"{s:7d} {100.0 * s / max(tot, 1):5.1f}%  #{i:5d} {r[col['Source']].strip()[:70]:70s} | prev: {ctx:50s} | "
              + ", ".join(f"{n}={c}" for c, n in reasons if c))


if __name__ == "__main__":
    main()
