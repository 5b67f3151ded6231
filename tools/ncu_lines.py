"""Per-source-line summary of an ncu report's source page (development aid).

    ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > X.csv
    python tools/ncu_lines.py X.csv [top_n] [lo-hi ...]

Prints, per CUDA source line: warp-instructions executed, stall samples, the dominant stall reasons; then the
totals for the given line ranges (e.g. the functions of one warp role).
"""
import csv
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    ranges = [tuple(int(x) for x in a.split("-")) for a in sys.argv[3:]]
    rows = list(csv.reader(open(path)))
    hdr = None
    lines = {}
    fname = ""
    for r in rows:
        if r and r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or not r or not r[0].strip().isdigit():
            continue
        d = dict(zip(hdr[4:], r[4:]))
        ln = int(r[0])

        def num(k):
            try:
                return float(d.get(k, "0"))
            except ValueError:
                return 0.0
        stalls = {k[6:]: num(k) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}
        key = (fname, ln)
        e = lines.setdefault(key, {"src": r[1].strip(), "inst": 0.0, "samples": 0.0, "stalls": {}})
        e["inst"] += num("Instructions Executed")
        e["samples"] += num("# Samples")
        for k, v in stalls.items():
            e["stalls"][k] = e["stalls"].get(k, 0.0) + v
    ti = sum(e["inst"] for e in lines.values()) or 1
    ts = sum(e["samples"] for e in lines.values()) or 1
    print(f"total warp-inst {ti:.3e}  samples {ts:.0f}")
    for (f, ln), e in sorted(lines.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        st = sorted(e["stalls"].items(), key=lambda kv: -kv[1])[:3]
        sts = " ".join(f"{k}:{v:.0f}" for k, v in st if v > 0)
        print(f"{f}:{ln:5d} inst {100*e['inst']/ti:5.1f}%  samp {100*e['samples']/ts:5.1f}%  [{sts}]  {e['src'][:90]}")
    for lo, hi in ranges:
        i = sum(e["inst"] for (f, ln), e in lines.items() if lo <= ln <= hi)
        s = sum(e["samples"] for (f, ln), e in lines.items() if lo <= ln <= hi)
        agg = {}
        for (f, ln), e in lines.items():
            if lo <= ln <= hi:
                for k, v in e["stalls"].items():
                    agg[k] = agg.get(k, 0) + v
        st = " ".join(f"{k}:{v:.0f}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:5])
        print(f"lines {lo}-{hi}: inst {100*i/ti:5.1f}%  samples {100*s/ts:5.1f}%  [{st}]")


if __name__ == "__main__":
    main()
