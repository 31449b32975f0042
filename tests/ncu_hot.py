"""Hot instructions of one kernel from `ncu --page source --csv --print-source sass` (whole-file dump).

    python tests/ncu_hot.py src.csv <kernel-index> [top]
"""
import csv
import sys


def main(path, which, top=24):
    lines = open(path).read().splitlines()
    starts = [i for i, l in enumerate(lines) if l.startswith('"Kernel Name"')]
    a = starts[which]
    b = starts[which + 1] if which + 1 < len(starts) else len(lines)
    rows = list(csv.reader(lines[a:b]))
    print(rows[0][1][:100])
    hdr = rows[1]
    iS, iE, iP = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    body = [(k, r) for k, r in enumerate(rows[2:]) if len(r) > iP and r[iP].strip()]
    tots = sum(int(r[iP] or 0) for _, r in body)
    tote = sum(int(r[iE] or 0) for _, r in body)
    agg = {}
    for _, r in body:
        for i in stall:
            agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
    print(f"samples {tots}, warp instructions {tote}; stall totals:",
          ", ".join(f"{k[6:]} {100 * v / max(tots, 1):.0f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]))
    hot = sorted(body, key=lambda kr: -int(kr[1][iP] or 0))[:top]
    for k, r in sorted(hot, key=lambda kr: kr[0]):
        st = sorted([(int(r[i] or 0), hdr[i][6:]) for i in stall], reverse=True)[:2]
        print(f"{k:5d} {100 * int(r[iP]) / tots:5.1f}%  ex {int(r[iE]):9d}  {r[iS].strip()[:64]:64s} {st}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 24)
