"""Per source line totals from `ncu -i rep --page source --csv --print-source cuda,sass`:
instructions executed, stall samples, shared-memory wavefronts; grouped by file:line.

    python tests/ncu_lines.py dump.csv [file-substring] [top]
"""
import csv
import sys


def main(path, want="x3_search_seg.cu", top=45):
    rows = list(csv.reader(open(path)))
    cur = None
    hdr = None
    agg = {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            iL, iS = 0, 1
            iE = hdr.index("Instructions Executed")
            iP = hdr.index("# Samples")
            iW = hdr.index("L1 Wavefronts Shared")
            iX = hdr.index("L1 Wavefronts Shared Excessive")
            stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
            continue
        if hdr is None or cur is None or want not in cur:
            continue
        if r[2] != "-":   # SASS row under a source line: skip (the source row carries the totals)
            continue
        try:
            key = (cur.split("/")[-1], int(r[iL]))
        except ValueError:
            continue
        a = agg.setdefault(key, dict(src=r[iS].strip(), ex=0, smp=0, wf=0, xs=0, st={}))
        a["ex"] += int(r[iE] or 0)
        a["smp"] += int(r[iP] or 0)
        a["wf"] += int(r[iW] or 0)
        a["xs"] += int(r[iX] or 0)
        for i in stall:
            a["st"][hdr[i][6:]] = a["st"].get(hdr[i][6:], 0) + int(r[i] or 0)
    tote = sum(a["ex"] for a in agg.values())
    tots = sum(a["smp"] for a in agg.values())
    print(f"{want}: warp instructions {tote}, samples {tots}")
    tot_st = {}
    for a in agg.values():
        for k, v in a["st"].items():
            tot_st[k] = tot_st.get(k, 0) + v
    print("stalls:", ", ".join(f"{k} {100 * v / max(tots, 1):.0f}%" for k, v in sorted(tot_st.items(), key=lambda kv: -kv[1])[:8]))
    for key, a in sorted(sorted(agg.items(), key=lambda kv: -kv[1]["smp"])[:top]):
        st = sorted(((v, k) for k, v in a["st"].items()), reverse=True)[:2]
        print(f"{key[1]:5d} smp {100 * a['smp'] / max(tots, 1):5.1f}%  ex {100 * a['ex'] / max(tote, 1):5.1f}%  wf {a['wf']:9d} (+{a['xs']:8d})  "
              f"{a['src'][:70]:70s} {[(k, v) for v, k in st]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "x3_search_seg.cu", int(sys.argv[3]) if len(sys.argv) > 3 else 45)
