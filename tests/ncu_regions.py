"""Aggregates an `ncu --page source --csv --print-source sass` dump by code region.

Regions are delimited by the SASS addresses of CALL / RET and by instruction density: the
script prints, per contiguous address range of `step` instructions, the executed warp
instructions and stall samples, so the shares of hot loop / drain / staging can be read off.
"""
import csv
import sys


def main(path, nbuckets=24):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ia, isrc, iex, ismp, ithr = (hdr.index(k) for k in ("Address", "Source", "Instructions Executed", "# Samples",
                                                       "Thread Instructions Executed"))
    ins = []
    for r in rows[2:]:
        if len(r) <= ithr or not r[iex].strip():
            continue
        try:
            ins.append((r[ia], r[isrc], int(r[iex]), int(r[ismp] or 0), int(r[ithr] or 0)))
        except ValueError:
            continue
    tot = sum(i[2] for i in ins)
    tots = sum(i[3] for i in ins)
    print(f"{len(ins)} instructions, {tot} executed, {tots} samples")
    # split at control-flow landmarks
    marks = [k for k, i in enumerate(ins) if any(t in i[1] for t in ("CALL", "RET", "EXIT"))]
    edges = sorted(set([0] + [m + 1 for m in marks] + [len(ins)]))
    for a, b in zip(edges[:-1], edges[1:]):
        ex = sum(i[2] for i in ins[a:b])
        sm = sum(i[3] for i in ins[a:b])
        th = sum(i[4] for i in ins[a:b])
        if ex == 0:
            continue
        print(f"[{a:5d},{b:5d}) {ins[a][0][-5:]}..{ins[b - 1][0][-5:]}  exec {100.0 * ex / tot:6.2f}%  samples {100.0 * sm / max(tots, 1):6.2f}%  "
              f"avg threads {th / ex:5.1f}  last: {ins[b - 1][1][:40]}")


if __name__ == "__main__":
    main(sys.argv[1])
