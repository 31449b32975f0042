"""Per-launch table of one rank search from an ncu launch list (the last of the repeated searches)."""
import csv
import sys


def main(path, reps=3):
    lines = open(path).read().splitlines()
    st = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rows = list(csv.DictReader(lines[st:]))
    n = len(rows) // reps
    rows = rows[-n:]
    tot = 0.0
    agg = {}
    for r in rows:
        v = float(r["Metric Value"]) / 1000
        tot += v
        name = r["Kernel Name"].split("(")[0].split("::")[-1]
        agg.setdefault(name, [0, 0.0])
        agg[name][0] += 1
        agg[name][1] += v
        if "-v" in sys.argv:
            print(f"  {name:32s} grid {r['Grid Size']:>14s} {v:8.1f} us")
    print(f"{path}: {n} launches per search, {tot:.1f} us of device time")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"  {k:32s} {v[0]:4d} x {v[1]:9.1f} us  {100 * v[1] / tot:5.1f} %")


if __name__ == "__main__":
    main(sys.argv[1])
