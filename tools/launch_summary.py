"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python tools/launch_summary.py launches.csv [skip_launches] [n_steps]"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    steps = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
    rows = rows[skip:]
    d = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        n = r[4]
        n = n.replace("pn2::<unnamed>::", "").replace("void ", "")
        d[n[:90]][0] += 1
        d[n[:90]][1] += float(r[14])
    tot = sum(v[1] for v in d.values())
    print("%d launches, %.3f ms total, %.3f ms/step, %.0f launches/step" % (len(rows), tot / 1e6, tot / 1e6 / steps, len(rows) / steps))
    for n, v in sorted(d.items(), key=lambda kv: -kv[1][1])[:60]:
        print("%7.1f %9.1f us/step %5.1f%%  %7.1f us/launch  %s" % (v[0] / steps, v[1] / 1e3 / steps, 100 * v[1] / tot, v[1] / 1e3 / v[0], n))


if __name__ == "__main__":
    main()
