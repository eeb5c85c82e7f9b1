"""Print the per-kernel durations of the last full training step in an `ncu --metrics gpu__time_duration.sum --csv` log."""
import csv
import sys


def main(path, end_marker='rows_zero'):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rows = [(x['Kernel Name'], float(x['Metric Value']) / 1000) for x in csv.DictReader(lines)]
    ends = [i for i, (k, _) in enumerate(rows) if end_marker in k]
    s, e = ends[-2] + 1, ends[-1]
    tot = 0.0
    for k, v in rows[s:e + 1]:
        k = k.replace('void ', '').replace('rpb::', '')
        print(f"{v:8.1f}  {k[:100]}")
        tot += v
    print(f"{tot:8.1f}  total ({e - s + 1} launches)")


if __name__ == '__main__':
    main(*sys.argv[1:])
