"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel (markdown table)."""
import collections, csv, re, sys

def summarise(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = float(row['Metric Value'].replace(',', '')); unit = row['Metric Unit']
        v = v / 1e3 if unit == 'ns' else (v * 1e3 if unit == 'ms' else v)
        agg.setdefault(re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', ''), []).append(v)
    tot = sum(sum(v) for v in agg.values())
    out = ["| kernel | launches | mean us | max us | sum us | share |", "|---|---|---|---|---|---|"]
    for k, v in agg.items():
        out.append("| %s | %d | %.1f | %.1f | %.1f | %.1f%% |" % (k, len(v), sum(v) / len(v), max(v), sum(v), 100 * sum(v) / tot))
    out.append("| total | | | | %.1f | |" % tot)
    return "\n".join(out)

if __name__ == "__main__":
    print(summarise(sys.argv[1]))
