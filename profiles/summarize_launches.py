"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) of `bench.py --no-graph`: per-kernel totals of ONE
training step (the span between two consecutive embedding_fwd launches), as a markdown table.

    python profiles/summarize_launches.py gpurun_out/launches.csv [step_index_from_end] > profiles/rNN_launches_step.md
"""
import collections
import csv
import re
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") == "gpu__time_duration.sum":
            v = float(row["Metric Value"].replace(",", ""))
            us = v / 1000 if row["Metric Unit"].startswith("n") else v
            rows.append((int(row["ID"]), row["Kernel Name"], us, row["Grid Size"], row["Block Size"]))
    return rows


def main():
    rows = load(sys.argv[1])
    back = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    emb = [i for i, r in enumerate(rows) if "embedding_fwd" in r[1] or "qenc_front" in r[1] or "embed" in r[1].lower() and "fwd" in r[1]]
    s, e = emb[-back - 1], emb[-back]
    step = rows[s:e]
    agg = collections.OrderedDict()
    tot = 0.0
    for _, name, us, g, b in step:
        short = re.sub(r"\(.*", "", name)
        short = re.sub(r"^void ", "", short)[:110]
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += us
        tot += us
    ours = sum(us for k, (n, us) in agg.items() if k.startswith("hca::"))
    print(f"# ncu launch list, one training step ({len(step)} launches, {tot:.1f} us summed device time; "
          f"{ours:.1f} us = {100 * ours / tot:.1f}% in this library's kernels)\n")
    print("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n")
    print("| us | share | launches | kernel |\n|---:|---:|---:|---|")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {us:.1f} | {100 * us / tot:.1f}% | {n} | `{k}` |")


if __name__ == "__main__":
    main()
