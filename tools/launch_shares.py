"""Per-kernel share of the device time in an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv`):
python tools/launch_shares.py profiles/<launches>.csv  -> one line per kernel family, percent of the summed time."""
import collections
import csv
import re
import sys


def shares(path):
    hdr, tot = None, collections.Counter()
    for r in csv.reader(open(path)):
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            try:
                v = float(d["Metric Value"].replace(",", ""))
            except ValueError:
                continue
            name = re.sub(r"^.*?(\w+)(<[^>]*>)?\(.*$", r"\1\2", d["Kernel Name"])
            tot[name] += v
    s = sum(tot.values())
    return {k: 100 * v / s for k, v in tot.most_common()}, s


if __name__ == "__main__":
    sh, total = shares(sys.argv[1])
    for k, v in sh.items():
        print(f"{v:6.1f} %  {k}")
    print(f"total {total:.0f} (units of the file)")
