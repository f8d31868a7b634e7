"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: total and share per kernel.

    python tools/summarize_launches.py gpurun_out/launches.csv <epochs in the run> > profiles/..._summary.txt
"""
import csv, sys
from collections import defaultdict

path, epochs = sys.argv[1], int(sys.argv[2])
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10 and r[0].isdigit()]
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows:
    name, unit, val = r[4], r[13], float(r[14].replace(",", ""))
    ms = val / 1e6 if unit in ("ns", "nsecond") else val / 1e3 if unit in ("us", "usecond") else val
    short = name.split("(")[0].replace("void ", "").replace("dbgsom::<unnamed>::", "").replace("dbgsom::", "")
    tot[short] += ms
    cnt[short] += 1
ours = {k: v for k, v in tot.items() if not k.startswith("at::") and "at::" not in k and "elementwise" not in k and "nccl" not in k.lower()}
per_epoch = sum(ours.values()) / epochs
print(f"cold-cache, serialised launches: compare SHARES, not absolutes. {epochs} epochs in the capture; torch kernels = data generation / glue.")
print(f"dbgsom kernels: {per_epoch:.3f} ms per epoch")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:40]:
    share = f"{100 * v / sum(ours.values()):5.1f}%" if k in ours else "  -   "
    print(f"{v:10.3f} ms total  n={cnt[k]:4d}  mean {v / cnt[k]:9.4f} ms  share {share}  {k[:110]}")
