"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (time share per kernel class)."""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i
        break
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg, cnt = collections.Counter(), collections.Counter()
for r in rows[start + 2:]:
    if len(r) <= iv:
        continue
    name = re.sub(r"\(.*", "", r[ik])
    name = re.sub(r"^void ", "", name)[:64]
    try:
        t = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    agg[name] += t
    cnt[name] += 1
tot = sum(agg.values())
mine = sum(v for k, v in agg.items() if k.startswith("hipie::"))
print(f"total {tot/1e6:.2f} ms over {sum(cnt.values())} launches; hipie:: kernels {mine/1e6:.2f} ms ({mine/tot*100:.1f} %), torch glue {(tot-mine)/1e6:.2f} ms")
for k, v in agg.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print(f"{v/1e6:8.2f} ms {v/tot*100:5.1f}% n={cnt[k]:4d} {k}")
