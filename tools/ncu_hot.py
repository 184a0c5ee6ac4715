"""Summarise the SASS page of one ncu launch: top instructions by stall samples.
usage: ncu -i rep --page source --csv --launch-skip N --launch-count 1 > x.csv; python tools/ncu_hot.py x.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print(rows[0][1][:100])
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_")]
data = []
for r in rows[2:]:
    try:
        s = int(r[ix["# Samples"]])
    except Exception:
        continue
    data.append((s, r))
tot = sum(s for s, _ in data)
print("total samples", tot, "instructions", len(data))
agg = {h: 0 for h in stalls}
for s, r in data:
    for h in stalls:
        try:
            agg[h] += int(r[ix[h]])
        except Exception:
            pass
print({k: v for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v})
order = sorted(range(len(data)), key=lambda i: -data[i][0])[:top]
for i in sorted(order):
    s, r = data[i]
    st = {h[6:]: int(r[ix[h]]) for h in stalls if r[ix[h]] not in ("", "0")}
    st = dict(sorted(st.items(), key=lambda x: -x[1])[:3])
    print("%5d %5.1f%%  %-4d %-70s %s" % (s, 100.0 * s / max(tot, 1), i, r[ix["Source"]][:70], st))
