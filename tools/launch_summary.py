"""Condenses an ncu launch-list CSV (gpu__time_duration + dram bytes) into profiles/*.txt and traffic.json."""
import csv, collections, json, sys
src, out_txt, out_json, title = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]
rows = [r for r in csv.reader(open(src)) if len(r) > 5]
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(lambda: collections.defaultdict(float)); cnt = collections.Counter()
for r in rows[1:]:
    name = r[ix['Kernel Name']].split('(')[0].replace('void ', '').replace('ehic::', '')
    m = r[ix['Metric Name']]; v = float(r[ix['Metric Value']].replace(',', '')); u = r[ix['Metric Unit']]
    if m == 'gpu__time_duration.sum':
        v = v / 1e3 if u == 'ns' else (v if u == 'us' else v * 1e3)
        cnt[name] += 1
    else:
        v = v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
    agg[name][m] += v
step_k = ('forward', 'prior', 'gather', 'backward', 'combine', 'pack', 'bbox')
tot = sum(a['gpu__time_duration.sum'] for k, a in agg.items() if any(s in k for s in step_k))
lines = [title, "(cold-cache, serialised launches: compare SHARES; share = of the kernels of one posterior-gradient step)",
         "%-40s %5s %11s %7s %14s %14s" % ("kernel", "n", "avg us", "share", "dram rd MB/l", "dram wr MB/l")]
out = {}
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]['gpu__time_duration.sum']):
    n = cnt[name]; t = a['gpu__time_duration.sum']
    instep = any(s in name for s in step_k)
    lines.append("%-40s %5d %11.1f %7s %14.2f %14.2f" % (name[:40], n, t / n, ("%.1f%%" % (100 * t / tot)) if instep else "-",
                                                          a['dram__bytes_read.sum'] / n / 1e6, a['dram__bytes_write.sum'] / n / 1e6))
    out[name] = {'launches': n, 'avg_us': t / n, 'dram_bytes_per_launch': (a['dram__bytes_read.sum'] + a['dram__bytes_write.sum']) / n}
open(out_txt, 'w').write("\n".join(lines) + "\n")
print("\n".join(lines))
dom = max((k for k in out if any(s in k for s in ('backward', 'gather'))), key=lambda k: out[k]['avg_us'])
step = sum(v['dram_bytes_per_launch'] for k, v in out.items() if any(s in k for s in step_k))
json.dump({'source': out_txt + ' (ncu dram__bytes_read.sum + dram__bytes_write.sum per launch; config D, 8 replicas per launch)',
           'dominant_kernel': dom, 'gather_dram_bytes_per_launch': out[dom]['dram_bytes_per_launch'], 'step_dram_bytes': step},
          open(out_json, 'w'), indent=1)
