import csv, json, re, collections, sys
csvf = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/launches2.csv'
calls = json.load(open(sys.argv[2] if len(sys.argv) > 2 else 'gpurun_out/gemm_calls.json'))
lines = [l for l in open(csvf) if not l.startswith('==')]
rows = list(csv.DictReader(lines))
def ms(row):
    v = float(row['Metric Value'].replace(',', '')); u = row['Metric Unit']
    return v * {'us': 1e-3, 'ns': 1e-6, 'ms': 1, 's': 1e3}.get(u, 1)
gem = [r for r in rows if 'gemm_bf16' in r['Kernel Name']]
print("gemm launches", len(gem), "calls", len(calls))
tot = sum(ms(r) for r in rows)
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for r, c in zip(gem, calls):
    fl = 2.0 * c['M'] * c['N'] * c['K'] * c['batch_i'] * c['batch_o'] * (2 if c['epi'] == 4 else 1)
    key = (c['M'], c['N'], c['K'], c['batch_i'] * c['batch_o'], c['a_major'], c['b_major'], c['epi'], c['f32'])
    agg[key][0] += 1; agg[key][1] += ms(r); agg[key][2] += fl
gt = sum(v[1] for v in agg.values()); gf = sum(v[2] for v in agg.values())
print(f"total kernel ms {tot:.1f}; gemm ms {gt:.1f} ({100*gt/tot:.1f}%), gemm avg {gf/gt/1e9:.0f} TF")
print(f"{'M':>6} {'N':>6} {'K':>6} {'bat':>5} maj epi f32 {'n':>4} {'ms':>8} {'TF/s':>7} {'lost ms vs 1400':>8}")
for key, (n, t, fl) in sorted(agg.items(), key=lambda x: -x[1][1])[:45]:
    M, N, K, bt, am, bm, epi, f32 = key
    print(f"{M:6d} {N:6d} {K:6d} {bt:5d} {am}{bm}  {epi}   {int(f32)}  {n:4d} {t:8.2f} {fl/t/1e9:7.0f} {t - fl/1.4e12:8.2f}")
other = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if 'gemm_bf16' in r['Kernel Name']: continue
    nm = re.sub(r'\(.*', '', r['Kernel Name']); other[nm][0] += 1; other[nm][1] += ms(r)
for k, (n, t) in sorted(other.items(), key=lambda x: -x[1][1])[:14]:
    print(f"{t:8.2f} ms n={n:4d} {k[:90]}")
