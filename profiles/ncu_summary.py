import csv, collections, re, sys
def summarize(path, top=22):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    rows = list(csv.DictReader(lines))
    tot = collections.Counter(); cnt = collections.Counter()
    for row in rows:
        name = re.sub(r'\(.*', '', row['Kernel Name'])
        v = float(row['Metric Value'].replace(',',''))
        unit = row['Metric Unit']
        v = v/1e3 if unit in ('ns','nsecond') else (v*1e3 if unit in ('ms','msecond') else v)
        tot[name]+=v; cnt[name]+=1
    T = sum(tot.values())
    out = ['%d launches, %.1f us total (cold-cache, serialised)' % (len(rows), T)]
    for k,v in tot.most_common(top):
        out.append('%-72s n=%4d %10.1f us %5.1f%%' % (k[:72], cnt[k], v, 100*v/T))
    return '\n'.join(out)
if __name__ == '__main__':
    for p in sys.argv[1:]:
        print(p); print(summarize(p)); print()
