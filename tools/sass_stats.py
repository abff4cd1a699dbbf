import re, sys, collections
# usage: sass_stats.py sass.txt kernel_substr
txt = open(sys.argv[1]).read().split('\n')
kern = sys.argv[2]
start = None
for i, l in enumerate(txt):
    if 'Function :' in l:
        if start is not None and end is None:
            end = i
        if kern in l and start is None:
            start = i; end = None
if end is None: end = len(txt)
ins = []
rx = re.compile(r'^\s+/\*([0-9a-f]+)\*/\s+(.*?);')
for l in txt[start:end]:
    m = rx.match(l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
targets = set()
for a, s in ins:
    m = re.search(r'CALL\.REL\.NOINC\s+(0x[0-9a-f]+)', s)
    if m: targets.add(int(m.group(1), 16))
targets = sorted(targets)
bounds = [0] + targets + [ins[-1][0] + 16]
print("kernel", kern, "instructions", len(ins), "call targets", len(targets))
for k in range(len(bounds) - 1):
    lo, hi = bounds[k], bounds[k + 1]
    seg = [s for a, s in ins if lo <= a < hi]
    c = collections.Counter()
    for s in seg:
        s2 = re.sub(r'^@!?U?P\d+\s+', '', s)
        op = s2.split()[0]
        c[op] += 1
    wide = sum(v for o, v in c.items() if o.startswith('IMAD.WIDE'))
    hi_ = sum(v for o, v in c.items() if o.startswith('IMAD.HI'))
    imadx = sum(v for o, v in c.items() if o.startswith('IMAD') and not o.startswith('IMAD.WIDE') and not o.startswith('IMAD.HI'))
    calls = sum(v for o, v in c.items() if o.startswith('CALL'))
    mem = sum(v for o, v in c.items() if o[:3] in ('LDL', 'STL', 'LDS', 'STS', 'LDG', 'STG') or o[:2] in ('LD', 'ST'))
    n = len(seg)
    print("  seg @%06x n=%5d wide=%4d imad.hi=%3d imad_other=%4d mem=%4d calls=%3d alu_other=%5d" % (lo, n, wide, hi_, imadx, mem, calls, n - wide - hi_ - imadx - mem))
