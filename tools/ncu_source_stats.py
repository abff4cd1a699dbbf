import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
nwarps = float(sys.argv[2]) if len(sys.argv) > 2 else 32768.0
def f(r, k):
    try: return float(r[idx[k]])
    except: return 0.0
tot_inst = sum(f(r, "Instructions Executed") for r in data)
tot_samp = sum(f(r, "# Samples") for r in data)
print("total warp-instr per warp: %.0f   samples %d" % (tot_inst / nwarps, tot_samp))
ops = collections.defaultdict(lambda: [0, 0])
for r in data:
    s = re.sub(r'^@!?U?P\d+\s+', '', r[idx["Source"]].strip())
    op = s.split()[0] if s else "?"
    ops[op][0] += f(r, "Instructions Executed")
    ops[op][1] += f(r, "# Samples")
print("%-22s %10s %7s %8s %s" % ("opcode", "inst/warp", "inst%", "samples%", "samples/inst (rel)"))
for op, (n, sm) in sorted(ops.items(), key=lambda kv: -kv[1][1])[:28]:
    print("%-22s %10.0f %6.1f%% %7.1f%% %6.2f" % (op, n / nwarps, 100 * n / tot_inst, 100 * sm / tot_samp, (sm / tot_samp) / (n / tot_inst) if n else 0))
# segments by call targets
addr = [int(r[idx["Address"]], 16) if r[idx["Address"]].startswith("0x") else int(r[idx["Address"]]) for r in data]
base = addr[0]
targets = set()
for r in data:
    m = re.search(r'CALL\.REL\.NOINC\s+(0x[0-9a-f]+)', r[idx["Source"]])
    if m: targets.add(int(m.group(1), 16))
print("segments:")
b = sorted(targets)
bounds = [0] + b + [1 << 60]
for k in range(len(bounds) - 1):
    lo, hi = bounds[k], bounds[k + 1]
    seg = [r for r, a in zip(data, addr) if lo <= a - base < hi]
    if not seg: 
        seg = [r for r, a in zip(data, addr) if lo <= a < hi]
    n = sum(f(r, "Instructions Executed") for r in seg); sm = sum(f(r, "# Samples") for r in seg)
    stall = collections.Counter()
    for r in seg:
        for h in hdr:
            if h.startswith("stall_") and "Not Issued" not in h:
                stall[h] += f(r, h)
    top = ", ".join("%s %.0f%%" % (k2[6:], 100 * v / max(sm, 1)) for k2, v in stall.most_common(5))
    print("  @%06x static %5d  inst/warp %9.0f (%4.1f%%)  samples %4.1f%%  cyc/inst %.2f | %s" % (lo, len(seg), n / nwarps, 100 * n / tot_inst, 100 * sm / tot_samp, (sm / tot_samp) / (n / tot_inst) if n else 0, top))
# per segment opcode detail for chosen opcode
want = sys.argv[3] if len(sys.argv) > 3 else None
if want:
    print("dynamic", want, "per segment:")
    for k in range(len(bounds) - 1):
        lo, hi = bounds[k], bounds[k + 1]
        seg = [r for r, a in zip(data, addr) if lo <= a < hi]
        n = sum(f(r, "Instructions Executed") for r in seg if re.sub(r'^@!?U?P\d+\s+', '', r[idx["Source"]].strip()).split()[0].startswith(want))
        print("  @%x %9.0f" % (lo, n / nwarps))
