"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name, launches / mean / total."""
import csv, sys, collections, re
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]
kn, mv, mu = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    name = re.sub(r"\(.*", "", r[kn]).replace("void ", "").replace("<unnamed>::", "")
    v = float(r[mv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[mu], 1.0)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
import signal
signal.signal(signal.SIGPIPE, signal.SIG_DFL)  # `| head` closes the pipe early
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%8d x %9.2f us = %10.1f us  %5.1f%%  %s" % (n, t / n, t, 100 * t / tot, name[:110]))
