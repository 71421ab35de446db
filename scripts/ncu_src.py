"""Top SASS instructions of an `ncu --page source --csv` dump by stall samples / executed instructions.
usage: ncu -i rep.ncu-rep --page source --csv [--kernel-name regex:x] > src.csv; python scripts/ncu_src.py src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
ci = {n: i for i, n in enumerate(h)}
body = [r for r in rows[hi + 1:] if len(r) >= len(h) - 2 and r[0] != "Address"]
def f(r, n):
    try: return float(r[ci[n]])
    except Exception: return 0.0
tot_i = sum(f(r, "Instructions Executed") for r in body)
tot_s = sum(f(r, "# Samples") for r in body)
print("warp instructions %.0f  samples %.0f  sass lines %d" % (tot_i, tot_s, len(body)))
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
agg = {n: sum(f(r, n) for r in body) for n in stalls}
print("stalls:", {k[6:]: int(v) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
for key in ("# Samples", "Instructions Executed"):
    print("---- top by", key)
    for r in sorted(body, key=lambda r: -f(r, key))[:top]:
        st = sorted(((f(r, n), n[6:]) for n in stalls), reverse=True)[:2]
        print("%5d %-70s inst %9.0f smp %6.0f thr %4.1f  %s" % (body.index(r), r[ci["Source"]][:70], f(r, "Instructions Executed"), f(r, "# Samples"),
                                                        f(r, "Avg. Threads Executed"), " ".join("%s:%d" % (n, v) for v, n in st if v)))
