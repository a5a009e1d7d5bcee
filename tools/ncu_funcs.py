"""Warp-instructions and stall samples of an .ncu-rep (captured with --import-source on) grouped by the CUDA function
of qp_kernel.cuh the source line belongs to:  python tools/ncu_funcs.py rep [source-file] [problems-per-launch]"""
import bisect, csv, io, re, subprocess, sys
rep = sys.argv[1]
srcf = sys.argv[2] if len(sys.argv) > 2 else "qppvm_b200/csrc/qp_kernel.cuh"
nprob = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
iS, iI = hdr.index("# Samples"), hdr.index("Instructions Executed")
data = {}
for r in rows[hi + 1:]:
    try:
        data[int(r[0])] = (int(r[iS]), int(r[iI]))
    except Exception:
        pass
starts = []
for i, t in enumerate(open(srcf).read().split("\n"), 1):
    m = re.search(r"__device__\s+(?:static\s+)?(?:__forceinline__\s+|__noinline__\s+)?[\w:<>\*&]+\s+(\w+)\s*\(", t)
    if m:
        starts.append((i, m.group(1)))
    elif re.match(r"\s*qp_(solve|factor)_kernel\(", t):
        starts.append((i, t.strip().split("(")[0]))
sl = [s[0] for s in starts]
agg, ti, ts = {}, 0, 0
for ln, (s, n) in data.items():
    i = bisect.bisect_right(sl, ln) - 1
    a = agg.setdefault(starts[i][1] if i >= 0 else "?", [0, 0])
    a[0] += s; a[1] += n; ti += n; ts += s
print("total warp-instructions %d (%.0f per problem), samples %d" % (ti, ti / nprob, ts))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if v[1] > 0.002 * ti:
        print("%-22s inst %5.1f%% (%7.0f /problem)  samples %5.1f%%" % (k, 100 * v[1] / ti, v[1] / nprob, 100 * v[0] / max(1, ts)))
