"""Bytes of SASS per device function inside one kernel's text section:  python tools/sass_sizes.py lib.so kernel-substring
(the solve kernel is bound by instruction fetch: the hot loop has to fit the 32 KB L1.5 with room for the excursions)"""
import re, subprocess, sys, tempfile, os, glob
lib, kern = sys.argv[1], sys.argv[2]
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
tot_all = 0
for cub in sorted(glob.glob(d + "/*.cubin")):
    if "-" in os.path.basename(cub).split(".sm_")[0]: continue
    sass = subprocess.run(["nvdisasm", "-c", cub], capture_output=True, text=True).stdout
    on = False; name = None; n = 0; res = []
    for line in sass.split("\n"):
        if line.startswith(".text."):
            if on: res.append((n * 16, name))
            on = kern in line; name = "kernel body"; n = 0; continue
        if not on: continue
        m = re.match(r"^(\$\S+):\s*$", line)
        if m:
            res.append((n * 16, name)); n = 0
            name = re.sub(r"^\d+", "", m.group(1).split("EE")[-1]) if "$_ZN" in m.group(1)[1:] else m.group(1); name = m.group(1).split("$")[-1][:60]; continue
        if re.match(r"\s+/\*[0-9a-f]{4}\*/\s+\S", line): n += 1
    if on: res.append((n * 16, name))
    if res:
        print("total %d B" % sum(r[0] for r in res))
        for b, nm in res: print("%7d  %s" % (b, nm))
