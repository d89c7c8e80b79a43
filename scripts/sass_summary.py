"""Per-kernel counts of the Blackwell-specific SASS instructions in the built library (tcgen05 MMA = UTCHMMA / UTCMMA
(tf32: UTCMMA? listed as found), TMA = UTMALDG / UTMASTG, TMEM loads = LDTM, TMEM alloc = UTCATOMSWS / UTCBAR, ...).

    python scripts/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "r3d_b200", "csrc", "libr3d_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTC[A-Z0-9]*MMA[A-Z0-9.]*|UTMALDG[A-Z0-9.]*|UTMASTG[A-Z0-9.]*|LDTM[A-Z0-9.]*|STTM[A-Z0-9.]*|UTCBAR[A-Z0-9.]*|UTCATOMSWS[A-Z0-9.]*|SYNCS[A-Z0-9.]*|MUFU\.TANH|HMMA[A-Z0-9.]*)")
cur, counts, total = None, collections.OrderedDict(), collections.Counter()
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for op in pat.findall(ln):
        base = op.split(".")[0] if not op.startswith("MUFU") else op
        counts[cur][base] += 1
        total[base] += 1
def demangle(n):
    try:
        d = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
        d = d.replace("(anonymous namespace)::", "")
        return re.split(r"\((?!anonymous)", d)[0][:110]
    except Exception:
        return n[:110]
print(f"# SASS summary of {os.path.relpath(so, ROOT)} (cuobjdump -sass, sm_100a), kernels with tcgen05 / TMA / TMEM instructions")
print(f"# totals: " + ", ".join(f"{k} x{v}" for k, v in sorted(total.items())))
for k, c in counts.items():
    if any(key.startswith(("UTC", "UTMA", "LDTM", "STTM")) for key in c):
        print(f"{demangle(k)}\n    " + ", ".join(f"{a} x{b}" for a, b in sorted(c.items())))
