"""Opcode histogram of every kernel in libbbme.so (cuobjdump -sass): what proves sm_100a-native code (UTMALDG = TMA tile loads,
SYNCS = mbarriers, VABSDIFF4 = the SAD instruction, UCGABAR / CCTL = cluster barrier).  usage: sass_histogram.py [lib] > profiles/..."""
import collections, os, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "blockbasedmotionestimation_b200", "libbbme.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
kern, hist, total = None, {}, collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
        total[m.group(1)] += 1
KEY = ["VABSDIFF4", "UTMALDG", "SYNCS", "UCGABAR_ARV", "UCGABAR_WAIT", "CCTL", "MATCH", "REDUX", "IDP", "ATOMS", "ATOMG", "LDG", "LD", "LDS", "STG", "SHFL", "SHF", "BAR"]
print(f"# SASS opcode histogram of {os.path.basename(lib)} ({', '.join(arch)} only), `python scripts/sass_histogram.py`\n")
print("Whole library: " + ", ".join(f"{k} {total[k]}" for k in KEY if total[k]) + f"; {sum(total.values())} instructions in {len(hist)} kernels.\n")
print("| kernel | instructions | " + " | ".join(KEY[:11]) + " |")
print("|---|---|" + "---|" * 11)
groups = collections.OrderedDict()
for k, h in hist.items():
    base = re.sub(r"<.*", "", k.replace("void ", "").replace("bbme::", ""))
    g = groups.setdefault(base, [0, collections.Counter(), 0])
    g[0] += sum(h.values()); g[1].update(h); g[2] += 1
for base, (n, h, cnt) in sorted(groups.items(), key=lambda kv: -kv[1][0]):
    name = base + (f" (x{cnt} instantiations)" if cnt > 1 else "")
    print(f"| `{name}` | {n} | " + " | ".join(str(h[k]) if h[k] else "" for k in KEY[:11]) + " |")
