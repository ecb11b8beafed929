"""Per-source-line view of an ncu capture: warp instructions executed and stall samples, aggregated by CUDA source line.
usage: python scripts/ncu_lines.py <report.ncu-rep> <cubin-from-cuobjdump -xelf> <mangled kernel name> [top N]
(ncu's --page source CSV carries SASS addresses only; the line table comes from `nvdisasm -g` on the same cubin.)"""
import csv, re, subprocess, sys, collections
rep, cubin, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
sass = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(sass) if l.startswith(".text." + kern + ":"))
line_of, cur = {}, ("?", 0)
for l in sass[start + 1:]:
    if l.startswith(".text.") or l.lstrip().startswith(".section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if "inlined at" not in l or True:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[h]
ia, ie, iss = H.index("Address"), H.index("Instructions Executed"), H.index("Warp Stall Sampling (All Samples)")
stall_cols = [i for i, c in enumerate(H) if c.startswith("stall_") and "Not Issued" not in c]
base = int(rows[h + 1][ia], 16)
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
tot_i = tot_s = 0
for r in rows[h + 1:]:
    try:
        off = int(r[ia], 16) - base; n = int(r[ie]); s = int(r[iss])
    except Exception:
        continue
    key = line_of.get(off, (("?", 0), ""))[0]
    a = agg[key]; a[0] += n; a[1] += s
    for i in stall_cols:
        try: a[2][H[i]] += int(r[i])
        except Exception: pass
    tot_i += n; tot_s += s
print(f"total warp instructions {tot_i:,}  stall samples {tot_s:,}")
src_cache = {}
def src(f, ln):
    import glob, os
    if f not in src_cache:
        c = glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "fpsample_b200", "csrc", f))
        src_cache[f] = open(c[0]).read().splitlines() if c else []
    L = src_cache[f]
    return L[ln - 1].strip()[:100] if 0 < ln <= len(L) else ""
for key, (n, s, st) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    why = ",".join(f"{k[6:]}:{v * 100 // max(s, 1)}" for k, v in st.most_common(3))
    print(f"{n / tot_i * 100:5.1f}% inst {s / max(tot_s, 1) * 100:5.1f}% samples [{why:38s}] {key[0]}:{key[1]:<4d} {src(*key)}")
