"""Annotate the hot SASS lines of an ncu source-page CSV with CUDA source lines, using nvdisasm line info of the
in-tree library.   usage: python tools/sass_lines.py <ncu_sass.csv> <cubin-name e.g. gemm_tc> <kernel-substr> [topN]"""
import csv, os, re, subprocess, sys, tempfile
csvf, cub, ksub = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "gotennet_b200/lib/libgotennet_b200.so")], cwd=tmp, capture_output=True)
out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub + ".sm_100a.cubin")], capture_output=True, text=True).stdout
# split per function
funcs, cur, name = {}, None, None
line_no = None
for ln in out.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", ln) or re.match(r"//-+ \.text\.(\S+) -+", ln)
    if m:
        name = m.group(1); cur = funcs.setdefault(name, []); line_no = None; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        line_no = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if cur is not None and re.match(r"\s+/\*[0-9a-f]{4}\*/", ln):
        cur.append(line_no)
rows = list(csv.reader(open(csvf)))
kname = rows[0][1]
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr)]
cands = [k for k in funcs if ksub in k and len(funcs[k]) == len(body)]
if not cands:
    print("no function with", len(body), "instructions matching", ksub, [(k, len(v)) for k, v in funcs.items() if ksub in k]); sys.exit(1)
lines = funcs[cands[0]]
tot = sum(int(r[ci['# Samples']]) for r in body)
src = {}
per_line = {}
for i, r in enumerate(body):
    n = int(r[ci['# Samples']])
    per_line[lines[i]] = per_line.get(lines[i], 0) + n
print(kname[:90], "samples", tot)
for (ln, n) in sorted(per_line.items(), key=lambda kv: -kv[1])[:top]:
    text = ""
    if ln:
        try:
            text = open(os.path.join(root, "gotennet_b200/csrc", ln[0])).read().splitlines()[ln[1] - 1].strip()
        except Exception:
            pass
    print(f"{100.0 * n / tot:5.1f}%  {ln}  {text[:100]}")
