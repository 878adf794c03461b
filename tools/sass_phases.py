"""Static SASS instruction count per `// [phase: …]` marker of one kernel (inline chains resolved to the frame in
fbank_fast.cuh / fbank_ws.cuh / melspec_fast.cuh that carries the marker).
With --ncu FILE (the `ncu --page source --csv --print-source sass` dump of the same kernel from the same build) the
per-instruction executed counts, stall samples and shared-memory wavefronts are joined in by instruction order.
usage: python tools/sass_phases.py lib.so 'mangled-name substring' [--ncu sass.csv] [--units N] [file.cuh ...]"""
import collections, os, re, subprocess, sys, tempfile
import csv
argv = sys.argv[1:]
ncu_csv, units = None, 1024 * 498
if "--ncu" in argv:
    i = argv.index("--ncu"); ncu_csv = argv[i + 1]; del argv[i:i + 2]
if "--units" in argv:
    i = argv.index("--units"); units = float(argv[i + 1]); del argv[i:i + 2]
lib, key = argv[0], argv[1]
files = argv[2:] or ["fbank_fast.cuh", "fbank_ws.cuh"]
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "dl_sound_classification_b200", "csrc")
marks = {}
for f in files:
    marks[f] = [(i, re.search(r"\[phase:\s*([^\]]+)\]", l).group(1).strip()) for i, l in enumerate(open(os.path.join(root, f)), 1) if "[phase:" in l]
def phase(f, line):
    n = f + ":preamble"
    for l, name in marks[f]:
        if l <= line: n = name
    return n
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cub = [os.path.join(tmp, x) for x in os.listdir(tmp) if x.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-gi", "-c", cub], capture_output=True, text=True).stdout
on, cur, inchain, chain = False, "?", False, []
seq = []
cnt, ops = collections.Counter(), collections.defaultdict(collections.Counter)
for line in txt.splitlines():
    if line.startswith("//---"):
        on = key in line
        continue
    if not on: continue
    if "//## File" in line:
        if not inchain: chain = []
        inchain = True
        chain.append(re.findall(r'"([^"]+)", line (\d+)', line)[0])
        continue
    if inchain:
        inchain = False
        cur = "other"
        for f, l in chain:                      # innermost first: the innermost frame in a marked file that is not a helper
            b = os.path.basename(f)
            if b in marks and phase(b, int(l)) != "-":
                cur = phase(b, int(l)); break
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
    if m:
        t = m.group(1).split()
        if t[0].startswith("@"): t = t[1:]
        cnt[cur] += 1
        seq.append((cur, t[0]))
        ops[cur][t[0].split(".")[0] + ("." + [x for x in t[0].split(".") if x in ("64", "128")][0] if any(x in ("64", "128") for x in t[0].split(".")) else "")] += 1
print(sum(cnt.values()), "instructions")
for k, v in cnt.most_common():
    print(f"{k:28s} {v:5d}  " + " ".join(f"{o}:{n}" for o, n in ops[k].most_common(12)))

if ncu_csv:
    rows = list(csv.reader(open(ncu_csv)))
    hdr = rows[1]
    ix = {n: hdr.index(n) for n in ("Instructions Executed", "# Samples", "L1 Wavefronts Shared", "stall_long_sb", "stall_short_sb", "stall_wait",
                                    "stall_math", "stall_not_selected", "stall_selected", "stall_mio", "stall_barrier", "stall_no_inst", "stall_dispatch", "stall_branch_resolving")}
    body = rows[2:]
    assert len(body) == len(seq), (len(body), len(seq))
    agg = collections.defaultdict(lambda: collections.Counter())
    for (ph, op), r in zip(seq, body):
        assert op.split(".")[0] in r[1], (op, r[1])
        for n, i in ix.items():
            agg[ph][n] += float(r[i] or 0)
    ti = sum(a["Instructions Executed"] for a in agg.values()); ts = sum(a["# Samples"] for a in agg.values())
    print(f"\ndynamic: {ti:.4e} warp-instr = {ti / units:.1f} per unit; {ts:.0f} stall samples")
    names = ["long_sb", "short_sb", "wait", "math", "not_selected", "selected", "mio", "barrier", "no_inst", "dispatch", "branch_resolving"]
    print(f"{'phase':24s} {'inst/unit':>9s} {'inst%':>6s} {'smp%':>6s} {'smemWF/u':>8s}  " + " ".join(f"{n[:7]:>7s}" for n in names))
    for ph, a in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"]):
        print(f"{ph:24s} {a['Instructions Executed'] / units:9.1f} {100 * a['Instructions Executed'] / ti:6.2f} {100 * a['# Samples'] / ts:6.2f} "
              f"{a['L1 Wavefronts Shared'] / units:8.1f}  " + " ".join(f"{100 * a['stall_' + n] / ts:7.2f}" for n in names))
