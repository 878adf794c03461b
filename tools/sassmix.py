"""Static SASS opcode mix of one kernel of a .so: python tools/sassmix.py lib.so 'substring of mangled name' [top]"""
import subprocess, sys, collections, re
lib, key = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 24
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, cnt = None, collections.Counter()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur and key in cur:
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m:
            t = m.group(1).split()
            if t[0].startswith("@"):
                t = t[1:]
            op = t[0].split(".")[0]
            if op in ("LDS", "STS", "LDG", "STG"):
                w = [x for x in t[0].split(".") if x in ("64", "128")]
                op += "." + w[0] if w else ""
            cnt[op] += 1
print(sum(cnt.values()), "instructions;", " ".join(f"{k}:{v}" for k, v in cnt.most_common(top)))
