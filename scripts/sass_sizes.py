"""Development helper: SASS instruction count per (sub)function of an nvdisasm listing.
usage: nvdisasm -g -c file.cubin > listing.txt; python scripts/sass_sizes.py listing.txt"""
import re
import sys

cur = None
cnt = {}
for ln in open(sys.argv[1]):
    m = re.match(r"^(\$?_Z\S+):", ln)
    if m:
        cur = m.group(1)
        cnt.setdefault(cur, 0)
        continue
    if re.match(r"\s+/\*[0-9a-f]+\*/\s+\S", ln) and cur:
        cnt[cur] += 1
for k, v in cnt.items():
    nm = re.sub(r"_ZN\d+_GLOBAL__N__\w+?_cu_\w{8}", "", k)
    print(v, nm[:60], "...", nm[-70:] if len(nm) > 60 else "")
