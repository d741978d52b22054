#!/usr/bin/env python
"""Per-kernel SASS hashes of a built library (addresses and encodings stripped): `python scripts/sass_hash.py xlumina_b200/libxlprop.so`.
Two builds are the same binary iff the lists are equal (profiles/sass_hashes_r01.txt = the library validated and timed in round 1)."""
import subprocess, hashlib, sys, re
out = subprocess.run(["cuobjdump","-sass",sys.argv[1]],capture_output=True,text=True).stdout
funcs={}; cur=None
for line in out.splitlines():
    m=re.match(r"\s*Function : (\S+)", line)
    if m: cur=m.group(1); funcs[cur]=hashlib.sha1(); continue
    if cur and "/*" in line:
        # strip addresses/encodings: keep the instruction text
        t=re.sub(r"/\*[0-9a-f]{4,}\*/","",line)
        t=re.sub(r"/\* 0x[0-9a-f]+ \*/","",t)
        funcs[cur].update(t.strip().encode())
for k in sorted(funcs): print(funcs[k].hexdigest()[:16], k)
