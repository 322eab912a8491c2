#!/usr/bin/env python
"""md5 of the SASS of every kernel in a shared library -- to show that a source change left the other instantiations
byte-identical.  usage: python tools/sass_hashes.py ppr_diffphys_b200/libppr_b200.so > hashes.txt"""
import hashlib
import re
import subprocess
import sys

out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
name, body, res = None, [], {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        if name:
            res[name] = hashlib.md5("\n".join(body).encode()).hexdigest()
        name, body = m.group(1), []
    elif name and re.match(r"\s*/\*[0-9a-f]{4}\*/", line):
        body.append(re.sub(r"/\* 0x[0-9a-f]+ \*/", "", line).strip())
if name:
    res[name] = hashlib.md5("\n".join(body).encode()).hexdigest()
for k in sorted(res):
    print(res[k], k)
