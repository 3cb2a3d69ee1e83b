"""Compare the SASS of two objects kernel by kernel (instruction text, encodings ignored):
the way to show that a source change is a refactoring as far as the GPU is concerned.

    python scripts/sass_diff.py old.o new.o
"""
import re
import subprocess
import sys


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    ks, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            ks[cur] = []
        elif cur and re.match(r"^\s+/\*[0-9a-f]{4,5}\*/", line):
            ks[cur].append(re.sub(r"/\* 0x[0-9a-f]+ \*/", "", line).strip())
    return ks


old, new = kernels(sys.argv[1]), kernels(sys.argv[2])
rc = 0
for k in sorted(set(old) | set(new)):
    if k not in new:
        print("removed   ", k)
    elif k not in old:
        print("added     ", k, len(new[k]), "instructions")
    elif old[k] == new[k]:
        print("identical ", k, len(new[k]), "instructions")
    else:
        rc = 1
        print("DIFFERENT ", k, len(old[k]), "->", len(new[k]), "instructions")
sys.exit(rc)
