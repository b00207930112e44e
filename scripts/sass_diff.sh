#!/bin/bash
# Device-code identity check without a GPU: builds the library at a given commit in a scratch worktree and compares the SASS of
# every kernel of that build (cuobjdump -sass, anonymous-namespace hashes normalised) with the same kernel in the library built
# from the working tree.  New kernels in the working tree are listed, not compared.
# Usage: scripts/sass_diff.sh <commit>
set -e
ref=${1:?commit}
root=$(git rev-parse --show-toplevel)
wt=$(mktemp -d /tmp/sassdiff.XXXXXX)
git worktree add -q "$wt" "$ref"
( cd "$wt" && python qgdsolver_b200/build.py --force > /dev/null 2>&1 )
( cd "$root" && python qgdsolver_b200/build.py > /dev/null 2>&1 )
python3 - "$wt/qgdsolver_b200/libqgd_b200.so" "$root/qgdsolver_b200/libqgd_b200.so" "$ref" <<'PY'
import re, subprocess, sys
def funcs(lib):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    txt = re.sub(r"_GLOBAL__N__[0-9a-f_]*qgd_[a-z0-9]*_cu_[0-9a-f]*", "ANON", txt)
    out, name, buf = {}, None, []
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name: out[name] = "\n".join(buf)
            name, buf = m.group(1), []
        elif name and line.strip() and not line.startswith("identifier") and not line.startswith("Fatbin") and not line.startswith("="):
            buf.append(line)
    if name: out[name] = "\n".join(buf)
    return out
a, b = funcs(sys.argv[1]), funcs(sys.argv[2])
changed = [k for k in a if k in b and a[k] != b[k]]
gone = [k for k in a if k not in b]
new = [k for k in b if k not in a]
print(f"{len(a)} kernels at {sys.argv[3]}: {len(a) - len(changed) - len(gone)} identical, {len(changed)} changed, {len(gone)} removed; {len(new)} new in the working tree")
for k in changed: print("  CHANGED", k[:140])
for k in gone: print("  REMOVED", k[:140])
for k in new: print("  new    ", k[:140])
PY
git worktree remove --force "$wt"
