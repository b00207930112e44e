#!/bin/bash
# Device-code identity check without a GPU: builds the library at a given commit in a scratch worktree and compares its SASS
# (cuobjdump -sass, path identifiers and anonymous-namespace hashes normalised) with the library built from the working tree.
# Usage: scripts/sass_diff.sh <commit>      -> prints "SASS identical" or the first differing lines
set -e
ref=${1:?commit}
root=$(git rev-parse --show-toplevel)
wt=$(mktemp -d /tmp/sassdiff.XXXXXX)
git worktree add -q "$wt" "$ref"
( cd "$wt" && python qgdsolver_b200/build.py --force > /dev/null 2>&1 )
( cd "$root" && python qgdsolver_b200/build.py > /dev/null 2>&1 )
norm() { cuobjdump -sass "$1" | grep -v '^\s*$' | grep -v '^identifier = ' | sed 's/_GLOBAL__N__[0-9a-f_]*qgd_[a-z]*_cu_[0-9a-f]*/ANON/g'; }
norm "$wt/qgdsolver_b200/libqgd_b200.so" > "$wt.a"
norm "$root/qgdsolver_b200/libqgd_b200.so" > "$wt.b"
if diff -q "$wt.a" "$wt.b" > /dev/null; then echo "SASS identical to $ref ($(wc -l < "$wt.a") lines)"; else diff "$wt.a" "$wt.b" | head -40; fi
git worktree remove --force "$wt"; rm -f "$wt.a" "$wt.b"
