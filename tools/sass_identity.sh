#!/bin/bash
# Compare the device code (SASS instruction lines) of every kernel object with the one built from another commit.
# Used to show that a refactor / an added opt-in kernel leaves the validated kernels byte-identical:
#   tools/sass_identity.sh db2323b        (db2323b = the state of the last full GPU validation, run 24)
set -e
ref=${1:?commit}
root=$(git rev-parse --show-toplevel)
wt=$root/gpurun_out/wt_sass_$ref
rm -rf "$wt"; git worktree add -f -q "$wt" "$ref"
(cd "$wt" && python - <<'PY'
import importlib.util, os
spec = importlib.util.spec_from_file_location("b", os.path.join(os.getcwd(), "nerf_slam_b200", "build.py"))
b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b); b.build(force=True)
PY
)
python -m nerf_slam_b200.build > /dev/null
sig() { cuobjdump -sass "$1" | grep -E "^\s+/\*[0-9a-f]{4}\*/" | md5sum | cut -c1-12; }
for f in "$wt"/nerf_slam_b200/_build/*.o; do
  n=$(basename "$f"); a=$(sig "$f"); b=$(sig "$root/nerf_slam_b200/_build/$n")
  echo "$n $a $b $([ "$a" = "$b" ] && echo same || echo DIFFERENT)"
done
for f in "$root"/nerf_slam_b200/_build/*.o; do
  n=$(basename "$f"); [ -e "$wt/nerf_slam_b200/_build/$n" ] || echo "$n (new object, not in $ref)"
done
git worktree remove --force "$wt"; git worktree prune
