#!/bin/bash
# Developer script: is the device code (SASS of every kernel) of the working tree's libpb2.so the same as that of <commit>?
# Builds <commit> in a temporary worktree and compares `cuobjdump -sass` (the translation unit's hash in the names of
# anonymous-namespace kernels is masked).  Used after host-only changes to pb2_cuda.cu once the GPU budget of a round is spent:
# the kernels that were verified on the GPU are then provably the ones that ship.
# usage: tools/same_device_code.sh <commit>
set -e
cd "$(dirname "$0")/.."
wt=$(mktemp -d /tmp/pb2_wt.XXXXXX)
git worktree add -q "$wt" "$1"
make -C "$wt/pbrt_v3_b200/csrc" -j8 > "$wt/build.log" 2>&1
a=$(cuobjdump -sass "$wt/pbrt_v3_b200/lib/libpb2.so" | sed 's/_GLOBAL__N__[0-9a-f]*_/_GLOBAL__N__X_/' | md5sum)
b=$(cuobjdump -sass pbrt_v3_b200/lib/libpb2.so | sed 's/_GLOBAL__N__[0-9a-f]*_/_GLOBAL__N__X_/' | md5sum)
git worktree remove --force "$wt"
git worktree prune
echo "$1: $a"; echo "working tree: $b"
[ "$a" = "$b" ] && echo "device code identical" || { echo "device code DIFFERS"; exit 1; }
