#!/bin/bash
# Apply tools/ref_patch/main.patch to a scratch copy of the reference's main.c and build it against
# libecloop_b200.so with the reference's own flags (Makefile:4-8,15-16) plus the two INTEGRATION.md §2 adds.
# Output: tools/ref_patch/_build/ecloop_patched (git-ignored; travels to the GPU box). No reference source enters the repo.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
REF="${REF:-/root/reference}"
[ -f "$REF/main.c" ] || { echo "no $REF: keeping the prebuilt binary"; exit 0; }
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
cp "$REF/main.c" "$TMP/main.c"
patch -s -p1 -d "$TMP" < "$HERE/main.patch"
mkdir -p "$HERE/_build"
"${CC:-gcc}" -O3 -ffast-math -w -march=x86-64-v3 -msha -pthread -I"$REF" -I"$ROOT/include" "$TMP/main.c" -o "$HERE/_build/ecloop_patched" \
  -L"$ROOT/ecloop_b200" -lecloop_b200 -Wl,-rpath,'$ORIGIN/../../../ecloop_b200' -lpthread -lm
echo "built $HERE/_build/ecloop_patched"
