#!/bin/sh
# Type-checks integration/backends/cuda.c against the reference's own headers (compiler-internals.h, mathmap.h,
# drawable.h, userval.h ...) and include/mathmap_b200.h.  integration/shim/ stands in for what this container lacks:
# glib, gtk, libgimp, GSL and the two headers the reference GENERATES with clisp (compiler_types.h, opdefs.h).
# Usage: sh integration/check.sh [/path/to/reference]      (exit 0 = cuda.c matches every interface it touches)
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
REF=${1:-/root/reference}
TMP=$(mktemp -d)
trap 'rm -rf "$TMP"' EXIT
mkdir -p "$TMP/backends"
# cuda.c includes "../mathmap.h" like backends/cc.c does: give it the place in a tree it would have
cp "$HERE/backends/cuda.c" "$TMP/backends/cuda.c"
for f in "$REF"/*.h; do ln -s "$f" "$TMP/$(basename "$f")"; done
for d in builtins designer native-filters lispreader; do [ -d "$REF/$d" ] && ln -s "$REF/$d" "$TMP/$d"; done
gcc -std=gnu99 -Wall -Wno-unused-function -fsyntax-only -I "$HERE/shim" -I "$TMP" -I "$HERE/../include" "$TMP/backends/cuda.c"
echo "integration/backends/cuda.c: type-checks against $REF"
