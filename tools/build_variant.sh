#!/bin/bash
# build a variant of libmrs_b200.so for same-box A/B timing:  tools/build_variant.sh NAME "-DFLAG ..."  ->  _ab/NAME.so
# (run it with MRS_LIB=$PWD/_ab/NAME.so; _ab/ is not tracked but travels with gpurun)
set -e
root=$(cd "$(dirname "$0")/.." && pwd)
name="$1"; flags="$2"
tmp=$(mktemp -d)
mkdir -p "$tmp/pkg/csrc" "$tmp/include" "$root/_ab"
cp "$root"/movie-recommender-system_b200/csrc/*.cu "$root"/movie-recommender-system_b200/csrc/*.cuh "$root"/movie-recommender-system_b200/csrc/Makefile "$tmp/pkg/csrc/"
cp "$root"/include/*.h "$tmp/include/"
make -C "$tmp/pkg/csrc" -j8 EXTRA="$flags" OUT="$root/_ab/$name.so" > /dev/null
rm -rf "$tmp"
echo "built _ab/$name.so"
