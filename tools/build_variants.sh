#!/bin/bash
# Builds tuning variants of libbm25x.so side by side (vectorchord-bm25_b200/variants/libbm25x_<name>.so) for
# tools/time_variants.py.  Usage: tools/build_variants.sh name1="-DFOO=1 -DBAR=2" name2="..."   (run from the repo root)
set -e
cd "$(dirname "$0")/../vectorchord-bm25_b200/csrc"
mkdir -p ../variants
for spec in "$@"; do
  name="${spec%%=*}"
  flags="${spec#*=}"
  [ "$flags" = "$spec" ] && flags=""
  ( make -s OUT=../variants/libbm25x_${name}.so OBJDIR=build_${name} EXTRA="$flags" > /dev/null 2>&1 \
      && echo "built $name [$flags]" || echo "FAILED $name [$flags]" ) &
  while [ "$(jobs -r | wc -l)" -ge 2 ]; do sleep 1; done
done
wait
