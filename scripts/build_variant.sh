#!/bin/bash
# Build the library as of git revision $1 into vipant_b200/_lib/libvipant_b200_$2.so (A/B measurements:
# select at run time with VIPANT_B200_LIB=<path>).
set -e
rev=$1; name=$2
tmp=$(mktemp -d)
git archive "$rev" vipant_b200 include | tar -x -C "$tmp"
(cd "$tmp" && python -c "
import sys; sys.path.insert(0, '.')
from vipant_b200 import build
print(build.build(force=True))")
cp "$tmp/vipant_b200/_lib/libvipant_b200.so" "vipant_b200/_lib/libvipant_b200_$name.so"
rm -rf "$tmp"
echo "vipant_b200/_lib/libvipant_b200_$name.so"
