#!/bin/bash
# Full SASS listings (gzipped; encodings stripped) and the resource usage of every kernel of the built library.
# Usage: bash tools/sass_listing.sh <round-prefix>      (after wgpu-sigops_b200/build.py; no GPU needed)
pre=${1:-r02}
mkdir -p profiles/sass
cd wgpu-sigops_b200/build || exit 1
for f in kern_k1 kern_r1 kern_ed kern_k1g kern_k1gc kern_r1g kern_edg kern_edgc; do
  cuobjdump -sass $f.o | sed 's/  *\/\* 0x[0-9a-f]* \*\///' | grep -v "^\s*$" | gzip -9 > ../../profiles/sass/${pre}_$f.sass.gz
done
(echo "# cuobjdump --dump-resource-usage, library sources $(cat ../libsigops.srchash)"
 for f in kern_k1 kern_r1 kern_ed kern_k1g kern_k1gc kern_r1g kern_edg kern_edgc kern_misc; do echo "== $f.o"; cuobjdump --dump-resource-usage $f.o | grep -E "Function|REG|STACK"; done) > ../../profiles/sass/${pre}_resource_usage.txt
