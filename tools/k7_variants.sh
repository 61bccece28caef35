#!/bin/bash
# Builds kernel variants of the library in tree (stereo-vision_b200/variants/<tag>/, git-ignored) so that they travel to
# the GPU box.  Usage: tools/k7_variants.sh tag1:"-DFLAG=1 ..." tag2:"..."   then on the box:
#   for t in stereo-vision_b200/variants/*/; do ELAS_B200_LIB=$t/libelas_b200.so python tools/k7_group_time.py; done
set -e
cd "$(dirname "$0")/.."
for spec in "$@"; do
  tag=${spec%%:*}; flags=${spec#*:}
  d=$PWD/stereo-vision_b200/variants/$tag
  mkdir -p $d
  make -s -j8 -C stereo-vision_b200 OBJ=$d/build LIB=$d/libelas_b200.so EXTRA="$flags" 2>&1 | grep -i "error\|spill" || true
  echo "built $tag ($flags)"
done
