#!/bin/sh
# copies the fixtures listed in README.md from the reference tree
set -e
REF=${REF:-/root/reference}
HERE=$(dirname "$0")
cp "$REF/src/V1/feat/features2.ft" "$REF/src/V1/feat/features2.txt" "$HERE/"
mkdir -p "$HERE/images_provided" "$HERE/images_traffic" "$HERE/images_laptops"
cp "$REF"/data/images_provided/img*.pgm "$HERE/images_provided/"
for i in 1 2 3 4; do
  cp "$REF/data/images_traffic/img$i.pgm" "$HERE/images_traffic/"
  cp "$REF/data/images_laptops/img$i.pgm" "$HERE/images_laptops/"
done
