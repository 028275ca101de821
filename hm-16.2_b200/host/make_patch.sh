#!/bin/sh
# Regenerates hm_gpume.patch from the patched working copy under build/src (see Makefile) against the
# unmodified reference.  Run after editing build/src/source/...; the patch is the only tracked form of the
# HM-side hook lines (the reference sources themselves are never copied into the repository).
set -e
REF=${REF:-/root/reference}
cd "$(dirname "$0")"
T=$(mktemp -d)
mkdir -p $T/a/source/Lib $T/a/source/App $T/b/source/Lib $T/b/source/App
ln -s $REF/source/Lib/TLibEncoder $T/a/source/Lib/TLibEncoder
ln -s $REF/source/App/TAppEncoder $T/a/source/App/TAppEncoder
ln -s $PWD/build/src/source/Lib/TLibEncoder $T/b/source/Lib/TLibEncoder
ln -s $PWD/build/src/source/App/TAppEncoder $T/b/source/App/TAppEncoder
(cd $T && diff -ru a/source b/source | sed -E 's/^(---|\+\+\+) ([^\t]+)\t.*/\1 \2/') > hm_gpume.patch || true
rm -rf $T
touch build/src/.patched
grep -c '^@@' hm_gpume.patch
