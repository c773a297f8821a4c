#!/bin/bash
# A/B of library builds on the same harness:  bash profiles/ab_libs.sh <tag> <lib name> [<lib name> ...]
# (profiles/ab_builds/libtb_<name>.so, built with tetris_b200._lib.build_variant; "current" = the in-tree library)
tag=$1; shift; out=gpurun_out; mkdir -p $out
for rep in 1 2; do
  for lib in "$@"; do
    if [ "$lib" = current ]; then so=""; else so=$PWD/profiles/ab_builds/libtb_$lib.so; fi
    echo -n "$lib " >> $out/ab_$tag.txt
    TB_SO_PATH=$so python profiles/ab_cfg.py ${AB_ARGS:---k1 0 --k3 0,7} >> $out/ab_$tag.txt 2>> $out/ab_$tag.err
  done
done
cat $out/ab_$tag.txt
