#!/bin/bash
# A/B of the variant builds (tarok_b200/build.py --variant): same measurements, one process per library.
for lib in "" f5 f6 s4 r5; do
  for n in 1048576 8388608; do
    if [ -z "$lib" ]; then python tools/parts.py $n; else TAROK_B200_LIB=$PWD/tarok_b200/libtarok_b200_$lib.so python tools/parts.py $n; fi
  done
done
