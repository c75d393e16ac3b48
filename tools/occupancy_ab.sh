#!/bin/bash
# A/B of the variant builds (tarok_b200/build.py --variant): same measurements, one process per library.
for n in 1048576 8388608; do
  python tools/parts.py $n
  for lib in r4 r6 c128; do
    TAROK_B200_LIB=$PWD/tarok_b200/libtarok_b200_$lib.so python tools/parts.py $n
  done
done
