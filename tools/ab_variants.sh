#!/bin/bash
# A/B of variant libraries (python -m tarok_b200.build --variant NAME -D...): tools/parts.py per variant, interleaved twice.
# usage: bash tools/ab_variants.sh "" _e6 _ei ...   (suffixes of tarok_b200/libtarok_b200<suffix>.so); GAMES=8388608 for the large state
for rep in 1 2; do
  for v in "$@"; do
    TAROK_B200_LIB=$PWD/tarok_b200/libtarok_b200$v.so python tools/parts.py ${GAMES:-1048576} | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['lib'].ljust(24), d['games'], 'step_random', d['step_random_us'], 'forced', d['step_forced_us'], 'rollout', d['stepwise_rollout_us'], 'fused', d['fused_us'])"
  done
done
