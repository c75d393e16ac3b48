# same-box A/B of libraries at a large (HBM-bound) and the headline batch size: bash tools/ab_large.sh "" _old _late ...
for n in 8388608 1048576; do
for rep in 1 2; do
for v in "$@"; do
TAROK_B200_LIB=$PWD/tarok_b200/libtarok_b200$v.so python tools/parts.py $n ${MODE:-17} | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['lib'].ljust(24), d['games'], 'step_random', d['step_random_us'], 'rollout', d['stepwise_rollout_us'], 'fused', d['fused_us'])"
done; done; done
