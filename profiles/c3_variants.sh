# C3 Metropolis-Hastings kernel, the three compiled CTA shapes (PEUQSE_MH_SHAPE=1: 3 CTAs x 4 warps, 2: 2 x 7, 3: 2 x 8; unset: the
# launcher's wave-aware choice) on the bench configuration (4096 chains x 20000 steps).  usage: bash profiles/c3_variants.sh
for v in ${SHAPES:-1 2 3 0}; do PEUQSE_MH_SHAPE=$v python bench.py --steps 20 --no-c4 --no-c5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('shape $v', d['secondary']['value'], d['secondary']['config']['acceptance_rate'], d['secondary']['e2e']['value'])"; done
