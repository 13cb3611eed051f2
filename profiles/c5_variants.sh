# C5 (1e10-point uniform sweep, logP not stored): the tensor-pipe pass-1 kernel with 3 and 4 CTAs per SM, and the scalar kernel.
for v in ${CTAS:-3 4 5}; do PEUQSE_SWEEP_DMMA_CTAS=$v python bench.py --steps 20 --no-c4 --no-mh --c5-points 2e9 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['secondary_c5']; print('dmma ctas/sm $v', c['pass1_evals_per_s'], c['value'], c['config']['map_index'])"; done
PEUQSE_SWEEP_NO_DMMA=1 python bench.py --steps 20 --no-c4 --no-mh --c5-points 2e9 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['secondary_c5']; print('scalar', c['pass1_evals_per_s'], c['value'], c['config']['map_index'])"
