# C3 Metropolis-Hastings kernel on the bench configuration (4096 chains x 20000 steps), one launcher shape per run
# (PEUQSE_MH_SHAPE, see launch_mh in csrc/mh.cu: 1 = general kernel k_mh_warp; k_mh_tpd with 4 = up to four chains per warp,
# 8 warps; 5 = up to two, 16 warps; 7 = one chain per warp, 12 warps; 8 = up to three, 12 warps; 0 / unset = the launcher's choice).
# usage: [SHAPES="4 7"] bash profiles/c3_variants.sh
for v in ${SHAPES:-1 4 5 7 8 0}; do PEUQSE_MH_SHAPE=$v python bench.py --steps 20 --no-c4 --no-c5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('shape $v', d['secondary']['value'], d['secondary']['config']['acceptance_rate'], d['secondary']['e2e']['value'])"; done
