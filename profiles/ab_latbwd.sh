run() { name=$1; shift; env "$@" timeout 60 python bench.py --no-cpu --no-also > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err; python -c "
import json,sys
d=json.load(open('gpurun_out/ab_$name.json')); print('$name', round(d['ms_per_step']*1e3,2), d['launches_per_step'])"; }
run f1_side64 VLA_FUSE_LATBWD=1
run f1_side48 VLA_FUSE_LATBWD=1 VLA_SIDE_CTAS=48
run f1_side32 VLA_FUSE_LATBWD=1 VLA_SIDE_CTAS=32
run f1_side0 VLA_FUSE_LATBWD=1 VLA_SIDE=0
run f0_side64 VLA_FUSE_LATBWD=0
run f0_side48 VLA_FUSE_LATBWD=0 VLA_SIDE_CTAS=48
run f0_side0 VLA_FUSE_LATBWD=0 VLA_SIDE=0
