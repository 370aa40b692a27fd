# whole-step A/B of library builds under the same conditions: tools/_ab.sh "default old cos ..." [steps]
LIBS=${1:-default}; STEPS=${2:-50}
for rep in 1 2 3; do for lib in $LIBS; do
  if [ $lib = default ]; then unset MLMCPI_LIB; else export MLMCPI_LIB=$PWD/gpurun_var/libmlmcpi_$lib.so; fi
  python bench.py --no-extra --ess-draws 0 --steps $STEPS 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', round(d['ms_per_step'],3), d['clocks']['sm_mhz'], round(d['roofline']['avg_launch_ms']*1e3,1), d['clocks']['reasons'])"
done; done
