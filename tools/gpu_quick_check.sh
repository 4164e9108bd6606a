timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -1
for c in ur10 tiago; do
FG_RQ_CALIBRATE=2 timeout 300 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/ab.err | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$c', 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'step', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['ms_per_step'],3), d['parity_check'].get('idx_base_equal'), d['parity_check'].get('torque_norm_err'))"
grep "rq calibration" gpurun_out/ab.err | cut -c1-110 | tail -2
done
