for i in 1 2; do
LP_LIB_PATH=$PWD/lit_parrot_b200/liblitparrot_b200_old.bin python bench.py --steps 64 --warmup 8 --no-cpu-baseline --no-extras 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('OLD 3b tok/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3))"
python bench.py --steps 64 --warmup 8 --no-cpu-baseline --no-extras 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('NEW 3b tok/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3))"
done
