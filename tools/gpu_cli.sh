df -h /tmp | tail -1; nproc; free -g | head -2
MP_BENCH_CLI_PAIRS=6291456 timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_cli.json 2> gpurun_out/bench_cli.err; grep -E "timing|Elapsed time on host" gpurun_out/bench_cli.err | tail -16
python - <<PY
import json
d=json.load(open('gpurun_out/bench_cli.json'))
print(d['value'], d['e2e']['value'], d.get('e2e_cli'), d.get('parity_at_scale'))
PY
