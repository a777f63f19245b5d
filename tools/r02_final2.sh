#!/bin/bash
# Final evidence after the narrowed index upload: parity suite, the contract's bench line, launch list of the bench command.
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/r02j_pytest.txt 2>&1; tail -2 $OUT/r02j_pytest.txt
python bench.py > $OUT/r02_bench_c3_n1.json 2> $OUT/r02_bench_c3_n1.err; python -c "
import json; d=json.loads(open('$OUT/r02_bench_c3_n1.json').read().strip().splitlines()[-1]); e=d['e2e']
print('bench c3: ms/step %.3f value %.4g e2e %.3f ms (host %.3f ms, h2d %.1f of %.1f MB) frac %.4f cpu %.3g' % (d['ms_per_step'], d['value'], e['ms_per_step'], e['host_enqueue_ms_per_step'], e['h2d_bytes_per_step']/1e6, e['input_bytes_per_step']/1e6, d['roofline']['frac'], d['cpu_baseline']['value']))"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/r02_launches_bench_c3.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu > /dev/null 2>&1
grep -c Kernel $OUT/r02_launches_bench_c3.csv
