#!/bin/bash
# Round-2 evidence on ONE B200 (run from the repo root under gpurun): default bench line, reference arm, ncu launch list of the
# quick bench command, ncu --set full table of the kernels either side of the forward.  Reports are summarised on the box.
timeout 600 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/bench_default.err; echo bench rc=$?
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/bench_ref.err; echo ref rc=$?
timeout 300 python bench.py --quick --no-cpu --steps 2 --warmup 1 > /tmp/q.json 2>/dev/null && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --quick --no-cpu --steps 2 --warmup 1 > gpurun_out/ncu_launches.log 2>&1; echo launches rc=$?
timeout 200 python profiles/run_forward.py fp32 8 1 aux > gpurun_out/aux_plain.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none -o /tmp/aux python profiles/run_forward.py fp32 8 1 aux > gpurun_out/ncu_aux.log 2>&1
python profiles/ncu_table.py /tmp/aux.ncu-rep > gpurun_out/r02_ncu_aux.txt 2> gpurun_out/ncu_table.err; echo aux rc=$?
wc -l gpurun_out/r02_launches.csv gpurun_out/r02_ncu_aux.txt
cut -c1-400 gpurun_out/r02_bench_default.json
