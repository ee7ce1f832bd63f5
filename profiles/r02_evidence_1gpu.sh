#!/bin/bash
# Round-2 evidence on ONE B200 (run from the repo root under gpurun): GPU tests, smoke, default bench line, reference arm, ncu
# launch list of the quick bench command, one ncu --set full line per kernel of a batch-64 ESPNet-C segment() call.
# Reports are summarised on the box (profiles/ncu_table.py); only text comes back.
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r02_gpu_tests.log 2>&1; tail -2 gpurun_out/r02_gpu_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/bench_default.err; echo bench rc=$?
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/bench_ref.err; echo ref rc=$?
timeout 300 python bench.py --quick --no-cpu --steps 2 --warmup 1 > /tmp/q.json 2>/dev/null && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --quick --no-cpu --steps 2 --warmup 1 > gpurun_out/ncu_launches.log 2>&1; echo launches rc=$?
timeout 200 python profiles/run_forward.py fp32 64 0 encoder > gpurun_out/enc_plain.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none -o /tmp/enc python profiles/run_forward.py fp32 64 0 encoder > gpurun_out/ncu_enc.log 2>&1
python profiles/ncu_table.py /tmp/enc.ncu-rep > gpurun_out/r02_ncu_espnet_c_fp32_b64.txt 2> gpurun_out/ncu_table.err; echo table rc=$?
wc -l gpurun_out/r02_launches.csv gpurun_out/r02_ncu_espnet_c_fp32_b64.txt
cut -c1-300 gpurun_out/r02_bench_default.json
