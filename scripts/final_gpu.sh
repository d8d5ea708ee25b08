#!/bin/bash
# One gpurun call: full GPU test suite, the bench line, the ncu launch list of the same bench command, and one
# `ncu --set full` capture of the decode / encode kernels (developer tool; outputs under gpurun_out/).
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/final_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/final_tests.log
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 2 --warmup 1 > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench.log 2>&1
timeout 200 python scripts/dev_bench_decode.py text 8192 1 > gpurun_out/plain_dec.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dec_ -s 28 -c 7 -o gpurun_out/prof_dec_final3 -f python scripts/dev_bench_decode.py text 8192 1 > gpurun_out/ncu_dec.log 2>&1
timeout 200 python scripts/dev_bench_encode.py silesia 8192 1 > gpurun_out/plain_enc.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:enc_ -s 4 -c 4 -o gpurun_out/prof_enc_final3 -f python scripts/dev_bench_encode.py silesia 8192 1 > gpurun_out/ncu_enc.log 2>&1
tail -3 gpurun_out/final_tests.log; cat gpurun_out/bench_final.json | head -c 3000
