#!/bin/bash
# Round-2 evidence job (one gpurun call, 1 GPU): GPU test suite, the default bench line, the ncu launch list of the same
# command, and one `--set full` capture per kernel family.  Outputs under gpurun_out/r2z_*.
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $O/r2z_clocks.csv 2>/dev/null &
SMI=$!
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2z_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2z_pytest.log
timeout 600 python bench.py > $O/r2z_bench_n1.json 2> $O/r2z_bench_n1.err; echo "bench rc=$?" >> $O/r2z_bench_n1.err
kill $SMI
# launch list of the bench command (short form: the timed region is the same; parity / cfg5 / cpu legs skipped)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2z_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cfg5 --no-cpu-baseline > $O/r2z_ncu_bench.log 2>&1
# full captures
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sage_layer_umma -s 3 -c 3 -f -o $O/r2z_sage \
    python scripts/profile_sage.py auto 2 > $O/r2z_ncu_sage.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"intern_hash|intern_assign|rs_scatter|rs_hist|node_acc_events" -s 20 -c 8 -f -o $O/r2z_stream \
    python scripts/profile_stream.py > $O/r2z_ncu_stream.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mcts_search -s 1 -c 1 -f -o $O/r2z_mcts \
    python scripts/mcts_timing.py > $O/r2z_ncu_mcts.log 2>&1
ls -la $O/r2z_*
