# one gpurun call: GPU tests, the default bench, ncu captures of K1 on both distributions, headline launch list
mkdir -p gpurun_out
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/s2d_gputest.log 2>&1; echo rc=$? >> gpurun_out/s2d_gputest.log; tail -3 gpurun_out/s2d_gputest.log
python bench.py > gpurun_out/s2d_bench.json 2> gpurun_out/s2d_bench.err; echo bench rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_stream -s 1 -c 2 -f -o gpurun_out/s2d_k1 python profiles/k1_profile.py > gpurun_out/s2d_ncu1.log 2>&1; echo ncu1 rc=$?
timeout 600 ncu --set full --clock-control none -k regex:step_stream -s 513 -c 2 -f -o gpurun_out/s2d_k1_steady python profiles/k1_profile.py steady > gpurun_out/s2d_ncu2.log 2>&1; echo ncu2 rc=$?
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s2d_launches_headline.csv python bench.py --headline-only --steps 20 --warmup 5 > gpurun_out/s2d_ncu3.log 2>&1; echo ncu3 rc=$?
python profiles/update_trace.py conv > gpurun_out/s2d_conv_trace.txt 2>&1; echo trace conv rc=$?
python profiles/update_trace.py dense > gpurun_out/s2d_dense_trace.txt 2>&1; echo trace dense rc=$?
