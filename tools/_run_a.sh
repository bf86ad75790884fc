mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02_gputest_k.log
cat gpurun_out/r02_gputest_k.log | tail -4
python tools/e2e_ab.py > gpurun_out/r02_e2e_ab2.json 2> gpurun_out/r02_e2e_ab2.err; cat gpurun_out/r02_e2e_ab2.json; tail -3 gpurun_out/r02_e2e_ab2.err
