mkdir -p gpurun_out
for c in 6 4 8 12; do WL_SCAN_CTAS_PER_SM=$c timeout 120 python tools/scan_ab.py 2>&1 | tail -1; done | tee gpurun_out/r02_scan_ab.jsonl
ENVS=65536 timeout 120 python tools/scan_ab.py 2>&1 | tail -1 | tee -a gpurun_out/r02_scan_ab.jsonl
timeout 200 ncu --set full --import-source on --clock-control none -k regex:"wl_scan" -s 8 -c 1 -o gpurun_out/r02_ncu_scan_pipe python tools/sweep.py --task elevation --sizes 4096 --steps 6 --warm 3 > gpurun_out/r02_ncu_m1.log 2>&1; tail -2 gpurun_out/r02_ncu_m1.log
