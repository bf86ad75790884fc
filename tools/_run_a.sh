mkdir -p gpurun_out
export KEXP_VARIANTS='{"base":[],"nofinal":["WL_EXP_NOFINAL"],"nofence":["WL_EXP_NOFENCE"],"nofinal_nofence":["WL_EXP_NOFINAL","WL_EXP_NOFENCE"],"bs128":["WL_QUAD_BS=128"],"bs128_nofinal_nofence":["WL_QUAD_BS=128","WL_EXP_NOFINAL","WL_EXP_NOFENCE"],"bs64":["WL_QUAD_BS=64"]}'
python tools/kexp.py run > gpurun_out/r02_kexp_a.jsonl 2> gpurun_out/r02_kexp_a.err
cat gpurun_out/r02_kexp_a.jsonl; tail -3 gpurun_out/r02_kexp_a.err
