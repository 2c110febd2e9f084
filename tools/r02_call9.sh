python -m pytest tests/test_gpu_fused.py -x -q 2>&1 | tail -5
python tools/prof_fused.py || exit 1
ncu --set full --clock-control none --import-source on -k regex:'gram_sym_kernel' -c 1 -o gpurun_out/r02_prof_fused -f python tools/prof_fused.py > gpurun_out/r02_prof_fused_ncu.log 2>&1
ncu -i gpurun_out/r02_prof_fused.ncu-rep --page raw --csv > gpurun_out/r02_prof_fused_raw.csv
ncu -i gpurun_out/r02_prof_fused.ncu-rep --page source --csv > gpurun_out/r02_prof_fused_source.csv
rm gpurun_out/r02_prof_fused.ncu-rep
