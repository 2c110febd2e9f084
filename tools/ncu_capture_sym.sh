set -e
export PROF_SINGLE=1 PROF_KERNELS=1,3
python tools/prof_driver.py gram 32768 > gpurun_out/prof_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gram_sym_kernel -o gpurun_out/r01_prof_g -f python tools/prof_driver.py gram 32768 > gpurun_out/ncu_g.log 2>&1
ncu -i gpurun_out/r01_prof_g.ncu-rep --page raw --csv > gpurun_out/r01_prof_g_raw.csv
ncu -i gpurun_out/r01_prof_g.ncu-rep --page source --csv > gpurun_out/r01_prof_g_source.csv
ncu -i gpurun_out/r01_prof_g.ncu-rep --page details > gpurun_out/r01_prof_g_details.txt
rm gpurun_out/r01_prof_g.ncu-rep
tail -3 gpurun_out/ncu_g.log
