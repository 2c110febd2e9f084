set -e
export PROF_SINGLE=1
python tools/prof_driver.py matvec 32768 131072 > gpurun_out/r01_prof_i_plain.log 2>&1
cat gpurun_out/r01_prof_i_plain.log
ncu --set full --clock-control none --import-source on -k regex:'matvec_mma_sym' -o gpurun_out/r01_prof_i -f python tools/prof_driver.py matvec 32768 131072 > gpurun_out/r01_prof_i_ncu.log 2>&1
ncu -i gpurun_out/r01_prof_i.ncu-rep --page raw --csv > gpurun_out/r01_prof_i_raw.csv
ncu -i gpurun_out/r01_prof_i.ncu-rep --page details > gpurun_out/r01_prof_i_details.txt
ncu -i gpurun_out/r01_prof_i.ncu-rep --page source --csv > gpurun_out/r01_prof_i_source.csv
rm gpurun_out/r01_prof_i.ncu-rep
tail -2 gpurun_out/r01_prof_i_ncu.log
