python tools/prof_matvec_k.py 65536 8 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'symk' -c 1 -o gpurun_out/r02_prof_symk -f python tools/prof_matvec_k.py 65536 8 > gpurun_out/r02_prof_symk_ncu.log 2>&1
ncu -i gpurun_out/r02_prof_symk.ncu-rep --page raw --csv > gpurun_out/r02_prof_symk_raw.csv
ncu -i gpurun_out/r02_prof_symk.ncu-rep --page source --csv > gpurun_out/r02_prof_symk_source.csv
rm gpurun_out/r02_prof_symk.ncu-rep
