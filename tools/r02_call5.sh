set -x
python -m pytest tests/test_gpu_gram_wide.py -x -q 2>&1 | tail -3
python tools/gram_wide_bench.py 16384 32 100 256 2>&1 | grep "ident=on"
python tools/prof_gram_wide.py 16384 100 && PROF_CROSS=1 python tools/prof_gram_wide.py 16384 100 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'gram_mma_kernel' -c 2 -o gpurun_out/r02_prof_gw_sym -f python tools/prof_gram_wide.py 16384 100 > gpurun_out/r02_prof_gw_ncu.log 2>&1
PROF_CROSS=1 ncu --set full --clock-control none --import-source on -k regex:'gram_mma_kernel' -c 2 -o gpurun_out/r02_prof_gw_cross -f python tools/prof_gram_wide.py 16384 100 >> gpurun_out/r02_prof_gw_ncu.log 2>&1
for part in sym cross; do
  ncu -i gpurun_out/r02_prof_gw_${part}.ncu-rep --page raw --csv > gpurun_out/r02_prof_gw_${part}_raw.csv
  ncu -i gpurun_out/r02_prof_gw_${part}.ncu-rep --page details > gpurun_out/r02_prof_gw_${part}_details.txt
  ncu -i gpurun_out/r02_prof_gw_${part}.ncu-rep --page source --csv > gpurun_out/r02_prof_gw_${part}_source.csv
  rm gpurun_out/r02_prof_gw_${part}.ncu-rep
done
tail -3 gpurun_out/r02_prof_gw_ncu.log
