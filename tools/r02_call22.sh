set -e
export PROF_SINGLE=1
python tools/prof_driver.py gram 32768 > gpurun_out/r02_prof_gram_plain.log 2>&1
PROF_KERNELS=1 PROF_CROSS=1 python tools/prof_driver.py gram 32768 >> gpurun_out/r02_prof_gram_plain.log 2>&1
cat gpurun_out/r02_prof_gram_plain.log
PROF_KERNELS=1,2 ncu --set full --clock-control none --import-source on -k regex:'gram_' -o gpurun_out/r02_prof_gram_a -f python tools/prof_driver.py gram 32768 > gpurun_out/r02_prof_gram_ncu.log 2>&1
PROF_KERNELS=1 PROF_CROSS=1 ncu --set full --clock-control none --import-source on -k regex:'gram_fast' -o gpurun_out/r02_prof_gram_b -f python tools/prof_driver.py gram 32768 >> gpurun_out/r02_prof_gram_ncu.log 2>&1
for part in a b; do
  ncu -i gpurun_out/r02_prof_gram_${part}.ncu-rep --page raw --csv > gpurun_out/r02_prof_gram_${part}_raw.csv
  ncu -i gpurun_out/r02_prof_gram_${part}.ncu-rep --page details > gpurun_out/r02_prof_gram_${part}_details.txt
  rm gpurun_out/r02_prof_gram_${part}.ncu-rep
done
