#!/bin/bash
# ncu --set full capture of the shipping kernels, one launch each (run on the GPU box through gpurun):
#   bash tools/ncu_capture.sh <tag>     ->  gpurun_out/<tag>_{raw.csv,details.txt,source.csv}
# The driver is first run WITHOUT ncu (must exit 0), as the profiling recipe asks.
set -e
tag=${1:-r01_prof_h}
export PROF_SINGLE=1
python tools/prof_driver.py gram 32768 > gpurun_out/${tag}_plain.log 2>&1
PROF_KERNELS=1 PROF_CROSS=1 python tools/prof_driver.py gram 32768 >> gpurun_out/${tag}_plain.log 2>&1
python tools/prof_driver.py extra 32768 >> gpurun_out/${tag}_plain.log 2>&1
python tools/prof_driver.py matvec 32768 65536 >> gpurun_out/${tag}_plain.log 2>&1
cat gpurun_out/${tag}_plain.log
ncu --set full --clock-control none --import-source on -k regex:'gram_|matvec_mma' -o gpurun_out/${tag}_a -f python tools/prof_driver.py all 32768 65536 > gpurun_out/${tag}_ncu.log 2>&1
PROF_KERNELS=1 PROF_CROSS=1 ncu --set full --clock-control none --import-source on -k regex:'gram_fast' -o gpurun_out/${tag}_b -f python tools/prof_driver.py gram 32768 >> gpurun_out/${tag}_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'gram_extra|combine' -o gpurun_out/${tag}_c -f python tools/prof_driver.py extra 32768 >> gpurun_out/${tag}_ncu.log 2>&1
for part in a b c; do
  ncu -i gpurun_out/${tag}_${part}.ncu-rep --page raw --csv > gpurun_out/${tag}_${part}_raw.csv
  ncu -i gpurun_out/${tag}_${part}.ncu-rep --page details > gpurun_out/${tag}_${part}_details.txt
  ncu -i gpurun_out/${tag}_${part}.ncu-rep --page source --csv > gpurun_out/${tag}_${part}_source.csv
  rm gpurun_out/${tag}_${part}.ncu-rep
done
tail -2 gpurun_out/${tag}_ncu.log
