set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench0.json 2> gpurun_out/r02_bench0.err
python tools/host_bw_probe.py --gib 8 > gpurun_out/r02_hostbw_1gpu.json 2> gpurun_out/r02_hostbw_1gpu.err
python tools/host_bw_probe.py --gib 8 --numa >> gpurun_out/r02_hostbw_1gpu.json 2>> gpurun_out/r02_hostbw_1gpu.err
export PROF_SINGLE=1
python tools/prof_driver.py matvec 32768 262144 > gpurun_out/r02_prof_mv_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'matvec_mma_sym' -o gpurun_out/r02_prof_mv -f python tools/prof_driver.py matvec 32768 262144 > gpurun_out/r02_prof_mv_ncu.log 2>&1
ncu -i gpurun_out/r02_prof_mv.ncu-rep --page raw --csv > gpurun_out/r02_prof_mv_raw.csv
ncu -i gpurun_out/r02_prof_mv.ncu-rep --page source --csv > gpurun_out/r02_prof_mv_source.csv
rm -f gpurun_out/r02_prof_mv.ncu-rep
cat gpurun_out/r02_prof_mv_plain.log gpurun_out/r02_hostbw_1gpu.json
