#!/bin/bash
# tools/capture_profiles.sh — run on a B200 box (gpurun): GPU test suite, the bench line, the ncu launch list of the bench command and
# one `ncu --set full` capture per dominant kernel; everything lands in gpurun_out/ (copy what is to be judged into profiles/).
TAG=${1:-r2}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s 2>&1 | grep -vE "^\s*$" | tail -60 > gpurun_out/${TAG}_gputests.txt
python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
python bench.py --impl reference > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-configs --no-cpu-baseline --no-pipelined > gpurun_out/${TAG}_bench_under_ncu.json 2> /dev/null
for A in 1 0; do
  NAME=$([ $A = 1 ] && echo fast || echo exact)
  ncu --set full --clock-control none --import-source on -k regex:pmb_kernel -s 2 -c 1 -o gpurun_out/${TAG}_sqp_solve_${NAME} \
      python tools/one_solve.py $A > /dev/null 2>&1
  ncu -i gpurun_out/${TAG}_sqp_solve_${NAME}.ncu-rep --page raw --csv > gpurun_out/${TAG}_sqp_solve_${NAME}_rawpage.csv 2> /dev/null
done
ncu --set full --clock-control none -k regex:pmb_kernel -s 5 -c 1 -o gpurun_out/${TAG}_kkt_assemble ./tools/ubench/kkt > /dev/null 2>&1
ncu -i gpurun_out/${TAG}_kkt_assemble.ncu-rep --page raw --csv > gpurun_out/${TAG}_kkt_assemble_rawpage.csv 2> /dev/null
./tools/ubench/kkt > gpurun_out/${TAG}_kkt_ubench.txt 2>&1
./tools/ubench/dmma > gpurun_out/${TAG}_dmma_ubench.txt 2>&1
./tools/ubench/fast_phases > gpurun_out/${TAG}_fast_phases.txt 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,driver_version --format=csv > gpurun_out/${TAG}_smi.txt
tail -3 gpurun_out/${TAG}_gputests.txt
