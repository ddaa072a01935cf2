#!/bin/bash
# Round-end measurements on one GPU: driver-style bench line, launch list and full ncu capture of the two dominant kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1_s20.json 2> gpurun_out/bench_n1_s20.err; echo "bench rc=$?"
python tools/profile_step.py --steps 3 > /dev/null 2>&1; echo "profile_step rc=$?"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct \
    --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py --steps 3 > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"pair_gemm_kernel|bwd_fused_kernel" --launch-skip 4 -c 2 -f -o gpurun_out/fwd_bwd_full \
    python tools/profile_step.py --steps 3 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/fwd_bwd_full.ncu-rep --page raw --csv > gpurun_out/fwd_bwd_full_raw.csv 2>/dev/null; echo "raw rc=$?"
ls -la gpurun_out/fwd_bwd_full.ncu-rep
