#!/bin/bash
# one GPU-box visit: tests, bench, ncu launch list, full captures of the SVF kernels (outputs under gpurun_out/)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt
python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; echo "pytest exit $?" >> gpurun_out/t_all.log
tail -3 gpurun_out/t_all.log
python bench.py > gpurun_out/bench_default.log 2>&1; echo "exit $?" >> gpurun_out/bench_default.log
tail -2 gpurun_out/bench_default.log | cut -c1-1500
if [ "$1" != "noprof" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:svf_step -s 96 -c 26 -o gpurun_out/prof_svf -f \
    python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
fi
if [ "$1" != "noprof" ]; then
ncu --set full --clock-control none --import-source on -k "regex:langevin|smooth|warp_vox|box_march|gmm_|reg_hyper|sgd_update" -s 150 -c 20 \
    -o gpurun_out/prof_other -f python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
fi
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-300
# secondary configuration (SVFFD_3D, spacing 4), the GPU-vs-GPU baseline, and the FFD kernels' metrics
python bench.py --cps 4 --steps 100 --warmup 5 --e2e-steps 20 --no-cpu-baseline > gpurun_out/bench_svffd4.log 2>&1; tail -1 gpurun_out/bench_svffd4.log | cut -c1-300
python bench.py --steps 50 --warmup 5 --no-cpu-baseline --aten-gpu-baseline > gpurun_out/bench_aten_gpu.log 2>&1; tail -1 gpurun_out/bench_aten_gpu.log | grep -o '"aten_gpu_baseline".*' | cut -c1-400
if [ "$1" != "noprof" ]; then
bash tools/gpu_ffd_ncu.sh
fi
