#!/bin/bash
# round-2 evidence run on ONE B200: tests, smoke, bench (default command of the driver), reference arm, ncu launch list of the
# same bench command, full captures of the transition's kernels, of the chain-walk kernel (64 chains at 64^3) and of the VI kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt
python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; echo "pytest exit $?" >> gpurun_out/t_all.log; tail -2 gpurun_out/t_all.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.log 2>&1; echo "exit $?" >> gpurun_out/bench_default.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1
B="python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-aten-gpu-baseline --no-configs --e2e-steps 1"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:svf_step -s 96 -c 14 -o gpurun_out/prof_svf -f $B > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:langevin|smooth|warp_vox|box_march|gmm_|reg_hyper|sgd_update" -s 150 -c 17 \
    -o gpurun_out/prof_other -f $B > gpurun_out/ncu_full2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gmm_chain_walk -s 2 -c 1 -o gpurun_out/prof_walk64 -f \
    python bench.py --size 64 --chains 64 --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-aten-gpu-baseline --no-configs --e2e-steps 1 > gpurun_out/ncu_walk.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:vi_" -s 4 -c 4 -o gpurun_out/prof_vi -f \
    python -c "
import torch
from irsgmcmc_b200.vi import VIWarmStart
from irsgmcmc_b200.sampler import SGLDConfig
from irsgmcmc_b200.data_loader.synthetic import make_pair
f, m, vp = make_pair(128)
w = VIWarmStart(f, m, vp, SGLDConfig(), device='cuda:0')
w.sampler.init_gmm()
w.step(6, use_graph=False)
torch.cuda.synchronize()
" > gpurun_out/ncu_vi.log 2>&1
ls -la gpurun_out/*.ncu-rep
bash tools/gpu_stages.sh > gpurun_out/stages.txt 2>&1
