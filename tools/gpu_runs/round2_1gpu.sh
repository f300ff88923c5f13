#!/bin/bash
# 1-GPU follow-ups for the next round: full validation of the final round-1 tree, then the training-path experiments.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- bash tools/gpu_runs/round2_1gpu.sh
mkdir -p gpurun_out
S=gpurun_out/summary_r2_1gpu.txt; : > $S
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r2_gpu_tests.log 2>&1; echo "pytest -m gpu exit=$?" | tee -a $S
tail -3 gpurun_out/r2_gpu_tests.log | cut -c1-300 | tee -a $S
PETALS_B200_RUN_UNVALIDATED=1 timeout 300 python -m pytest tests/test_l2_prefetch_gpu.py -q -x > gpurun_out/r2_l2pf_tests.log 2>&1; echo "l2 prefetch tests exit=$?" | tee -a $S
tail -3 gpurun_out/r2_l2pf_tests.log | cut -c1-300 | tee -a $S
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke exit=$?" | tee -a $S
# LoRA on the kernels (merged-weight views): the LoRA engine tests with the opt-in switched on
PETALS_B200_RUN_UNVALIDATED=1 timeout 400 python -m pytest tests/test_lora_engine_gpu.py -q -x > gpurun_out/r2_lora_engine.log 2>&1; echo "lora-on-engine tests exit=$?" | tee -a $S
tail -3 gpurun_out/r2_lora_engine.log | cut -c1-300 | tee -a $S
cd benchmarks
for sdpa in 0 1; do
  PETALS_B200_SDPA_BACKWARD=$sdpa timeout 400 python benchmark_training.py --model llama-3-8b --n_steps 8 --warmup_steps 3 --batch_size 8 --seq_len 128 \
      > ../gpurun_out/r2_train_sdpa$sdpa.log 2>&1
  echo "training SDPA_BACKWARD=$sdpa exit=$?: $(grep 'Final result' ../gpurun_out/r2_train_sdpa$sdpa.log | tail -1)" | tee -a ../$S
done
cd ..
# where does a backward step spend its time? (torch profiler table of one rpc_backward on one block span)
timeout 400 python - > gpurun_out/r2_backward_profile.log 2>&1 <<'PY'
import torch, sys
sys.path.insert(0, ".")
from torch.profiler import profile, ProfilerActivity
from petals_b200.parallel.swarm import Swarm
from petals_b200.utils.random_model import launch_random_stage, write_config_only
path = write_config_only("llama-3-8b", {"torch_dtype": "bfloat16"})
swarm = Swarm("r2-prof")
stage = launch_random_stage(path, range(4), swarm, "cuda:0", dtype=torch.bfloat16, attn_cache_tokens=4096, inference_max_length=2048, peer_id="s")
st = stage.stage
x = torch.randn(8, 144, 4096, device="cuda", dtype=torch.bfloat16); g = torch.randn_like(x)
for _ in range(3): st.backward(x, g)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    st.backward(x, g); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40))
stage.shutdown()
PY
echo "backward profile exit=$?" | tee -a $S
head -60 gpurun_out/r2_backward_profile.log | cut -c1-200
