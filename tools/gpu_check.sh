#!/bin/bash
# First-contact GPU diagnostics: every group runs in its own process (a CUDA fault poisons the context)
# under its own timeout; logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() {  # name, timeout, command...
  local name=$1; shift; local t=$1; shift
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1
  echo "exit=$? $(tail -n 3 gpurun_out/$name.log | tr '\n' ' ')" | tee -a gpurun_out/summary.txt
}
run bandwidth 300 python -m pytest tests/test_ops_gpu.py -q -x -k "layernorm or patch_merge or errors"
run linear_f32 300 python -m pytest tests/test_ops_gpu.py -q -k "linear_fp32"
run attn_f32 300 python -m pytest tests/test_ops_gpu.py -q -k "window_attention_fp32 or no_qkv_bias"
run linear_tc 300 python -m pytest tests/test_ops_gpu.py -q -k "linear_bf16"
run attn_simt16 300 python -m pytest tests/test_ops_gpu.py -q -k "window_attention_bf16 and simt"
run attn_tc 300 python -m pytest tests/test_ops_gpu.py -q -k "window_attention_bf16 and not simt"
run stem 300 python -m pytest tests/test_ops_gpu.py -q -k "stem"
run diag_linear 200 python tools/gpu_diag.py linear
run diag_attn 200 python tools/gpu_diag.py attn
run diag_attnsimt 200 python tools/gpu_diag.py attnsimt
run backbone 600 python -m pytest tests/test_backbone_gpu.py -q
run smoke 300 python __graft_entry__.py smoke
grep -h "" gpurun_out/diag_*.log
cat gpurun_out/summary.txt
